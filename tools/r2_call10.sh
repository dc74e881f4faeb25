#!/bin/bash
# round-2 GPU call 10 (1 GPU): two-threads-per-row attention kernel (16 softmax warps) against the product kernel
O=gpurun_out/c10; mkdir -p $O
timeout 300 python tools/gpu_check_kernels.py attmodes > $O/attmodes.log 2>&1; echo "rc=$?" >> $O/attmodes.log
cut -c1-230 $O/attmodes.log
