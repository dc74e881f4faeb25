#!/bin/bash
# round-2 GPU call 8 (1 GPU): softmax instruction-mix variants of the attention kernel (scalar exponent math, deferred TMEM-store wait)
O=gpurun_out/c8; mkdir -p $O
timeout 400 python tools/gpu_check_kernels.py attmodes > $O/attmodes.log 2>&1
timeout 300 python tools/gpu_check_kernels.py atttrace > $O/atttrace.log 2>&1
cut -c1-230 $O/attmodes.log; cut -c1-330 $O/atttrace.log
