#!/bin/bash
# round-2 GPU call 12 (1 GPU): two-threads-per-row attention variants with cycle traces
O=gpurun_out/c12; mkdir -p $O
timeout 300 python tools/gpu_check_kernels.py attmodes > $O/attmodes.log 2>&1; echo "rc=$?" >> $O/attmodes.log
timeout 300 python tools/gpu_check_kernels.py atttrace > $O/atttrace.log 2>&1; echo "rc=$?" >> $O/atttrace.log
cut -c1-220 $O/attmodes.log; grep -v "raw rows" $O/atttrace.log | cut -c1-330
