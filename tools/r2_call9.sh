#!/bin/bash
# round-2 GPU call 9 (1 GPU): parked-warp interference microbenchmark; attention variants incl. the split S issue
O=gpurun_out/c9; mkdir -p $O
timeout 300 tools/microbench/instr_rate > $O/instr_rate.log 2>&1; echo "rc=$?" >> $O/instr_rate.log
timeout 600 python tools/gpu_check_kernels.py attmodes > $O/attmodes.log 2>&1
timeout 300 python tools/gpu_check_kernels.py atttrace > $O/atttrace.log 2>&1
tail -n 6 $O/instr_rate.log | cut -c1-200; cut -c1-200 $O/attmodes.log; grep -v "raw rows" $O/atttrace.log | cut -c1-330
