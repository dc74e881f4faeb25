#!/bin/bash
# round-2 GPU call 13 (1 GPU): MUFU / FMA issue rate while the tensor pipe of the same SM is busy
O=gpurun_out/c13; mkdir -p $O
timeout 120 tools/microbench/mufu_under_mma > $O/mufu_under_mma.log 2>&1; echo "rc=$?" >> $O/mufu_under_mma.log
cat $O/mufu_under_mma.log | cut -c1-200
