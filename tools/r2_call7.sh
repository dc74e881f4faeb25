#!/bin/bash
# round-2 GPU call 7 (1 GPU): instruction-rate microbenchmark for the softmax inner loop, attention comparators
O=gpurun_out/c7; mkdir -p $O
timeout 300 tools/microbench/instr_rate > $O/instr_rate.log 2>&1; echo "rc=$?" >> $O/instr_rate.log
timeout 600 python tools/attention_comparators.py > $O/attention_comparators.json 2> $O/attention_comparators.err; echo "rc=$?" >> $O/attention_comparators.err
cat $O/instr_rate.log; tail -c 3000 $O/attention_comparators.json; tail -n 3 $O/attention_comparators.err
