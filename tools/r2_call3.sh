#!/bin/bash
# round-2 GPU call 3: pipelined attention schedule, SM-pair GEMM / conv, fused QKV+all-to-all (one-GPU emulation), full suite, bench
O=gpurun_out/c3; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
timeout 600 python -m pytest tests -m gpu -q -x -k "attention or gemm or fused_qkv or conv3d" > $O/t_kernels.log 2>&1; echo "rc=$?" >> $O/t_kernels.log
timeout 300 python tools/gpu_check_kernels.py attmodes atttrace > $O/attmodes.log 2>&1
timeout 300 python tools/gpu_check_kernels.py gemmpair > $O/gemmpair.log 2>&1
timeout 300 python tools/gpu_check_kernels.py convpair > $O/convpair.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
kill $SMI
tail -n 30 $O/t_kernels.log $O/attmodes.log $O/gemmpair.log $O/convpair.log; tail -n 15 $O/t_all.log; tail -c 2000 $O/bench.json; tail -5 $O/bench.err
