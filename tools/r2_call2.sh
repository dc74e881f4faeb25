#!/bin/bash
# round-2 GPU call 2: new attention softmax schedules + glue kernels: parity, then timing; then the whole GPU suite
O=gpurun_out/c2; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
timeout 600 python -m pytest tests -m gpu -q -x -k "attention or elementwise" > $O/t_att.log 2>&1; echo "rc=$?" >> $O/t_att.log
timeout 300 python tools/gpu_check_kernels.py attmodes elementwise > $O/attmodes.log 2>&1
timeout 300 python tools/gpu_check_kernels.py gemmpair > $O/gemmpair.log 2>&1
timeout 300 python tools/gpu_check_kernels.py convpair > $O/convpair.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-vae > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
kill $SMI
tail -n 40 $O/t_att.log $O/attmodes.log $O/gemmpair.log $O/convpair.log; tail -n 15 $O/t_all.log; tail -c 3000 $O/bench.json; tail -5 $O/bench.err
