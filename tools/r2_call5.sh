#!/bin/bash
# round-2 GPU call 5: lookahead attention (fixed), pipelined gate+residual epilogue, hyvideo tile assembly, full suite, bench
O=gpurun_out/c5; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x -k "attention" > $O/t_att.log 2>&1; echo "rc=$?" >> $O/t_att.log
timeout 300 python tools/gpu_check_kernels.py attmodes > $O/attmodes.log 2>&1
timeout 300 python tools/gpu_check_kernels.py gemmpair > $O/gemmpair.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -n 12 $O/t_att.log; cat $O/attmodes.log | cut -c1-220; cut -c1-330 $O/gemmpair.log; tail -n 15 $O/t_all.log; tail -c 1500 $O/bench.json; tail -5 $O/bench.err
