#!/bin/bash
# round-2 GPU call 4: attention schedules (F2FP / ALU pack x round-1 / lookahead), conv auto-pair, full suite
O=gpurun_out/c4; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "attention or conv3d" > $O/t_kernels.log 2>&1; echo "rc=$?" >> $O/t_kernels.log
timeout 300 python tools/gpu_check_kernels.py attmodes atttrace > $O/attmodes.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
tail -n 30 $O/t_kernels.log $O/attmodes.log; tail -n 15 $O/t_all.log
