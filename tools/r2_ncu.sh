#!/bin/bash
# ncu evidence for round 2: (1) launch list of one denoise step (device time per launch: compare SHARES), (2) --set full captures of the
# kernels of one block (SM-pair GEMM plain / gate+residual, attention, fused q+k norm/RoPE, LayerNorm+modulate)
O=gpurun_out/ncu; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 440 -c 450 --csv --log-file $O/launches_step.csv python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline --no-supplementary > $O/launch_run.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel|attention_kernel|qk_norm_rope|ln_modulate_warp" -s 360 -c 13 -o $O/prof_block python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline --no-supplementary > $O/full_run.log 2>&1
ls -la $O; tail -3 $O/launch_run.log $O/full_run.log
