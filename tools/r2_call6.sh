#!/bin/bash
# round-2 GPU call 6 (1 GPU): VAE encoders + strided convs, reproducibility of the hyvideo decode, full suite, bench, ncu evidence
O=gpurun_out/c6; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "stride2 or encode" > $O/t_enc.log 2>&1; echo "rc=$?" >> $O/t_enc.log
timeout 300 python tools/vae_determinism.py > $O/vae_determinism.log 2>&1; echo "rc=$?" >> $O/vae_determinism.log
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
N=$O/ncu; mkdir -p $N
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:.*yb::.*" -c 1400 --csv --log-file $N/launches.csv python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline --no-supplementary > $N/launch_run.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:.*yb::(gemm_pair_kernel|attention_kernel|qk_norm_rope|ln_modulate_warp).*" -s 372 -c 13 -o $N/prof_block python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline --no-supplementary > $N/full_run.log 2>&1
ncu -i $N/prof_block.ncu-rep --page raw --csv > $N/prof_block_raw.csv 2>/dev/null
tail -n 8 $O/t_enc.log; cat $O/vae_determinism.log | tail -4; tail -n 12 $O/t_all.log; tail -c 2500 $O/bench.json; tail -3 $O/bench.err; ls -la $N; tail -3 $N/launch_run.log $N/full_run.log
