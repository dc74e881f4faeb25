#!/bin/bash
# round-2 GPU call 21 (1 GPU): headline bench line with the isolated CPU baseline, then the reference arm, back to back
O=gpurun_out/c21; mkdir -p $O
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
timeout 900 python bench.py --no-supplementary > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -c 700 $O/bench_ref.json; echo; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/c21/bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["cpu_baseline"])
PY
tail -n 2 $O/bench.err
