#!/bin/bash
# round-2 GPU call 15 (1 GPU): full GPU suite and the default bench line on the final code; smoke()
O=gpurun_out/c15; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
tail -n 6 $O/t_all.log; tail -n 3 $O/smoke.log; tail -c 1500 $O/bench.json; tail -2 $O/bench.err; tail -c 600 $O/bench_ref.json; tail -2 $O/bench_ref.err
