#!/bin/bash
# round-2 GPU call 20 (1 GPU): final state — build(), smoke(), the whole GPU suite, default bench line
O=gpurun_out/c20; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('build+smoke ok')" > $O/entry.log 2>&1; echo "rc=$?" >> $O/entry.log
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -n 3 $O/entry.log; tail -n 4 $O/t_all.log; tail -c 400 $O/bench.json; echo; tail -n 2 $O/bench.err
