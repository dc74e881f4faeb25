#!/bin/bash
# round-2 GPU call 18 (1 GPU): sanity of the last library change (GEMM workspace query): build(), smoke(), GEMM tests, headline bench line
O=gpurun_out/c18; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('build+smoke ok')" > $O/entry.log 2>&1; echo "rc=$?" >> $O/entry.log
timeout 600 python -m pytest tests -m gpu -q -k "gemm or attention" > $O/t_gemm_att.log 2>&1; echo "rc=$?" >> $O/t_gemm_att.log
timeout 900 python bench.py --no-supplementary > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -n 3 $O/entry.log; tail -n 4 $O/t_gemm_att.log; tail -c 500 $O/bench.json; tail -n 2 $O/bench.err
