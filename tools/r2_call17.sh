#!/bin/bash
# round-2 GPU call 17 (4 GPUs): bench at N = 4 with and without the GEMM tail split-K (the per-rank shape it was built for)
O=gpurun_out/c17; mkdir -p $O
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) "$@"; }
run bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench4.json 2> $O/bench4.err; echo "rc=$?" >> $O/bench4.err
run bench.py --gpus 4 --steps 5 --warmup 3 --no-supplementary --gemm-split-k 1 > $O/bench4_nosplit.json 2> $O/bench4_nosplit.err; echo "rc=$?" >> $O/bench4_nosplit.err
run bench.py --gpus 4 --steps 5 --warmup 3 --no-supplementary > $O/bench4_b.json 2> $O/bench4_b.err; echo "rc=$?" >> $O/bench4_b.err
tail -c 700 $O/bench4.json; echo; tail -c 400 $O/bench4_nosplit.json; echo; tail -c 400 $O/bench4_b.json; tail -2 $O/*.err
