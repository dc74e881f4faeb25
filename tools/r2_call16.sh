#!/bin/bash
# round-2 GPU call 16 (2 GPUs): the Ulysses pytest cases at world 2, tile-parallel hyvideo decode (reproducible GroupNorm statistics), bench at N = 2
O=gpurun_out/c16; mkdir -p $O
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) "$@"; }
timeout 900 python -m pytest tests -m gpu -q -k "ulysses" > $O/t_ulysses.log 2>&1; echo "rc=$?" >> $O/t_ulysses.log
run tools/vae_tile_parallel.py --full > $O/vae_tile_parallel.log 2>&1; echo "rc=$?" >> $O/vae_tile_parallel.log
run bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench2.json 2> $O/bench2.err; echo "rc=$?" >> $O/bench2.err
tail -n 5 $O/t_ulysses.log; grep -h "full 720p\|rc=" $O/vae_tile_parallel.log | tail -3; tail -c 900 $O/bench2.json; tail -2 $O/bench2.err
