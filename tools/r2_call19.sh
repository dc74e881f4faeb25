#!/bin/bash
# round-2 GPU call 19 (8 GPUs): the default bench line at N = 8 on the final code (NVML clock samples, supplementary with parity)
O=gpurun_out/c19; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench8.json 2> $O/bench8.err; echo "rc=$?" >> $O/bench8.err
tail -c 1200 $O/bench8.json; tail -n 3 $O/bench8.err
