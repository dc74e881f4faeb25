#!/bin/bash
# multi-GPU call: Ulysses parity for every transport at world N (and the forced KV split), then the bench at N with both p2p transports
N=${1:-2}
O=gpurun_out/mg$N; mkdir -p $O
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) "$@"; }
for T in p2p p2p_gemm nccl; do
  YB_SP_TRANSPORT=$T run tools/sp_parity.py > $O/parity_$T.log 2>&1; echo "rc=$?" >> $O/parity_$T.log
done
YB_SP_TRANSPORT=p2p_gemm YB_ATT_FORCE_SPLIT=2 run tools/sp_parity.py > $O/parity_p2p_gemm_split2.log 2>&1; echo "rc=$?" >> $O/parity_p2p_gemm_split2.log
run tools/vae_tile_parallel.py --full > $O/vae_tile_parallel.log 2>&1; echo "rc=$?" >> $O/vae_tile_parallel.log
run bench.py --gpus $N --steps 5 --warmup 3 --no-14b --sp-transport p2p > $O/bench_p2p.json 2> $O/bench_p2p.err; echo "rc=$?" >> $O/bench_p2p.err
run bench.py --gpus $N --steps 5 --warmup 3 --no-supplementary --sp-transport p2p_gemm > $O/bench_p2p_gemm.json 2> $O/bench_p2p_gemm.err; echo "rc=$?" >> $O/bench_p2p_gemm.err
tail -4 $O/vae_tile_parallel.log; grep -h "rc=\|MISMATCH" $O/parity_*.log | sort | uniq -c; tail -c 600 $O/bench_p2p.json; echo; tail -c 600 $O/bench_p2p_gemm.json; tail -3 $O/*.err
