#!/bin/bash
# 8-GPU call: Ulysses parity at world 8 (every transport, forced KV split) and world 4, tile-parallel VAE, bench at 8 with both p2p transports
O=gpurun_out/mg8; mkdir -p $O
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) "$@"; }
for T in p2p p2p_gemm nccl; do
  YB_SP_TRANSPORT=$T run 8 tools/sp_parity.py > $O/parity8_$T.log 2>&1; echo "rc=$?" >> $O/parity8_$T.log
done
YB_SP_TRANSPORT=p2p_gemm YB_ATT_FORCE_SPLIT=2 run 8 tools/sp_parity.py > $O/parity8_p2p_gemm_split2.log 2>&1; echo "rc=$?" >> $O/parity8_p2p_gemm_split2.log
YB_SP_TRANSPORT=p2p run 4 tools/sp_parity.py > $O/parity4_p2p.log 2>&1; echo "rc=$?" >> $O/parity4_p2p.log
YB_SP_TRANSPORT=p2p_gemm run 4 tools/sp_parity.py > $O/parity4_p2p_gemm.log 2>&1; echo "rc=$?" >> $O/parity4_p2p_gemm.log
run 8 bench.py --gpus 8 --steps 10 --warmup 3 --no-supplementary --sp-transport p2p > $O/bench8_p2p.json 2> $O/bench8_p2p.err; echo "rc=$?" >> $O/bench8_p2p.err
run 8 bench.py --gpus 8 --steps 10 --warmup 3 --sp-transport p2p_gemm > $O/bench8_p2p_gemm.json 2> $O/bench8_p2p_gemm.err; echo "rc=$?" >> $O/bench8_p2p_gemm.err
run 4 bench.py --gpus 4 --steps 10 --warmup 3 --no-supplementary --sp-transport p2p_gemm > $O/bench4_p2p_gemm.json 2> $O/bench4_p2p_gemm.err; echo "rc=$?" >> $O/bench4_p2p_gemm.err
run 8 tools/vae_tile_parallel.py --full > $O/vae_tile_parallel8.log 2>&1; echo "rc=$?" >> $O/vae_tile_parallel8.log
grep -h "rc=\|MISMATCH\|full 720p" $O/*.log | sort | uniq -c; for f in $O/bench*.json; do grep '^{' $f | tail -c 400; echo; done
