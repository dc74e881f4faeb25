#!/bin/bash
# round-2 GPU call 11 (1 GPU): tail split-K of the SM-pair gate+residual GEMM — tests, timings at the per-rank shapes
O=gpurun_out/c11; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "gemm" > $O/t_gemm.log 2>&1; echo "rc=$?" >> $O/t_gemm.log
timeout 300 python tools/gpu_check_kernels.py gemmsplit > $O/gemmsplit.log 2>&1
tail -n 8 $O/t_gemm.log; cut -c1-260 $O/gemmsplit.log
