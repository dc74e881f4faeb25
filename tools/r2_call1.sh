#!/bin/bash
# round-2 GPU call 1: validate / measure the experimental kernels left by round 1 (each under its own short timeout)
O=gpurun_out/c1; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
timeout 60 tools/experimental/build/probe_2cta > $O/probe_2cta.log 2>&1; echo "probe rc=$?" >> $O/probe_2cta.log
YB_RUN_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q -x -k "experimental_attention_q64" > $O/t_att64.log 2>&1; echo "rc=$?" >> $O/t_att64.log
YB_RUN_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q -x -k "experimental_gemm_2cta" > $O/t_gemm2cta.log 2>&1; echo "rc=$?" >> $O/t_gemm2cta.log
timeout 200 python tools/gpu_check_kernels.py att64 > $O/att64.log 2>&1
YB_ATT64_HALVES=1 timeout 200 python tools/gpu_check_kernels.py att64 > $O/att64_halves.log 2>&1
timeout 200 python tools/gpu_check_kernels.py gemm2cta > $O/gemm2cta.log 2>&1
timeout 200 python tools/gpu_check_kernels.py atttrace atttune > $O/atttrace.log 2>&1
kill $SMI
tail -n 30 $O/*.log
