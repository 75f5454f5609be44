#!/usr/bin/env bash
# First-contact GPU run: tcgen05 GEMM self-test under a short timeout, then the rest of the parity suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "blackwell or gemm" > gpurun_out/t_gemm.log 2>&1
rc=$?
echo "gemm rc=$rc" >> gpurun_out/t_gemm.log
tail -15 gpurun_out/t_gemm.log
if [ $rc -eq 0 ]; then
  timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not gemm and not blackwell" > gpurun_out/t_rest.log 2>&1
else
  timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nmf or svd or baselines or mf_predict or knn_kernel or knn_long" > gpurun_out/t_rest.log 2>&1
fi
echo "rest rc=$?" >> gpurun_out/t_rest.log
tail -60 gpurun_out/t_rest.log
