#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "blackwell or gemm" > gpurun_out/t_gemm.log 2>&1; rc=$?; echo "gemm rc=$rc"; tail -6 gpurun_out/t_gemm.log
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests -m gpu -q -k "simil or knn or smoke or toy" > gpurun_out/t_sim.log 2>&1; echo "sim rc=$?"; tail -6 gpurun_out/t_sim.log
timeout 300 python tools/profile_sim.py; timeout 300 python tools/profile_sim.py 8192 32768 4000000 pearson_baseline
fi
