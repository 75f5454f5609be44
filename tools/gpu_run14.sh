#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "blackwell or gemm" > gpurun_out/t_gemm.log 2>&1; rc=$?; echo "gemm rc=$rc"; tail -12 gpurun_out/t_gemm.log
[ $rc -eq 0 ] || exit 1
timeout 300 python -m pytest tests -m gpu -q -k "simil or knn or smoke or toy" > gpurun_out/t_sim.log 2>&1; echo "sim rc=$?"; tail -6 gpurun_out/t_sim.log
SB2_SIM_TIMING=1 timeout 120 python tools/profile_sim.py 2>&1 | tail -3
SB2_SIM_TIMING=1 timeout 120 python tools/profile_sim.py 8192 32768 4000000 pearson_baseline 2>&1 | tail -2
SB2_GEMM_CG2=0 SB2_SIM_TIMING=1 timeout 120 python tools/profile_sim.py 2>&1 | tail -2
