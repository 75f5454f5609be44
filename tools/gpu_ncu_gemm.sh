#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/profile_sim.py > gpurun_out/sim_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_u8_tc -s 2 -c 1 -f -o gpurun_out/gemm_prof python tools/profile_sim.py > gpurun_out/gemm_ncu_full.log 2>&1
echo "gemm ncu rc=$?"; tail -3 gpurun_out/gemm_ncu_full.log
