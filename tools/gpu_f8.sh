#!/usr/bin/env bash
SB2_GEMM_CG2=0 SB2_SIM_TIMING=1 timeout 120 python tools/profile_sim.py 2>&1 | grep "sb2" | tail -1
SB2_GEMM_CG2=0 SB2_GEMM_TIMING_F8=1 SB2_SIM_TIMING=1 timeout 120 python tools/profile_sim.py 2>&1 | grep "sb2" | tail -1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
