#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_all.log
SB2_SIM_TIMING=1 timeout 900 python tools/bench_configs.py c3 > gpurun_out/cfg_c3.log 2>&1; echo "c3 rc=$?"; grep "sb2" gpurun_out/cfg_c3.log | tail -3; tail -1 gpurun_out/cfg_c3.log
