#!/usr/bin/env bash
mkdir -p gpurun_out
free -g | head -2
timeout 1500 python tools/bench_configs.py c4 > gpurun_out/cfg_c4.log 2>&1; echo "c4 rc=$?"; tail -3 gpurun_out/cfg_c4.log
timeout 1500 python tools/bench_configs.py c3 > gpurun_out/cfg_c3.log 2>&1; echo "c3 rc=$?"; tail -3 gpurun_out/cfg_c3.log
timeout 2400 python tools/bench_configs.py c5 > gpurun_out/cfg_c5.log 2>&1; echo "c5 rc=$?"; tail -3 gpurun_out/cfg_c5.log
