#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "nmf or smoke" > gpurun_out/t_nmf.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_nmf.log
timeout 2400 python tools/bench_configs.py c5 scale=0.5 > gpurun_out/cfg_c5h.log 2>&1; echo "c5 rc=$?"; tail -2 gpurun_out/cfg_c5h.log
