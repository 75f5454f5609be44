#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "nmf" > gpurun_out/t_nmf.log 2>&1; rc=$?; echo "nmf rc=$rc"; tail -3 gpurun_out/t_nmf.log
[ $rc -eq 0 ] || exit 1
python tools/profile_nmf.py scale=0.3 epochs=3 2>&1 | tail -2
timeout 900 python tools/bench_configs.py c5 scale=0.5 2>&1 | tail -1
