#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -x -k "svd_conflict or svd_zero or svdpp" > gpurun_out/t_svd0.log 2>&1; rc=$?; echo "svd0 rc=$rc"; tail -3 gpurun_out/t_svd0.log
[ $rc -eq 0 ] || exit 1
timeout 300 python -m pytest tests -m gpu -q -k "svd or smoke or unknown" > gpurun_out/t_svd.log 2>&1; echo "svd rc=$?"; tail -3 gpurun_out/t_svd.log
for CH in 0 60 30 15; do echo "== CHUNKS=$CH"; if [ $CH -eq 0 ]; then timeout 600 python tools/bench_configs.py c4 2>&1 | tail -1; else SB2_SVDPP_CHUNKS=$CH timeout 600 python tools/bench_configs.py c4 2>&1 | tail -1; fi; done
