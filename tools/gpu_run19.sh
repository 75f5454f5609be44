#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x -k "svdpp" > gpurun_out/t_svd0.log 2>&1; rc=$?; echo "svdpp rc=$rc"; tail -3 gpurun_out/t_svd0.log
for SC in 0.5 1.0; do for CH in 1 2 4 8 16 32; do echo "== scale=$SC CHUNKS=$CH"; SB2_SVDPP_CHUNKS=$CH timeout 600 python tools/bench_configs.py c4 scale=$SC 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['fit_s'], d['heldout_rmse'], d['svd_f20_heldout_rmse'])"; done; done
