#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -q -x -k "svd_conflict or svd_zero or split_launches or lane_shapes" > gpurun_out/t_svd0.log 2>&1; rc=$?; echo "svd0 rc=$rc"; tail -3 gpurun_out/t_svd0.log
[ $rc -eq 0 ] || exit 1
QUIET=1 timeout 120 python tools/profile_svd.py 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -q -k "svd or smoke or skewed_svd" > gpurun_out/t_svd.log 2>&1; echo "svd rc=$?"; tail -2 gpurun_out/t_svd.log
