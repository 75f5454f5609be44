#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests -m gpu -q -x -k "svd_conflict or svd_zero" > gpurun_out/t_svd0.log 2>&1; rc=$?; echo "svd0 rc=$rc"; tail -3 gpurun_out/t_svd0.log
[ $rc -eq 0 ] || exit 1
QUIET=1 timeout 90 python tools/profile_svd.py 2>&1 | grep -E "dsgd kernel|per stratum|waves|grid|rror"; rc=${PIPESTATUS[0]}; echo "profile rc=$rc"
[ $rc -eq 0 ] || exit 1
timeout 300 python -m pytest tests -m gpu -q -x -k "svd or smoke or unknown" > gpurun_out/t_svd.log 2>&1; echo "svd rc=$?"; tail -4 gpurun_out/t_svd.log
