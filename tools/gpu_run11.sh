#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "svd_conflict or svd_zero" > gpurun_out/t_svd0.log 2>&1; rc=$?; echo "svd0 rc=$rc"; tail -15 gpurun_out/t_svd0.log
if [ $rc -eq 0 ]; then
timeout 600 python -m pytest tests -m gpu -q -k "svd or smoke or unknown" > gpurun_out/t_svd.log 2>&1; echo "svd rc=$?"; tail -8 gpurun_out/t_svd.log
for C in 8 1; do for G in 8 16 32; do echo "== CLUSTER=$C LANES=$G"; SB2_DSGD_LANES=$G SB2_DSGD_CLUSTER=$C QUIET=1 timeout 120 python tools/profile_svd.py 2>&1 | grep -E "dsgd kernel|per stratum|group0|grid|rror"; done; done
fi
