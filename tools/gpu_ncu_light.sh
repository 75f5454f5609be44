#!/usr/bin/env bash
mkdir -p gpurun_out
QUIET=1 timeout 300 python tools/profile_svd.py > gpurun_out/svd_plain.log 2>&1 && \
QUIET=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/svd_light.csv python tools/profile_svd.py > gpurun_out/svd_ncu_light.log 2>&1
echo "ncu light rc=$?"; tail -5 gpurun_out/svd_ncu_light.log; grep -c dsgd_svd gpurun_out/svd_light.csv; grep dsgd_svd gpurun_out/svd_light.csv | cut -c1-60,200-400 | head -3
