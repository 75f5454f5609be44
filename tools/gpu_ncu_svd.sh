#!/usr/bin/env bash
mkdir -p gpurun_out
QUIET=1 timeout 300 python tools/profile_svd.py > gpurun_out/svd_plain.log 2>&1 && \
QUIET=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dsgd_svd -c 1 -f -o gpurun_out/svd_prof python tools/profile_svd.py > gpurun_out/svd_ncu_full.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/svd_ncu_full.log; ls -la gpurun_out/*.ncu-rep
