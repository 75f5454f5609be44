#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/time_phases.py > gpurun_out/phases.log 2>&1; cat gpurun_out/phases.log
timeout 300 python tools/profile_sim.py > gpurun_out/sim_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/sim_launches.csv python tools/profile_sim.py > gpurun_out/sim_ncu.log 2>&1
echo "sim ncu rc=$?"
timeout 300 python tools/profile_svd.py > gpurun_out/svd_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/svd_launches.csv python tools/profile_svd.py > gpurun_out/svd_ncu.log 2>&1
echo "svd ncu rc=$?"
