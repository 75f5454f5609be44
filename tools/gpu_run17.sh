#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests -m gpu -q -x -k "svd_conflict or svd_zero" > gpurun_out/t_svd0.log 2>&1; rc=$?; echo "svd0 rc=$rc"; tail -3 gpurun_out/t_svd0.log
[ $rc -eq 0 ] || exit 1
QUIET=1 timeout 90 python tools/profile_svd.py 2>&1 | grep -E "dsgd kernel|per stratum|waves|grid|rror"
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','heldout_rmse','gpu_launches','secondary')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout 900 python tools/bench_configs.py c4 > gpurun_out/cfg_c4.log 2>&1; tail -1 gpurun_out/cfg_c4.log
