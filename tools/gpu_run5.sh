#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "svd or smoke" > gpurun_out/t_svd.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/t_svd.log
timeout 300 python tools/profile_svd.py > gpurun_out/svd_plain.log 2>&1; cat gpurun_out/svd_plain.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err; cut -c1-260 gpurun_out/bench_b.json; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print({k:d[k] for k in ('value','ms_per_step','heldout_rmse','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])"
