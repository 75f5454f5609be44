#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -3 gpurun_out/bench_c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c.json')); print({k:d[k] for k in ('value','ms_per_step','heldout_rmse','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])"
timeout 900 python tools/bench_configs.py c4 > gpurun_out/cfg_c4b.log 2>&1; tail -1 gpurun_out/cfg_c4b.log
