#!/usr/bin/env bash
# round-end refresh: tests, bench, launch list, kernel captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "bench ncu rc=$?"
QUIET=1 timeout 300 python tools/profile_svd.py > gpurun_out/svd_plain.log 2>&1 && \
QUIET=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dsgd_svd -c 1 -f -o gpurun_out/svd_prof python tools/profile_svd.py > gpurun_out/svd_ncu_full.log 2>&1
echo "svd ncu rc=$?"; cat gpurun_out/svd_plain.log
timeout 1200 python tools/bench_configs.py c5 > gpurun_out/cfg_c5.log 2>&1; tail -1 gpurun_out/cfg_c5.log
