#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python tools/bench_predict.py 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
echo "bench rc=$?"; cat gpurun_out/bench_n1.json
