#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/profile_svd.py > gpurun_out/svd_plain.log 2>&1; cat gpurun_out/svd_plain.log
BENCH_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; tail -8 gpurun_out/bench_dbg.err; cut -c1-400 gpurun_out/bench_dbg.json
BENCH_NO_SAMPLER=1 BENCH_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dbg2.json 2> gpurun_out/bench_dbg2.err; tail -4 gpurun_out/bench_dbg2.err; cut -c1-300 gpurun_out/bench_dbg2.json
timeout 600 python tools/profile_sim.py 8192 32768 4000000 pearson_baseline > gpurun_out/sim_pb_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/simpb_launches.csv python tools/profile_sim.py 8192 32768 4000000 pearson_baseline > gpurun_out/simpb_ncu.log 2>&1
cat gpurun_out/sim_pb_plain.log
timeout 300 python tools/profile_sim.py; 
