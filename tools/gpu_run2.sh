#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log; tail -5 gpurun_out/t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 python tools/profile_svd.py > gpurun_out/svd_plain.log 2>&1; echo "svd rc=$?"; cat gpurun_out/svd_plain.log
timeout 600 python tools/profile_sim.py > gpurun_out/sim_plain.log 2>&1; echo "sim rc=$?"; cat gpurun_out/sim_plain.log
timeout 600 python tools/profile_sim.py 8192 32768 4000000 pearson_baseline > gpurun_out/sim_pb_plain.log 2>&1; cat gpurun_out/sim_pb_plain.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
