#!/usr/bin/env bash
# Round-2 ncu evidence, one gpurun call: every capture only after the same command exited 0 without ncu.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile_r2.sh'
# Reports land in gpurun_out/r2_*; tools/summarize_profiles_r2.py turns them into profiles/r2_summary.md.
# ONLY="svd knn" limits the captures to the named sections (bench svd gemm simrows knn nmf).
set -u
mkdir -p gpurun_out
want() { [ -z "${ONLY:-}" ] || [[ " $ONLY " == *" $1 "* ]]; }
FULL="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 3"
want bench && $B > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/r2_bench_ncu.log 2>&1
echo "bench launches rc=$?"
want svd && QUIET=1 python tools/profile_svd.py > gpurun_out/r2_svd_plain.log 2>&1 && \
  QUIET=1 $FULL -k regex:dsgd_svd -c 1 -o gpurun_out/r2_svd_prof python tools/profile_svd.py > gpurun_out/r2_svd_ncu.log 2>&1
echo "svd rc=$?"
want gemm && SB2_SIM_PATH=digit python tools/profile_sim.py > gpurun_out/r2_gemm_plain.log 2>&1 && \
  SB2_SIM_PATH=digit $FULL -k regex:gemm_u8_tc -c 1 -o gpurun_out/r2_gemm_prof python tools/profile_sim.py > gpurun_out/r2_gemm_ncu.log 2>&1
echo "gemm rc=$?"
want simrows && SB2_SIM_PATH=general python tools/profile_sim.py 8192 32768 4000000 pearson_baseline > gpurun_out/r2_simrows_plain.log 2>&1 && \
  SB2_SIM_PATH=general $FULL -k regex:sim_rows -c 1 -o gpurun_out/r2_simrows_prof python tools/profile_sim.py 8192 32768 4000000 pearson_baseline > gpurun_out/r2_simrows_ncu.log 2>&1
echo "sim_rows rc=$?"
want knn && python tools/profile_knn.py > gpurun_out/r2_knn_plain.log 2>&1 && \
  $FULL -k regex:knn_predict -c 1 -o gpurun_out/r2_knn_prof python tools/profile_knn.py > gpurun_out/r2_knn_ncu.log 2>&1
echo "knn rc=$?"
want nmf && python tools/profile_nmf.py scale=0.6 epochs=3 > gpurun_out/r2_nmf_plain.log 2>&1 && \
  $FULL -k regex:nmf_pass_fused -c 2 -o gpurun_out/r2_nmf_prof python tools/profile_nmf.py scale=0.6 epochs=3 > gpurun_out/r2_nmf_ncu.log 2>&1
echo "nmf rc=$?"
tail -2 gpurun_out/r2_*_plain.log
ls -la gpurun_out/r2_*
