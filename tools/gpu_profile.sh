#!/usr/bin/env bash
# One ncu capture per gpurun call (pool rule), each only after the same command exited 0 without ncu.
#   gpurun -- 'bash tools/gpu_profile.sh svd|gemm|nmf|knn'
# Reports land in gpurun_out/*.ncu-rep; tools/summarize_profiles.py turns them into profiles/r1_summary.md.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
case "${1:-svd}" in
  svd)  QUIET=1 python tools/profile_svd.py > gpurun_out/svd_plain.log 2>&1 && QUIET=1 $NCU -k regex:dsgd_svd -c 1 -o gpurun_out/svd_prof python tools/profile_svd.py > gpurun_out/svd_ncu_full.log 2>&1 ;;
  gemm) python tools/profile_sim.py > gpurun_out/sim_plain.log 2>&1 && $NCU -k regex:gemm_u8_tc -c 1 -o gpurun_out/gemm_prof python tools/profile_sim.py > gpurun_out/gemm_ncu_full.log 2>&1 ;;
  nmf)  python tools/profile_nmf.py scale=0.6 epochs=3 > gpurun_out/nmf_plain.log 2>&1 && $NCU -k regex:nmf_pass_fused -c 2 -o gpurun_out/nmf_prof python tools/profile_nmf.py scale=0.6 epochs=3 > gpurun_out/nmf_ncu_full.log 2>&1 ;;
  knn)  python tools/bench_predict.py > gpurun_out/predict_plain.log 2>&1 && $NCU -k regex:knn_predict -c 1 -o gpurun_out/knn_prof python tools/bench_predict.py > gpurun_out/knn_ncu_full.log 2>&1 ;;
esac
echo "rc=$?"
