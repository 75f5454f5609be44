#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "svd_conflict or u1_svdpp or toy_similarities or synthetic_nmf" > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/sanitize.log | head -20; tail -5 gpurun_out/sanitize.log
