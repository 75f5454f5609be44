#!/usr/bin/env bash
mkdir -p gpurun_out
for W in 32 16 8; do for B in 55 148; do
  echo "== GROUPS=$W BLOCKS=$B"; SB2_DSGD_GROUPS=$W SB2_DSGD_BLOCKS=$B QUIET=1 timeout 120 python tools/profile_svd.py 2>&1 | grep -E "dsgd kernel|per stratum|group0"
done; done 2>&1 | tee gpurun_out/sweep3.log
