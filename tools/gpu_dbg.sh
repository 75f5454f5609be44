#!/usr/bin/env bash
for D in 0 1 2 3; do echo "== DEBUG=$D"; SB2_DSGD_DEBUG=$D QUIET=1 timeout 90 python tools/profile_svd.py 2>&1 | grep -E "dsgd kernel|per stratum|waves|rror"; done
