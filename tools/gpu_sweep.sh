#!/usr/bin/env bash
for C in 16 8; do echo "== CLUSTER=$C"; SB2_DSGD_CLUSTER=$C QUIET=1 timeout 90 python tools/profile_svd.py 2>&1 | grep -E "dsgd kernel|per stratum|waves|grid|rror"; done
