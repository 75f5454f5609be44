#!/usr/bin/env bash
for B in 32 48 64 80 96 112; do echo "== BLOCKS=$B"; SB2_DSGD_BLOCKS=$B QUIET=1 timeout 120 python tools/profile_svd.py 2>&1 | tail -4; done
for B in 64 96; do echo "== BLOCKS=$B CLUSTER=8"; SB2_DSGD_CLUSTER=8 SB2_DSGD_BLOCKS=$B QUIET=1 timeout 120 python tools/profile_svd.py 2>&1 | tail -4; done
