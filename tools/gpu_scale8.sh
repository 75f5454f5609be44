#!/usr/bin/env bash
mkdir -p gpurun_out
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/scale_configs.py c3 2>&1 | grep -v Warning | tail -2
