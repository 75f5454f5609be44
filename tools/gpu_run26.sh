#!/usr/bin/env bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests -m gpu -q -x -k "sim or knn" 2>&1 | tail -2
timeout 1200 python tools/scale_configs.py c3 c5 2>&1 | grep -v Warning | tail -1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/scale_configs.py c3 c5 2>&1 | grep -v Warning | tail -1
