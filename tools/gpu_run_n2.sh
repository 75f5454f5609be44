#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "nmf" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_check.py 2>&1 | tail -3
