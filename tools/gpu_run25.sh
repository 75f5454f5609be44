#!/usr/bin/env bash
# 2-GPU: upper-shard test on rank-0 GPU, then multi-GPU checks and scaling configs
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "upper_shards or sim_build_sharded or nmf" 2>&1 | tail -2
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py 2>&1 | grep -v Warning | tail -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/scale_configs.py c3 c5 2>&1 | grep -v Warning | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n$N.json; cat gpurun_out/bench_n$N.json
