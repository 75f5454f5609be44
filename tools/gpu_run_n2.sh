#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 rc=$?"; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1b.json 2>/dev/null; cut -c1-200 gpurun_out/bench_n1b.json
