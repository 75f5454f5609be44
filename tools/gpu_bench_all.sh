#!/usr/bin/env bash
# bench.py at N GPUs of this box (N = number of visible GPUs), both arms; writes gpurun_out/r2_bench_n<N>.json
N=$(python -c "import torch; print(torch.cuda.device_count())")
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
fi
tail -c 600 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
for ln in open("gpurun_out/r2_bench_n$N.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("N=%d value %.4g ms_per_step %.3f frac %.3f e2e_ms %.2f rmse %.5f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["heldout_rmse"]))
        print({k: v for k, v in (d["secondary"] or {}).items() if not isinstance(v, (dict, str))})
PY
