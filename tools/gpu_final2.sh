#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python - <<'PY' 2>&1 | tail -3
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch, surprise_b200 as sb
from surprise_b200 import synth
for skew in (False, True):
    d = synth.ratings(6040, 3706, 1_000_000, seed=0, skew=skew)
    u, i, r = d["train"]
    ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
    a = sb.SVD(random_state=0); a.fit(ts)
    torch.cuda.synchronize(); t0 = time.perf_counter(); a.fit(ts); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tu, ti, tr = d["test"]; est, _ = a._estimate_batch(tu, ti)
    print("SVD f=100 20 epochs ml-1M shape skew=%s: max item raters %d, fit %.1f ms (host API), heldout rmse %.4f" % (
        skew, np.bincount(i).max(), dt * 1e3, np.sqrt(np.mean((np.clip(est, 1, 5) - tr) ** 2))))
PY
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
echo "bench rc=$?"; cat gpurun_out/bench_n1.json
