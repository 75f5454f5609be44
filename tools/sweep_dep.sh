#!/usr/bin/env bash
# sweep of the dependency-driven cell mode of dsgd_svd_kernel at config 2 (kernel ms per 20-epoch fit, RMSE)
# usage (GPU box): bash tools/sweep_dep.sh > gpurun_out/sweep_dep.log
run() {
  echo "== $*"
  env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 10 --warmup 3 2>&1 | python -c '
import sys, json
for ln in sys.stdin:
    if ln.startswith("{"):
        d = json.loads(ln)
        print("kernel_ms %.3f ms_per_step %.3f frac %.3f rmse %.5f e2e_ms %.2f" % (d["roofline"]["kernel_ms_per_launch"], d["ms_per_step"], d["roofline"]["frac"], d["heldout_rmse"], d["e2e"]["ms_per_step"]))
    elif "rror" in ln: print(ln.strip()[:300])
'
}
for sync in 0 1 2; do
  run SB2_DSGD_DEP=1 SB2_DSGD_DEP_SYNC=$sync
  run SB2_DSGD_DEP=1 SB2_DSGD_DEP_SYNC=$sync SB2_DSGD_GROUPS=16
done
