#!/usr/bin/env bash
# kernel ms per 20-epoch fit at config 2 for a list of env settings: bash tools/quick_svd.sh "A=1 B=2" "C=3" ...
run() {
  echo "== $*"
  env $* python bench.py --no-secondary --no-cpu-baseline --steps 10 --warmup 3 2>&1 | python -c '
import sys, json
for ln in sys.stdin:
    if ln.startswith("{"):
        d = json.loads(ln)
        print("kernel_ms %.3f ms_per_step %.3f frac %.3f rmse %.5f e2e_ms %.2f" % (d["roofline"]["kernel_ms_per_launch"], d["ms_per_step"], d["roofline"]["frac"], d["heldout_rmse"], d["e2e"]["ms_per_step"]))
    elif "rror" in ln: print(ln.strip()[:300])
'
}
if [ $# -eq 0 ]; then run X=1; else for v in "$@"; do run $v; done; fi
