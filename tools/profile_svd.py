"""One SVD fit of the bench workload (for ncu)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from surprise_b200 import _native as nat  # noqa: E402

ts, uu, ii, rr, pu0, qi0, test = bench.load_workload()
prm = bench.sgd_params(nat, float(ts.global_mean))
lib = nat.lib()
plan = C.c_void_p()
nat.check(lib.sb2_svd_plan_create(ts.n_users, ts.n_items, len(rr), nat.hptr(uu), nat.hptr(ii), nat.hptr(rr), C.byref(prm),
                                  0, C.byref(plan)))
for it in range(2 if not os.environ.get('QUIET') else 1):
    nat.check(lib.sb2_svd_plan_reset(plan, nat.hptr(pu0), nat.hptr(qi0), None))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat.check(lib.sb2_svd_plan_run(plan, 20, nat.stream()))
    e1.record()
    torch.cuda.synchronize()
    print("dsgd kernel 20 epochs: %.3f ms -> %.1f M updates/s" % (e0.elapsed_time(e1), 2e7 / e0.elapsed_time(e1) / 1e3))
b, w = C.c_int(), C.c_int()
lib.sb2_svd_plan_grid(plan, C.byref(b), C.byref(w))
print("grid B=%d W=%d" % (b.value, w.value))
prof = np.zeros((b.value, 8), dtype=np.int64)
nat.check(lib.sb2_svd_plan_profile(plan, nat.hptr(prof)))
tot = prof[:, :4].sum(1)
m = prof.mean(0)
print("per-CTA cycles (mean over CTAs) wait %.0f load %.0f update %.0f writeback %.0f  total %.0f; per stratum: %s"
      % (*m[:4], tot.mean(), np.round(m[:4] / (20 * b.value))))
print("%.1f waves/stratum; update-phase cycles per wave %.0f; DSMEM hop cycles per stratum %.0f"
      % (m[6] / (20 * b.value), m[2] / max(m[6], 1), m[7] / (20 * b.value)))
if not os.environ.get("WAVE_PROF"):
    print("of the hop (thread 0): %.0f + %.0f cycles before waiting, %.0f waiting for the right neighbour's block  [SB2_DSGD_HOP=%s]"
          % (m[4] / (20 * b.value), m[5] / (20 * b.value), (m[7] - m[4] - m[5]) / (20 * b.value), os.environ.get("SB2_DSGD_HOP", "bulk")))
else:
    print("warp 0 per wave: update path %.0f cycles, barrier %.0f cycles" % (m[4] / max(m[6], 1), m[5] / max(m[6], 1)))
lib.sb2_svd_plan_destroy(plan)
