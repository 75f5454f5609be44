"""Wall-clock of each phase of one SVD fit step and of a similarity build (host timers around synchronised calls)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from surprise_b200 import _native as nat  # noqa: E402

ts, uu, ii, rr, pu0, qi0, test = bench.load_workload()
prm = bench.sgd_params(nat, float(ts.global_mean))
lib = nat.lib()
nu, ni, n = ts.n_users, ts.n_items, len(rr)
d_u, d_i, d_r = nat.to_dev(uu, np.int32), nat.to_dev(ii, np.int32), nat.to_dev(rr, np.float64)
d_pu0, d_qi0 = nat.to_dev(pu0, np.float64), nat.to_dev(qi0, np.float64)
d_pu, d_qi = torch.empty_like(d_pu0), torch.empty_like(d_qi0)
d_bu, d_bi = nat.empty_dev((nu,), np.float64), nat.empty_dev((ni,), np.float64)
st = nat.stream()


def timed(label, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    acc.setdefault(label, []).append((time.perf_counter() - t0) * 1e3)


acc = {}
for it in range(4):
    plan = C.c_void_p()
    timed("create_dev", lambda: nat.check(lib.sb2_svd_plan_create_dev(nu, ni, n, nat.ptr(d_u), nat.ptr(d_i), nat.ptr(d_r),
                                                                     C.byref(prm), 0, None, None, st, C.byref(plan))), acc)
    timed("reset_dev", lambda: nat.check(lib.sb2_svd_plan_reset_dev(plan, nat.ptr(d_pu0), nat.ptr(d_qi0), None, st)), acc)
    timed("run20", lambda: nat.check(lib.sb2_svd_plan_run(plan, 20, st)), acc)
    timed("run1", lambda: nat.check(lib.sb2_svd_plan_run(plan, 1, st)), acc)
    timed("read_dev", lambda: nat.check(lib.sb2_svd_plan_read_dev(plan, nat.ptr(d_pu), nat.ptr(d_qi), nat.ptr(d_bu),
                                                                 nat.ptr(d_bi), None, st)), acc)
    timed("destroy", lambda: lib.sb2_svd_plan_destroy(plan), acc)
for k, v in acc.items():
    print("%-12s %s" % (k, " ".join("%8.3f" % x for x in v)))
