"""A few NMF epochs at a Netflix-shaped scale (for ncu / timing of nmf_pass_fused_kernel).
usage: python tools/profile_nmf.py [scale=0.3] [epochs=3]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import surprise_b200 as sb  # noqa: E402
from surprise_b200 import _native as nat, synth  # noqa: E402

scale = float(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("scale=")), 0.3))
epochs = int(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("epochs=")), 3))
d = synth.shaped("netflix", seed=0, scale=scale)
u, i, r = d["train"]
ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
uu, ii, rr = ts.coo()
f = 15
rng = np.random.RandomState(0)
pu0 = rng.uniform(0, 1, (ts.n_users, f)); qi0 = rng.uniform(0, 1, (ts.n_items, f))
d_u, d_i, d_r = nat.to_dev(uu, np.int32), nat.to_dev(ii, np.int32), nat.to_dev(rr, np.float64)
d_bu, d_bi = nat.empty_dev((ts.n_users,), np.float64), nat.empty_dev((ts.n_items,), np.float64)
d_pu, d_qi = nat.to_dev(pu0, np.float64), nat.to_dev(qi0, np.float64)
prm = nat.NmfParams(n_factors=f, n_epochs=epochs, biased=0, reserved=0, global_mean=0.0, reg_pu=.06, reg_qi=.06,
                    reg_bu=.02, reg_bi=.02, lr_bu=.005, lr_bi=.005)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nat.check(nat.lib().sb2_nmf_fit_dev(ts.n_users, ts.n_items, len(rr), nat.ptr(d_u), nat.ptr(d_i), nat.ptr(d_r), C.byref(prm),
                                        nat.ptr(d_pu), nat.ptr(d_qi), nat.ptr(d_bu), nat.ptr(d_bi), nat.stream()))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%d users x %d items, %d ratings, %d epochs (+plan): %.1f ms" % (ts.n_users, ts.n_items, len(rr), epochs, dt * 1e3))
