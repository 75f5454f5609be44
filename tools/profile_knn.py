"""One k-NN estimate batch at the ml-1M shape (for ncu): 500k (item, user) pairs, k = 40, KNNBasic mode on an msd matrix."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import surprise_b200 as sb  # noqa: E402
from surprise_b200 import _native as nat, similarities as sims, synth  # noqa: E402

d = synth.shaped("ml-1m", seed=0)
u, i, r = d["train"]
ts = sb.Trainset.from_coo(u, i, r, d["n_users"], d["n_items"])
yr = ts.user_csr()
sim = sims.build_device("msd", ts.n_items, yr, 1)
rng = np.random.RandomState(0)
n = 500_000
x = nat.to_dev(rng.randint(0, ts.n_items, n), np.int32); y = nat.to_dev(rng.randint(0, ts.n_users, n), np.int32)
d_ptr, d_idx, d_val = nat.to_dev(yr[0], np.int64), nat.to_dev(yr[1], np.int32), nat.to_dev(yr[2], np.float64)
est = nat.empty_dev((n,), np.float64); ak = nat.empty_dev((n,), np.int32); imp = nat.empty_dev((n,), np.uint8)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat.check(nat.lib().sb2_knn_predict_dev(n, nat.ptr(x), nat.ptr(y), ts.n_items, nat.ptr(sim), ts.n_items, nat.ptr(d_ptr),
                                            nat.ptr(d_idx), nat.ptr(d_val), 40, 1, 0, 0.0, None, None, nat.ptr(est), nat.ptr(ak),
                                            nat.ptr(imp), nat.stream()))
    e1.record()
    torch.cuda.synchronize()
    print("knn_predict_kernel %d pairs: %.3f ms = %.3g pairs/s" % (n, e0.elapsed_time(e1), n / e0.elapsed_time(e1) * 1e3))
