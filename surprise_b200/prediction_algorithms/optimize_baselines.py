"""baseline_als / baseline_sgd (reference: prediction_algorithms/optimize_baselines.pyx:14-84), run by
the order-preserving segmented kernels (sb2_baseline_als_dev / sb2_baseline_sgd_dev): bit-exact."""
import numpy as np

from .. import _native as nat


def baseline_als(self):
    ts = self.trainset
    opts = self.bsl_options
    n_epochs = int(opts.get("n_epochs", 10))
    reg_u = float(opts.get("reg_u", 15))
    reg_i = float(opts.get("reg_i", 10))
    up, ui, ur = ts.user_csr()
    ip, iu, ir = ts.item_csr()
    d = [nat.to_dev(up, np.int64), nat.to_dev(ui, np.int32), nat.to_dev(ur, np.float64),
         nat.to_dev(ip, np.int64), nat.to_dev(iu, np.int32), nat.to_dev(ir, np.float64)]
    bu = nat.empty_dev((ts.n_users,), np.float64)
    bi = nat.empty_dev((ts.n_items,), np.float64)
    rc = nat.lib().sb2_baseline_als_dev(ts.n_users, ts.n_items, *[nat.ptr(t) for t in d], float(ts.global_mean),
                                        n_epochs, reg_u, reg_i, nat.ptr(bu), nat.ptr(bi), nat.stream())
    nat.check(rc)
    return bu.cpu().numpy(), bi.cpu().numpy()


def baseline_sgd(self):
    ts = self.trainset
    opts = self.bsl_options
    n_epochs = int(opts.get("n_epochs", 20))
    reg = float(opts.get("reg", 0.02))
    lr = float(opts.get("learning_rate", 0.005))
    u, i, r = ts.coo()
    d = [nat.to_dev(u, np.int32), nat.to_dev(i, np.int32), nat.to_dev(r, np.float64)]
    bu = nat.empty_dev((ts.n_users,), np.float64)
    bi = nat.empty_dev((ts.n_items,), np.float64)
    rc = nat.lib().sb2_baseline_sgd_dev(ts.n_users, ts.n_items, len(r), *[nat.ptr(t) for t in d],
                                        float(ts.global_mean), n_epochs, reg, lr, nat.ptr(bu), nat.ptr(bi),
                                        nat.stream())
    nat.check(rc)
    return bu.cpu().numpy(), bi.cpu().numpy()
