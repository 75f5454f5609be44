"""CPU prototype: candidate parallelisable SVD++ update schedules vs the reference's per-rating schedule
(oracle).  Used to choose the relaxation implemented in csrc/sgd.cu (DESIGN.md 'SVD++')."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import oracle
from surprise_b200 import synth
import surprise_b200 as sb


def rmse(pu, qi, yj, bu, bi, mu, ptr, idx, tu, ti, tr):
    est, _ = oracle.mf_estimate(tu, ti, True, mu, pu, qi, bu, bi, yj, ptr, idx)
    return float(np.sqrt(np.mean((np.clip(est, 1, 5) - tr) ** 2)))


def variant_A(u, i, r, ptr, idx, pu, qi, yj, n_epochs, mu, lr, reg, order_seed=None, chunks=1, expint=False):
    """per-epoch: z_u from y; pass over ratings with own-effect z update and per-user gradient accumulation;
    y applied item-side in `chunks` instalments per epoch."""
    nu, f = pu.shape; ni = qi.shape[0]
    bu = np.zeros(nu); bi = np.zeros(ni)
    n_u = np.diff(ptr).astype(float); sq = np.sqrt(n_u)
    n = len(r)
    users_of_rating = np.repeat(np.arange(nu), np.diff(ptr))
    rng = np.random.RandomState(0 if order_seed is None else order_seed)
    for ep in range(n_epochs):
        order = np.arange(n) if order_seed is None else rng.permutation(n)
        z = np.zeros((nu, f)); np.add.at(z, users_of_rating, yj[idx]); z /= sq[:, None]
        bounds = np.linspace(0, n, chunks + 1).astype(int)
        for c in range(chunks):
            g = np.zeros((nu, f)); cnt = np.zeros(nu)
            for k in order[bounds[c]:bounds[c + 1]]:
                uu, ii = u[k], i[k]
                p, q = pu[uu].copy(), qi[ii].copy()
                err = r[k] - (mu + bu[uu] + bi[ii] + q @ (p + z[uu]))
                bu[uu] += lr * (err - reg * bu[uu]); bi[ii] += lr * (err - reg * bi[ii])
                pu[uu] = p + lr * (err * q - reg * p)
                qi[ii] = q + lr * (err * (p + z[uu]) - reg * q)
                g[uu] += err * q / sq[uu]; cnt[uu] += 1
                z[uu] += lr * (err * q - reg * z[uu])
            # item side: decay by the number of (rating of u) events seen by each y_j, then add the gradients
            cj = np.zeros(ni); np.add.at(cj, idx, cnt[users_of_rating])
            gj = np.zeros((ni, f)); np.add.at(gj, idx, g[users_of_rating])
            dec = (1 - lr * reg) ** cj
            yj *= dec[:, None]
            if expint:  # gradients spread uniformly over the c_j decays: sum_k d^k = (1 - d^c) / (1 - d)
                fac = np.where(cj > 0, (1 - dec) / np.maximum(cj * lr * reg, 1e-300), 1.0)
                yj += lr * gj * fac[:, None]
            else:
                yj += lr * gj
            if c + 1 < chunks:
                z = np.zeros((nu, f)); np.add.at(z, users_of_rating, yj[idx]); z /= sq[:, None]
    return pu, qi, yj, bu, bi


def run(name, u, i, r, nu, ni, tu, ti, tr, f=20, n_epochs=20):
    ts = sb.Trainset.from_coo(u, i, r, nu, ni)
    uu, ii, rr = ts.coo(); ptr, idx, _ = ts.user_csr(); mu = float(ts.global_mean)
    def init():
        rng = np.random.RandomState(0)
        return rng.normal(0, .1, (nu, f)), rng.normal(0, .1, (ni, f)), rng.normal(0, .1, (ni, f))
    pu, qi, yj = init()
    t = time.time()
    ref = oracle.svdpp_sgd(uu, ii, rr, ptr, idx, pu, qi, yj, n_epochs, mu, *([.007] * 5), *([.02] * 5))
    print(name, "reference schedule rmse %.5f (%.1fs)" % (rmse(*ref[:3], ref[3], ref[4], mu, ptr, idx, tu, ti, tr), time.time() - t))
    pu, qi, _ = init()
    s = oracle.svd_sgd(uu, ii, rr, pu, qi, n_epochs, True, mu, *([.007] * 4), *([.02] * 4))
    print(name, "  plain SVD same lr      rmse %.5f" % rmse(s[0], s[1], np.zeros_like(s[1]), s[2], s[3], mu, ptr, idx, tu, ti, tr))
    for chunks in (1, 4, 16):
        for seed in (None, 1):
            pu, qi, yj = init()
            a = variant_A(uu, ii, rr, ptr, idx, pu, qi, yj, n_epochs, mu, .007, .02, seed, chunks)
            print(name, "  A chunks=%2d order=%s rmse %.5f" % (chunks, "file" if seed is None else "perm", rmse(*a[:3], a[3], a[4], mu, ptr, idx, tu, ti, tr)), flush=True)
        pu, qi, yj = init()
        a = variant_A(uu, ii, rr, ptr, idx, pu, qi, yj, n_epochs, mu, .007, .02, 1, chunks, True)
        print(name, "  A chunks=%2d order=perm EXPINT rmse %.5f" % (chunks, rmse(*a[:3], a[3], a[4], mu, ptr, idx, tu, ti, tr)), flush=True)


if __name__ == "__main__":
    from conftest import inner_pairs
    from surprise_b200.model_selection import PredefinedKFold
    G = os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden")
    data = sb.Dataset.load_from_folds([(os.path.join(G, "u1_ml100k_train"), os.path.join(G, "u1_ml100k_test"))], sb.Reader("ml-100k"))
    ts, te = next(PredefinedKFold().split(data))
    u, i, r = ts.coo(); iu, ii = inner_pairs(ts, te); tr = np.array([t[2] for t in te])
    if "dense" not in sys.argv:
        run("u1", u, i, r, ts.n_users, ts.n_items, iu, ii, tr)
        d = synth.ratings(600, 400, 30000, seed=2)
        run("synth", *d["train"], d["n_users"], d["n_items"], *d["test"], n_epochs=10)
    if len(sys.argv) > 1 and sys.argv[1] == "dense":
        d = synth.ratings(3000, 300, 200000, seed=5)
        run("dense", *d["train"], d["n_users"], d["n_items"], *d["test"], n_epochs=10)
