// Batched estimate kernels: factor models (gather + dot) and k-NN (gather + stable top-k select +
// ordered weighted sum).  These replace the per-pair Python estimate() calls that AlgoBase.test makes
// (algo_base.py:191-218 -> matrix_factorization.pyx:269-299 / :506-522 / :737-761, knns.py:99-123 /
// :274-309).
#include "common.cuh"

namespace sb2 {

// ------------------------------------------------------------------------------------------------
// Factor models.  One warp per (u, i) pair; lanes stride the factor dimension, partial dots are
// combined with a shuffle tree (the reference uses np.dot, whose summation order is BLAS-defined, so
// the contract here is fp64 accuracy, not bit equality).
// ------------------------------------------------------------------------------------------------
__global__ void mf_predict_kernel(int64_t n_pairs, const int32_t* __restrict__ u, const int32_t* __restrict__ i, int f,
                                  int biased, double mu, const double* __restrict__ pu, const double* __restrict__ qi,
                                  const double* __restrict__ bu, const double* __restrict__ bi,
                                  const double* __restrict__ yj, const int64_t* __restrict__ u_ptr,
                                  const int32_t* __restrict__ ui_idx, double* __restrict__ est,
                                  uint8_t* __restrict__ impossible) {
    const int lane = threadIdx.x & 31;
    const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (k >= n_pairs) return;
    const int32_t uu = u[k], ii = i[k];
    const bool ku = uu >= 0, ki = ii >= 0;
    double e = 0.0;
    uint8_t imp = 0;
    if (biased) {
        e = mu;
        if (ku) e += bu[uu];
        if (ki) e += bi[ii];
    } else if (!(ku && ki)) {
        imp = 1;  // PredictionImpossible('User and item are unkown.')
    }
    if (ku && ki) {
        const double* p = pu + (size_t)uu * f;
        const double* q = qi + (size_t)ii * f;
        double part = 0.0;
        if (yj) {
            const int64_t b = u_ptr[uu], en = u_ptr[uu + 1];
            const double sq = sqrt((double)(en - b));
            for (int j = lane; j < f; j += 32) {
                double s = 0.0;
                for (int64_t a = b; a < en; ++a) s += yj[(size_t)ui_idx[a] * f + j];
                part += q[j] * (p[j] + s / sq);
            }
        } else {
            for (int j = lane; j < f; j += 32) part += q[j] * p[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        e = biased ? e + part : part;
    }
    if (lane == 0) {
        est[k] = e;
        impossible[k] = imp;
    }
}

int mf_predict_dev(int64_t n_pairs, const int32_t* u, const int32_t* i, int f, int biased, double mu, const double* pu,
                   const double* qi, const double* bu, const double* bi, const double* yj, const int64_t* u_ptr,
                   const int32_t* ui_idx, double* est, uint8_t* impossible, cudaStream_t st) {
    if (n_pairs <= 0) return SB2_OK;
    const int threads = 256;
    mf_predict_kernel<<<(unsigned)ceil_div(n_pairs * 32, threads), threads, 0, st>>>(
        n_pairs, u, i, f, biased, mu, pu, qi, bu, bi, yj, u_ptr, ui_idx, est, impossible);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

// ------------------------------------------------------------------------------------------------
// k-NN.  One warp per (x, y) pair.
//   neighbors = [(sim[x, x2], r) for (x2, r) in yr[y]]; heapq.nlargest(k, key=sim)
// heapq.nlargest == sorted(reverse=True)[:k]: descending by sim, ties keep list order.  The warp
// repeats k rounds of "smallest element after the previously selected one in the total order
// (sim desc, position asc)" with a shuffle arg-max; the gathered sims live in shared memory when the
// list fits, otherwise they are re-gathered from the sim row.  The weighted sums are accumulated by
// every lane identically, in selection order, in round-to-nearest fp64 without contraction -- the
// same sequence of operations as the reference, hence the same bits.
// ------------------------------------------------------------------------------------------------
constexpr int KNN_WARPS = 8;
constexpr int KNN_CACHE = 768;  // sims cached per warp (doubles)

__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_predict_kernel(int64_t n_pairs, const int32_t* __restrict__ x, const int32_t* __restrict__ y, int64_t n_x,
                   const double* __restrict__ sim, int64_t sim_ld, const int64_t* __restrict__ y_ptr,
                   const int32_t* __restrict__ x_idx, const double* __restrict__ r, int k, int min_k, int mode,
                   double mu, const double* __restrict__ bx, const double* __restrict__ by, double* __restrict__ est,
                   int32_t* __restrict__ actual_k, uint8_t* __restrict__ impossible) {
    __shared__ double cache[KNN_WARPS][KNN_CACHE];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t p = blockIdx.x * (int64_t)KNN_WARPS + w; p < n_pairs; p += (int64_t)gridDim.x * KNN_WARPS) {
        const int32_t xx = x[p], yy = y[p];
        const bool kx = xx >= 0, ky = yy >= 0;
        double e = 0.0;
        int ak = -1;
        uint8_t imp = 0;
        if (mode == 1 || mode == 2) {
            e = mu;
            // est += bu[u] first, then bi[i] (knns.py:276-280); mode 1: x is the user, mode 2: y is
            if (mode == 1) { if (kx) e = __dadd_rn(e, bx[xx]); if (ky) e = __dadd_rn(e, by[yy]); }
            else { if (ky) e = __dadd_rn(e, by[yy]); if (kx) e = __dadd_rn(e, bx[xx]); }
        }
        if (!(kx && ky)) {
            if (mode == 0 || mode >= 3) imp = 1;  // KNNBasic / WithMeans / WithZScore raise PredictionImpossible
        } else {
            if (mode >= 3) e = bx[xx];  // means[x] (knns.py:187, :382)
            const int64_t b = y_ptr[yy], len = y_ptr[yy + 1] - b;
            const double* srow = sim + (size_t)xx * (size_t)sim_ld;
            const bool cached = len <= KNN_CACHE;
            if (cached) {
                for (int64_t a = lane; a < len; a += 32) cache[w][a] = srow[x_idx[b + a]];
                __syncwarp();
            }
            double last_s = 0.0;
            int64_t last_pos = -1;
            bool first = true;
            double sum_sim = 0.0, sum_r = 0.0;
            ak = 0;
            const int64_t rounds = len < (int64_t)k ? len : (int64_t)k;
            for (int64_t t = 0; t < rounds; ++t) {
                // best candidate strictly after (last_s, last_pos) in (sim desc, pos asc) order
                double bs = 0.0;
                int64_t bp = -1;
                for (int64_t a = lane; a < len; a += 32) {
                    const double s = cached ? cache[w][a] : srow[x_idx[b + a]];
                    const bool after = first || s < last_s || (s == last_s && a > last_pos);
                    if (after && (bp < 0 || s > bs)) { bs = s; bp = a; }  // ascending a: first max wins
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double os = __shfl_xor_sync(0xFFFFFFFFu, bs, o);
                    const int64_t op = __shfl_xor_sync(0xFFFFFFFFu, bp, o);
                    if (op >= 0 && (bp < 0 || os > bs || (os == bs && op < bp))) { bs = os; bp = op; }
                }
                if (bp < 0) break;  // only NaNs left
                first = false;
                last_s = bs;
                last_pos = bp;
                if (!(bs > 0.0)) break;  // everything that follows is <= 0 and contributes nothing
                const double rr = r[b + bp];
                sum_sim = __dadd_rn(sum_sim, bs);
                if (mode == 3) {
                    sum_r = __dadd_rn(sum_r, __dmul_rn(bs, __dsub_rn(rr, bx[x_idx[b + bp]])));
                } else if (mode == 4) {
                    const int32_t nb = x_idx[b + bp];
                    sum_r = __dadd_rn(sum_r, __ddiv_rn(__dmul_rn(bs, __dsub_rn(rr, bx[nb])), by[nb]));
                } else if (mode != 0) {
                    const double nb_bsl = __dadd_rn(__dadd_rn(mu, bx[x_idx[b + bp]]), by[yy]);
                    sum_r = __dadd_rn(sum_r, __dmul_rn(bs, __dsub_rn(rr, nb_bsl)));
                } else {
                    sum_r = __dadd_rn(sum_r, __dmul_rn(bs, rr));
                }
                ++ak;
            }
            if (mode != 0) {
                if (ak < min_k) sum_r = 0.0;
                if (ak > 0) {  // ZeroDivisionError swallowed otherwise
                    if (mode == 4) e = __dadd_rn(e, __dmul_rn(__ddiv_rn(sum_r, sum_sim), by[xx]));
                    else e = __dadd_rn(e, __ddiv_rn(sum_r, sum_sim));
                }
            } else {
                if (ak < min_k) imp = 1;       // PredictionImpossible('Not enough neighbors.')
                else if (ak == 0) imp = 2;     // min_k <= 0: the reference divides 0 / 0
                else e = __ddiv_rn(sum_r, sum_sim);
            }
            __syncwarp();
        }
        if (lane == 0) {
            est[p] = e;
            actual_k[p] = ak;
            impossible[p] = imp;
        }
    }
}

int knn_predict_dev(int64_t n_pairs, const int32_t* x, const int32_t* y, int64_t n_x, const double* sim,
                    int64_t sim_ld, const int64_t* y_ptr, const int32_t* x_idx, const double* r, int k, int min_k,
                    int mode, double mu, const double* bx, const double* by, double* est, int32_t* actual_k,
                    uint8_t* impossible, cudaStream_t st) {
    if (n_pairs <= 0) return SB2_OK;
    if (mode < 0 || mode > 4 || (mode != 0 && !bx) || ((mode == 1 || mode == 2 || mode == 4) && !by)) {
        set_error("knn_predict: invalid mode / missing baselines");
        return SB2_ERR_INVALID;
    }
    int64_t blocks = ceil_div(n_pairs, KNN_WARPS);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    knn_predict_kernel<<<(unsigned)blocks, KNN_WARPS * 32, 0, st>>>(n_pairs, x, y, n_x, sim, sim_ld, y_ptr, x_idx, r, k,
                                                                    min_k, mode, mu, bx, by, est, actual_k, impossible);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

// ------------------------------------------------------------------------------------------------
// SlopeOne.estimate (slope_one.pyx:82-97).  One warp per (u, i):
//   Ri = [j for (j, _) in ur[u] if freq[i, j] > 0];  est = user_mean[u] + sum(dev[i, j] for j in Ri) / len(Ri)
// Python's sum() adds left to right in ur[u] order; lanes gather 32 entries of row i at a time and the
// warp folds them in lane order (every lane performs the same additions), so the bits match.
// ------------------------------------------------------------------------------------------------
__global__ void slope_one_predict_kernel(int64_t n_pairs, const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                         int64_t n_items, const int64_t* __restrict__ freq,
                                         const double* __restrict__ dev, const int64_t* __restrict__ u_ptr,
                                         const int32_t* __restrict__ i_idx, const double* __restrict__ user_mean,
                                         double* __restrict__ est, uint8_t* __restrict__ impossible) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; p < n_pairs; p += nwarps) {
        const int32_t uu = u[p], ii = i[p];
        double e = 0.0;
        uint8_t imp = 0;
        if (uu < 0 || ii < 0) {
            imp = 1;  // PredictionImpossible('User and/or item is unkown.')
        } else {
            const int64_t b = u_ptr[uu], len = u_ptr[uu + 1] - b;
            const int64_t* frow = freq + (size_t)ii * (size_t)n_items;
            const double* drow = dev + (size_t)ii * (size_t)n_items;
            double sum = 0.0;
            int64_t cnt = 0;
            for (int64_t a0 = 0; a0 < len; a0 += 32) {
                const int64_t a = a0 + lane;
                bool rel = false;
                double d = 0.0;
                if (a < len) {
                    const int32_t j = i_idx[b + a];
                    rel = frow[j] > 0;
                    if (rel) d = drow[j];
                }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, rel);
                cnt += __popc(m);
                for (unsigned rest = m; rest; rest &= rest - 1) {
                    const int src = __ffs(rest) - 1;
                    sum = __dadd_rn(sum, __shfl_sync(0xFFFFFFFFu, d, src));
                }
            }
            e = user_mean[uu];
            if (cnt > 0) e = __dadd_rn(e, __ddiv_rn(sum, (double)cnt));
        }
        if (lane == 0) {
            est[p] = e;
            impossible[p] = imp;
        }
    }
}

int slope_one_predict_dev(int64_t n_pairs, const int32_t* u, const int32_t* i, int64_t n_items, const int64_t* freq,
                          const double* dev, const int64_t* u_ptr, const int32_t* i_idx, const double* user_mean,
                          double* est, uint8_t* impossible, cudaStream_t st) {
    if (n_pairs <= 0) return SB2_OK;
    int64_t blocks = ceil_div(n_pairs * 32, 256);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    slope_one_predict_kernel<<<(unsigned)blocks, 256, 0, st>>>(n_pairs, u, i, n_items, freq, dev, u_ptr, i_idx, user_mean,
                                                              est, impossible);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

// ------------------------------------------------------------------------------------------------
// AlgoBase.get_neighbors (algo_base.py:303-334): others = [(x, sim[iid, x]) for x != iid];
// others.sort(key = sim, reverse = True)  -- stable, so ties keep ascending x -- and the first k ids.
// One block per requested row; k rounds of "best element strictly after the previous pick in
// (sim desc, x asc) order", block-wide arg-max through warp shuffles + one shared-memory stage.
// ------------------------------------------------------------------------------------------------
constexpr int TOPK_THREADS = 1024;

__global__ void __launch_bounds__(TOPK_THREADS)
row_topk_kernel(const double* __restrict__ sim, int64_t sim_ld, int64_t n_x, const int32_t* __restrict__ rows, int k,
                int32_t* __restrict__ out) {
    __shared__ double ws[32];
    __shared__ int wi[32];
    __shared__ double pick_s;
    __shared__ int pick_i;
    const int32_t row = rows[blockIdx.x];
    int32_t* o = out + (size_t)blockIdx.x * k;
    if (row < 0 || row >= n_x) {
        for (int t = threadIdx.x; t < k; t += blockDim.x) o[t] = -1;
        return;
    }
    const double* srow = sim + (size_t)row * (size_t)sim_ld;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double last_s = 0.0;
    int last_i = -1;
    for (int t = 0; t < k; ++t) {
        double bs = 0.0;
        int bi = -1;
        for (int64_t x = threadIdx.x; x < n_x; x += blockDim.x) {
            if (x == row) continue;
            const double s = srow[x];
            const bool after = last_i < 0 || s < last_s || (s == last_s && x > last_i);
            if (after && (bi < 0 || s > bs)) { bs = s; bi = (int)x; }  // ascending x: first max wins
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xFFFFFFFFu, bs, off);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, off);
            if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
        }
        if (lane == 0) { ws[w] = bs; wi[w] = bi; }
        __syncthreads();
        if (w == 0) {
            bs = lane < (blockDim.x >> 5) ? ws[lane] : 0.0;
            bi = lane < (blockDim.x >> 5) ? wi[lane] : -1;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double os = __shfl_xor_sync(0xFFFFFFFFu, bs, off);
                const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, off);
                if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
            }
            if (lane == 0) { pick_s = bs; pick_i = bi; o[t] = bi; }
        }
        __syncthreads();
        last_s = pick_s;
        last_i = pick_i;
        if (last_i < 0) {  // fewer than k other entries (or only NaNs left): pad with -1
            for (int r = t + 1 + threadIdx.x; r < k; r += blockDim.x) o[r] = -1;
            return;
        }
    }
}

int get_neighbors_dev(int64_t n_x, const double* sim, int64_t sim_ld, int64_t n_rows, const int32_t* rows, int k,
                      int32_t* out, cudaStream_t st) {
    if (n_rows <= 0 || k <= 0) return SB2_OK;
    if (!sim || !rows || !out || sim_ld < n_x) {
        set_error("get_neighbors: invalid argument");
        return SB2_ERR_INVALID;
    }
    row_topk_kernel<<<(unsigned)n_rows, TOPK_THREADS, 0, st>>>(sim, sim_ld, n_x, rows, k, out);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

}  // namespace sb2
