// Order-preserving segmented kernels: baselines (ALS / SGD) and NMF.
//
// Bit-exactness discipline: the reference is strict IEEE fp64 without FMA contraction and with a
// fixed summation order, so every accumulation below is a *sequential* chain of __dadd_rn /
// __dmul_rn / __ddiv_rn in the reference's order; parallelism is across segments (users, items) and
// across factors, never inside one ordered sum.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <vector>

#include "common.cuh"

struct sb2_nmf_plan;

namespace sb2 {

enum { SEG_ZERODIV = 0 };

// =================================================================================================
// baseline_als  (optimize_baselines.pyx:14-54)
//   bi[i] = (sum_{(u,r) in ir[i]} r - mu - bu[u]) / (reg_i + |ir[i]|)   then the same for users.
// One thread per segment walks its list in order.
// =================================================================================================
__global__ void als_pass_kernel(int64_t n_seg, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                const double* __restrict__ r, const double* other, double* mine, double mu,
                                double reg, int* status, int64_t seg_begin = 0) {
    const int64_t s = seg_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const int64_t b = ptr[s], e = ptr[s + 1];
    double dev = 0.0;
    for (int64_t a = b; a < e; ++a) dev = __dadd_rn(dev, __dsub_rn(__dsub_rn(r[a], mu), other[idx[a]]));
    const double den = __dadd_rn(reg, (double)(e - b));
    if (den == 0.0) {
        atomicExch(&status[SEG_ZERODIV], 1);
        return;
    }
    mine[s] = __ddiv_rn(dev, den);
}

int baseline_als_dev(int64_t n_users, int64_t n_items, const int64_t* u_ptr, const int32_t* ui_idx, const double* u_r,
                     const int64_t* i_ptr, const int32_t* iu_idx, const double* i_r, double mu, int n_epochs,
                     double reg_u, double reg_i, double* bu, double* bi, cudaStream_t st) {
    DevBuf status;
    SB2_TRY(status.alloc(sizeof(int), st));
    SB2_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), st));
    SB2_CUDA(cudaMemsetAsync(bu, 0, (size_t)n_users * sizeof(double), st));
    SB2_CUDA(cudaMemsetAsync(bi, 0, (size_t)n_items * sizeof(double), st));
    for (int ep = 0; ep < n_epochs; ++ep) {
        if (n_items > 0) {
            als_pass_kernel<<<(unsigned)ceil_div(n_items, 128), 128, 0, st>>>(n_items, i_ptr, iu_idx, i_r, bu, bi, mu,
                                                                              reg_i, status.as<int>());
            SB2_LAUNCH_CHECK();
        }
        if (n_users > 0) {
            als_pass_kernel<<<(unsigned)ceil_div(n_users, 128), 128, 0, st>>>(n_users, u_ptr, ui_idx, u_r, bi, bu, mu,
                                                                              reg_u, status.as<int>());
            SB2_LAUNCH_CHECK();
        }
    }
    int h = 0;
    SB2_CUDA(cudaMemcpyAsync(&h, status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB2_CUDA(cudaStreamSynchronize(st));
    if (h) {
        set_error("float division");
        return SB2_ERR_ZERO_DIVISION;
    }
    return SB2_OK;
}

// One half-epoch of baseline_als restricted to the segments [seg_begin, seg_end) -- the unit of the multi-rank ALS
// (SURVEY.md 8e "baselines (ALS)": every ordered sum is evaluated by exactly one rank, the new biases are all-gathered
// between the two passes of an epoch; surprise_b200/distributed.py baseline_als_sharded).  status_dev: one int,
// set to 1 where the reference divides by zero.
int baseline_als_pass_dev(int64_t seg_begin, int64_t seg_end, const int64_t* ptr, const int32_t* idx, const double* r,
                          const double* other, double* mine, double mu, double reg, int* status_dev, cudaStream_t st) {
    if (seg_end <= seg_begin) return SB2_OK;
    als_pass_kernel<<<(unsigned)ceil_div(seg_end - seg_begin, 128), 128, 0, st>>>(seg_end, ptr, idx, r, other, mine, mu, reg,
                                                                                  status_dev, seg_begin);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

// baseline_sgd (optimize_baselines.pyx:57-84): a sequential recursion through bu/bi over
// all_ratings() -- executed as such by one thread (bit-exact; the reference's non-default method).
__global__ void baseline_sgd_kernel(int64_t n, const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                    const double* __restrict__ r, double mu, int n_epochs, double reg, double lr,
                                    double* bu, double* bi) {
    if (blockIdx.x || threadIdx.x) return;
    for (int ep = 0; ep < n_epochs; ++ep)
        for (int64_t k = 0; k < n; ++k) {
            const int32_t uu = u[k], ii = i[k];
            const double b_u = bu[uu], b_i = bi[ii];
            const double err = __dsub_rn(r[k], __dadd_rn(__dadd_rn(mu, b_u), b_i));
            bu[uu] = __dadd_rn(b_u, __dmul_rn(lr, __dsub_rn(err, __dmul_rn(reg, b_u))));
            bi[ii] = __dadd_rn(b_i, __dmul_rn(lr, __dsub_rn(err, __dmul_rn(reg, b_i))));
        }
}

int baseline_sgd_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i, const double* r,
                     double mu, int n_epochs, double reg, double lr, double* bu, double* bi, cudaStream_t st) {
    SB2_CUDA(cudaMemsetAsync(bu, 0, (size_t)n_users * sizeof(double), st));
    SB2_CUDA(cudaMemsetAsync(bi, 0, (size_t)n_items * sizeof(double), st));
    baseline_sgd_kernel<<<1, 32, 0, st>>>(n, u, i, r, mu, n_epochs, reg, lr, bu, bi);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

// =================================================================================================
// NMF  (matrix_factorization.pyx:684-730)
//
// Per epoch the reference makes one pass over all_ratings() accumulating, for every rating k=(u,i,r):
//   est_k = mu + bu[u] + bi[i] + sum_f qi[i,f]*pu[u,f]            (f ascending, mul then add)
//   user_num[u,f] += qi[i,f]*r      user_denom[u,f] += qi[i,f]*est_k
//   item_num[i,f] += pu[u,f]*r      item_denom[i,f] += pu[u,f]*est_k
// then  pu[u,f] *= user_num / (user_denom + |ur[u]|*reg_pu*pu[u,f])  and the same for items, both
// from the OLD pu / qi (the accumulators were built before either update).
//
// Each accumulator element is an ordered sum over one user's (one item's) ratings in all_ratings()
// order.  est_k is computed once per epoch by one thread per rating (factor-ordered mul/add chain, so it
// carries the reference's bits), stored in all_ratings() order for the user pass and gathered into item
// order for the item pass; a group of G = 2^ceil(log2 f) (<= 32) lanes then owns one segment, lane l owns
// factors l, l+G, ...  The biased model makes bu/bi a sequential recursion over all ratings (:707-709); that
// part runs on one thread between the dot kernel and the passes.
// =================================================================================================
// One pass: segments are users (fixed row = pu[u], gathered rows = qi[i_a]) or items (the transpose).
// seg_ptr / other_idx / r_seg / est_seg: CSR of the side in all_ratings() order.  A group of G lanes owns one
// segment, lane l owns factors l, l+G, ...; the entry loop is unrolled so that the gathers of the next
// entries are in flight while the two ordered fp64 accumulation chains (num, den) advance.
template <int G>
__global__ void nmf_pass_kernel(int64_t n_seg, int f, const int64_t* __restrict__ seg_ptr,
                                const int32_t* __restrict__ other_idx, const double* __restrict__ r_seg,
                                const double* __restrict__ est_seg, const double* __restrict__ mine_old,
                                const double* __restrict__ other_old, double* __restrict__ mine_new, double reg,
                                int* status) {
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x % G;
    const int64_t seg = blockIdx.x * (int64_t)gpb + threadIdx.x / G;
    if (seg >= n_seg) return;
    const int64_t b = seg_ptr[seg], e = seg_ptr[seg + 1];
    const double* myrow = mine_old + (size_t)seg * f;
    constexpr int MAXS = 8;  // factors per lane: f <= 32*8
    constexpr int UN = 4;
    double num[MAXS], den[MAXS];
#pragma unroll
    for (int s = 0; s < MAXS; ++s) num[s] = den[s] = 0.0;
    if (f <= G) {
        // common case (one factor per lane): keep everything in registers
        const bool act = gl < f;
        double n0 = 0.0, d0 = 0.0;
        int64_t a = b;
        for (; a + UN <= e; a += UN) {
            double o[UN], rr[UN], ee[UN];
#pragma unroll
            for (int t = 0; t < UN; ++t) {
                const int32_t oi = other_idx[a + t];
                rr[t] = r_seg[a + t];
                ee[t] = est_seg[a + t];
                o[t] = act ? other_old[(size_t)oi * f + gl] : 0.0;
            }
#pragma unroll
            for (int t = 0; t < UN; ++t) {
                n0 = __dadd_rn(n0, __dmul_rn(o[t], rr[t]));
                d0 = __dadd_rn(d0, __dmul_rn(o[t], ee[t]));
            }
        }
        for (; a < e; ++a) {
            const double o = act ? other_old[(size_t)other_idx[a] * f + gl] : 0.0;
            n0 = __dadd_rn(n0, __dmul_rn(o, r_seg[a]));
            d0 = __dadd_rn(d0, __dmul_rn(o, est_seg[a]));
        }
        num[0] = n0;
        den[0] = d0;
    } else {
        for (int64_t a = b; a < e; ++a) {
            const double* orow = other_old + (size_t)other_idx[a] * f;
            const double r = r_seg[a], est = est_seg[a];
#pragma unroll
            for (int s = 0; s < MAXS; ++s) {
                const int j = gl + s * G;
                if (j < f) {
                    const double o = orow[j];
                    num[s] = __dadd_rn(num[s], __dmul_rn(o, r));
                    den[s] = __dadd_rn(den[s], __dmul_rn(o, est));
                }
            }
        }
    }
    const double nr = (double)(e - b);
#pragma unroll
    for (int s = 0; s < MAXS; ++s) {
        const int j = gl + s * G;
        if (j < f) {
            const double p = myrow[j];
            const double d = __dadd_rn(den[s], __dmul_rn(__dmul_rn(nr, reg), p));
            if (d == 0.0) atomicExch(&status[SEG_ZERODIV], 1);
            mine_new[(size_t)seg * f + j] = __dmul_rn(p, __ddiv_rn(num[s], d));
        }
    }
}

// Unbiased model, f <= G: the same pass with est = dot(q_i, p_u) recomputed on the fly instead of read from a
// materialised array (which costs one extra gather of both factor rows per rating plus an 8-byte scatter into item
// order, all through L2 / HBM).  The G lanes of a group take G consecutive entries of their segment:
//   load   (lane = factor): the G gathered rows of the other side go to a shared-memory tile, one coalesced
//           120-byte read per row -- the only global traffic of the entry;
//   step A (lane = entry): est = ((0 + q[0] p[0]) + q[1] p[1]) + ... in factor order -- the multiply / add chain of
//           nmf_dot_kernel, i.e. the reference's `dot` loop (matrix_factorization.pyx:699-702); the lane walks its
//           row in the tile against the fixed row (broadcast reads); (r, est) of the entry go to shared memory;
//   step B (lane = factor): for each of the G entries in order, (r, est) is one broadcast read, the row element
//           comes from the tile, and the two ordered accumulation chains advance.
// The kernel is bound by the L1TEX data pipe (shared-memory wavefronts; shuffles ride the same pipe), see
// profiles/r1_summary.md "NMF": values that every lane of the group needs are therefore broadcast reads, not
// 64-bit shuffles (two SHFL each).  A variant that kept the G row elements in registers and tiled only the
// products was slower (80 registers, half the resident warps).  Same bits as the three-kernel path (est is the
// same chain; the accumulation order is unchanged).
constexpr int NMF_THREADS = 128;

template <int G>
__global__ void __launch_bounds__(NMF_THREADS)
nmf_pass_fused_kernel(int64_t n_seg, int f, const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ other_idx,
                      const double* __restrict__ r_seg, const double* __restrict__ mine_old,
                      const double* __restrict__ other_old, double* __restrict__ mine_new, double reg, int* status) {
    __shared__ double tile_s[NMF_THREADS / G][G][G + 1];
    __shared__ double2 re_s[NMF_THREADS / G][G];
    __shared__ double mrow_s[NMF_THREADS / G][G];
    const int gl = threadIdx.x % G, g = threadIdx.x / G;
    const int64_t seg = blockIdx.x * (int64_t)(NMF_THREADS / G) + g;
    if (seg >= n_seg) return;  // whole groups leave; the shuffles below name only the lanes of one group
    const unsigned gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    const int64_t b = seg_ptr[seg], e = seg_ptr[seg + 1];
    const bool act = gl < f;
    const double m_l = act ? mine_old[(size_t)seg * f + gl] : 0.0;
    double (*tile)[G + 1] = tile_s[g];
    double2* re = re_s[g];
    double* mrow = mrow_s[g];
    mrow[gl] = m_l;
    double n0 = 0.0, d0 = 0.0;
    for (int64_t a = b; a < e; a += G) {
        const int64_t mine = a + gl;
        const bool has = mine < e;
        const int32_t oi = has ? other_idx[mine] : 0;
        const double rr = has ? r_seg[mine] : 0.0;
        const int cnt = (int)((e - a) < (int64_t)G ? (e - a) : (int64_t)G);
        for (int t = 0; t < cnt; ++t) {
            const int32_t oit = __shfl_sync(gmask, oi, t, G);
            if (act) tile[t][gl] = other_old[(size_t)oit * f + gl];
        }
        __syncwarp(gmask);
        double est = 0.0;
        if (has)
            for (int j = 0; j < f; ++j) est = __dadd_rn(est, __dmul_rn(tile[gl][j], mrow[j]));
        re[gl] = make_double2(rr, est);
        __syncwarp(gmask);
        for (int t = 0; t < cnt; ++t) {
            const double2 x = re[t];
            const double o = act ? tile[t][gl] : 0.0;
            n0 = __dadd_rn(n0, __dmul_rn(o, x.x));
            d0 = __dadd_rn(d0, __dmul_rn(o, x.y));
        }
        __syncwarp(gmask);  // the next tile overwrites the rows
    }
    if (act) {
        const double nr = (double)(e - b);
        const double d = __dadd_rn(d0, __dmul_rn(__dmul_rn(nr, reg), m_l));
        if (d == 0.0) atomicExch(&status[SEG_ZERODIV], 1);
        mine_new[(size_t)seg * f + gl] = __dmul_rn(m_l, __ddiv_rn(n0, d));
    }
}

// biased model: sequential bias recursion + est per rating (all_ratings order), one thread.
__global__ void nmf_dot_kernel(int64_t n, int f, const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                               const double* __restrict__ pu, const double* __restrict__ qi, double* __restrict__ dot) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* p = pu + (size_t)u[k] * f;
    const double* q = qi + (size_t)i[k] * f;
    double d = 0.0;
    for (int j = 0; j < f; ++j) d = __dadd_rn(d, __dmul_rn(q[j], p[j]));
    dot[k] = d;
}

__global__ void nmf_bias_scan_kernel(int64_t n, const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                     const double* __restrict__ r, const double* __restrict__ dot, double mu,
                                     double lr_bu, double lr_bi, double reg_bu, double reg_bi, double* bu, double* bi,
                                     double* __restrict__ est) {
    if (blockIdx.x || threadIdx.x) return;
    for (int64_t k = 0; k < n; ++k) {
        const int32_t uu = u[k], ii = i[k];
        const double b_u = bu[uu], b_i = bi[ii];
        const double e = __dadd_rn(__dadd_rn(__dadd_rn(mu, b_u), b_i), dot[k]);
        est[k] = e;
        const double err = __dsub_rn(r[k], e);
        bu[uu] = __dadd_rn(b_u, __dmul_rn(lr_bu, __dsub_rn(err, __dmul_rn(reg_bu, b_u))));
        bi[ii] = __dadd_rn(b_i, __dmul_rn(lr_bi, __dsub_rn(err, __dmul_rn(reg_bi, b_i))));
    }
}

__global__ void iota_kernel(int64_t n, int64_t* v) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) v[k] = k;
}
__global__ void count_kernel(int64_t n, const int32_t* __restrict__ key, unsigned long long* cnt) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) atomicAdd(&cnt[key[k]], 1ull);
}
__global__ void gather_item_side_kernel(int64_t n, const int64_t* __restrict__ perm, const int32_t* __restrict__ u,
                                        const double* __restrict__ r, int32_t* __restrict__ u_out,
                                        double* __restrict__ r_out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) {
        u_out[k] = u[perm[k]];
        r_out[k] = r[perm[k]];
    }
}
__global__ void gather_f64_kernel(int64_t n, const int64_t* __restrict__ perm, const double* __restrict__ src,
                                  double* __restrict__ dst) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) dst[k] = src[perm[k]];
}
__global__ void check_grouped_kernel(int64_t n, const int32_t* __restrict__ u, int* status) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k + 1 < n && u[k + 1] < u[k]) atomicExch(&status[1], 1);
}

template <int G>
static int launch_pass(int64_t n_seg, int f, const int64_t* ptr, const int32_t* idx, const double* r, const double* est,
                       const double* mine_old, const double* other_old, double* mine_new, double reg, int* status,
                       cudaStream_t st) {
    const int threads = 128, gpb = threads / G;
    nmf_pass_kernel<G><<<(unsigned)ceil_div(n_seg, gpb), threads, 0, st>>>(n_seg, f, ptr, idx, r, est, mine_old,
                                                                           other_old, mine_new, reg, status);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

static int nmf_pass(int64_t n_seg, int f, const int64_t* ptr, const int32_t* idx, const double* r, const double* est,
                    const double* mine_old, const double* other_old, double* mine_new, double reg, int* status,
                    cudaStream_t st) {
    if (n_seg <= 0) return SB2_OK;
    if (f <= 4) return launch_pass<4>(n_seg, f, ptr, idx, r, est, mine_old, other_old, mine_new, reg, status, st);
    if (f <= 8) return launch_pass<8>(n_seg, f, ptr, idx, r, est, mine_old, other_old, mine_new, reg, status, st);
    if (f <= 16) return launch_pass<16>(n_seg, f, ptr, idx, r, est, mine_old, other_old, mine_new, reg, status, st);
    return launch_pass<32>(n_seg, f, ptr, idx, r, est, mine_old, other_old, mine_new, reg, status, st);
}

template <int G>
static int launch_pass_fused(int64_t n_seg, int f, const int64_t* ptr, const int32_t* idx, const double* r,
                             const double* mine_old, const double* other_old, double* mine_new, double reg, int* status,
                             cudaStream_t st) {
    const int gpb = NMF_THREADS / G;
    nmf_pass_fused_kernel<G><<<(unsigned)ceil_div(n_seg, gpb), NMF_THREADS, 0, st>>>(n_seg, f, ptr, idx, r, mine_old,
                                                                                     other_old, mine_new, reg, status);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

// f <= 32 only
static int nmf_pass_fused(int64_t n_seg, int f, const int64_t* ptr, const int32_t* idx, const double* r,
                          const double* mine_old, const double* other_old, double* mine_new, double reg, int* status,
                          cudaStream_t st) {
    if (n_seg <= 0) return SB2_OK;
    if (f <= 4) return launch_pass_fused<4>(n_seg, f, ptr, idx, r, mine_old, other_old, mine_new, reg, status, st);
    if (f <= 8) return launch_pass_fused<8>(n_seg, f, ptr, idx, r, mine_old, other_old, mine_new, reg, status, st);
    if (f <= 16) return launch_pass_fused<16>(n_seg, f, ptr, idx, r, mine_old, other_old, mine_new, reg, status, st);
    return launch_pass_fused<32>(n_seg, f, ptr, idx, r, mine_old, other_old, mine_new, reg, status, st);
}

// est for a range of CSC (item-ordered) entries, recomputed from the factors: dot(q[item], p[user]) in factor
// order -- the same multiply/add chain as nmf_dot_kernel, hence the same bits as est[perm[k]].
__global__ void nmf_dot_item_kernel(int64_t e0, int64_t e1, int f, const int32_t* __restrict__ i_it,
                                    const int32_t* __restrict__ u_it, const double* __restrict__ pu,
                                    const double* __restrict__ qi, double* __restrict__ est_it) {
    const int64_t k = e0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= e1) return;
    const double* p = pu + (size_t)u_it[k] * f;
    const double* q = qi + (size_t)i_it[k] * f;
    double d = 0.0;
    for (int j = 0; j < f; ++j) d = __dadd_rn(d, __dmul_rn(q[j], p[j]));
    est_it[k] = d;
}

}  // namespace sb2

// NMF plan: everything that depends only on the rating structure (user CSR offsets, item-side CSC by a
// stable sort of positions) plus per-epoch scratch.  One epoch can be restricted to a user range and an item
// range: the multi-GPU formulation shards the ACCUMULATORS (each ordered sum lives on exactly one rank, so the
// result stays bit-exact) and all-gathers the new factors between epochs (surprise_b200/distributed.py).
struct sb2_nmf_plan {
    int64_t n_users = 0, n_items = 0, n = 0;
    int f = 0;
    const int32_t* u = nullptr;  // borrowed: the all_ratings COO must outlive the plan
    const int32_t* i = nullptr;
    const double* r = nullptr;
    sb2::DevBuf status, ptr_u, ptr_i, perm, u_it, i_it, r_it, dot, est, est_it;
    std::vector<int64_t> ptr_u_h, ptr_i_h;
};

namespace sb2 {

int nmf_plan_create_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i, const double* r,
                        int n_factors, cudaStream_t st, sb2_nmf_plan** out) {
    const int f = n_factors;
    if (f <= 0 || f > 256 || n_users <= 0 || n_items <= 0 || n < 0 || n > 0x7FFFFFF0ll) {
        set_error("nmf_plan: invalid shape (n_factors must be in [1, 256])");
        return SB2_ERR_INVALID;
    }
    sb2_nmf_plan* p = new sb2_nmf_plan();
    struct Guard {
        sb2_nmf_plan* p;
        ~Guard() { delete p; }
    } guard{p};
    p->n_users = n_users; p->n_items = n_items; p->n = n; p->f = f; p->u = u; p->i = i; p->r = r;
    DevBuf cnt_u, cnt_i, perm_in, tmp;
    SB2_TRY(p->status.alloc(2 * sizeof(int), st));
    SB2_CUDA(cudaMemsetAsync(p->status.p, 0, 2 * sizeof(int), st));
    const size_t n1 = (size_t)std::max<int64_t>(n, 1);
    const unsigned nb = (unsigned)ceil_div((int64_t)n1, 256);
    // user CSR = the COO itself (grouped by u); item CSC = stable sort of positions by item
    SB2_TRY(cnt_u.alloc((size_t)(n_users + 1) * 8, st));
    SB2_TRY(cnt_i.alloc((size_t)(n_items + 1) * 8, st));
    SB2_TRY(p->ptr_u.alloc((size_t)(n_users + 1) * 8, st));
    SB2_TRY(p->ptr_i.alloc((size_t)(n_items + 1) * 8, st));
    SB2_CUDA(cudaMemsetAsync(cnt_u.p, 0, (size_t)(n_users + 1) * 8, st));
    SB2_CUDA(cudaMemsetAsync(cnt_i.p, 0, (size_t)(n_items + 1) * 8, st));
    if (n > 0) {
        check_grouped_kernel<<<nb, 256, 0, st>>>(n, u, p->status.as<int>());
        SB2_LAUNCH_CHECK();
        count_kernel<<<nb, 256, 0, st>>>(n, u, cnt_u.as<unsigned long long>());
        SB2_LAUNCH_CHECK();
        count_kernel<<<nb, 256, 0, st>>>(n, i, cnt_i.as<unsigned long long>());
        SB2_LAUNCH_CHECK();
    }
    size_t tb1 = 0, tb2 = 0, tb3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb1, cnt_u.as<int64_t>(), p->ptr_u.as<int64_t>(), (int)(n_users + 1), st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, cnt_i.as<int64_t>(), p->ptr_i.as<int64_t>(), (int)(n_items + 1), st);
    SB2_TRY(p->perm.alloc(n1 * 8, st));
    SB2_TRY(perm_in.alloc(n1 * 8, st));
    SB2_TRY(p->i_it.alloc(n1 * 4, st));
    cub::DeviceRadixSort::SortPairs(nullptr, tb3, i, p->i_it.as<int32_t>(), perm_in.as<int64_t>(), p->perm.as<int64_t>(),
                                    (int)n, 0, 32, st);
    size_t tb = std::max(tb1, std::max(tb2, tb3));
    SB2_TRY(tmp.alloc(tb + 16, st));
    SB2_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt_u.as<int64_t>(), p->ptr_u.as<int64_t>(), (int)(n_users + 1), st));
    SB2_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt_i.as<int64_t>(), p->ptr_i.as<int64_t>(), (int)(n_items + 1), st));
    launch_counter() += 2;
    SB2_TRY(p->u_it.alloc(n1 * 4, st));
    SB2_TRY(p->r_it.alloc(n1 * 8, st));
    if (n > 0) {
        iota_kernel<<<nb, 256, 0, st>>>(n, perm_in.as<int64_t>());
        SB2_LAUNCH_CHECK();
        // LSD radix sort is stable: equal items keep all_ratings() order (u ascending, then ur[u] order)
        SB2_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, i, p->i_it.as<int32_t>(), perm_in.as<int64_t>(),
                                                 p->perm.as<int64_t>(), (int)n, 0, 32, st));
        launch_counter() += 1;
        gather_item_side_kernel<<<nb, 256, 0, st>>>(n, p->perm.as<int64_t>(), u, r, p->u_it.as<int32_t>(),
                                                    p->r_it.as<double>());
        SB2_LAUNCH_CHECK();
    }
    // dot / est_it (8 bytes per rating each) are only needed by the biased / f > 32 path: allocated on first use
    p->ptr_u_h.resize((size_t)n_users + 1);
    p->ptr_i_h.resize((size_t)n_items + 1);
    SB2_CUDA(cudaMemcpyAsync(p->ptr_u_h.data(), p->ptr_u.p, (size_t)(n_users + 1) * 8, cudaMemcpyDeviceToHost, st));
    SB2_CUDA(cudaMemcpyAsync(p->ptr_i_h.data(), p->ptr_i.p, (size_t)(n_items + 1) * 8, cudaMemcpyDeviceToHost, st));
    SB2_CUDA(cudaStreamSynchronize(st));
    guard.p = nullptr;
    *out = p;
    return SB2_OK;
}

// One epoch restricted to users [u0, u1) and items [i0, i1): reads pu_cur / qi_cur (full), writes the rows of
// pu_new / qi_new in the ranges.  bu / bi are updated in place when biased (sequential recursion, replicated).
int nmf_plan_epoch_dev(sb2_nmf_plan* p, const sb2_nmf_params* prm, const double* pu_cur, const double* qi_cur,
                       double* pu_new, double* qi_new, double* bu, double* bi, int64_t u0, int64_t u1, int64_t i0,
                       int64_t i1, cudaStream_t st) {
    if (u0 < 0 || u1 > p->n_users || u0 > u1 || i0 < 0 || i1 > p->n_items || i0 > i1 || prm->n_factors != p->f) {
        set_error("nmf_epoch: invalid range / n_factors");
        return SB2_ERR_INVALID;
    }
    const int f = p->f;
    const int64_t n = p->n;
    const bool biased = prm->biased != 0;
    const double mu = biased ? prm->global_mean : 0.0;
    const bool full = (u0 == 0 && u1 == p->n_users && i0 == 0 && i1 == p->n_items);
    const int64_t eu0 = p->ptr_u_h[(size_t)u0], eu1 = p->ptr_u_h[(size_t)u1];
    const int64_t ei0 = p->ptr_i_h[(size_t)i0], ei1 = p->ptr_i_h[(size_t)i1];
    const double* est_u = p->dot.as<double>();
    if (!biased && f <= 32 && getenv("SB2_NMF_UNFUSED") == nullptr) {
        // unbiased: est == dot exactly (mu = bu = bi = 0 and 0 + x == x); recomputed inside the passes
        SB2_TRY(nmf_pass_fused(u1 - u0, f, p->ptr_u.as<int64_t>() + u0, p->i, p->r, pu_cur + (size_t)u0 * f, qi_cur,
                               pu_new + (size_t)u0 * f, prm->reg_pu, p->status.as<int>(), st));
        SB2_TRY(nmf_pass_fused(i1 - i0, f, p->ptr_i.as<int64_t>() + i0, p->u_it.as<int32_t>(), p->r_it.as<double>(),
                               qi_cur + (size_t)i0 * f, pu_cur, qi_new + (size_t)i0 * f, prm->reg_qi,
                               p->status.as<int>(), st));
        return SB2_OK;
    }
    if (n > 0) {
        if (!p->dot.p) SB2_TRY(p->dot.alloc((size_t)n * 8, st));
        if (!p->est_it.p) SB2_TRY(p->est_it.alloc((size_t)n * 8, st));
        est_u = p->dot.as<double>();
        if (biased) {
            if (!p->est.p) SB2_TRY(p->est.alloc((size_t)n * 8, st));
            const unsigned nb = (unsigned)ceil_div(n, 256);
            nmf_dot_kernel<<<nb, 256, 0, st>>>(n, f, p->u, p->i, pu_cur, qi_cur, p->dot.as<double>());
            SB2_LAUNCH_CHECK();
            nmf_bias_scan_kernel<<<1, 32, 0, st>>>(n, p->u, p->i, p->r, p->dot.as<double>(), mu, prm->lr_bu, prm->lr_bi,
                                                   prm->reg_bu, prm->reg_bi, bu, bi, p->est.as<double>());
            SB2_LAUNCH_CHECK();
            est_u = p->est.as<double>();
            gather_f64_kernel<<<nb, 256, 0, st>>>(n, p->perm.as<int64_t>(), est_u, p->est_it.as<double>());
            SB2_LAUNCH_CHECK();
        } else {
            // unbiased: est == dot exactly (mu = bu = bi = 0 and 0 + x == x)
            if (eu1 > eu0) {
                nmf_dot_kernel<<<(unsigned)ceil_div(eu1 - eu0, 256), 256, 0, st>>>(eu1 - eu0, f, p->u + eu0, p->i + eu0,
                                                                                  pu_cur, qi_cur, p->dot.as<double>() + eu0);
                SB2_LAUNCH_CHECK();
            }
            if (full) {
                gather_f64_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(n, p->perm.as<int64_t>(), est_u,
                                                                              p->est_it.as<double>());
                SB2_LAUNCH_CHECK();
            } else if (ei1 > ei0) {
                nmf_dot_item_kernel<<<(unsigned)ceil_div(ei1 - ei0, 256), 256, 0, st>>>(
                    ei0, ei1, f, p->i_it.as<int32_t>(), p->u_it.as<int32_t>(), pu_cur, qi_cur, p->est_it.as<double>());
                SB2_LAUNCH_CHECK();
            }
        }
    }
    // segment ranges: shift the CSR offsets / row pointers so that segment 0 of the launch is u0 (i0)
    SB2_TRY(nmf_pass(u1 - u0, f, p->ptr_u.as<int64_t>() + u0, p->i, p->r, est_u, pu_cur + (size_t)u0 * f, qi_cur,
                     pu_new + (size_t)u0 * f, prm->reg_pu, p->status.as<int>(), st));
    SB2_TRY(nmf_pass(i1 - i0, f, p->ptr_i.as<int64_t>() + i0, p->u_it.as<int32_t>(), p->r_it.as<double>(),
                     p->est_it.as<double>(), qi_cur + (size_t)i0 * f, pu_cur, qi_new + (size_t)i0 * f, prm->reg_qi,
                     p->status.as<int>(), st));
    return SB2_OK;
}

// synchronises the stream; SB2_ERR_ZERO_DIVISION where the reference raises (:723, :730)
int nmf_plan_status(sb2_nmf_plan* p, cudaStream_t st) {
    int h[2] = {0, 0};
    SB2_CUDA(cudaMemcpyAsync(h, p->status.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    SB2_CUDA(cudaStreamSynchronize(st));
    if (h[1]) {
        set_error("nmf: (u, i, r) must be grouped by ascending u (all_ratings() order)");
        return SB2_ERR_INVALID;
    }
    if (h[0]) {
        set_error("float division");
        return SB2_ERR_ZERO_DIVISION;
    }
    return SB2_OK;
}

void nmf_plan_destroy(sb2_nmf_plan* p) { delete p; }

int nmf_fit_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i, const double* r,
                const sb2_nmf_params* prm, double* pu, double* qi, double* bu, double* bi, cudaStream_t st) {
    sb2_nmf_plan* plan = nullptr;
    SB2_TRY(nmf_plan_create_dev(n_users, n_items, n, u, i, r, prm->n_factors, st, &plan));
    struct Guard {
        sb2_nmf_plan* p;
        ~Guard() { delete p; }
    } guard{plan};
    const int f = prm->n_factors;
    DevBuf pu2, qi2;
    SB2_TRY(pu2.alloc((size_t)n_users * f * 8, st));
    SB2_TRY(qi2.alloc((size_t)n_items * f * 8, st));
    SB2_CUDA(cudaMemsetAsync(bu, 0, (size_t)n_users * sizeof(double), st));
    SB2_CUDA(cudaMemsetAsync(bi, 0, (size_t)n_items * sizeof(double), st));
    double* pu_cur = pu;
    double* pu_nxt = pu2.as<double>();
    double* qi_cur = qi;
    double* qi_nxt = qi2.as<double>();
    for (int ep = 0; ep < prm->n_epochs; ++ep) {
        SB2_TRY(nmf_plan_epoch_dev(plan, prm, pu_cur, qi_cur, pu_nxt, qi_nxt, bu, bi, 0, n_users, 0, n_items, st));
        std::swap(pu_cur, pu_nxt);
        std::swap(qi_cur, qi_nxt);
    }
    if (pu_cur != pu) {
        SB2_CUDA(cudaMemcpyAsync(pu, pu_cur, (size_t)n_users * f * 8, cudaMemcpyDeviceToDevice, st));
        SB2_CUDA(cudaMemcpyAsync(qi, qi_cur, (size_t)n_items * f * 8, cudaMemcpyDeviceToDevice, st));
    }
    return nmf_plan_status(plan, st);
}

}  // namespace sb2
