// Stratified (DSGD) SGD for SVD -- replaces SVD.sgd (matrix_factorization.pyx:241-262).
//
// The reference walks all_ratings() sequentially; every update reads the latest pu[u], qi[i], so two
// ratings can run concurrently only if they share neither user nor item.  We therefore block the
// rating matrix B x B (B = one persistent CTA per SM): user u lives in user block u % B, item i in
// item block i % B.  An epoch is B strata; in stratum s CTA `ub` owns user block ub (its pu/bu rows
// stay in shared memory for the whole fit) and item block (ub + s) % B, whose qi/bi rows it pulls
// into shared memory, updates, and writes back.  Item blocks move CTA -> CTA along a ring, so instead
// of a grid-wide barrier per stratum each CTA waits on a per-item-block step counter that its ring
// neighbour publishes (st.release / ld.acquire) -- every block is touched by exactly one CTA at a time.
// Inside a block the same construction is repeated across the W lane-groups of the CTA (W x W
// sub-blocks, W sub-strata separated by __syncthreads()), so no two lane-groups ever hold the same
// user or item row: the schedule is conflict-free by construction, with no atomics on factors.
//
// One lane-group (G = 4..32 lanes, chosen from n_factors) performs one rating update: 128-bit loads
// of the pu / qi rows, shuffle-tree dot product, bias update by the group leader, factor update and
// 128-bit stores.  Arithmetic is fp32 (the contract is held-out RMSE within 0.005 of the reference).
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace sb2 {

struct DsgdArgs {
    int n_users, n_items, B, W, f, FP;  // FP = n_factors rounded up to 4
    int max_ul, max_il;                 // rows per user / item block (ceil)
    const int* ul;                      // records sorted by (stratum, user block, sub-stratum, user sub-block)
    const int* il;
    const float* r;
    const int* off;                     // B*B*W*W + 1 offsets into the records
    float* pu;                          // n_users x FP
    float* qi;                          // n_items x FP
    float* bu;
    float* bi;
    int* flags;                         // B step counters (ring hand-off of item blocks)
    float mu, lr_bu, lr_bi, lr_pu, lr_qi, reg_bu, reg_bi, reg_pu, reg_qi;
    int n_epochs;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool SMEM>
__device__ __forceinline__ float4 row_ld4(const float* p) {
    if (SMEM) return *reinterpret_cast<const float4*>(p);
    return __ldcg(reinterpret_cast<const float4*>(p));
}
template <bool SMEM>
__device__ __forceinline__ void row_st4(float* p, float4 v) {
    if (SMEM) *reinterpret_cast<float4*>(p) = v;
    else __stcg(reinterpret_cast<float4*>(p), v);
}
template <bool SMEM>
__device__ __forceinline__ float sc_ld(const float* p) {
    if (SMEM) return *p;
    return __ldcg(p);
}
template <bool SMEM>
__device__ __forceinline__ void sc_st(float* p, float v) {
    if (SMEM) *p = v;
    else __stcg(p, v);
}

// G lanes per rating, SU / SI: user / item block staged in shared memory, BIASED: SVD(biased=True)
template <int G, bool SU, bool SI, bool BIASED>
__global__ void __launch_bounds__(256, 1) dsgd_svd_kernel(const DsgdArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    const int B = a.B, W = a.W, FP = a.FP, F4 = a.FP >> 2;
    const int ub = blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int gid = tid / G, gl = tid % G;
    const int gbase = (tid & 31) / G * G;
    const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << gbase);

    // shared memory carve-up
    float* pu_s = smem_f;
    float* qi_s = pu_s + (SU ? (size_t)a.max_ul * FP : 0);
    float* bu_s = qi_s + (SI ? (size_t)a.max_il * FP : 0);
    float* bi_s = bu_s + (SU ? a.max_ul : 0);
    int* off_s = reinterpret_cast<int*>(bi_s + (SI ? a.max_il : 0));

    const int nu_local = (a.n_users - ub + B - 1) / B;
    if (SU) {
        for (int t = tid; t < nu_local * F4; t += nthr) {
            const int l = t / F4, c = t % F4;
            reinterpret_cast<float4*>(pu_s)[l * F4 + c] =
                __ldcg(reinterpret_cast<const float4*>(a.pu + ((size_t)(ub + (size_t)l * B)) * FP) + c);
        }
        for (int l = tid; l < nu_local; l += nthr) bu_s[l] = __ldcg(a.bu + ub + (size_t)l * B);
    }
    __syncthreads();

    for (int ep = 0; ep < a.n_epochs; ++ep) {
        for (int s = 0; s < B; ++s) {
            const int ib = (ub + s) % B;
            const int step = ep * B + s;
            if (tid == 0) {
                while (ld_acquire(a.flags + ib) != step) { /* ring neighbour still owns the block */ }
            }
            __syncthreads();
            const int ni_local = (a.n_items - ib + B - 1) / B;
            if (SI) {
                for (int t = tid; t < ni_local * F4; t += nthr) {
                    const int l = t / F4, c = t % F4;
                    reinterpret_cast<float4*>(qi_s)[l * F4 + c] =
                        __ldcg(reinterpret_cast<const float4*>(a.qi + ((size_t)(ib + (size_t)l * B)) * FP) + c);
                }
                for (int l = tid; l < ni_local; l += nthr) bi_s[l] = __ldcg(a.bi + ib + (size_t)l * B);
            }
            const int* offg = a.off + ((size_t)s * B + ub) * (size_t)(W * W);
            for (int t = tid; t <= W * W; t += nthr) off_s[t] = offg[t];
            __syncthreads();

            for (int t = 0; t < W; ++t) {
                const int k0 = off_s[t * W + gid], k1 = off_s[t * W + gid + 1];
                for (int kb = k0; kb < k1; kb += G) {
                    // the group fetches up to G records at once, then replays them one by one
                    const int kk = kb + gl;
                    int my_ul = 0, my_il = 0;
                    float my_r = 0.f;
                    if (kk < k1) { my_ul = a.ul[kk]; my_il = a.il[kk]; my_r = a.r[kk]; }
                    const int cnt = min(G, k1 - kb);
                    for (int j = 0; j < cnt; ++j) {
                        const int ul = __shfl_sync(gmask, my_ul, gbase + j);
                        const int il = __shfl_sync(gmask, my_il, gbase + j);
                        const float r = __shfl_sync(gmask, my_r, gbase + j);
                        float* prow = SU ? pu_s + (size_t)ul * FP : a.pu + ((size_t)(ub + (size_t)ul * B)) * FP;
                        float* qrow = SI ? qi_s + (size_t)il * FP : a.qi + ((size_t)(ib + (size_t)il * B)) * FP;
                        float dot = 0.f;
                        for (int c = gl; c < F4; c += G) {
                            const float4 p = row_ld4<SU>(prow + 4 * c), q = row_ld4<SI>(qrow + 4 * c);
                            dot += p.x * q.x + p.y * q.y + p.z * q.z + p.w * q.w;
                        }
#pragma unroll
                        for (int o = G / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(gmask, dot, o);
                        float err;
                        if (BIASED) {
                            float* bup = SU ? bu_s + ul : a.bu + ub + (size_t)ul * B;
                            float* bip = SI ? bi_s + il : a.bi + ib + (size_t)il * B;
                            const float b_u = sc_ld<SU>(bup), b_i = sc_ld<SI>(bip);
                            err = r - (a.mu + b_u + b_i + dot);
                            if (gl == 0) {
                                sc_st<SU>(bup, b_u + a.lr_bu * (err - a.reg_bu * b_u));
                                sc_st<SI>(bip, b_i + a.lr_bi * (err - a.reg_bi * b_i));
                            }
                        } else {
                            err = r - dot;
                        }
                        for (int c = gl; c < F4; c += G) {
                            const float4 p = row_ld4<SU>(prow + 4 * c), q = row_ld4<SI>(qrow + 4 * c);
                            float4 pn, qn;
                            pn.x = p.x + a.lr_pu * (err * q.x - a.reg_pu * p.x);
                            pn.y = p.y + a.lr_pu * (err * q.y - a.reg_pu * p.y);
                            pn.z = p.z + a.lr_pu * (err * q.z - a.reg_pu * p.z);
                            pn.w = p.w + a.lr_pu * (err * q.w - a.reg_pu * p.w);
                            qn.x = q.x + a.lr_qi * (err * p.x - a.reg_qi * q.x);
                            qn.y = q.y + a.lr_qi * (err * p.y - a.reg_qi * q.y);
                            qn.z = q.z + a.lr_qi * (err * p.z - a.reg_qi * q.z);
                            qn.w = q.w + a.lr_qi * (err * p.w - a.reg_qi * q.w);
                            row_st4<SU>(prow + 4 * c, pn);
                            row_st4<SI>(qrow + 4 * c, qn);
                        }
                        __syncwarp(gmask);
                    }
                }
                __syncthreads();
            }

            if (SI) {
                for (int t = tid; t < ni_local * F4; t += nthr) {
                    const int l = t / F4, c = t % F4;
                    __stcg(reinterpret_cast<float4*>(a.qi + ((size_t)(ib + (size_t)l * B)) * FP) + c,
                           reinterpret_cast<const float4*>(qi_s)[l * F4 + c]);
                }
                if (BIASED)
                    for (int l = tid; l < ni_local; l += nthr) __stcg(a.bi + ib + (size_t)l * B, bi_s[l]);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(a.flags + ib, step + 1);
        }
    }
    if (SU) {
        for (int t = tid; t < nu_local * F4; t += nthr) {
            const int l = t / F4, c = t % F4;
            __stcg(reinterpret_cast<float4*>(a.pu + ((size_t)(ub + (size_t)l * B)) * FP) + c,
                   reinterpret_cast<const float4*>(pu_s)[l * F4 + c]);
        }
        if (BIASED)
            for (int l = tid; l < nu_local; l += nthr) __stcg(a.bu + ub + (size_t)l * B, bu_s[l]);
    }
}

// ------------------------------------------------------------------------------------------------
// preparation kernels
// ------------------------------------------------------------------------------------------------
__global__ void dsgd_key_kernel(int64_t n, const int32_t* __restrict__ u, const int32_t* __restrict__ i, int B, int W,
                                int n_users, int n_items, unsigned* __restrict__ key, int* __restrict__ val,
                                int* status) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int uu = u[k], ii = i[k];
    if (uu < 0 || uu >= n_users || ii < 0 || ii >= n_items) {
        atomicExch(status, 1);
        key[k] = 0;
        val[k] = (int)k;
        return;
    }
    const int ub = uu % B, ibk = ii % B;
    const int uw = (uu / B) % W, iw = (ii / B) % W;
    const int s = (ibk - ub + B) % B, t = (iw - uw + W) % W;
    key[k] = (unsigned)((((size_t)s * B + ub) * W + t) * W + uw);
    val[k] = (int)k;
}

__global__ void dsgd_gather_kernel(int64_t n, const int* __restrict__ val, const unsigned* __restrict__ key_sorted,
                                   const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                   const double* __restrict__ r, int B, int* __restrict__ ul, int* __restrict__ il,
                                   float* __restrict__ rr, int* __restrict__ cnt) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int src = val[k];
    ul[k] = u[src] / B;
    il[k] = i[src] / B;
    rr[k] = (float)r[src];
    atomicAdd(&cnt[key_sorted[k]], 1);
}

__global__ void f64_to_rows_kernel(int64_t rows, int f, int FP, const double* __restrict__ src, float* __restrict__ dst) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= rows * FP) return;
    const int64_t row = t / FP;
    const int c = (int)(t % FP);
    dst[t] = c < f ? (float)src[row * f + c] : 0.f;
}
__global__ void rows_to_f64_kernel(int64_t rows, int f, int FP, const float* __restrict__ src, double* __restrict__ dst) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= rows * f) return;
    const int64_t row = t / f;
    const int c = (int)(t % f);
    dst[t] = (double)src[row * FP + c];
}
__global__ void f32_to_f64_kernel(int64_t n, const float* __restrict__ src, double* __restrict__ dst) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) dst[t] = (double)src[t];
}

}  // namespace sb2

// ------------------------------------------------------------------------------------------------
// plan object (C-ABI handle)
// ------------------------------------------------------------------------------------------------
struct sb2_svd_plan {
    int64_t n_users = 0, n_items = 0, n = 0;
    sb2_sgd_params prm;
    int B = 0, W = 0, G = 0, FP = 0;
    bool stage_u = false, stage_i = false;
    size_t smem = 0;
    int *ul = nullptr, *il = nullptr, *off = nullptr, *flags = nullptr;
    float *r = nullptr, *pu = nullptr, *qi = nullptr, *bu = nullptr, *bi = nullptr;
    bool owns_factors = true;
};

namespace sb2 {

template <int G, bool SU, bool SI, bool BIASED>
static int dsgd_launch_t(const sb2_svd_plan* p, const DsgdArgs& a, cudaStream_t st) {
    auto kern = dsgd_svd_kernel<G, SU, SI, BIASED>;
    SB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    void* args[] = {(void*)&a};
    SB2_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(p->B), dim3(p->W * G), args, p->smem, st));
    launch_counter()++;
    return SB2_OK;
}

template <int G>
static int dsgd_launch_g(const sb2_svd_plan* p, const DsgdArgs& a, cudaStream_t st) {
    const bool b = p->prm.biased != 0;
    if (p->stage_u && p->stage_i) return b ? dsgd_launch_t<G, true, true, true>(p, a, st) : dsgd_launch_t<G, true, true, false>(p, a, st);
    if (!p->stage_u && p->stage_i) return b ? dsgd_launch_t<G, false, true, true>(p, a, st) : dsgd_launch_t<G, false, true, false>(p, a, st);
    if (p->stage_u && !p->stage_i) return b ? dsgd_launch_t<G, true, false, true>(p, a, st) : dsgd_launch_t<G, true, false, false>(p, a, st);
    return b ? dsgd_launch_t<G, false, false, true>(p, a, st) : dsgd_launch_t<G, false, false, false>(p, a, st);
}

static void plan_free(sb2_svd_plan* p) {
    if (!p) return;
    cudaFree(p->ul); cudaFree(p->il); cudaFree(p->off); cudaFree(p->flags); cudaFree(p->r);
    if (p->owns_factors) { cudaFree(p->pu); cudaFree(p->qi); cudaFree(p->bu); cudaFree(p->bi); }
    delete p;
}

// u, i, r: DEVICE arrays (all_ratings COO)
int svd_plan_create_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                        const double* r, const sb2_sgd_params* prm, int with_yj, const int64_t* u_ptr,
                        const int32_t* ui_idx, cudaStream_t st, sb2_svd_plan** out) {
    if (with_yj) {
        set_error("svdpp: not implemented in this build");
        return SB2_ERR_UNSUPPORTED;
    }
    if (n_users <= 0 || n_items <= 0 || n < 0 || n > 0x7FFFFFF0ll || prm->n_factors <= 0 || prm->n_factors > 1024) {
        set_error("svd_plan: invalid shape");
        return SB2_ERR_INVALID;
    }
    sb2_svd_plan* p = new sb2_svd_plan();
    p->n_users = n_users; p->n_items = n_items; p->n = n; p->prm = *prm;
    const int f = prm->n_factors;
    p->FP = (int)round_up(f, 4);
    const int F4 = p->FP / 4;
    p->G = F4 <= 4 ? 4 : F4 <= 8 ? 8 : F4 <= 16 ? 16 : 32;
    p->W = p->G == 32 ? 8 : 16;
    if (p->W * p->G > 256) p->W = 256 / p->G;
    // one CTA per SM, but keep >= ~32 ratings per (user block, item block) so that a stratum's work is
    // not dwarfed by its hand-off latency (matters for the small sub-matrices of the multi-GPU ring)
    int B = sm_count();
    const int b_work = std::max(8, (int)sqrt((double)std::max<int64_t>(n, 1) / 32.0));
    if (B > b_work) B = b_work;
    if (B > n_users) B = (int)n_users;
    if (B > n_items) B = (int)n_items;
    p->B = B;
    const int max_ul = (int)ceil_div(n_users, B), max_il = (int)ceil_div(n_items, B);
    // shared-memory staging plan: items first (they move every stratum), then users
    const size_t off_bytes = (size_t)(p->W * p->W + 1) * sizeof(int) + 16;
    const size_t need_i = (size_t)max_il * (p->FP + 1) * sizeof(float) + 16;
    const size_t need_u = (size_t)max_ul * (p->FP + 1) * sizeof(float) + 16;
    const size_t budget = 200 * 1024;
    p->stage_i = off_bytes + need_i <= budget;
    p->stage_u = off_bytes + (p->stage_i ? need_i : 0) + need_u <= budget;
    p->smem = off_bytes + (p->stage_i ? need_i : 0) + (p->stage_u ? need_u : 0);

    auto fail = [&](int rc) { plan_free(p); return rc; };
#define PLAN_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return fail(SB2_ERR_CUDA);                                                         \
        }                                                                                      \
    } while (0)
    const size_t n1 = (size_t)std::max<int64_t>(n, 1);
    const size_t nkeys = (size_t)B * B * p->W * p->W;
    PLAN_CUDA(cudaMalloc(&p->ul, n1 * 4));
    PLAN_CUDA(cudaMalloc(&p->il, n1 * 4));
    PLAN_CUDA(cudaMalloc(&p->r, n1 * 4));
    PLAN_CUDA(cudaMalloc(&p->off, (nkeys + 1) * 4));
    PLAN_CUDA(cudaMalloc(&p->flags, (size_t)B * 4));
    PLAN_CUDA(cudaMalloc(&p->pu, (size_t)n_users * p->FP * 4));
    PLAN_CUDA(cudaMalloc(&p->qi, (size_t)n_items * p->FP * 4));
    PLAN_CUDA(cudaMalloc(&p->bu, (size_t)n_users * 4));
    PLAN_CUDA(cudaMalloc(&p->bi, (size_t)n_items * 4));

    // stratify: key -> stable radix sort -> gather records -> offsets
    unsigned *key = nullptr, *key2 = nullptr;
    int *val = nullptr, *val2 = nullptr, *cnt = nullptr, *status = nullptr;
    void* tmp = nullptr;
    auto cleanup = [&]() { cudaFree(key); cudaFree(key2); cudaFree(val); cudaFree(val2); cudaFree(cnt); cudaFree(status); cudaFree(tmp); };
#define PREP_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            cleanup();                                                                         \
            return fail(SB2_ERR_CUDA);                                                         \
        }                                                                                      \
    } while (0)
    PREP_CUDA(cudaMalloc(&key, n1 * 4));
    PREP_CUDA(cudaMalloc(&key2, n1 * 4));
    PREP_CUDA(cudaMalloc(&val, n1 * 4));
    PREP_CUDA(cudaMalloc(&val2, n1 * 4));
    PREP_CUDA(cudaMalloc(&cnt, (nkeys + 1) * 4));
    PREP_CUDA(cudaMalloc(&status, 4));
    PREP_CUDA(cudaMemsetAsync(cnt, 0, (nkeys + 1) * 4, st));
    PREP_CUDA(cudaMemsetAsync(status, 0, 4, st));
    int end_bit = 1;
    while (((size_t)1 << end_bit) < nkeys) ++end_bit;
    size_t tb1 = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb1, key, key2, val, val2, (int)n, 0, end_bit, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, cnt, p->off, (int)(nkeys + 1), st);
    const size_t tb = std::max(tb1, tb2);
    PREP_CUDA(cudaMalloc(&tmp, tb + 16));
    if (n > 0) {
        const unsigned nb = (unsigned)ceil_div(n, 256);
        dsgd_key_kernel<<<nb, 256, 0, st>>>(n, u, i, B, p->W, (int)n_users, (int)n_items, key, val, status);
        launch_counter()++;
        size_t t1 = tb;
        PREP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t1, key, key2, val, val2, (int)n, 0, end_bit, st));
        launch_counter()++;
        dsgd_gather_kernel<<<nb, 256, 0, st>>>(n, val2, key2, u, i, r, B, p->ul, p->il, p->r, cnt);
        launch_counter()++;
    }
    size_t t2 = tb;
    PREP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, t2, cnt, p->off, (int)(nkeys + 1), st));
    launch_counter()++;
    int status_h = 0;
    PREP_CUDA(cudaMemcpyAsync(&status_h, status, 4, cudaMemcpyDeviceToHost, st));
    PREP_CUDA(cudaStreamSynchronize(st));
    PREP_CUDA(cudaGetLastError());
    cleanup();
    if (status_h) {
        set_error("svd_plan: user / item index out of range");
        return fail(SB2_ERR_INVALID);
    }
    *out = p;
    return SB2_OK;
}

// pu0 / qi0: DEVICE fp64 (n_users x f), (n_items x f)
int svd_plan_reset_dev(sb2_svd_plan* p, const double* pu0, const double* qi0, const double* yj0, cudaStream_t st) {
    const int f = p->prm.n_factors;
    f64_to_rows_kernel<<<(unsigned)ceil_div(p->n_users * p->FP, 256), 256, 0, st>>>(p->n_users, f, p->FP, pu0, p->pu);
    SB2_LAUNCH_CHECK();
    f64_to_rows_kernel<<<(unsigned)ceil_div(p->n_items * p->FP, 256), 256, 0, st>>>(p->n_items, f, p->FP, qi0, p->qi);
    SB2_LAUNCH_CHECK();
    SB2_CUDA(cudaMemsetAsync(p->bu, 0, (size_t)p->n_users * 4, st));
    SB2_CUDA(cudaMemsetAsync(p->bi, 0, (size_t)p->n_items * 4, st));
    return SB2_OK;
}

int svd_plan_run(sb2_svd_plan* p, int n_epochs, cudaStream_t st) {
    if (n_epochs <= 0 || p->n == 0) return SB2_OK;
    DsgdArgs a;
    memset(&a, 0, sizeof(a));
    a.n_users = (int)p->n_users; a.n_items = (int)p->n_items; a.B = p->B; a.W = p->W;
    a.f = p->prm.n_factors; a.FP = p->FP;
    a.max_ul = (int)ceil_div(p->n_users, p->B); a.max_il = (int)ceil_div(p->n_items, p->B);
    a.ul = p->ul; a.il = p->il; a.r = p->r; a.off = p->off;
    a.pu = p->pu; a.qi = p->qi; a.bu = p->bu; a.bi = p->bi; a.flags = p->flags;
    const sb2_sgd_params& q = p->prm;
    a.mu = q.biased ? (float)q.global_mean : 0.f;
    a.lr_bu = (float)q.lr_bu; a.lr_bi = (float)q.lr_bi; a.lr_pu = (float)q.lr_pu; a.lr_qi = (float)q.lr_qi;
    a.reg_bu = (float)q.reg_bu; a.reg_bi = (float)q.reg_bi; a.reg_pu = (float)q.reg_pu; a.reg_qi = (float)q.reg_qi;
    a.n_epochs = n_epochs;
    SB2_CUDA(cudaMemsetAsync(p->flags, 0, (size_t)p->B * 4, st));
    switch (p->G) {
        case 4: return dsgd_launch_g<4>(p, a, st);
        case 8: return dsgd_launch_g<8>(p, a, st);
        case 16: return dsgd_launch_g<16>(p, a, st);
        default: return dsgd_launch_g<32>(p, a, st);
    }
}

// outputs: DEVICE fp64
int svd_plan_read_dev(sb2_svd_plan* p, double* pu, double* qi, double* bu, double* bi, double* yj, cudaStream_t st) {
    const int f = p->prm.n_factors;
    if (pu) {
        rows_to_f64_kernel<<<(unsigned)ceil_div(p->n_users * f, 256), 256, 0, st>>>(p->n_users, f, p->FP, p->pu, pu);
        SB2_LAUNCH_CHECK();
    }
    if (qi) {
        rows_to_f64_kernel<<<(unsigned)ceil_div(p->n_items * f, 256), 256, 0, st>>>(p->n_items, f, p->FP, p->qi, qi);
        SB2_LAUNCH_CHECK();
    }
    if (bu) {
        f32_to_f64_kernel<<<(unsigned)ceil_div(p->n_users, 256), 256, 0, st>>>(p->n_users, p->bu, bu);
        SB2_LAUNCH_CHECK();
    }
    if (bi) {
        f32_to_f64_kernel<<<(unsigned)ceil_div(p->n_items, 256), 256, 0, st>>>(p->n_items, p->bi, bi);
        SB2_LAUNCH_CHECK();
    }
    return SB2_OK;
}

void svd_plan_destroy(sb2_svd_plan* p) { plan_free(p); }
// Use caller-owned fp32 factor buffers (rows x FP, FP = n_factors rounded up to 4): the multi-GPU ring
// keeps one user block and a rotating item block per rank in torch tensors and binds them per sub-epoch.
void svd_plan_bind(sb2_svd_plan* p, float* pu, float* qi, float* bu, float* bi) {
    if (p->owns_factors) { cudaFree(p->pu); cudaFree(p->qi); cudaFree(p->bu); cudaFree(p->bi); }
    p->owns_factors = false;
    p->pu = pu; p->qi = qi; p->bu = bu; p->bi = bi;
}
int svd_plan_stride(const sb2_svd_plan* p) { return p->FP; }
// algorithmic bytes per rating update: read + write pu[u], qi[i], bu[u], bi[i] at fp32 + (u, i, r)
int64_t svd_plan_bytes_per_update(const sb2_svd_plan* p) { return 2ll * (2ll * p->prm.n_factors + 2) * 4 + 12; }
void svd_plan_grid(const sb2_svd_plan* p, int* b, int* w) {
    if (b) *b = p->B;
    if (w) *w = p->W;
}
void svd_plan_dims(const sb2_svd_plan* p, int64_t* n_users, int64_t* n_items, int* f, int* with_yj) {
    *n_users = p->n_users; *n_items = p->n_items; *f = p->prm.n_factors; *with_yj = 0;
}

}  // namespace sb2
