/*
 * surprise_b200 -- C-ABI of the B200-native fit-time hot path of Surprise (nickmvincent/Surprise).
 *
 * The reference has no FFI of its own: its "operator API" for this path is (i) the module-level
 * callables of surprise/similarities.pyx looked up by name in
 * surprise/prediction_algorithms/algo_base.py:269-272 and (ii) the sgd()/estimate() methods of the
 * AlgoBase subclasses in surprise/prediction_algorithms/matrix_factorization.pyx.  Every entry point
 * below replaces one of those Cython functions and says which (file:line, relative to the reference
 * root).  INTEGRATION.md shows the ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  No torch / numpy types.
 *   - Every function comes in two forms:
 *       sb2_<name>_dev(...,  void* stream)  all array arguments are DEVICE pointers; work is enqueued
 *                                           on `stream` (a cudaStream_t; NULL = default stream) and the
 *                                           call returns after the stream has been synchronised only
 *                                           where a status must be read back (documented per call).
 *       sb2_<name>(...)                     all array arguments are HOST pointers; the library does
 *                                           the H2D / D2H copies itself (this is the e2e boundary).
 *   - Return value: SB2_OK or an error code; sb2_last_error() gives the message of the calling
 *     thread's last failure.  There is NO CPU fallback: without a usable CUDA device every compute
 *     entry point fails with SB2_ERR_CUDA.
 *   - Index types: int32 ids, int64 CSR offsets, fp64 ratings / factors at the boundary (the reference
 *     hands Python floats / float64 ndarrays).
 *   - "yr CSR": y_ptr[n_y+1], x_idx[nnz], r[nnz] -- the flattening of the reference's
 *     `yr` dict-of-lists in iteration order (for y in yr: for (x, r) in yr[y]).
 *   - "all_ratings COO": u[n], i[n], r[n] in trainset.all_ratings() order (trainset.py:180-190).
 */
#ifndef SURPRISE_B200_H
#define SURPRISE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    SB2_OK = 0,
    SB2_ERR_CUDA = 1,          /* CUDA runtime / driver failure (includes "no device") */
    SB2_ERR_INVALID = 2,       /* bad argument */
    SB2_ERR_ZERO_DIVISION = 3, /* the reference would raise ZeroDivisionError (cdivision=False) */
    SB2_ERR_DUPLICATE = 4,     /* duplicate (x, y) pair in yr: not representable in the dense panels */
    SB2_ERR_UNSUPPORTED = 5    /* ratings not representable on the exact integer-digit path */
};

enum { SB2_SIM_COSINE = 0, SB2_SIM_MSD = 1, SB2_SIM_PEARSON = 2, SB2_SIM_PEARSON_BASELINE = 3 };

/* library / device introspection */
const char* sb2_last_error(void);
int sb2_version(void);
int sb2_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem_bytes);
/* kernels launched by this library on the calling thread since the last reset (bench: gpu_launches) */
int64_t sb2_launch_count(void);
void sb2_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Similarity matrices.  Replaces similarities.pyx:28-97 (cosine), :100-166 (msd), :169-258
 * (pearson), :261-361 (pearson_baseline), as dispatched by algo_base.py:256-301.
 *
 * Ratings must be exact multiples of 1/rating_denom with r*rating_denom an integer in [0, 65535]
 * (MovieLens stars: denom 1, half-stars: 2, Jester two-decimals: 100); they are split into base-256
 * digits and contracted on the int8 tensor cores with exact int32 accumulation.  For cosine / msd /
 * pearson the result is bit-identical to the reference whenever all the reference's partial sums
 * are exactly representable (true for every integer / half-integer scale); pearson_baseline is
 * floating point in the reference too and matches it to ~1e-12 absolute (contract: 1e-9).
 *
 * Rows [row_begin, row_end) of the n_x x n_x matrix are produced into sim_out (row-major,
 * (row_end-row_begin) x n_x): the row-block shard of one rank (row_begin a multiple of 256, the tile of the
 * CTA-pair MMA kernel).  row_begin = 0, row_end = n_x builds
 * the whole matrix and exploits symmetry.  x_biases / y_biases / global_mean / shrinkage are only
 * read for SB2_SIM_PEARSON_BASELINE (min_support is clamped to >= 2 there, similarities.pyx:334).
 * The _dev form synchronises `stream` once to read the duplicate / zero-division status.
 * ------------------------------------------------------------------------------------------------ */
int sb2_sim_build_dev(int kind, int64_t n_x, int64_t n_y, const int64_t* y_ptr, const int32_t* x_idx,
                      const double* r, int64_t nnz, int rating_denom, int min_support, double global_mean,
                      const double* x_biases, const double* y_biases, double shrinkage, int64_t row_begin,
                      int64_t row_end, double* sim_out, void* stream);
int sb2_sim_build(int kind, int64_t n_x, int64_t n_y, const int64_t* y_ptr, const int32_t* x_idx,
                  const double* r, int64_t nnz, int rating_denom, int min_support, double global_mean,
                  const double* x_biases, const double* y_biases, double shrinkage, int64_t row_begin,
                  int64_t row_end, double* sim_out);

/* Test hook: C[m x n] (int32, row-major, ld = n) = A[m x k] * B[n x k]^T for u8 operands through the
 * tcgen05 kernel (use_tensor_cores = 1) or through the scalar dp4a cross-check kernel (0).
 * m, n multiples of 256, k multiple of 128.  Device pointers. */
int sb2_gemm_u8_selftest_dev(int use_tensor_cores, int64_t m, int64_t n, int64_t k, const uint8_t* a,
                             const uint8_t* b, int32_t* c, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Baselines.  Replaces prediction_algorithms/optimize_baselines.pyx:14-54 (baseline_als) and
 * :57-84 (baseline_sgd).  ALS is bit-exact (order-preserving segmented sums); SGD is a sequential
 * recursion and is executed as such (one thread), bit-exact but slow -- non-default in the reference.
 * ur CSR: u_ptr[n_users+1], ui_idx[n], u_r[n] (ur[u] list order); ir CSR likewise (ir[i] list order).
 * ------------------------------------------------------------------------------------------------ */
int sb2_baseline_als_dev(int64_t n_users, int64_t n_items, const int64_t* u_ptr, const int32_t* ui_idx,
                         const double* u_r, const int64_t* i_ptr, const int32_t* iu_idx, const double* i_r,
                         double global_mean, int n_epochs, double reg_u, double reg_i, double* bu, double* bi,
                         void* stream);
int sb2_baseline_als(int64_t n_users, int64_t n_items, const int64_t* u_ptr, const int32_t* ui_idx,
                     const double* u_r, const int64_t* i_ptr, const int32_t* iu_idx, const double* i_r,
                     double global_mean, int n_epochs, double reg_u, double reg_i, double* bu, double* bi);
/* One half-epoch of baseline_als (optimize_baselines.pyx:41-46 items / :48-53 users) over the segments
 * [seg_begin, seg_end) of one side's CSR: mine[s] = sum_{a in seg s} (r[a] - mu - other[idx[a]]) / (reg + |seg s|).
 * The unit of the multi-rank ALS: every ordered sum is evaluated on exactly one rank (bit-identical to the
 * single-GPU result), the caller all-gathers `mine` between the passes.  status_dev: DEVICE int, set to 1 where the
 * reference would raise ZeroDivisionError. */
int sb2_baseline_als_pass_dev(int64_t seg_begin, int64_t seg_end, const int64_t* ptr, const int32_t* idx, const double* r,
                              const double* other, double* mine, double global_mean, double reg, int* status_dev,
                              void* stream);
int sb2_baseline_sgd_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                         const double* r, double global_mean, int n_epochs, double reg, double lr, double* bu,
                         double* bi, void* stream);
int sb2_baseline_sgd(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                     const double* r, double global_mean, int n_epochs, double reg, double lr, double* bu,
                     double* bi);

/* ------------------------------------------------------------------------------------------------
 * SVD.  Replaces SVD.sgd, matrix_factorization.pyx:172-267 (hot loop :241-262).
 * pu (n_users x f) and qi (n_items x f) hold the rng.normal initialisation on entry (:233-236, made
 * on the host with numpy so that seeds agree) and the fitted factors on exit; bu / bi are outputs.
 * `global_mean` is the trainset mean; when biased == 0 it is ignored and biases stay 0 (:238, :253).
 * Update order is stratified (DSGD) and arithmetic is fp32: results are judged on held-out
 * RMSE / MAE (|diff| <= 0.005), not element-wise.
 * ------------------------------------------------------------------------------------------------ */
typedef struct sb2_sgd_params {
    int32_t n_factors;
    int32_t n_epochs;
    int32_t biased;
    int32_t reserved;
    double global_mean;
    double lr_bu, lr_bi, lr_pu, lr_qi, lr_yj;
    double reg_bu, reg_bi, reg_pu, reg_qi, reg_yj;
} sb2_sgd_params;

int sb2_svd_fit_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                    const double* r, const sb2_sgd_params* prm, double* pu, double* qi, double* bu, double* bi,
                    void* stream);
int sb2_svd_fit(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                const double* r, const sb2_sgd_params* prm, double* pu, double* qi, double* bu, double* bi);

/* Resident-input form used by the benchmark: prepare once (stratify + upload), then run epochs with
 * everything already in HBM.  handle lifetime: create -> (reset ->) run* -> read -> destroy. */
typedef struct sb2_svd_plan sb2_svd_plan;
int sb2_svd_plan_create(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u_host,
                        const int32_t* i_host, const double* r_host, const sb2_sgd_params* prm, int with_yj,
                        sb2_svd_plan** out);
int sb2_svd_plan_reset(sb2_svd_plan* plan, const double* pu_host, const double* qi_host,
                       const double* yj_host);
/* device-pointer forms of create / reset / read (u_ptr / ui_idx: ur CSR, only read when with_yj) */
int sb2_svd_plan_create_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                            const double* r, const sb2_sgd_params* prm, int with_yj, const int64_t* u_ptr,
                            const int32_t* ui_idx, void* stream, sb2_svd_plan** out);
int sb2_svd_plan_reset_dev(sb2_svd_plan* plan, const double* pu, const double* qi, const double* yj,
                           void* stream);
int sb2_svd_plan_read_dev(sb2_svd_plan* plan, double* pu, double* qi, double* bu, double* bi, double* yj,
                          void* stream);
int sb2_svd_plan_run(sb2_svd_plan* plan, int n_epochs, void* stream); /* async on stream */
int sb2_svd_plan_read(sb2_svd_plan* plan, double* pu, double* qi, double* bu, double* bi, double* yj);
void sb2_svd_plan_destroy(sb2_svd_plan* plan);
/* algorithmic bytes per rating update of the plan's kernel (2*(2f+2)*4 + 12 for SVD) */
int64_t sb2_svd_plan_bytes_per_update(const sb2_svd_plan* plan);
int sb2_svd_plan_grid(const sb2_svd_plan* plan, int* n_blocks, int* n_sub);
/* Synchronises `stream` and reports whether the persistent kernel ran to completion: every wait of a CTA for a
 * neighbour CTA / rank is bounded (~6 s); if a partner was never scheduled (SMs held by another persistent kernel,
 * a dead peer rank) the launch drains and this returns SB2_ERR_CUDA instead of the GPU hanging.  The host forms
 * (sb2_svd_fit, sb2_svdpp_fit) check it themselves. */
int sb2_svd_plan_status(sb2_svd_plan* plan, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU ring (DSGD strata across ranks, one process per GPU; SURVEY.md 8e "SVD (DSGD)").  Replaces the same
 * loop, matrix_factorization.pyx:241-262 (SVD) / :466-498 (SVD++), for a trainset sharded by user.
 * Rank g of `world` owns the users u % world == g for the whole fit and, in sub-epoch E, the item super-block
 * (g + E) % world; the kernel itself hands every finished item block to rank g - 1 over NVLink (stores through the
 * peer mapping + system-scope release / acquire flags) -- ONE persistent launch per rank for a whole SVD fit.
 *   create:  every rank passes the SAME global COO (DEVICE arrays) and its (rank, world); it keeps its own ratings.
 *   connect: exchange the 64-byte handles between the processes (e.g. torch.distributed.all_gather) and pass the
 *            left (rank - 1) and right (rank + 1) neighbours' handles; or, inside one process, the plans themselves.
 *   reset:   sb2_svd_plan_reset_dev with the WHOLE initial matrices; then synchronise the ranks (any collective)
 *            before run: reset clears the mailboxes the neighbours write into.
 *   run:     sb2_svd_plan_run (SVD), or per epoch sb2_svd_ring_epoch_dev(phase 0) -> all-reduce(SUM) of `exchange`
 *            (n_items x (row_stride + 1) fp32, DEVICE) over the ranks -> sb2_svd_ring_epoch_dev(phase 1) (SVD++: the
 *            per-item sums of the y_j application span the users of all ranks).
 *   read:    sb2_svd_plan_read_dev returns the rank's OWN rows (n_users_local x f, n_items_local x f, local order:
 *            global id = rank + local * world); synchronise the ranks before reading / destroying.
 * ------------------------------------------------------------------------------------------------ */
int sb2_svd_ring_create_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                            const double* r, const sb2_sgd_params* prm, int with_yj, const int64_t* u_ptr,
                            const int32_t* ui_idx, int rank, int world, void* stream, sb2_svd_plan** out);
int sb2_svd_ring_ipc_handle(const sb2_svd_plan* plan, unsigned char* handle64);
int sb2_svd_ring_connect_ipc(sb2_svd_plan* plan, const unsigned char* left64, const unsigned char* right64);
int sb2_svd_ring_connect_local(sb2_svd_plan* plan, sb2_svd_plan* left, sb2_svd_plan* right);
int sb2_svd_ring_epoch_dev(sb2_svd_plan* plan, int phase, float* exchange, void* stream);
/* All ranks of a ring in ONE process on ONE GPU (the single-GPU tests of the ring): plans[0 .. n) are the ranks
 * 0 .. n-1 (n <= 4), connected with sb2_svd_ring_connect_local.  Kernels that wait for one another must never be
 * issued as separate launches on one GPU, so all ranks run as one launch of n x B co-resident CTAs -- the same
 * kernel body, the "peer" buffers being the other ranks' buffers in the same memory.  sb2_svd_ring_run_local: n_epochs
 * SVD epochs; sb2_svd_ring_epoch_local: phase 0 of one SVD++ epoch for every rank (exchange[g]: rank g's partial
 * sums; add them up, then call sb2_svd_ring_epoch_dev(plans[g], 1, sum) per rank). */
int sb2_svd_ring_run_local(sb2_svd_plan** plans, int n_plans, int n_epochs, void* stream);
int sb2_svd_ring_epoch_local(sb2_svd_plan** plans, int n_plans, float** exchange, void* stream);
int sb2_svd_ring_info(const sb2_svd_plan* plan, int64_t* n_users_local, int64_t* n_items_local,
                      int64_t* n_ratings_local, int* row_stride);
/* Per-CTA counters of the last run, cycles_host[n_blocks][8] (n_blocks from sb2_svd_plan_grid): SM cycles in
 * {ring wait, item-block load, rating updates, write-back, lane-group-0 updates}, then the number of
 * lane-group-0 updates, the number of waves, 0. */
int sb2_svd_plan_profile(const sb2_svd_plan* plan, int64_t* cycles_host);

/* SVD++.  Replaces SVDpp.sgd, matrix_factorization.pyx:420-504.  yj (n_items x f) in/out like qi.
 * u_ptr / ui_idx: the ur CSR (I_u).  See DESIGN.md for the per-user batching of the y_j update. */
int sb2_svdpp_fit_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                      const double* r, const int64_t* u_ptr, const int32_t* ui_idx, const sb2_sgd_params* prm,
                      double* pu, double* qi, double* yj, double* bu, double* bi, void* stream);
int sb2_svdpp_fit(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                  const double* r, const int64_t* u_ptr, const int32_t* ui_idx, const sb2_sgd_params* prm,
                  double* pu, double* qi, double* yj, double* bu, double* bi);

/* ------------------------------------------------------------------------------------------------
 * NMF.  Replaces NMF.sgd, matrix_factorization.pyx:646-735.  Bit-exact (fp64, the reference's
 * summation order, no FMA) for biased == 0 and biased != 0.  (u, i, r) is the all_ratings COO, which
 * must be grouped by u (it is, by construction of all_ratings()).  pu / qi: rng.uniform init in,
 * factors out.  Returns SB2_ERR_ZERO_DIVISION where the reference raises (:723, :730).
 * ------------------------------------------------------------------------------------------------ */
typedef struct sb2_nmf_params {
    int32_t n_factors;
    int32_t n_epochs;
    int32_t biased;
    int32_t reserved;
    double global_mean;
    double reg_pu, reg_qi, reg_bu, reg_bi, lr_bu, lr_bi;
} sb2_nmf_params;

int sb2_nmf_fit_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                    const double* r, const sb2_nmf_params* prm, double* pu, double* qi, double* bu, double* bi,
                    void* stream);
int sb2_nmf_fit(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                const double* r, const sb2_nmf_params* prm, double* pu, double* qi, double* bu, double* bi);

/* Epoch-at-a-time form for the multi-GPU fit (device pointers).  The plan holds what depends only on the rating
 * structure; the COO arrays are borrowed and must outlive it.  One epoch restricted to users
 * [user_begin, user_end) and items [item_begin, item_end) reads the full pu_cur / qi_cur and writes those rows
 * of pu_new / qi_new: every ordered accumulator sum is evaluated on exactly one rank, so sharded results stay
 * bit-identical; the caller all-gathers the new factors between epochs.  bu / bi (biased model) are advanced in
 * place by the sequential recursion, which every rank replays.  sb2_nmf_plan_status synchronises the stream and
 * reports SB2_ERR_ZERO_DIVISION / SB2_ERR_INVALID (input not grouped by user). */
typedef struct sb2_nmf_plan sb2_nmf_plan;
int sb2_nmf_plan_create_dev(int64_t n_users, int64_t n_items, int64_t n, const int32_t* u, const int32_t* i,
                            const double* r, int n_factors, void* stream, sb2_nmf_plan** out);
int sb2_nmf_plan_epoch_dev(sb2_nmf_plan* plan, const sb2_nmf_params* prm, const double* pu_cur, const double* qi_cur,
                           double* pu_new, double* qi_new, double* bu, double* bi, int64_t user_begin,
                           int64_t user_end, int64_t item_begin, int64_t item_end, void* stream);
int sb2_nmf_plan_status(sb2_nmf_plan* plan, void* stream);
void sb2_nmf_plan_destroy(sb2_nmf_plan* plan);

/* ------------------------------------------------------------------------------------------------
 * Batched estimate for the factor models.  Replaces SVD.estimate (matrix_factorization.pyx:269-299),
 * NMF.estimate (:737-761) and SVDpp.estimate (:506-522; yj != NULL) called once per pair by
 * AlgoBase.test (algo_base.py:191-218).  u[k] < 0 / i[k] < 0 encode an unknown user / item.
 * impossible[k] = 1 where the reference raises PredictionImpossible.
 * n_factors = 0 with biased = 1 is BaselineOnly.estimate (baseline_only.py:33-46): (mu + bu[u]) + bi[i], pu / qi unused.
 * ------------------------------------------------------------------------------------------------ */
int sb2_mf_predict_dev(int64_t n_pairs, const int32_t* u, const int32_t* i, int n_factors, int biased,
                       double global_mean, const double* pu, const double* qi, const double* bu,
                       const double* bi, const double* yj, const int64_t* u_ptr, const int32_t* ui_idx,
                       double* est, uint8_t* impossible, void* stream);
int sb2_mf_predict(int64_t n_pairs, const int32_t* u, const int32_t* i, int64_t n_users, int64_t n_items,
                   int n_factors, int biased, double global_mean, const double* pu, const double* qi,
                   const double* bu, const double* bi, const double* yj, const int64_t* u_ptr,
                   const int32_t* ui_idx, double* est, uint8_t* impossible);

/* ------------------------------------------------------------------------------------------------
 * Batched k-NN estimate.  Replaces KNNBasic.estimate (knns.py:99-123) and KNNBaseline.estimate
 * (:274-309): gather sim[x, x2] over yr[y], stable top-k (heapq.nlargest semantics), ordered fp64
 * weighted sum -- bit-identical to the reference.
 * mode: 0 = KNNBasic; 1 = KNNBaseline with x = user (user_based); 2 = KNNBaseline with x = item;
 *       3 = KNNWithMeans (knns.py:178-208; bx = means[n_x], by unused); 4 = KNNWithZScore (:372-403; bx = means,
 *       by = sigmas, both indexed by x).
 * x[k] / y[k] < 0 = unknown.  actual_k[k] = -1 where the reference reports no actual_k.
 * impossible[k]: 1 = PredictionImpossible, 2 = the reference would raise ZeroDivisionError.
 * sim: n_x x n_x row-major with leading dimension sim_ld (>= n_x).
 * ------------------------------------------------------------------------------------------------ */
int sb2_knn_predict_dev(int64_t n_pairs, const int32_t* x, const int32_t* y, int64_t n_x, const double* sim,
                        int64_t sim_ld, const int64_t* y_ptr, const int32_t* x_idx, const double* r, int k,
                        int min_k, int mode, double global_mean, const double* bx, const double* by,
                        double* est, int32_t* actual_k, uint8_t* impossible, void* stream);
int sb2_knn_predict(int64_t n_pairs, const int32_t* x, const int32_t* y, int64_t n_x, int64_t n_y,
                    const double* sim, const int64_t* y_ptr, const int32_t* x_idx, const double* r, int k,
                    int min_k, int mode, double global_mean, const double* bx, const double* by, double* est,
                    int32_t* actual_k, uint8_t* impossible);

/* ------------------------------------------------------------------------------------------------
 * SlopeOne ("next" row 4 of the hot-path table).  sb2_slope_one_fit replaces the two loops of
 * SlopeOne.fit (prediction_algorithms/slope_one.pyx:59-70): freq[i][j] = number of users who rated both
 * items (n_items x n_items int64; diagonal = raters of i), dev[i][j] = mean over those users of
 * r_ui - r_uj with the ratings TRUNCATED to C ints exactly as the reference's `cdef int r_ui, r_uj` does
 * (slope_one.pyx:52); 0 / 0 = NaN where freq is 0; dev[j][i] = -dev[i][j]; dev[i][i] = 0.  Both are
 * by-products of the similarity contractions (freq = M M^T, sum r_ui = R M^T) on the tensor cores and are
 * bit-identical to the reference.  u_ptr / i_idx / r: the ur CSR (trainset.ur flattened, user-major).
 * sb2_slope_one_predict replaces SlopeOne.estimate (:82-97); user_mean[u] is computed by the host mirror.
 * impossible[k] = 1 where the reference raises PredictionImpossible (u[k] < 0 or i[k] < 0).
 * ------------------------------------------------------------------------------------------------ */
int sb2_slope_one_fit_dev(int64_t n_items, int64_t n_users, const int64_t* u_ptr, const int32_t* i_idx,
                          const double* r, int64_t nnz, int64_t* freq_out, double* dev_out, void* stream);
int sb2_slope_one_fit(int64_t n_items, int64_t n_users, const int64_t* u_ptr, const int32_t* i_idx,
                      const double* r, int64_t* freq_out, double* dev_out);
int sb2_slope_one_predict_dev(int64_t n_pairs, const int32_t* u, const int32_t* i, int64_t n_items,
                              const int64_t* freq, const double* dev, const int64_t* u_ptr, const int32_t* i_idx,
                              const double* user_mean, double* est, uint8_t* impossible, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AlgoBase.get_neighbors (algo_base.py:303-334) on the device-resident similarity matrix: for each
 * requested row x = rows[b], the k other ids with the largest sim[x, .], in the order of the reference's
 * stable descending sort (ties by ascending id).  out: n_rows x k int32, padded with -1 when fewer than
 * k other ids exist (the reference returns a shorter list) or rows[b] is out of range.
 * ------------------------------------------------------------------------------------------------ */
int sb2_get_neighbors_dev(int64_t n_x, const double* sim, int64_t sim_ld, int64_t n_rows, const int32_t* rows,
                          int k, int32_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row shard of a SYMMETRIC multi-rank similarity build (SURVEY.md 8e "cyclic upper-triangular tile sharding").
 * Same arguments as sb2_sim_build_dev, but only sim[i][j] with row_begin <= i < row_end and j >= row_begin is
 * computed (tiles at or above the block diagonal; the shard's own diagonal square is mirrored like the
 * reference mirrors, similarities.pyx:95).  Columns j < row_begin of sim_out are left untouched: they are
 * the transposes of blocks computed by the shards before this one (surprise_b200/distributed.py exchanges
 * them over NCCL).  With row ranges of equal triangular area every rank does 1/N of the single-GPU work.
 * ------------------------------------------------------------------------------------------------ */
int sb2_sim_build_upper_dev(int kind, int64_t n_x, int64_t n_y, const int64_t* y_ptr, const int32_t* x_idx,
                            const double* r, int64_t nnz, int rating_denom, int min_support, double global_mean,
                            const double* x_biases, const double* y_biases, double shrinkage, int64_t row_begin,
                            int64_t row_end, double* sim_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * rating_denom for sb2_sim_build*: the smallest d in {1, 2, 4, 5, 10, 20, 100, 1000} such that every rating
 * (DEVICE array) is an integer multiple of 1/d within [0, 65535/d] (relative tolerance 1e-9, the same test
 * sb2_sim_build applies); *denom_out (HOST int) = 0 when the ratings lie on none of these grids.
 * ------------------------------------------------------------------------------------------------ */
int sb2_rating_denominator_dev(const double* r, int64_t nnz, int* denom_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SURPRISE_B200_H */
