// tcgen05 / TMA / TMEM u8 GEMM for the similarity contractions (see sim_gemm.cuh).
//
// Roles (one CTA per SM, persistent over a host-built tile list):
//   warp 0 lane 0 : TMA producer  -- cp.async.bulk.tensor.2d tiles of 128 rows x 64 B (SWIZZLE_64B)
//   warp 1 lane 0 : MMA issuer    -- tcgen05.mma.cta_group::1.kind::i8, M=128 N=128 K=32, D in TMEM
//   warps 2..5    : epilogue      -- tcgen05.ld 32x32b.x32 -> fp64 fold into the output planes
// Pipelines: smem ring (full/empty mbarriers, TMA <-> MMA) and one TMEM hand-off (MMA <-> epilogue).
#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>

#include "sim_gemm.cuh"

namespace sb2 {

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

// K-major, SWIZZLE_64B shared-memory matrix descriptor (sm_100 format, cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 = 8 rows * 64 B
//   [46,48) version = 1 | [61,64) layout = 4 (SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) @[4,6), a/b format U8 (0)
// @[7,10)/[10,13), K-major both, N>>3 @[17,23), M>>4 @[24,29).
__host__ __device__ constexpr uint32_t make_idesc_u8(int m, int n) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct GemmSmemLayout {
    // [stages][na+nb] tiles of 8 KB, then barriers
    static __host__ __device__ size_t tiles_bytes(int stages, int npanels) {
        return (size_t)stages * npanels * GEMM_TILE_BYTES;
    }
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_u8_tc_kernel(const __grid_constant__ GemmParams p, const int2* __restrict__ tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem base is only guaranteed 16 B aligned: round up to 1024 for the swizzled tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int npanels = p.na + p.nb;
    const int stages = p.stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GemmSmemLayout::tiles_bytes(stages, npanels));
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* tmem_full_bar = empty_bar + stages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        // 512 columns: 4 accumulators x 128 columns (fp32/int32 lanes)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int nk = p.num_k_blocks;
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                const int2 tile = tiles[t];
                for (int kb = 0; kb < nk; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(npanels * GEMM_TILE_BYTES));
                    uint8_t* sbase = smem + (size_t)stage * npanels * GEMM_TILE_BYTES;
                    for (int a = 0; a < p.na; ++a)
                        tma_load_2d(&p.a_maps[a], &full_bar[stage], sbase + a * GEMM_TILE_BYTES, kb * GEMM_BK,
                                    tile.x * GEMM_BM);
                    for (int b = 0; b < p.nb; ++b)
                        tma_load_2d(&p.b_maps[b], &full_bar[stage], sbase + (p.na + b) * GEMM_TILE_BYTES,
                                    kb * GEMM_BK, tile.y * GEMM_BN);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_u8(GEMM_BM, GEMM_BN);
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                mbar_wait(tmem_empty_bar, tphase ^ 1);  // epilogue has drained the accumulators
                tc_fence_after();
                for (int kb = 0; kb < nk; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(smem + (size_t)stage * npanels * GEMM_TILE_BYTES);
                    for (int j = 0; j < p.nacc; ++j) {
                        const uint32_t sa = sbase + p.acc_a[j] * GEMM_TILE_BYTES;
                        const uint32_t sb = sbase + (p.na + p.acc_b[j]) * GEMM_TILE_BYTES;
#pragma unroll
                        for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
                            const uint64_t ad = make_desc_sw64(sa + k * GEMM_UMMA_K);
                            const uint64_t bd = make_desc_sw64(sb + k * GEMM_UMMA_K);
                            tc_mma_i8(tmem_base + j * GEMM_BN, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    tc_commit(&empty_bar[stage]);  // frees the smem slot when the MMAs above retire
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(tmem_full_bar);
                tphase ^= 1;
            }
        }
    } else {
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        uint32_t tphase = 0;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            const int2 tile = tiles[t];
            mbar_wait(tmem_full_bar, tphase);
            tc_fence_after();
            const long long row = (long long)tile.x * GEMM_BM + q * 32 + lane - p.plane_row0;
            for (int j = 0; j < p.nacc; ++j) {
                double* orow = p.out[j] + row * p.ld + (long long)tile.y * GEMM_BN;
                const double alpha = p.alpha[j];
                const int beta = p.beta[j];
#pragma unroll 1
                for (int c0 = 0; c0 < GEMM_BN; c0 += 32) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * GEMM_BN + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    double2* o2 = reinterpret_cast<double2*>(orow + c0);
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        double2 w;
                        w.x = alpha * (double)(int)v[2 * e];
                        w.y = alpha * (double)(int)v[2 * e + 1];
                        if (beta) {
                            const double2 old = o2[e];
                            w.x = __dadd_rn(old.x, w.x);
                            w.y = __dadd_rn(old.y, w.y);
                        }
                        o2[e] = w;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
            tphase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// ----------------------------------------------------------------------------------------------
// host: tensor maps + batching + launch
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

static int make_panel_map(CUtensorMap* map, const uint8_t* panel, int64_t rows, int64_t k_pad) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return SB2_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)k_pad};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)GEMM_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(panel), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld k_pad=%lld", (int)r, (long long)rows, (long long)k_pad);
        return SB2_ERR_CUDA;
    }
    return SB2_OK;
}

static int launch_batch(GemmParams& prm, const int2* tiles_dev, cudaStream_t st) {
    const int npanels = prm.na + prm.nb;
    int stages = (int)((200 * 1024) / ((size_t)npanels * GEMM_TILE_BYTES));
    if (stages > 8) stages = 8;
    if (stages < 2) {
        set_error("gemm: too many panels per stage");
        return SB2_ERR_INVALID;
    }
    prm.stages = stages;
    const size_t smem = GemmSmemLayout::tiles_bytes(stages, npanels) + (2 * stages + 2) * sizeof(uint64_t) + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        SB2_CUDA(cudaFuncSetAttribute(gemm_u8_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    int grid = sm_count();
    if (grid > prm.n_tiles) grid = prm.n_tiles;
    if (grid < 1) return SB2_OK;
    gemm_u8_tc_kernel<<<grid, GEMM_THREADS, smem, st>>>(prm, tiles_dev);
    SB2_LAUNCH_CHECK();
    return SB2_OK;
}

int gemm_u8_tc_run2(const GemmJob* jobs, int n_jobs, int64_t a_rows, int64_t b_rows, int64_t k_pad,
                    const int2* tiles_dev, int n_tiles, int64_t ld, int64_t plane_row0, cudaStream_t st) {
    if (k_pad % GEMM_BK || a_rows % GEMM_BM || b_rows % GEMM_BN) {
        set_error("gemm: unpadded shape");
        return SB2_ERR_INVALID;
    }
    int j = 0;
    while (j < n_jobs) {
        GemmParams prm;
        memset(&prm, 0, sizeof(prm));
        const uint8_t* ap[GEMM_MAX_PANELS];
        const uint8_t* bp[GEMM_MAX_PANELS];
        int na = 0, nb = 0, nacc = 0;
        while (j < n_jobs && nacc < GEMM_MAX_ACC) {
            int ia = -1, ib = -1;
            for (int t = 0; t < na; ++t)
                if (ap[t] == jobs[j].a) ia = t;
            for (int t = 0; t < nb; ++t)
                if (bp[t] == jobs[j].b) ib = t;
            if ((ia < 0 && na == GEMM_MAX_PANELS) || (ib < 0 && nb == GEMM_MAX_PANELS)) break;
            // keep a stage at <= 6 tiles (48 KB) so the ring stays >= 4 deep
            const int add = (ia < 0) + (ib < 0);
            if (nacc > 0 && na + nb + add > 6) break;
            if (ia < 0) { ia = na; ap[na++] = jobs[j].a; }
            if (ib < 0) { ib = nb; bp[nb++] = jobs[j].b; }
            prm.acc_a[nacc] = ia;
            prm.acc_b[nacc] = ib;
            prm.out[nacc] = jobs[j].out;
            prm.alpha[nacc] = jobs[j].alpha;
            prm.beta[nacc] = jobs[j].beta;
            ++nacc;
            ++j;
        }
        for (int t = 0; t < na; ++t) SB2_TRY(make_panel_map(&prm.a_maps[t], ap[t], a_rows, k_pad));
        for (int t = 0; t < nb; ++t) SB2_TRY(make_panel_map(&prm.b_maps[t], bp[t], b_rows, k_pad));
        prm.na = na;
        prm.nb = nb;
        prm.nacc = nacc;
        prm.num_k_blocks = (int)(k_pad / GEMM_BK);
        prm.n_tiles = n_tiles;
        prm.ld = ld;
        prm.plane_row0 = plane_row0;
        SB2_TRY(launch_batch(prm, tiles_dev, st));
    }
    return SB2_OK;
}

int gemm_u8_tc_run(const GemmJob* jobs, int n_jobs, int64_t n_pad_rows, int64_t k_pad, const int2* tiles_dev,
                   int n_tiles, int64_t ld, int64_t plane_row0, cudaStream_t st) {
    return gemm_u8_tc_run2(jobs, n_jobs, n_pad_rows, n_pad_rows, k_pad, tiles_dev, n_tiles, ld, plane_row0, st);
}

// ----------------------------------------------------------------------------------------------
// test hooks
// ----------------------------------------------------------------------------------------------
__global__ void gemm_u8_dp4a_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int32_t* __restrict__ c,
                                    int64_t m, int64_t n, int64_t k) {
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.y;
    if (col >= n || row >= m) return;
    const uint32_t* ar = reinterpret_cast<const uint32_t*>(a + row * k);
    const uint32_t* br = reinterpret_cast<const uint32_t*>(b + col * k);
    unsigned acc = 0;
    for (int64_t t = 0; t < k / 4; ++t) acc = __dp4a(ar[t], br[t], acc);
    c[row * n + col] = (int32_t)acc;
}

__global__ void plane_to_i32_kernel(const double* __restrict__ p, int32_t* __restrict__ c, int64_t n) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) c[t] = (int32_t)p[t];
}

int gemm_u8_selftest(int use_tc, int64_t m, int64_t n, int64_t k, const uint8_t* a, const uint8_t* b, int32_t* c,
                     cudaStream_t st) {
    if (m % GEMM_BM || n % GEMM_BN || k % GEMM_BK) {
        set_error("selftest: m, n multiples of 128 and k multiple of 64 required");
        return SB2_ERR_INVALID;
    }
    if (!use_tc) {
        dim3 grid((unsigned)ceil_div(n, 128), (unsigned)m);
        gemm_u8_dp4a_kernel<<<grid, 128, 0, st>>>(a, b, c, m, n, k);
        SB2_LAUNCH_CHECK();
        return SB2_OK;
    }
    std::vector<int2> tiles;
    for (int r = 0; r < m / GEMM_BM; ++r)
        for (int cc = 0; cc < n / GEMM_BN; ++cc) tiles.push_back(make_int2(r, cc));
    DevBuf tiles_d, plane;
    SB2_TRY(tiles_d.alloc(tiles.size() * sizeof(int2), st));
    SB2_CUDA(cudaMemcpyAsync(tiles_d.p, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    SB2_TRY(plane.alloc((size_t)m * n * sizeof(double), st));
    GemmJob job{a, b, plane.as<double>(), 1.0, 0};
    SB2_TRY(gemm_u8_tc_run2(&job, 1, m, n, k, tiles_d.as<int2>(), (int)tiles.size(), n, 0, st));
    plane_to_i32_kernel<<<(unsigned)ceil_div(m * n, 256), 256, 0, st>>>(plane.as<double>(), c, m * n);
    SB2_LAUNCH_CHECK();
    SB2_CUDA(cudaStreamSynchronize(st));  // tiles vector must outlive the async copy
    return SB2_OK;
}

}  // namespace sb2
