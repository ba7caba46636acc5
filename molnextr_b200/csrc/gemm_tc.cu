// bf16 GEMM on the 5th-generation tensor cores of sm_100a, hand-written:
//   * operands staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 5-stage ring,
//   * one elected thread issues tcgen05.mma (UMMA 128 x 128 x 16, kind::f16, bf16 in / fp32 out),
//   * the accumulator lives in TMEM (128 lanes x 128 columns) and is read back with tcgen05.ld,
//   * four epilogue warps apply bias / GELU / layer-scale / residual-add (optionally through a
//     row map that undoes Swin's window partition + cyclic shift) and write straight to HBM.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue
// (two warps per TMEM lane quarter, 64 accumulator columns each).
// Persistent: one CTA per SM walks over 128 x 128 output tiles (n fastest, so CTAs working at the same
// time share A rows through L2); the accumulator is double-buffered in TMEM (2 x 128 columns) so the
// epilogue of tile i overlaps the TMA / MMA main loop of tile i+1; a 5-stage operand ring (160 KB).
#include "gemm_tc.cuh"

#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace mnx {

// Two tile shapes: 128 x 128 (5 stages) and 128 x 256 (4 stages; used when N % 256 == 0).  With 128 x 128 tiles a CTA
// pulls 256 B of operands per 8192 MACs -- 128 B/clk/SM at the tensor peak, more than L2 delivers with all SMs
// pulling (~50 B/clk/SM) -- so the wide tile (384 B per 16384 MACs, 96 B/clk) is ~1.3x faster on the L2-bound GEMMs.
static constexpr int BM = 128, BK = 64;
static constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
template <int BN_> struct GemmCfg {
    static constexpr int BN = BN_;
    static constexpr int STAGES = (BN_ == 128) ? 5 : 4;
    static constexpr int B_STAGE_BYTES = BN_ * BK * 2;   // 16 / 32 KB
    // operand ring | barriers (1 KB slot) | 8 x 4 KB epilogue staging tiles (TMA stores), all 1024-byte aligned
    static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 1024 /*barriers*/ + 8 * 4096;
    static constexpr uint32_t TMEM_COLS = 2 * BN_;       // two accumulators
};
static constexpr int GEMM_THREADS = 320;   // producer warp + MMA warp + 8 epilogue warps

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]; both operands K-major, described by 64-bit shared-memory descriptors
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile whose rows are 128 bytes (64 bf16): 8-row groups are
// 1024 bytes apart (SBO), LBO is unused for swizzled K-major layouts, descriptor version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, 16-byte units     bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (ignored)    bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset = 1024 B      bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)   bits [46,48)
    d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B        bits [61,64)
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M x N tile
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct GemmKernelArgs {
    int M, N, K;
    int epilogue;
    const float* bias;
    const float* gamma;
    const int* row_map;
    void* out;
    const float* colsum;
    const float* ln_stats;
    int ln_splits;
    float ln_eps;
    int tma_out;   // 1: results leave through shared memory and TMA (bf16 tile stores / fp32 reduce-add into the residual stream)
};

// ---- epilogue through TMA -------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// GELU(erf) for the bf16-output epilogue, two elements per instruction on packed fp32 pairs and WITHOUT the special
// function unit: gelu(x) = 0.5 x + 0.5 |x| erf(|x| / sqrt 2), erf(z) = z P(2 z^2 / 9 - 1) on z <= 3 (degree-8 least-squares
// fit constrained to erf(3) := 1, so the far negative tail is exactly 0), z clamped at 3.  |gelu error| < 5e-5 (maximum
// at |x| = 4.2), far below the 2^-9 relative rounding of the bf16 result.  The previous Abramowitz-Stegun form cost one
// MUFU.RCP and one MUFU.EX2 per element (16 lanes per clock per SM): 128 x 128 outputs took ~5100 cycles per tile
// against ~2050 cycles of tcgen05 main loop at K = 512, i.e. the fc1 GEMMs ran at epilogue speed.
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ unsigned long long f2_splat(float v) { return f2_pack(v, v); }
__device__ __forceinline__ void gelu_erf_pair(float& x0, float& x1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const unsigned long long z = f2_pack(fminf(a0 * 0.70710678118654752440f, 3.0f), fminf(a1 * 0.70710678118654752440f, 3.0f));
    const unsigned long long u = f2_fma(f2_mul(z, z), f2_splat(2.0f / 9.0f), f2_splat(-1.0f));
    unsigned long long p = f2_splat(6.902202582e-03f);
    p = f2_fma(p, u, f2_splat(-1.792530945e-02f));
    p = f2_fma(p, u, f2_splat(2.403749889e-02f));
    p = f2_fma(p, u, f2_splat(-3.965777946e-02f));
    p = f2_fma(p, u, f2_splat(7.213525925e-02f));
    p = f2_fma(p, u, f2_splat(-1.111660958e-01f));
    p = f2_fma(p, u, f2_splat(1.576062722e-01f));
    p = f2_fma(p, u, f2_splat(-2.287270545e-01f));
    p = f2_fma(p, u, f2_splat(4.701283396e-01f));
    const unsigned long long e = f2_mul(z, p);                                     // erf(|x| / sqrt 2)
    const unsigned long long r = f2_fma(f2_pack(0.5f * a0, 0.5f * a1), e, f2_pack(0.5f * x0, 0.5f * x1));
    x0 = __uint_as_float((unsigned)(r & 0xffffffffull));
    x1 = __uint_as_float((unsigned)(r >> 32));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_out, GemmKernelArgs g) {
    constexpr int STAGES = GemmCfg<BN>::STAGES, B_STAGE_BYTES = GemmCfg<BN>::B_STAGE_BYTES;
    constexpr uint32_t TMEM_COLS = GemmCfg<BN>::TMEM_COLS;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled TMA / UMMA tiles need 1024-byte aligned bases
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;     // [2] accumulator ready for the epilogue
    uint64_t* acc_empty = acc_full + 2;          // [2] accumulator drained by the 4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint8_t* stage_out = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024;     // [8 epilogue warps][32 rows][128 B], 128-B swizzle

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = g.K / BK;
    const int n_tiles = g.N / BN;
    const int m_tiles = (g.M + BM - 1) / BM;
    const int total_tiles = n_tiles * m_tiles;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
        if (g.tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_out)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;   // running k-block counter over all tiles of this CTA
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile % n_tiles, m_tile = tile / n_tiles;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);     // slot free (first lap passes immediately)
                    mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + B_STAGE_BYTES);
                    tma_load_2d(smem_a + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BK, m_tile * BM);
                    tma_load_2d(smem_b + s * B_STAGE_BYTES, &tmap_w, &full_bar[s], kb * BK, n_tile * BN);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, BN);
            uint32_t it = 0, j = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
                const uint32_t acc = j & 1u;
                mbar_wait(&acc_empty[acc], ((j >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
                    const uint64_t b_desc = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advancing 16 bf16 (32 bytes) along K inside the swizzle atom = +2 in 16-byte units
                        umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);            // frees the smem slot when these MMAs retire
                }
                umma_commit(&acc_full[acc]);               // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> HBM =====================
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int chalf = (warp - 2) >> 2;             // which half of the accumulator columns
        constexpr int CPW = BN / 64;                   // 32-column chunks per warp
        // staging tile of this warp: row r = lane, 128 bytes, 16-byte chunk j stored at chunk (j ^ (r & 7))
        uint8_t* stg_row = stage_out + (warp - 2) * 4096 + lane * 128;
        const int sw = lane & 7;
        const bool tma_bf16 = g.tma_out && g.epilogue != GEMM_EPI_RESADD_F32;
        const bool tma_add = g.tma_out && g.epilogue == GEMM_EPI_RESADD_F32;
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
            const int n_tile = tile % n_tiles, m_tile = tile / n_tiles;
            const uint32_t acc = j & 1u;
            const int row_in_tile = q * 32 + lane;
            const int m = m_tile * BM + row_in_tile;
            int dst_row = m;
            if (g.row_map != nullptr && m < g.M) dst_row = g.row_map[m];
            const bool live = (m < g.M) && (dst_row >= 0);
            float ln_rstd = 1.f, ln_nmr = 0.f;          // rstd and -mean * rstd of this row (LN-fold epilogue)
            if (g.epilogue == GEMM_EPI_LNFOLD_GELU_BF16 && m < g.M) {
                float S = 0.f, Q = 0.f;
                for (int z = 0; z < g.ln_splits; ++z) {
                    const float2 p2 = *reinterpret_cast<const float2*>(g.ln_stats + ((size_t)z * g.M + m) * 2);
                    S += p2.x; Q += p2.y;
                }
                const float inv_k = 1.0f / (float)g.K;
                const float mean = S * inv_k;
                const float var = fmaxf(Q * inv_k - mean * mean, 0.f);
                ln_rstd = 1.0f / sqrtf(var + g.ln_eps);
                ln_nmr = -mean * ln_rstd;
            }
            mbar_wait(&acc_full[acc], (j >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c = chalf * CPW; c < chalf * CPW + CPW; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
                if (!live && !g.tma_out) continue;       // (TMA stores clip rows >= M themselves; every lane takes part)
                const int n0 = n_tile * BN + c * 32;
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                if (g.epilogue == GEMM_EPI_LNFOLD_GELU_BF16) {
                    // v = rstd * (acc - mean * colsum[n]) + bias[n]
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 s4 = *reinterpret_cast<const float4*>(g.colsum + n0 + i);
                        const float4 b4 = *reinterpret_cast<const float4*>(g.bias + n0 + i);
                        v[i] = fmaf(v[i], ln_rstd, fmaf(ln_nmr, s4.x, b4.x));
                        v[i + 1] = fmaf(v[i + 1], ln_rstd, fmaf(ln_nmr, s4.y, b4.y));
                        v[i + 2] = fmaf(v[i + 2], ln_rstd, fmaf(ln_nmr, s4.z, b4.z));
                        v[i + 3] = fmaf(v[i + 3], ln_rstd, fmaf(ln_nmr, s4.w, b4.w));
                    }
                } else if (g.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(g.bias + n0 + i);
                        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                    }
                }
                if (g.epilogue == GEMM_EPI_BF16 || g.epilogue == GEMM_EPI_GELU_BF16 || g.epilogue == GEMM_EPI_LNFOLD_GELU_BF16) {
                    if (g.epilogue != GEMM_EPI_BF16) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) gelu_erf_pair(v[i], v[i + 1]);
                    }
                    if (tma_bf16) {
                        // two 32-column chunks fill the warp's [32 rows][64 bf16] staging tile, then one TMA store
                        const int half = (c - chalf * CPW) & 1;
                        if (half == 0) {
                            if (lane == 0) bulk_wait_read0();        // the previous store has read the tile
                            __syncwarp();
                        }
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            uint4 pk;
                            pk.x = pack_bf16(v[i], v[i + 1]); pk.y = pack_bf16(v[i + 2], v[i + 3]);
                            pk.z = pack_bf16(v[i + 4], v[i + 5]); pk.w = pack_bf16(v[i + 6], v[i + 7]);
                            *reinterpret_cast<uint4*>(stg_row + (((half * 4 + (i >> 3)) ^ sw) << 4)) = pk;
                        }
                        if (half == 1) {
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmap_out, stage_out + (warp - 2) * 4096, n0 - 32, m_tile * BM + q * 32);
                                bulk_commit();
                            }
                        }
                        continue;
                    }
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out) + (size_t)dst_row * g.N + n0;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 pk;
                        pk.x = pack_bf16(v[i], v[i + 1]); pk.y = pack_bf16(v[i + 2], v[i + 3]);
                        pk.z = pack_bf16(v[i + 4], v[i + 5]); pk.w = pack_bf16(v[i + 6], v[i + 7]);
                        *reinterpret_cast<uint4*>(o + i) = pk;
                    }
                } else if (g.epilogue == GEMM_EPI_RESADD_F32) {
                    float* o = reinterpret_cast<float*>(g.out) + (size_t)dst_row * g.N + n0;
                    if (g.gamma != nullptr) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 g4 = *reinterpret_cast<const float4*>(g.gamma + n0 + i);
                            v[i] *= g4.x; v[i + 1] *= g4.y; v[i + 2] *= g4.z; v[i + 3] *= g4.w;
                        }
                    }
                    if (tma_add) {
                        // [32 rows][32 fp32] staging tile, added into the residual stream by the TMA unit (no read-modify-write
                        // through the SM; every element is added exactly once per GEMM, so the result is deterministic)
                        if (lane == 0) bulk_wait_read0();
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4*>(stg_row + (((i >> 2) ^ sw) << 4)) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_reduce_add_2d(&tmap_out, stage_out + (warp - 2) * 4096, n0, m_tile * BM + q * 32);
                            bulk_commit();
                        }
                        continue;
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 x4 = *reinterpret_cast<float4*>(o + i);
                        x4.x += v[i]; x4.y += v[i + 1]; x4.z += v[i + 2]; x4.w += v[i + 3];
                        *reinterpret_cast<float4*>(o + i) = x4;
                    }
                } else {
                    float* o = reinterpret_cast<float*>(g.out) + (size_t)dst_row * g.N + n0;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            // all TMEM reads of this accumulator by this warp have completed (tcgen05.wait::ld inside tmem_ld32)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
    }
    if (g.tma_out && warp >= 2 && lane == 0) bulk_wait_all();   // the staging tiles outlive the CTA otherwise
    // ---- teardown: every role is done with TMEM before it is released ----
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static bool g_tma_epilogue = true;     // MNX_GEMM_TMA_EPILOGUE=0: per-lane stores (A/B timing on the GPU box)

cudaError_t gemm_tc_configure() {
    if (const char* env = getenv("MNX_GEMM_TMA_EPILOGUE")) g_tma_epilogue = env[0] != '0';
    if (g_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || fn == nullptr) return cudaErrorNotSupported;
        g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256>::SMEM_BYTES);
}

// 2-D bf16 row-major [rows][cols] tensor, box = [box_rows][64 cols], 128-byte swizzle, zero OOB fill
static bool make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// output map for the TMA epilogue: [rows][cols] row-major, box = [32 rows][128 bytes], 128-byte swizzle
static bool make_out_map(CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, bool f32) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * (f32 ? 4u : 2u)};
    const cuuint32_t box[2] = {f32 ? 32u : 64u, 32u};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Tensor maps depend only on (pointers, shapes): the encoders call the same ~100 GEMMs on the same engine-owned workspaces
// for every batch, so the three cuTensorMapEncodeTiled calls per GEMM (~300 driver calls per forward, enough to make the
// host the bottleneck of a 6 ms encoder on a busy box) are cached per process.
struct MapKey {
    const void *a, *w, *out;
    int M, N, K, kind;
    bool operator==(const MapKey& o) const { return a == o.a && w == o.w && out == o.out && M == o.M && N == o.N && K == o.K && kind == o.kind; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.a) * 0x9E3779B97F4A7C15ull;
        h ^= reinterpret_cast<size_t>(k.w) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= reinterpret_cast<size_t>(k.out) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= ((size_t)k.M << 32 | (size_t)k.N << 12 | (size_t)k.K << 2 | (size_t)k.kind) + (h << 6) + (h >> 2);
        return h;
    }
};
struct MapTriple { CUtensorMap a, w, o; };
static std::unordered_map<MapKey, MapTriple, MapKeyHash> g_map_cache;
static std::mutex g_map_mutex;

cudaError_t gemm_tc_launch(const GemmParams& p, cudaStream_t s) {
    if (g_encode == nullptr) return cudaErrorNotReady;
    if (p.M < 1 || p.N % 128 != 0 || p.K % BK != 0 || p.K < BK) return cudaErrorInvalidValue;
    const int BN = (p.N % 256 == 0) ? 256 : 128;
    if (p.epilogue == GEMM_EPI_LNFOLD_GELU_BF16 && (!p.colsum || !p.ln_stats || !p.bias || p.ln_splits < 1)) return cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.W) & 15)) return cudaErrorInvalidValue;
    CUtensorMap ma, mw;
    // TMA epilogue: bf16 outputs, and the residual add when rows are not scattered through a row map
    const bool bf16_out = p.epilogue == GEMM_EPI_BF16 || p.epilogue == GEMM_EPI_GELU_BF16 || p.epilogue == GEMM_EPI_LNFOLD_GELU_BF16;
    const bool add_out = p.epilogue == GEMM_EPI_RESADD_F32 && p.row_map == nullptr;
    const int tma_out = (g_tma_epilogue && (bf16_out || add_out) && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) ? 1 : 0;
    CUtensorMap mo;
    {
        const MapKey key{p.A, p.W, tma_out ? p.out : nullptr, p.M, p.N, p.K, tma_out ? (add_out ? 2 : 1) : 0};
        std::lock_guard<std::mutex> lock(g_map_mutex);
        auto it = g_map_cache.find(key);
        if (it == g_map_cache.end()) {
            MapTriple t;
            if (!make_map(&t.a, p.A, (uint64_t)p.M, (uint64_t)p.K, BM)) return cudaErrorInvalidValue;
            if (!make_map(&t.w, p.W, (uint64_t)p.N, (uint64_t)p.K, BN)) return cudaErrorInvalidValue;
            t.o = t.a;
            if (tma_out && !make_out_map(&t.o, p.out, (uint64_t)p.M, (uint64_t)p.N, add_out)) return cudaErrorInvalidValue;
            if (g_map_cache.size() > 8192) g_map_cache.clear();      // (callers with ever-changing pointers: bounded memory)
            it = g_map_cache.emplace(key, t).first;
        }
        ma = it->second.a; mw = it->second.w; mo = it->second.o;
    }
    GemmKernelArgs g{p.M, p.N, p.K, p.epilogue, p.bias, p.gamma, p.row_map, p.out, p.colsum, p.ln_stats, p.ln_splits, p.ln_eps, tma_out};
    const int total_tiles = (p.N / BN) * ((p.M + BM - 1) / BM);
    int dev = 0, num_sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);     // cached by the runtime
    const int cap = (p.cta_limit > 0 && p.cta_limit < num_sms) ? p.cta_limit : num_sms;
    const int grid = total_tiles < cap ? total_tiles : cap;
    if (BN == 256) gemm_tc_kernel<256><<<grid, GEMM_THREADS, GemmCfg<256>::SMEM_BYTES, s>>>(ma, mw, mo, g);
    else gemm_tc_kernel<128><<<grid, GEMM_THREADS, GemmCfg<128>::SMEM_BYTES, s>>>(ma, mw, mo, g);
    return cudaGetLastError();
}

// ---- stand-alone test entry (include/molnextr_b200.h: mnx_test_gemm_bf16) ---------------------
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(in[i]);
}

}  // namespace mnx

extern "C" int mnx_test_gemm_bf16(const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N,
                                  int32_t K, int32_t epilogue, void* cuda_stream) {
    using namespace mnx;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (gemm_tc_configure() != cudaSuccess) return -2;
    __nv_bfloat16 *a = nullptr, *w = nullptr, *o = nullptr;
    const size_t na = (size_t)M * K, nw = (size_t)N * K, nc = (size_t)M * N;
    if (cudaMalloc(&a, na * 2) != cudaSuccess || cudaMalloc(&w, nw * 2) != cudaSuccess || cudaMalloc(&o, nc * 2) != cudaSuccess) return -2;
    f32_to_bf16_kernel<<<(unsigned)((na + 255) / 256), 256, 0, s>>>(A, a, na);
    f32_to_bf16_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, s>>>(W, w, nw);
    GemmParams p{};
    p.A = a; p.W = w; p.M = M; p.N = N; p.K = K; p.epilogue = epilogue; p.bias = bias;
    const bool bf16_out = (epilogue == GEMM_EPI_BF16 || epilogue == GEMM_EPI_GELU_BF16);
    p.out = bf16_out ? (void*)o : (void*)C;
    cudaError_t e = gemm_tc_launch(p, s);
    if (e == cudaSuccess && bf16_out) bf16_to_f32_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(o, C, nc);
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(a); cudaFree(w); cudaFree(o);
    if (e != cudaSuccess) return e == cudaErrorInvalidValue ? -1 : -2;
    return e2 == cudaSuccess ? 0 : -2;
}
