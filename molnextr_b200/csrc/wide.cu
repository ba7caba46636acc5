// Throughput-oriented persistent decode kernel: 8-CTA clusters that each own up to 16 rows.
//
// mega16_impl.cuh minimises the latency of ONE batch (112 SMs, G <= 5 rows per cluster, every phase a
// latency chain: FMA pipe 13 %, warps active 16 %).  This kernel optimises images per SM-second instead,
// so that several batches decode concurrently next to the encoder of the following ones:
//   * a cluster of 8 CTAs owns G <= 16 rows for the whole greedy decode; CTA h is attention head h
//     (q/k/v of a row never leave the CTA) and owns the 32-column slice h of the residual stream;
//   * every Linear is a register-tiled fp32 GEMM over all 16 rows at once.  Activations live TRANSPOSED and
//     k-pair interleaved in shared memory ([k/2][16 rows][2]), weights arrive in the same pairing
//     ([k/2][cols][2], 32 KB slots through a bulk-copy ring), so a lane's 4-row x 4-column block advances
//     two k per step with four conflict-free LDS.128 and sixteen packed fma.rn.f32x2 (even and odd k
//     accumulate in the two halves and are added at the end) -- scalar FFMA issues at half the fp32 rate on
//     sm_100 and the G <= 5 kernels are shared-memory-bound instead;
//   * Megatron-style split: q|k|v, Wq_ctx, W1 and the vocabulary are column-parallel (warps split K,
//     cross-warp reduction through shared memory); Wo, Wo_ctx and W2 are row-parallel (K = this head's
//     context / this CTA's 128 FFN columns; warps split the 256 output columns) followed by a
//     reduce-scatter (partials to the CTA that owns the column slice) and an all-gather of the new
//     residual slice, both as 16-byte st.async into distributed shared memory with mbarrier
//     complete_tx signalling.  Six exchanges per layer instead of eight; the FFN hidden never leaves its CTA;
//   * LayerNorm statistics ride with the all-gather (per-slice sum and centred M2, combined with the
//     parallel-variance formula), so the normalise pass is elementwise;
//   * attention: warp w serves rows w and w + 8.  Every warp is its own K/V producer: lane 0 keeps a ring
//     of 32-key sub-tiles (bulk copies, 4 KB) filled along the warp's fixed tile sequence of the whole step, so
//     the memory-bank K/V of a layer streams in while the previous GEMM phases still run and no block barrier
//     or producer hand-off sits inside the attention phase; scores stay in registers (two passes, the
//     reference's softmax: exp(s - max) / sum);
//   * log-softmax / grammar mask / argmax of row r run only in CTA r / 2 (targeted logits exchange),
//     which then broadcasts the chosen token; the row-rank positional-encoding rule uses
//     the cluster's own flags plus `row_state` of lower clusters, whose order is fixed by a ticket
//     taken at kernel start (whichever cluster is scheduled first owns the lowest rows).
// Arithmetic is fp32 and follows the same reference lines as decoder.cu / mega.cu.
#include "mega.cuh"

#include <math.h>

namespace mnx {
namespace {

#define W_G 16                      // rows per cluster
#define W_CT 256                    // compute threads (8 warps)
#define W_THREADS 288               // + weight producer warp
#define W_SLOT_FLOATS 8192          // 32 KB weight slot = two 16 KB halves
#define W_HALF_BYTES 16384
#define W_NWS 2                     // weight ring depth (slots)
#define W_SLOTS_PER_LAYER 14
#define W_TK 32                     // keys per staged K/V sub-tile (4 KB)
#define W_NSUB 2                    // sub-tiles in flight per warp
#define W_KV_BYTES (W_TK * 128)
#define W_QSCALE 5.656854152679443f

// parameter block offsets (floats) inside ppack[h][l][MG_PARAM_FLOATS_H] (engine.cu finalize_decoder)
enum { WP_LN1W = 0, WP_LN1B = 256, WP_LN2W = 512, WP_LN2B = 768, WP_LNFW = 1024, WP_LNFB = 1280,
       WP_BQ = 1536, WP_BK = 1568, WP_BV = 1600, WP_BO = 1632, WP_BQC = 1664, WP_BOC = 1696, WP_B1 = 1728, WP_B2 = 1856 };

// Transposed activation buffers hold 16-byte units: unit(kp, rp) = {x(2kp, 2rp), x(2kp+1, 2rp), x(2kp, 2rp+1),
// x(2kp+1, 2rp+1)} at index kp * 8 + rp  (x(k, row); kp = feature pair, rp = row pair).
#define W_UNIT(kp, rp) (((kp) << 3) + (rp))
#define W_ELEM(k, row) (((((k) >> 1) * 16 + (row)) << 1) + ((k) & 1))      // float index of x(k, row)

struct WSmem {
    static constexpr int wring = 0;
    static constexpr int kv = wring + W_NWS * W_SLOT_FLOATS * 4;          // [8 warps][W_NSUB][32 keys][32]
    static constexpr int xT = kv + 8 * W_NSUB * W_KV_BYTES;               // residual stream, 128 x 8 units
    static constexpr int nT = xT + 256 * W_G * 4;                         // LayerNorm output, same layout
    static constexpr int prtT = nT + 256 * W_G * 4;                       // [8 src][8 rp][16 cp] row-parallel partials
    static constexpr int red = prtT + 8 * 32 * W_G * 4;                   // [8 warps][8 rp][16 cp] K-split partials
    static constexpr int hT = red + 8 * 32 * W_G * 4;                     // FFN hidden slice, 64 x 8 units
    static constexpr int ctxT = hT + 128 * W_G * 4;                       // attention context of this head, 16 x 8 units
    static constexpr int qs = ctxT + 32 * W_G * 4;                        // [16 rows][32] scaled query
    static constexpr int ks = qs + W_G * 32 * 4;
    static constexpr int vs = ks + W_G * 32 * 4;
    static constexpr int stat = vs + W_G * 32 * 4;                        // [8 src][8 rp] {S(2rp), S(2rp+1), M2(2rp), M2(2rp+1)}
    static constexpr int mr = stat + 8 * 8 * 16;                          // [2][16]: mean, rstd
    static constexpr int lg = mr + 2 * W_G * 4;                           // [2 rows][256] logits of the rows this CTA decides
    static constexpr int tf = lg + 2 * 256 * 4;                           // [16] {token, finished}
    static constexpr int rank = tf + W_G * 8;                             // [16] PE rank
    static constexpr int bars = rank + W_G * 4;                           // mbarriers
    static constexpr int misc = bars + 256;
    static constexpr int total = misc + 64;
};
static_assert(WSmem::total <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ uint32_t w_mapa(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void w_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_W:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_W;\n"
        "bra WAIT_LOOP_W;\n"
        "DONE_W:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void w_cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void w_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned w_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void w_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// packed fp32 pair FMA: d = a * b + c on both halves (sm_100 issues scalar FFMA at half this rate)
__device__ __forceinline__ unsigned long long w_ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float w_pair_sum(unsigned long long v) {
    return __uint_as_float((unsigned)(v & 0xffffffffull)) + __uint_as_float((unsigned)(v >> 32));
}

// one single-query attention of a warp: n keys of K and of V rows (128 B each) in global memory; n = 0: nothing to do
struct WAtt {
    const float *K, *V;
    int n;
};

struct WCtx {
    uint8_t* sm;
    int h;                          // cluster rank = attention head = owner of residual columns [32h, 32h+32)
    int tid, lane, warp;
    int rg, cg;                     // accumulator block of this lane: row pairs {rg, 4+rg}, column pairs {cg, 8+cg}
    int G;
    uint64_t *wfull, *wempty, *kvfull, *xbar, *stepbar;
    uint32_t w_seq, x_seq;
    uint32_t xbar_base;
    // K/V ring of this warp: tile number n (since kernel start) lives in sub-slot n % W_NSUB
    uint32_t kv_cons;                   // tiles consumed so far
    unsigned alive;                     // bit r: row r decodes in this step
    int t, S, T, B, row0;
    const float *selfK, *selfV, *crossK, *crossV;
    size_t kv_layer;
    uint64_t kv_policy;
};

// ---- exchanges: st.async into a peer's shared memory, counted on the peer's current exchange barrier ----------
__device__ __forceinline__ void w_send4(const WCtx& c, int byte_off, uint32_t dst_cta, float4 v) {
    const uint32_t local = smem_u32(c.sm + byte_off);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     w_mapa(local, dst_cta)),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
                 "r"(w_mapa(c.xbar_base + 8u * (c.x_seq & 1u), dst_cta))
                 : "memory");
}
__device__ __forceinline__ void w_send1(const WCtx& c, int byte_off, uint32_t dst_cta, uint32_t v) {
    const uint32_t local = smem_u32(c.sm + byte_off);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(w_mapa(local, dst_cta)),
                 "r"(v), "r"(w_mapa(c.xbar_base + 8u * (c.x_seq & 1u), dst_cta))
                 : "memory");
}
// finish an exchange in which this CTA receives `bytes_in` bytes.  Two barriers alternate: a CTA can send
// exchange e + 1 while a peer still waits on e, never e + 2 (e + 1 cannot complete without that peer's data).
__device__ __forceinline__ void w_exchange(WCtx& c, uint32_t bytes_in) {
    uint64_t* bar = c.xbar + (c.x_seq & 1u);
    if (c.tid == 0) mbar_arrive_expect_tx(bar, bytes_in);
    w_wait_cluster(bar, (c.x_seq >> 1) & 1u);
    ++c.x_seq;
}

// ---- K/V ring ---------------------------------------------------------------------------------------------------
// A warp runs up to 24 attentions per step, in the order a = (layer * 2 + cross) * 2 + rr (row = warp + 8 rr); each
// consumes the tile sequence K0..K(nt-1), V0..V(nt-1).  The warp is its own producer: when it has finished
// reading a tile, lane 0 refills that sub-slot with the tile W_NSUB further down the sequence, which may belong to
// the warp's NEXT attention -- so the memory-bank K/V of a layer streams in while the GEMM phases in between run.
static_assert(W_NSUB == 2, "the look-ahead below assumes that every attention has at least W_NSUB tiles (nt >= 1)");
__device__ __forceinline__ WAtt w_att_desc(const WCtx& c, int a) {
    WAtt d;
    const int l = a >> 2, cross = (a >> 1) & 1, row = c.warp + 8 * (a & 1);
    d.n = (a < 4 * MNX_DEC_L && ((c.alive >> row) & 1u)) ? (cross ? c.S : c.t) : 0;
    const size_t off = cross ? ((((size_t)l * c.B + c.row0 + row) * 8 + c.h) * (size_t)c.S) * 32
                             : l * c.kv_layer + (((size_t)(c.row0 + row) * 8 + c.h) * c.T) * 32;
    d.K = (cross ? c.crossK : c.selfK) + off;
    d.V = (cross ? c.crossV : c.selfV) + off;
    return d;
}
// first attention after `a` that has keys to fetch (4 * MNX_DEC_L = none)
__device__ __forceinline__ int w_att_next(const WCtx& c, int a) {
#pragma unroll
    for (int d = 1; d <= 4; ++d) {
        const int b = a + d;
        if (b >= 4 * MNX_DEC_L) return 4 * MNX_DEC_L;
        const int row = c.warp + 8 * (b & 1);
        if (((c.alive >> row) & 1u) && (((b >> 1) & 1) || c.t > 0)) return b;
    }
    return 4 * MNX_DEC_L;
}
// lane 0: bulk copy of flat tile `idx` (K tiles then V tiles) of attention `d` into sub-slot `slot`
__device__ __forceinline__ void w_kv_issue(const WCtx& c, const WAtt& d, int idx, uint32_t slot) {
    const int nt = (d.n + W_TK - 1) / W_TK;
    const int part = idx >= nt ? 1 : 0, tile = idx - part * nt;
    const uint32_t bytes = (uint32_t)min(W_TK, d.n - tile * W_TK) * 128u;
    uint64_t* bar = &c.kvfull[c.warp * W_NSUB + slot];
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s_hint(c.sm + WSmem::kv + (c.warp * W_NSUB + slot) * W_KV_BYTES, (part ? d.V : d.K) + (size_t)tile * W_TK * 32, bytes, bar, c.kv_policy);
}
// wait for the next tile of the sequence; returns its shared-memory address
__device__ __forceinline__ const float* w_kv_acquire(WCtx& c) {
    const uint32_t slot = c.kv_cons % W_NSUB;
    mbar_wait(&c.kvfull[c.warp * W_NSUB + slot], (c.kv_cons / W_NSUB) & 1u);
    return reinterpret_cast<const float*>(c.sm + WSmem::kv + (c.warp * W_NSUB + slot) * W_KV_BYTES);
}
// the warp is done reading flat tile `idx` of attention `d`: refill its sub-slot W_NSUB tiles ahead
__device__ __forceinline__ void w_kv_release(WCtx& c, const WAtt& d, const WAtt& nx, int idx) {
    __syncwarp();
    if (c.lane == 0) {
        const int nt2 = 2 * ((d.n + W_TK - 1) / W_TK), j = idx + W_NSUB;
        if (j < nt2) w_kv_issue(c, d, j, c.kv_cons % W_NSUB);
        else if (nx.n > 0) w_kv_issue(c, nx, j - nt2, c.kv_cons % W_NSUB);
    }
    ++c.kv_cons;
}

// ---- register-tiled GEMM block on packed pairs --------------------------------------------------------------------
// acc[ri][cj] (+)= sum over NKP feature pairs of x2[ri] * w2[cj]  (both halves: even and odd k separately).
// A4: activation units (16 bytes), first pair kp0; Wu: this lane's first weight unit, `WS` units per pair.
template <int NKP, int WS>
__device__ __forceinline__ void w_fma_pairs(const WCtx& c, const ulonglong2* __restrict__ A4, int kp0,
                                            const ulonglong2* __restrict__ Wu, unsigned long long (&acc)[4][4]) {
    const ulonglong2* a = A4 + W_UNIT(kp0, c.rg);
#pragma unroll
    for (int kk = 0; kk < NKP; ++kk) {
        const ulonglong2 x01 = a[kk * 8], x23 = a[kk * 8 + 4];
        const ulonglong2 w01 = Wu[kk * WS], w23 = Wu[kk * WS + 8];
        acc[0][0] = w_ffma2(x01.x, w01.x, acc[0][0]); acc[0][1] = w_ffma2(x01.x, w01.y, acc[0][1]);
        acc[0][2] = w_ffma2(x01.x, w23.x, acc[0][2]); acc[0][3] = w_ffma2(x01.x, w23.y, acc[0][3]);
        acc[1][0] = w_ffma2(x01.y, w01.x, acc[1][0]); acc[1][1] = w_ffma2(x01.y, w01.y, acc[1][1]);
        acc[1][2] = w_ffma2(x01.y, w23.x, acc[1][2]); acc[1][3] = w_ffma2(x01.y, w23.y, acc[1][3]);
        acc[2][0] = w_ffma2(x23.x, w01.x, acc[2][0]); acc[2][1] = w_ffma2(x23.x, w01.y, acc[2][1]);
        acc[2][2] = w_ffma2(x23.x, w23.x, acc[2][2]); acc[2][3] = w_ffma2(x23.x, w23.y, acc[2][3]);
        acc[3][0] = w_ffma2(x23.y, w01.x, acc[3][0]); acc[3][1] = w_ffma2(x23.y, w01.y, acc[3][1]);
        acc[3][2] = w_ffma2(x23.y, w23.x, acc[3][2]); acc[3][3] = w_ffma2(x23.y, w23.y, acc[3][3]);
    }
}
// the four output units of a lane: (cpi, rpi) -> unit (cp = cg + 8 cpi, rp = rg + 4 rpi)
__device__ __forceinline__ float4 w_out_unit(const unsigned long long (&acc)[4][4], int cpi, int rpi) {
    return make_float4(w_pair_sum(acc[2 * rpi][2 * cpi]), w_pair_sum(acc[2 * rpi][2 * cpi + 1]),
                       w_pair_sum(acc[2 * rpi + 1][2 * cpi]), w_pair_sum(acc[2 * rpi + 1][2 * cpi + 1]));
}

// weight ring, consumer side: every warp reads one 16 KB half of every slot (warps 0-3 half 0, warps 4-7 half 1)
__device__ __forceinline__ const ulonglong2* w_slot_acquire(WCtx& c) {
    const uint32_t slot = c.w_seq % W_NWS, ph = (c.w_seq / W_NWS) & 1u;
    mbar_wait(&c.wfull[slot * 2 + (c.warp >> 2)], ph);
    return reinterpret_cast<const ulonglong2*>(c.sm + WSmem::wring + slot * W_SLOT_FLOATS * 4 + (c.warp >> 2) * W_HALF_BYTES);
}
__device__ __forceinline__ void w_slot_release(WCtx& c) {
    const uint32_t slot = c.w_seq % W_NWS;
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.wempty[slot]);
    ++c.w_seq;
}

// column-parallel tile: out[16 rows][32 cols] = nT[16][256] . slot[128 kp][32 cols][2]; warp w owns the 16 pairs
// [16w, 16w + 16) (slot half w / 4); the eight K-split partials are combined through `red`, then
// f(cp, rp, unit) runs once per (column pair, row pair) on threads 0..127: unit = {o(2cp, 2rp), o(2cp+1, 2rp),
// o(2cp, 2rp+1), o(2cp+1, 2rp+1)}.  Two block barriers.
__device__ __forceinline__ void w_col_gemm(WCtx& c) {
    unsigned long long acc[4][4] = {};
    const ulonglong2* half = w_slot_acquire(c);
    w_fma_pairs<16, 16>(c, reinterpret_cast<const ulonglong2*>(c.sm + WSmem::nT), 16 * c.warp, half + (16 * (c.warp & 3)) * 16 + c.cg, acc);
    w_slot_release(c);
    float4* red4 = reinterpret_cast<float4*>(c.sm + WSmem::red) + c.warp * 128;
#pragma unroll
    for (int rpi = 0; rpi < 2; ++rpi)
#pragma unroll
        for (int cpi = 0; cpi < 2; ++cpi) red4[(c.rg + 4 * rpi) * 16 + c.cg + 8 * cpi] = w_out_unit(acc, cpi, rpi);
    w_sync();
}
template <class F>
__device__ __forceinline__ void w_col_tile(WCtx& c, F f) {
    w_col_gemm(c);
    if (c.tid < 128) {
        const float4* r4 = reinterpret_cast<const float4*>(c.sm + WSmem::red) + c.tid;    // unit (rp = tid / 16, cp = tid % 16)
        float4 v = r4[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            const float4 p = r4[w * 128];
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        f(c.tid & 15, c.tid >> 4, v);
    }
    w_sync();
}

// row-parallel product + reduce-scatter + all-gather.  A = NK * 16 feature pairs of this CTA's K slice, NK (1 or 4) weight
// slots of [2 halves][16 kp][128 cols][2]; warp w computes output columns [32w, 32w + 32) and sends the
// partial block to CTA w.  After the exchange the owner adds the eight partials in source order, the bias
// and the residual, computes the LayerNorm statistics of its 32-column slice and all-gathers both.
__device__ __forceinline__ void w_row_gemm_allreduce(WCtx& c, const ulonglong2* AT4, int NK, const float* __restrict__ bias) {
    {
        unsigned long long acc[4][4] = {};
#pragma unroll 1
        for (int q = 0; q < NK; ++q) {
            const ulonglong2* half = w_slot_acquire(c);
            w_fma_pairs<16, 64>(c, AT4, 16 * q, half + 16 * (c.warp & 3) + c.cg, acc);
            w_slot_release(c);
        }
        // partial [16 rows][32 cols] of output slice `warp` -> prtT[src = h] of CTA `warp`
#pragma unroll
        for (int rpi = 0; rpi < 2; ++rpi)
#pragma unroll
            for (int cpi = 0; cpi < 2; ++cpi)
                w_send4(c, WSmem::prtT + ((c.h * 8 + c.rg + 4 * rpi) * 16 + c.cg + 8 * cpi) * 16, (uint32_t)c.warp, w_out_unit(acc, cpi, rpi));
    }
    const int cp = c.tid & 15, rp = c.tid >> 4;
    const float2 b = reinterpret_cast<const float2*>(bias)[cp];   // in flight during the wait
    w_exchange(c, 8u * 32u * W_G * 4u);
    if (c.tid < 128) {
        const float4* prt4 = reinterpret_cast<const float4*>(c.sm + WSmem::prtT) + c.tid;
        float4 v = prt4[0];
#pragma unroll
        for (int s = 1; s < 8; ++s) {
            const float4 p = prt4[s * 128];
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        const int unit = W_UNIT(c.h * 16 + cp, rp);
        const float4 x = reinterpret_cast<const float4*>(c.sm + WSmem::xT)[unit];
        v.x = (v.x + b.x) + x.x; v.y = (v.y + b.y) + x.y; v.z = (v.z + b.x) + x.z; v.w = (v.w + b.y) + x.w;
        // slice statistics of rows 2rp (x, y) and 2rp + 1 (z, w): sum and M2 around the slice mean, over the
        // 16 lanes that share rp
        float s0 = v.x + v.y, s1 = v.z + v.w;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        const float m0 = s0 * (1.f / 32.f), m1 = s1 * (1.f / 32.f);
        float q0 = (v.x - m0) * (v.x - m0) + (v.y - m0) * (v.y - m0), q1 = (v.z - m1) * (v.z - m1) + (v.w - m1) * (v.w - m1);
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o); }
#pragma unroll
        for (uint32_t dst = 0; dst < 8; ++dst) w_send4(c, WSmem::xT + unit * 16, dst, v);
        if (cp < 8) w_send4(c, WSmem::stat + (c.h * 8 + rp) * 16, (uint32_t)cp, make_float4(s0, s1, q0, q1));
    }
    w_exchange(c, 8u * (32u * W_G * 4u + 8u * 16u));
}

// LayerNorm (eps 1e-6) of the residual stream: statistics from the eight slices (Chan's parallel variance),
// then an elementwise pass xT -> nT.  Thread tid handles units tid + 256 i: row pair tid % 8, feature pairs
// tid / 8 + 32 i; gamma / beta come from global memory (L1 / L2 resident).
__device__ __forceinline__ void w_layer_norm(WCtx& c, const float* __restrict__ gw, const float* __restrict__ gb) {
    float2 g[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        g[i] = reinterpret_cast<const float2*>(gw)[(c.tid >> 3) + 32 * i];
        b[i] = reinterpret_cast<const float2*>(gb)[(c.tid >> 3) + 32 * i];
    }
    const float4* stat = reinterpret_cast<const float4*>(c.sm + WSmem::stat);
    float* mr = reinterpret_cast<float*>(c.sm + WSmem::mr);
    if (c.tid < W_G) {
        const int rp = c.tid >> 1, odd = c.tid & 1;
        float sl[8], ml[8];
#pragma unroll
        for (int src = 0; src < 8; ++src) {
            const float4 s4 = stat[src * 8 + rp];
            sl[src] = odd ? s4.y : s4.x;
            ml[src] = odd ? s4.w : s4.z;
        }
        float s = 0.f;
#pragma unroll
        for (int src = 0; src < 8; ++src) s += sl[src];
        const float mean = s * (1.0f / 256.0f);
        float m2 = 0.f;
#pragma unroll
        for (int src = 0; src < 8; ++src) {
            const float dm = sl[src] * (1.0f / 32.0f) - mean;
            m2 += ml[src] + 32.0f * dm * dm;
        }
        mr[c.tid] = mean;
        mr[W_G + c.tid] = 1.0f / sqrtf(m2 * (1.0f / 256.0f) + 1e-6f);
    }
    w_sync();
    const float4* x4 = reinterpret_cast<const float4*>(c.sm + WSmem::xT);
    float4* n4 = reinterpret_cast<float4*>(c.sm + WSmem::nT);
    const int rp = c.tid & 7;
    const float m0 = mr[2 * rp], m1 = mr[2 * rp + 1], r0 = mr[W_G + 2 * rp], r1 = mr[W_G + 2 * rp + 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 x = x4[c.tid + 256 * i];
        n4[c.tid + 256 * i] = make_float4((x.x - m0) * r0 * g[i].x + b[i].x, (x.y - m0) * r0 * g[i].y + b[i].y,
                                          (x.z - m1) * r1 * g[i].x + b[i].x, (x.w - m1) * r1 * g[i].y + b[i].y);
    }
    w_sync();
}

// single-query attention of head c.h for `row`, by one warp on its own K/V ring.  n keys come from global
// memory (K tiles, then V tiles, in the warp's fixed tile sequence); with `extra` the row's new key / value
// (ks / vs) is key n.  Softmax as onmt MultiHeadedAttention: fp32, exp(s - max) / sum.
__device__ __forceinline__ void w_attend(WCtx& c, int row, const WAtt& d, const WAtt& nx, bool extra) {
    const int n = d.n;
    const float4* q4 = reinterpret_cast<const float4*>(c.sm + WSmem::qs) + row * 8;
    // scores of this warp: 512 floats of the `red` area (no column-parallel GEMM runs during an attention phase)
    float* sc = reinterpret_cast<float*>(c.sm + WSmem::red) + c.warp * 512;
    const int nt = (n + W_TK - 1) / W_TK;
    float m = -INFINITY;
    // lane j scores key j of a tile, reading its 128-byte row in a lane-rotated order (conflict free); the rotation
    // does not depend on the tile, so the matching query chunks are loaded once per attention
    float4 qr[8];
#pragma unroll
    for (int cc0 = 0; cc0 < 8; ++cc0) qr[cc0] = q4[(cc0 + c.lane) & 7];
#pragma unroll 1
    for (int i = 0; i < nt; ++i) {
        const float* tile = w_kv_acquire(c);
        float s = -INFINITY;
        if (i * W_TK + c.lane < n) {
            const float4* kr = reinterpret_cast<const float4*>(tile + c.lane * 32);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int cc0 = 0; cc0 < 8; cc0 += 4) {
                const float4 ka = kr[(cc0 + c.lane) & 7], kb = kr[(cc0 + 1 + c.lane) & 7];
                const float4 kc = kr[(cc0 + 2 + c.lane) & 7], kd = kr[(cc0 + 3 + c.lane) & 7];
                const float4 qa = qr[cc0], qb = qr[cc0 + 1], qc = qr[cc0 + 2], qd = qr[cc0 + 3];
                s0 = fmaf(qa.x, ka.x, s0); s0 = fmaf(qa.y, ka.y, s0); s0 = fmaf(qa.z, ka.z, s0); s0 = fmaf(qa.w, ka.w, s0);
                s1 = fmaf(qb.x, kb.x, s1); s1 = fmaf(qb.y, kb.y, s1); s1 = fmaf(qb.z, kb.z, s1); s1 = fmaf(qb.w, kb.w, s1);
                s2 = fmaf(qc.x, kc.x, s2); s2 = fmaf(qc.y, kc.y, s2); s2 = fmaf(qc.z, kc.z, s2); s2 = fmaf(qc.w, kc.w, s2);
                s3 = fmaf(qd.x, kd.x, s3); s3 = fmaf(qd.y, kd.y, s3); s3 = fmaf(qd.z, kd.z, s3); s3 = fmaf(qd.w, kd.w, s3);
            }
            s = (s0 + s1) + (s2 + s3);
        }
        sc[i * W_TK + c.lane] = s;
        m = fmaxf(m, s);
        w_kv_release(c, d, nx, i);
    }
    float sx = -INFINITY;
    if (extra) {
        const float4* k4 = reinterpret_cast<const float4*>(c.sm + WSmem::ks) + row * 8;
        sx = 0.f;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const float4 kv = k4[cc], qv = q4[cc];
            sx = fmaf(qv.x, kv.x, sx); sx = fmaf(qv.y, kv.y, sx);
            sx = fmaf(qv.z, kv.z, sx); sx = fmaf(qv.w, kv.w, sx);
        }
    }
    m = warp_max(fmaxf(m, sx));
    // exp(s - max) as exp2((s - max) * log2 e) (MUFU.EX2; within 2 ulp of expf) and the 1 / sum applied once to the context
    // instead of to every probability: the same softmax-weighted sum up to fp32 rounding (the tests pin ids exactly and
    // log-probs / hidden states to 5e-4), ~20 % fewer instructions per attention than expf + an IEEE division per key
    float sum = 0.f;
#pragma unroll 1
    for (int i = 0; i < nt; ++i) {
        const float e = exp2f((sc[i * W_TK + c.lane] - m) * 1.4426950408889634f);      // exp2(-inf) = 0 for the padding
        sc[i * W_TK + c.lane] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    const float ex = extra ? exp2f((sx - m) * 1.4426950408889634f) : 0.f;
    sum += ex;
    const float inv_sum = 1.0f / sum;
    __syncwarp();
    // context: lane = (key sub-index, 4-dim group); four keys per 16-byte-per-lane load
    const int ksub = c.lane >> 3, dq = c.lane & 7;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int i = 0; i < nt; ++i) {
        const float* tile = w_kv_acquire(c);
        const int nk = min(W_TK, n - i * W_TK);
        const float* pr = sc + i * W_TK + ksub;
#pragma unroll
        for (int l4 = 0; l4 < W_TK / 4; ++l4) {
            if (l4 * 4 < nk) {                                    // warp-uniform; p = 0 beyond the last key, stale rows are finite
                const float p = pr[l4 * 4];
                const float4 v = reinterpret_cast<const float4*>(tile + (l4 * 4 + ksub) * 32)[dq];
                acc.x = fmaf(p, v.x, acc.x); acc.y = fmaf(p, v.y, acc.y);
                acc.z = fmaf(p, v.z, acc.z); acc.w = fmaf(p, v.w, acc.w);
            }
        }
        w_kv_release(c, d, nx, nt + i);
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (c.lane < 8) {
        if (extra) {
            const float4 v = reinterpret_cast<const float4*>(c.sm + WSmem::vs)[row * 8 + dq];
            acc.x = fmaf(ex, v.x, acc.x); acc.y = fmaf(ex, v.y, acc.y);
            acc.z = fmaf(ex, v.z, acc.z); acc.w = fmaf(ex, v.w, acc.w);
        }
        acc.x *= inv_sum; acc.y *= inv_sum; acc.z *= inv_sum; acc.w *= inv_sum;
        float* ctx = reinterpret_cast<float*>(c.sm + WSmem::ctxT);
        *reinterpret_cast<float2*>(ctx + W_ELEM(4 * dq, row)) = make_float2(acc.x, acc.y);
        *reinterpret_cast<float2*>(ctx + W_ELEM(4 * dq + 2, row)) = make_float2(acc.z, acc.w);
    }
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(W_THREADS, 1) decode_wide_kernel(MegaArgs a) {
    extern __shared__ __align__(128) uint8_t sm[];
    WCtx c;
    c.sm = sm;
    c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    c.rg = c.lane >> 3; c.cg = c.lane & 7;
    {
        uint32_t r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        c.h = (int)r;
    }
    c.w_seq = 0; c.x_seq = 0; c.kv_cons = 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + WSmem::bars);
    c.wfull = bars; c.wempty = bars + 2 * W_NWS; c.kvfull = c.wempty + W_NWS;
    c.xbar = c.kvfull + 8 * W_NSUB; c.stepbar = c.xbar + 2;
    static_assert((2 * W_NWS + W_NWS + 8 * W_NSUB + 3) * 8 <= 256, "barrier area too small");
    c.xbar_base = smem_u32(&c.xbar[0]);
    int* s_tf = reinterpret_cast<int*>(sm + WSmem::tf);         // [16][2]: token, finished
    int* s_rank = reinterpret_cast<int*>(sm + WSmem::rank);
    int* s_go = reinterpret_cast<int*>(sm + WSmem::misc);
    int* s_ticket = s_go + 1;

    // zero every activation / staging buffer: rows >= G and stale K/V slots must hold finite values
    for (int i = c.tid; i < (WSmem::bars - WSmem::kv) / 16; i += W_THREADS)
        reinterpret_cast<float4*>(sm + WSmem::kv)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c.tid == 0) {
        for (int i = 0; i < 2 * W_NWS; ++i) mbar_init(&c.wfull[i], 1);
        for (int i = 0; i < W_NWS; ++i) mbar_init(&c.wempty[i], 8);
        for (int i = 0; i < 8 * W_NSUB; ++i) mbar_init(&c.kvfull[i], 1);
        mbar_init(&c.xbar[0], 1); mbar_init(&c.xbar[1], 1);
        mbar_init(c.stepbar, 1);
        fence_barrier_init();
        *s_go = 1;
        // cluster order = scheduling order: the first cluster that runs owns the lowest rows, so the spin on
        // `row_state` of lower rows below can never wait for a cluster that is not resident yet
        if (c.h == 0) *s_ticket = atomicAdd(a.ticket, 1);
    }
    __syncthreads();
    w_cluster_sync_all();
    int cluster;
    {
        uint32_t v;
        asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(w_mapa(smem_u32(s_ticket), 0)));
        cluster = (int)v;
    }
    const int row0 = a.row_base + cluster * a.G;    // rows below row_base belong to earlier launches: their row_state is final
    c.G = min(a.G, a.B - row0);
    if (c.tid < W_G) { s_tf[2 * c.tid] = a.g.sos; s_tf[2 * c.tid + 1] = (c.tid < c.G) ? 0 : 1; s_rank[c.tid] = 0; }
    __syncthreads();

    const size_t kv_layer = (size_t)a.B * 8 * a.T * 32;
    const float* wbase = a.wpackW + (size_t)c.h * (MNX_DEC_L * W_SLOTS_PER_LAYER + 1) * W_SLOT_FLOATS;
    const float* pbase = a.ppack + (size_t)c.h * MNX_DEC_L * MG_PARAM_FLOATS_H;
    c.S = a.S; c.T = a.T; c.B = a.B; c.row0 = row0;
    c.selfK = a.selfK; c.selfV = a.selfV; c.crossK = a.crossK; c.crossV = a.crossV;
    c.kv_layer = kv_layer;
    c.kv_policy = l2_policy_evict_first();

    if (c.warp == 8) {
        // ======================= weight producer: 85 slots per step, in consumption order =======================
        if (c.lane == 0) {
            const uint64_t keep = l2_policy_evict_last();
            uint32_t seq = 0, step = 0;
            for (;;) {
                mbar_wait(c.stepbar, step & 1u);
                if (*reinterpret_cast<volatile int*>(s_go) == 0) break;
                for (int u = 0; u < MNX_DEC_L * W_SLOTS_PER_LAYER + 1; ++u, ++seq) {
                    const uint32_t slot = seq % W_NWS, ph = (seq / W_NWS) & 1u;
                    mbar_wait(&c.wempty[slot], ph ^ 1u);
                    uint8_t* dst = sm + WSmem::wring + slot * W_SLOT_FLOATS * 4;
                    const float* src = wbase + (size_t)u * W_SLOT_FLOATS;
                    mbar_arrive_expect_tx(&c.wfull[slot * 2 + 0], W_HALF_BYTES);
                    bulk_g2s_hint(dst, src, W_HALF_BYTES, &c.wfull[slot * 2 + 0], keep);
                    mbar_arrive_expect_tx(&c.wfull[slot * 2 + 1], W_HALF_BYTES);
                    bulk_g2s_hint(dst + W_HALF_BYTES, src + 4096, W_HALF_BYTES, &c.wfull[slot * 2 + 1], keep);
                }
                ++step;
            }
        }
        __syncwarp();
    } else {
        // ======================= compute warps =======================
        float4* xT4 = reinterpret_cast<float4*>(sm + WSmem::xT);
        float* qs = reinterpret_cast<float*>(sm + WSmem::qs);
        float* ks = reinterpret_cast<float*>(sm + WSmem::ks);
        float* vs = reinterpret_cast<float*>(sm + WSmem::vs);
        float4* stat4 = reinterpret_cast<float4*>(sm + WSmem::stat);
        float* lg = reinterpret_cast<float*>(sm + WSmem::lg);
        int pm = 0;
#define W_MARK() do { if (a.prof && t == 100 && l == 1 && cluster == 0 && c.h == 0 && c.tid == 0 && pm < 64) a.prof[pm++] = clock64(); } while (0)
        for (int t = 0;; ++t) {
            int n_alive = 0;
            unsigned alive = 0;
            for (int g = 0; g < c.G; ++g)
                if (s_tf[2 * g + 1] == 0) { ++n_alive; alive |= 1u << g; }
            if (c.tid == 0) {
                *reinterpret_cast<volatile int*>(s_go) = n_alive > 0 ? 1 : 0;
                __threadfence_block();
                mbar_arrive(c.stepbar);
            }
            if (n_alive == 0) break;
            // ---- K/V ring of this warp: prime the first W_NSUB tiles of this step's sequence (positions < t are final) ----
            c.t = t; c.alive = alive;
            int att = w_att_next(c, -1);              // the warp's next attention with keys in global memory
            WAtt att_d = w_att_desc(c, att);
            if (c.lane == 0 && att_d.n > 0) {
                asm volatile("fence.proxy.async;" ::: "memory");     // K/V stored by the epilogue threads in earlier steps
                w_kv_issue(c, att_d, 0, c.kv_cons % W_NSUB);
                w_kv_issue(c, att_d, 1, (c.kv_cons + 1) % W_NSUB);
            }
            // ---- rank of each alive row among all alive rows of the batch (row-rank PE rule) ----
            if (c.warp == 0) {
                int finished_before = 0;
                for (int r = c.lane; r < row0; r += 32) {
                    unsigned s;
                    do { s = w_ld_acquire(a.row_state + r); } while ((s >> 1) < (unsigned)t && (s & 1u) == 0u);
                    if ((s & 1u) && (s >> 1) <= (unsigned)t) ++finished_before;
                }
                finished_before = (int)warp_sum((float)finished_before);
                if (c.lane == 0) {
                    int alive_lower = row0 - finished_before;
                    for (int g = 0; g < W_G; ++g) {
                        s_rank[g] = alive_lower;
                        if (s_tf[2 * g + 1] == 0) ++alive_lower;
                    }
                }
            }
            w_sync();
            // ---- embedding: x = emb[tok] * 16 + pe[rank]; thread = feature k, warp w = column slice w ----
            {
                float x[W_G];
#pragma unroll
                for (int r = 0; r < W_G; ++r)
                    x[r] = (s_tf[2 * r + 1] == 0) ? a.emb[s_tf[2 * r] * 256 + c.tid] * 16.0f + a.pe[(size_t)s_rank[r] * 256 + c.tid] : 0.f;
                float4 st[8];
#pragma unroll
                for (int rp = 0; rp < 8; ++rp) {
                    const float s0 = warp_sum(x[2 * rp]), s1 = warp_sum(x[2 * rp + 1]);
                    const float d0 = x[2 * rp] - s0 * (1.0f / 32.0f), d1 = x[2 * rp + 1] - s1 * (1.0f / 32.0f);
                    st[rp] = make_float4(s0, s1, warp_sum(d0 * d0), warp_sum(d1 * d1));
                }
                if (c.lane < 8) {
                    float4 mine = st[0];
#pragma unroll
                    for (int rp = 1; rp < 8; ++rp) if (c.lane == rp) mine = st[rp];
                    stat4[c.warp * 8 + c.lane] = mine;
                }
                // pair the features: even threads store row pairs 0..3, odd threads 4..7 of feature pair tid / 2
                const bool odd = c.tid & 1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rlo = 2 * i, rhi = 8 + 2 * i;      // rows handled by the even / the odd thread
                    const float pa = __shfl_xor_sync(0xffffffffu, odd ? x[rlo] : x[rhi], 1);
                    const float pb = __shfl_xor_sync(0xffffffffu, odd ? x[rlo + 1] : x[rhi + 1], 1);
                    // even thread: unit(kp, i) = {x_even(rlo), x_odd(rlo), x_even(rlo+1), x_odd(rlo+1)}
                    const float4 u = odd ? make_float4(pa, x[rhi], pb, x[rhi + 1]) : make_float4(x[rlo], pa, x[rlo + 1], pb);
                    xT4[W_UNIT(c.tid >> 1, (odd ? 4 : 0) + i)] = u;
                }
            }
            w_sync();

            for (int l = 0; l < MNX_DEC_L; ++l) {
                const float* P = pbase + (size_t)l * MG_PARAM_FLOATS_H;
                float* Kc = a.selfK + l * kv_layer;
                float* Vc = a.selfV + l * kv_layer;
                // three sub-layers with the same shape: LayerNorm -> column-parallel tiles -> (attention) -> row-parallel
                // product + all-reduce.  One rolled loop, so each phase routine exists once in the instruction stream.
#pragma unroll 1
                for (int sub = 0; sub < 3; ++sub) {
                    W_MARK();
                    w_layer_norm(c, P + sub * 512, P + sub * 512 + 256);          // LN1 | LN2 | LN_ff: weight, bias
                    W_MARK();
                    const int ntile = (sub == 0) ? 3 : (sub == 1) ? 1 : 4;       // q k v | q_ctx | W1 x 4
                    const float* bias0 = P + ((sub == 0) ? WP_BQ : (sub == 1) ? WP_BQC : WP_B1);
#pragma unroll 1
                    for (int j = 0; j < ntile; ++j) {
                        const float2 bias = reinterpret_cast<const float2*>(bias0 + j * 32)[c.tid & 15];     // every thread: no branch on the load
                        w_col_tile(c, [&](int cp, int rp, float4 v) {
                            float o[4] = {v.x + bias.x, v.y + bias.y, v.z + bias.x, v.w + bias.y};   // (col, row): (0,0) (1,0) (0,1) (1,1)
                            if (sub == 2) {
                                reinterpret_cast<float4*>(sm + WSmem::hT)[W_UNIT(16 * j + cp, rp)] =
                                    make_float4(gelu_erf(o[0]), gelu_erf(o[1]), gelu_erf(o[2]), gelu_erf(o[3]));
                                return;
                            }
                            const bool is_q = (sub == 1) || (j == 0);
                            float* dst = is_q ? qs : (j == 1) ? ks : vs;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (is_q) o[i] = o[i] / W_QSCALE;
                                dst[(2 * rp + (i >> 1)) * 32 + 2 * cp + (i & 1)] = o[i];
                            }
                            if (!is_q) {
                                float* dc = (j == 1) ? Kc : Vc;
#pragma unroll
                                for (int r = 0; r < 2; ++r) {
                                    const int g = 2 * rp + r;
                                    if ((alive >> g) & 1u)
                                        *reinterpret_cast<float2*>(dc + (((size_t)(row0 + g) * 8 + c.h) * a.T + t) * 32 + 2 * cp) = make_float2(o[2 * r], o[2 * r + 1]);
                                }
                                asm volatile("fence.proxy.async;" ::: "memory");
                            }
                        });
                    }
                    W_MARK();
                    if (sub < 2) {
#pragma unroll 1
                        for (int rr = 0; rr < 2; ++rr) {
                            const int row = c.warp + 8 * rr;
                            if (!((alive >> row) & 1u)) continue;
                            const int me = (l * 2 + sub) * 2 + rr;
                            // me != att: step-0 self-attention, only the row's own key (no tiles in global memory)
                            WAtt d, nx;
                            d.K = d.V = nx.K = nx.V = nullptr; d.n = nx.n = 0;
                            if (me == att) {
                                d = att_d;
                                att = w_att_next(c, me);
                                nx = w_att_desc(c, att);
                                att_d = nx;
                            }
                            w_attend(c, row, d, nx, sub == 0);
                        }
                        w_sync();
                    }
                    W_MARK();
                    w_row_gemm_allreduce(c, reinterpret_cast<const ulonglong2*>(sm + ((sub == 2) ? WSmem::hT : WSmem::ctxT)), (sub == 2) ? 4 : 1,
                                         P + ((sub == 0) ? WP_BO : (sub == 1) ? WP_BOC : WP_B2));
                }
                W_MARK();
            }
            // ---------- final LayerNorm, vocabulary slice; the logits of row r go to CTA r / 2 ----------
            w_layer_norm(c, a.finalp, a.finalp + 256);
            {
                const float2 bias = reinterpret_cast<const float2*>(a.finalp + 512 + c.h * 32)[c.tid & 15];
                w_col_tile(c, [&](int cp, int rp, float4 v) {
                    // rows 2rp and 2rp + 1 are both decided by CTA rp
                    const int col = c.h * 32 + 2 * cp;
                    w_send1(c, WSmem::lg + (0 * 256 + col) * 4, (uint32_t)rp, __float_as_uint(v.x + bias.x));
                    w_send1(c, WSmem::lg + (0 * 256 + col + 1) * 4, (uint32_t)rp, __float_as_uint(v.y + bias.y));
                    w_send1(c, WSmem::lg + (1 * 256 + col) * 4, (uint32_t)rp, __float_as_uint(v.z + bias.x));
                    w_send1(c, WSmem::lg + (1 * 256 + col + 1) * 4, (uint32_t)rp, __float_as_uint(v.w + bias.y));
                });
            }
            w_exchange(c, 2u * 256u * 4u);
            // ---------- log_softmax, grammar mask, argmax of rows 2h and 2h + 1 (warps 0 and 1) ----------
            if (c.warp < 2) {
                const int g = 2 * c.h + c.warp, row = row0 + g;
                int bi = s_tf[2 * g], fin = 1;
                if ((alive >> g) & 1u) {
                    float lgv[8];
                    float m = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int v = i * 32 + c.lane;
                        lgv[i] = (v < a.g.vocab) ? lg[c.warp * 256 + v] : -INFINITY;
                        m = fmaxf(m, lgv[i]);
                    }
                    m = warp_max(m);
                    float se = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) se += (i * 32 + c.lane < a.g.vocab) ? expf(lgv[i] - m) : 0.f;
                    se = warp_sum(se);
                    const float lse = logf(se);
                    const int tok_in = s_tf[2 * g];
                    const bool in_x = tok_in >= a.g.offset && tok_in < a.g.offset + a.g.maxx;
                    const bool in_y = tok_in >= a.g.offset + a.g.maxx;
                    float bv = -INFINITY;
                    bi = 1 << 30;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int v = i * 32 + c.lane;
                        float lp = (lgv[i] - m) - lse;
                        if (in_x && v < a.g.offset + a.g.maxx) lp = -10000.0f;
                        if (in_y && v >= a.g.offset) lp = -10000.0f;
                        if (t == 0 && v == a.g.eos) lp = -1e20f;
                        if (v >= a.g.vocab) lp = -INFINITY;
                        if (lp > bv) { bv = lp; bi = v; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                    }
                    fin = (bi == a.g.eos) || (t == a.g.max_len - 1);
                    // hidden state of this step = final LayerNorm output (greedy_search.py:93-97)
                    float* hd = a.hidden + ((size_t)row * a.T + t) * 256;
                    const float* nT = reinterpret_cast<const float*>(sm + WSmem::nT);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k = i * 32 + c.lane;
                        hd[k] = nT[W_ELEM(k, g)];
                    }
                    if (c.lane == 0) {
                        a.ids[(size_t)row * a.T + t] = bi;
                        a.logp[(size_t)row * a.T + t] = bv;
                        if (fin) { a.lens[row] = t + 1; atomicMax(a.steps_run, t + 1); }
                        w_st_release(a.row_state + row, ((unsigned)(t + 1) << 1) | (fin ? 1u : 0u));
                    }
                }
                // token + finished flag of the row to every CTA (finished / absent rows re-send their state)
                if (c.lane < 8) {
                    w_send1(c, WSmem::tf + (2 * g) * 4, (uint32_t)c.lane, (uint32_t)bi);
                    w_send1(c, WSmem::tf + (2 * g + 1) * 4, (uint32_t)c.lane, (uint32_t)fin);
                }
            }
            w_exchange(c, W_G * 8u);
        }
    }
    // nobody may exit while peers can still write into its shared memory or arrive on its barriers
    __syncthreads();
    w_cluster_sync_all();
}

}  // anonymous namespace

cudaError_t wide_configure(int* max_clusters) {
    cudaError_t e = cudaFuncSetAttribute(decode_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WSmem::total);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(8 * 18);
    cfg.blockDim = dim3(W_THREADS);
    cfg.dynamicSmemBytes = WSmem::total;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, decode_wide_kernel, &cfg);
    if (e != cudaSuccess) return e;
    *max_clusters = n;
    return cudaSuccess;
}

cudaError_t wide_launch(const MegaArgs& a, int clusters, cudaStream_t s) {
    decode_wide_kernel<<<dim3(8 * clusters), W_THREADS, WSmem::total, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mnx
