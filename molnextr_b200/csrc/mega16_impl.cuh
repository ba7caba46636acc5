// Persistent decode kernel, 16-CTA clusters (non-portable cluster size): the same algorithm as mega.cu
// with every per-SM byte stream halved.
//
// What limits mega.cu is the ~40 B/clk a single SM can pull from L2: per decode step each CTA of an
// 8-CTA cluster ingests 2.7 MB of fp32 weights plus the K/V rows of its (row, head) pairs.  Here a
// cluster has 16 CTAs and owns G <= 5 rows: CTA i owns a 1/16 column slice of every Linear (16 KB tiles
// [256 k][16 cols], 1.35 MB per step) and the attention of head i/2 for rows {i%2, i%2 + 2, i%2 + 4}.
// All global -> shared traffic (weight tiles, parameter blocks, K/V tiles) is issued by ONE producer
// thread in the order the math consumes it, so latency-critical K/V tiles never queue behind weight
// tiles that are only needed later.  Each CTA runs up to three attention groups of three warps (rows
// half, half + 2, half + 4 of the cluster), so a cluster serves G <= 5 rows and seven co-resident clusters
// cover a batch of 32.  Exchanges between CTAs use st.async into distributed shared
// memory with mbarrier complete_tx signalling; q/k/v slices go only to the CTA that owns the head.
// Arithmetic is fp32 and follows the same reference lines as decoder.cu / mega.cu.
#include "mega.cuh"
#if !defined(H_NG) || !defined(H_GW) || !defined(H_GMAX) || !defined(H_RING) || !defined(H_KERNEL)
#error "include from mega16.cu / mega16s.cu"
#endif

namespace mnx {
namespace {

#define H_CS 16
// configuration (set by the including .cu): H_NG attention groups of H_GW warps per CTA, G <= H_GMAX rows per
// cluster, H_RING weight-tile slots.  mega16.cu: 3 x 3 warps, G <= 5, ring 3 (bs 29..35); mega16s.cu:
// 2 x 4 warps, G <= 4, ring 6 (bs <= 28: four-warp attention groups and no idle ninth warp are ~9 % faster).
#define H_CW (H_NG * H_GW)        // compute warps; warps 0..7 split K in the GEMMs, a ninth only serves attention
#define H_CT (32 * H_CW)          // compute threads
#define H_THREADS (H_CT + 32)     // + 1 producer warp
#define H_GT (32 * H_GW)          // threads per attention group
#define H_STR2(x) #x
#define H_STR(x) H_STR2(x)
#define H_TILE_FLOATS (256 * 16)
#define H_TILE_BYTES (H_TILE_FLOATS * 4)
#define H_TILES_PER_LAYER 14
#define H_PARAM_FLOATS 1728
#define H_TK 144
#define H_QSCALE 5.656854152679443f

enum { HP_LN1W = 0, HP_LN1B = 256, HP_LN2W = 512, HP_LN2B = 768, HP_LNFW = 1024, HP_LNFB = 1280,
       HP_BQ = 1536, HP_BK = 1552, HP_BV = 1568, HP_BO = 1584, HP_BQC = 1600, HP_BOC = 1616, HP_B2 = 1632, HP_B1 = 1648 };

struct HSmem {
    static constexpr int ring = 0;
    static constexpr int kv = ring + H_RING * H_TILE_BYTES;                   // [3 groups][2 bufs][144][32]
    static constexpr int params = kv + H_NG * 2 * H_TK * 128;
    static constexpr int finalp = params + 2 * H_PARAM_FLOATS * 4;
    static constexpr int xbuf = finalp + 768 * 4;
    static constexpr int nbuf = xbuf + H_GMAX * 256 * 4;
    static constexpr int ctxbuf = nbuf + H_GMAX * 256 * 4;
    static constexpr int hbuf = ctxbuf + H_GMAX * 256 * 4;
    // the logits alias the FFN hidden buffer: they are sent after the x3 exchange of the last layer (every
    // CTA has finished reading hbuf) and hbuf is next written after the x1 exchange of the following step
    // (every CTA has finished its argmax)
    static constexpr int lgbuf = hbuf;
    static constexpr int qkvs = hbuf + H_GMAX * 1024 * 4;                      // [3 groups][3][32]
    static constexpr int red = qkvs + H_NG * 3 * 32 * 4;                       // [4 sets][8 warps][G][16]; scores alias
    static constexpr int scores = red;                                         // [3 groups][1024]
    static constexpr int ared = red + H_NG * 1024 * 4;
    static constexpr int misc = ared + H_NG * 160 * 4;
    static constexpr int total = misc + 512;
};
static_assert(4 * 8 * H_GMAX * 16 * 4 <= H_NG * 1024 * 4, "reduction scratch must fit in the scores area");
static_assert(HSmem::total <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ uint32_t h_mapa(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void h_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_H:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_H;\n"
        "bra WAIT_LOOP_H;\n"
        "DONE_H:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void h_cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void h_sync() { asm volatile("bar.sync 1, " H_STR(H_CT) ";" ::: "memory"); }
__device__ __forceinline__ unsigned h_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void h_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct HCtx {
    uint8_t* sm;
    int rank, head, half;           // cluster rank i, head = i / 2, half = i % 2
    int tid, lane, warp;
    int G;
    int grp, gtid, gwarp;           // attention group, thread and warp index inside it (H_GW warps per group)
    uint64_t *full, *empty, *kvfull, *kvempty, *pbar, *xbar, *stepbar;
    uint32_t tile_seq, x_seq, kv_seq;
    uint32_t xbar_base;             // shared::cta address of xbar[0] (same offset in every CTA)
};

__device__ __forceinline__ const float* h_tile_acquire(HCtx& c) {
    const uint32_t slot = c.tile_seq % H_RING, ph = (c.tile_seq / H_RING) & 1u;
    mbar_wait(&c.full[slot], ph);
    return reinterpret_cast<const float*>(c.sm + HSmem::ring + slot * H_TILE_BYTES);
}
__device__ __forceinline__ void h_tile_release(HCtx& c) {
    const uint32_t slot = c.tile_seq % H_RING;
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[slot]);
    ++c.tile_seq;
}
// ---- sliced GEMM: out[g][col] = sum_k X[g][k] * tile[k][col], tile = [256 k][16 cols] fp32 ---------------
// The shared-memory pipe (not the FMA pipe) bounds these products, so the mapping minimises shared-memory
// wavefronts: warp w owns k in [32w, 32w+32); lane = j*8 + r*4 + cg owns the four columns 4cg..4cg+3 and
// the four rows k_i = 32w + 8j + 2i + r (i = 0..3).  Per tile a lane issues four 16-byte weight loads --
// each quarter-warp reads two adjacent rows = 128 contiguous bytes, conflict free, i.e. exactly the tile's
// bytes once -- and takes its 4 x G activations from registers, loaded once per phase and reused by every
// tile that shares them (q|k|v, the four W1 tiles).  The eight lanes that share cg then combine their
// partial sums with a halving butterfly (4G shuffles) that leaves lane (j1, j0, r, cg) holding the warp
// total of column 4cg + 2r + j0 for every row g; the eight warps are combined through shared memory.
__device__ __forceinline__ void h_load_x(const HCtx& c, const float* Xs, int ldx, int koff, float (&xr)[4][H_GMAX]) {
    const int kb = koff + 32 * c.warp + 8 * (c.lane >> 3) + ((c.lane >> 2) & 1);
    // rows g >= G are computed too (their buffers exist and hold finite values; nobody reads the results):
    // keeping the loops free of G-dependent branches keeps the shuffles below convergent
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int g = 0; g < H_GMAX; ++g) xr[i][g] = Xs[g * ldx + kb + 2 * i];
}
__device__ __forceinline__ void h_tile_fma(const HCtx& c, const float* tile, const float (&xr)[4][H_GMAX],
                                           float (&acc)[H_GMAX][4]) {
    const float* tw = tile + (32 * c.warp + 8 * (c.lane >> 3) + ((c.lane >> 2) & 1)) * 16 + 4 * (c.lane & 3);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(tw + 32 * i);
#pragma unroll
        for (int g = 0; g < H_GMAX; ++g) {
            acc[g][0] = fmaf(xr[i][g], w.x, acc[g][0]);
            acc[g][1] = fmaf(xr[i][g], w.y, acc[g][1]);
            acc[g][2] = fmaf(xr[i][g], w.z, acc[g][2]);
            acc[g][3] = fmaf(xr[i][g], w.w, acc[g][3]);
        }
    }
}
// halving butterfly over the 8 lanes that share cg: tot[g] = warp total of column 4cg + 2r + j0
__device__ __forceinline__ void h_warp_reduce(const HCtx& c, const float (&acc)[H_GMAX][4], float (&tot)[H_GMAX]) {
    const bool r = (c.lane >> 2) & 1, j0 = (c.lane >> 3) & 1;
#pragma unroll
    for (int g = 0; g < H_GMAX; ++g) {
            // bit r: lanes with r = 0 keep columns {0,1}, lanes with r = 1 keep {2,3}
            const float s0 = __shfl_xor_sync(0xffffffffu, r ? acc[g][0] : acc[g][2], 4);
            const float s1 = __shfl_xor_sync(0xffffffffu, r ? acc[g][1] : acc[g][3], 4);
            const float u0 = (r ? acc[g][2] : acc[g][0]) + s0;
            const float u1 = (r ? acc[g][3] : acc[g][1]) + s1;
            // bit j0: keep column 2r + j0
            const float s2 = __shfl_xor_sync(0xffffffffu, j0 ? u0 : u1, 8);
            float v = (j0 ? u1 : u0) + s2;
            // bit j1: plain sum (both lanes end up with the total)
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            tot[g] = v;
        }
}
// combine the 8 warps of NS output sets, then call f(set, row, col, value) once per output element.
// Two h_syncs.
template <int NS, class F>
__device__ __forceinline__ void h_reduce_apply(const HCtx& c, float (&tot)[NS][H_GMAX], F f) {
    float* red = reinterpret_cast<float*>(c.sm + HSmem::red);
    if (c.lane < 16 && c.warp < 8) {
        const int col = 4 * (c.lane & 3) + 2 * ((c.lane >> 2) & 1) + ((c.lane >> 3) & 1);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int g = 0; g < H_GMAX; ++g) red[((s * 8 + c.warp) * H_GMAX + g) * 16 + col] = tot[s][g];
    }
    h_sync();
    const int n_out = NS * c.G * 16;
    for (int idx = c.tid; idx < n_out; idx += H_CT) {
        const int col = idx & 15, sg = idx >> 4;
        const int s = sg / c.G, g = sg - s * c.G;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[((s * 8 + w) * H_GMAX + g) * 16 + col];
        f(s, g, col, v);
    }
    h_sync();
}
// NS tiles that share the activations Xs[g][0..255] (ldx floats apart): tot[s][g] per lane
template <int NS>
__device__ __forceinline__ void h_gemm_shared_x(HCtx& c, const float* Xs, int ldx, float (&tot)[NS][H_GMAX]) {
    if (c.warp >= 8) {          // the ninth compute warp only exists for the attention groups
        c.tile_seq += NS;
        return;
    }
    float xr[4][H_GMAX];
    h_load_x(c, Xs, ldx, 0, xr);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        float acc[H_GMAX][4] = {};
        const float* tile = h_tile_acquire(c);
        h_tile_fma(c, tile, xr, acc);
        h_tile_release(c);
        h_warp_reduce(c, acc, tot[s]);
    }
}
// asynchronous remote store that signals the destination CTA's current exchange barrier with its bytes
__device__ __forceinline__ void h_send(const HCtx& c, int byte_off, uint32_t dst_cta, float v) {
    const uint32_t local = smem_u32(c.sm + byte_off);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(h_mapa(local, dst_cta)),
                 "r"(__float_as_uint(v)), "r"(h_mapa(c.xbar_base + 8u * (c.x_seq & 3u), dst_cta))
                 : "memory");
}
__device__ __forceinline__ void h_send4(const HCtx& c, int byte_off, uint32_t dst_cta, float4 v) {
    const uint32_t local = smem_u32(c.sm + byte_off);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     h_mapa(local, dst_cta)),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
                 "r"(h_mapa(c.xbar_base + 8u * (c.x_seq & 3u), dst_cta))
                 : "memory");
}
__device__ __forceinline__ void h_bcast(const HCtx& c, int byte_off, float v) {
#pragma unroll
    for (uint32_t d = 0; d < H_CS; ++d) h_send(c, byte_off, d, v);
}
// all-gather form of h_reduce_apply: element (s, g, col) becomes val(s, g, col, sum) and is written to byte
// offset off(s, g) + 4 * col of EVERY CTA of the cluster with 16-byte remote stores (the DSMEM store path
// moves ~20 B/clk per SM but only ~1 request/clk, so scalar stores were 4x slower).  One task = four adjacent
// columns x 16 / DS destination CTAs; DS is chosen so that one pass over the threads covers all tasks.
// Measured alternatives that were SLOWER on B200: staging the epilogue through shared memory to avoid the
// DS-fold recomputation (one more block barrier per phase costs more than the recomputation), DSMEM bulk
// copies (cp.async.bulk shared::cta -> shared::cluster takes ~1000 cycles to issue).
template <int NS, int DS, class V, class O>
__device__ __forceinline__ void h_reduce_bcast(const HCtx& c, float (&tot)[NS][H_GMAX], V val, O off) {
    float* red = reinterpret_cast<float*>(c.sm + HSmem::red);
    if (c.lane < 16 && c.warp < 8) {
        const int col = 4 * (c.lane & 3) + 2 * ((c.lane >> 2) & 1) + ((c.lane >> 3) & 1);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int g = 0; g < H_GMAX; ++g) red[((s * 8 + c.warp) * H_GMAX + g) * 16 + col] = tot[s][g];
    }
    h_sync();
    const int n_task = NS * c.G * 4 * DS;      // (set, row, column quad, destination group)
    for (int idx = c.tid; idx < n_task; idx += H_CT) {
        const int dq = idx % DS, q = idx / DS;
        const int c4 = q & 3, sg = q >> 2;
        const int s = sg / c.G, g = sg - s * c.G;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float4 p = *reinterpret_cast<const float4*>(red + ((s * 8 + w) * H_GMAX + g) * 16 + 4 * c4);
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        const int col = 4 * c4;
        const float4 o = make_float4(val(s, g, col, v.x), val(s, g, col + 1, v.y), val(s, g, col + 2, v.z), val(s, g, col + 3, v.w));
        const int byte_off = off(s, g) + 16 * c4;
#pragma unroll
        for (uint32_t d = 0; d < 16 / DS; ++d) h_send4(c, byte_off, (uint32_t)dq * (16 / DS) + d, o);
    }
    // no barrier here: the caller waits on the exchange this all-gather feeds, which cannot complete before
    // every thread of THIS CTA has issued its stores (each CTA is one of its own destinations), i.e. before
    // every thread is done reading `red`
}
// finish an exchange in which this CTA receives `bytes_in` bytes in total.  Four barriers rotate: the
// targeted q/k/v exchanges only synchronise a head's two CTAs, so a fast CTA can run up to two
// exchanges ahead of a slow one -- its traffic must never land in a phase the slow CTA still waits on.
__device__ __forceinline__ void h_exchange(HCtx& c, uint32_t bytes_in) {
    uint64_t* bar = c.xbar + (c.x_seq & 3u);
    if (c.tid == 0) mbar_arrive_expect_tx(bar, bytes_in);
    h_wait_cluster(bar, (c.x_seq >> 2) & 1u);
    ++c.x_seq;
}
__device__ __forceinline__ void h_layer_norm(const HCtx& c, const float* w, const float* b) {
    if (c.warp < c.G) {
        const float* x = reinterpret_cast<const float*>(c.sm + HSmem::xbuf) + c.warp * 256;
        float* n = reinterpret_cast<float*>(c.sm + HSmem::nbuf) + c.warp * 256;
        float4 v0 = reinterpret_cast<const float4*>(x)[c.lane];
        float4 v1 = reinterpret_cast<const float4*>(x)[c.lane + 32];
        const float sum = ((v0.x + v0.y) + (v0.z + v0.w)) + ((v1.x + v1.y) + (v1.z + v1.w));
        const float mean = warp_sum(sum) * (1.0f / 256.0f);
        v0.x -= mean; v0.y -= mean; v0.z -= mean; v0.w -= mean;
        v1.x -= mean; v1.y -= mean; v1.z -= mean; v1.w -= mean;
        const float sq = ((v0.x * v0.x + v0.y * v0.y) + (v0.z * v0.z + v0.w * v0.w)) +
                         ((v1.x * v1.x + v1.y * v1.y) + (v1.z * v1.z + v1.w * v1.w));
        const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / 256.0f) + 1e-6f);
        const float4 g0 = reinterpret_cast<const float4*>(w)[c.lane], g1 = reinterpret_cast<const float4*>(w)[c.lane + 32];
        const float4 c0 = reinterpret_cast<const float4*>(b)[c.lane], c1 = reinterpret_cast<const float4*>(b)[c.lane + 32];
        v0.x = v0.x * rstd * g0.x + c0.x; v0.y = v0.y * rstd * g0.y + c0.y;
        v0.z = v0.z * rstd * g0.z + c0.z; v0.w = v0.w * rstd * g0.w + c0.w;
        v1.x = v1.x * rstd * g1.x + c1.x; v1.y = v1.y * rstd * g1.y + c1.y;
        v1.z = v1.z * rstd * g1.z + c1.z; v1.w = v1.w * rstd * g1.w + c1.w;
        reinterpret_cast<float4*>(n)[c.lane] = v0;
        reinterpret_cast<float4*>(n)[c.lane + 32] = v1;
    }
    h_sync();
}
__device__ __forceinline__ void h_group_sync(const HCtx& c) {
    asm volatile("bar.sync %0, " H_STR(H_GT) ";" ::"r"(2 + c.grp) : "memory");
}

// single-query attention of this CTA's head for the row of thread group c.grp (cluster row `g`).
// K/V tiles are delivered by the producer into kv[grp][seq & 1]; with `extra` the row's new key / value
// (already in qkvs) is appended as key index nglobal.  The context slice goes to all 16 CTAs.
__device__ void h_attend(HCtx& c, int g, int nglobal, bool extra) {
    float* scores = reinterpret_cast<float*>(c.sm + HSmem::scores) + c.grp * 1024;
    float* ared = reinterpret_cast<float*>(c.sm + HSmem::ared) + c.grp * 160;   // [H_GW warps][32] partials | 128.. max | 136.. sum
    const float* qs = reinterpret_cast<const float*>(c.sm + HSmem::qkvs) + (c.grp * 3 + 0) * 32;
    const float* ks = reinterpret_cast<const float*>(c.sm + HSmem::qkvs) + (c.grp * 3 + 1) * 32;
    const float* vs = reinterpret_cast<const float*>(c.sm + HSmem::qkvs) + (c.grp * 3 + 2) * 32;
    const int nkeys = nglobal + (extra ? 1 : 0);
    const int ntiles = (nkeys + H_TK - 1) / H_TK;
    float acc = 0.f;
    float mloc = -INFINITY;      // running maximum of the scores this thread computed (saves a pass and a barrier)
    for (int i = 0; i < 2 * ntiles; ++i) {
        const uint32_t seq = c.kv_seq + (uint32_t)i, slot = seq & 1u;
        mbar_wait(&c.kvfull[c.grp * 2 + slot], (seq >> 1) & 1u);
        float* tb = reinterpret_cast<float*>(c.sm + HSmem::kv + (c.grp * 2 + slot) * H_TK * 128);
        const int tile = (i < ntiles) ? i : i - ntiles;
        const int nk = min(H_TK, nkeys - tile * H_TK);
        if (extra && tile == ntiles - 1) {
            if (c.gtid < 32) tb[(nkeys - 1 - tile * H_TK) * 32 + c.gtid] = (i < ntiles) ? ks[c.gtid] : vs[c.gtid];
            h_group_sync(c);
        }
        if (i < ntiles) {
            for (int j = c.gtid; j < nk; j += H_GT) {
                const float4* kr = reinterpret_cast<const float4*>(tb + j * 32);
                float s = 0.f;
#pragma unroll
                for (int cc0 = 0; cc0 < 8; ++cc0) {
                    const int cc = (cc0 + j) & 7;
                    const float4 kv = kr[cc];
                    const float4 qv = reinterpret_cast<const float4*>(qs)[cc];
                    s = fmaf(qv.x, kv.x, s); s = fmaf(qv.y, kv.y, s);
                    s = fmaf(qv.z, kv.z, s); s = fmaf(qv.w, kv.w, s);
                }
                scores[tile * H_TK + j] = s;
                mloc = fmaxf(mloc, s);
            }
        } else {
            if (i == ntiles) {
                // softmax over all keys by the whole group: p = exp(s - max) / sum (fp32, as onmt MultiHeadedAttention)
                float m = warp_max(mloc);
                if (c.lane == 0) ared[128 + c.gwarp] = m;
                h_group_sync(c);                       // also: every score is in shared memory
                m = ared[128];
#pragma unroll
                for (int w = 1; w < H_GW; ++w) m = fmaxf(m, ared[128 + w]);
                float sum = 0.f;
                for (int j = c.gtid; j < nkeys; j += H_GT) {
                    const float e = expf(scores[j] - m);
                    scores[j] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                if (c.lane == 0) ared[136 + c.gwarp] = sum;
                h_group_sync(c);
                sum = (H_GW == 4) ? (ared[136] + ared[137]) + (ared[138] + ared[139]) : (ared[136] + ared[137]) + ared[138];
                for (int j = c.gtid; j < nkeys; j += H_GT) scores[j] = scores[j] / sum;
                h_group_sync(c);
            }
            const float* ps = scores + tile * H_TK;
#pragma unroll 4
            for (int j = c.gwarp; j < nk; j += H_GW) acc = fmaf(ps[j], tb[j * 32 + c.lane], acc);
        }
        h_group_sync(c);   // tile consumed
        if (c.gtid == 0) mbar_arrive(&c.kvempty[c.grp * 2 + slot]);
    }
    c.kv_seq += (uint32_t)(2 * ntiles);
    ared[c.gwarp * 32 + c.lane] = acc;
    h_group_sync(c);
    if (c.gwarp == 0)
        h_bcast(c, HSmem::ctxbuf + (g * 256 + c.head * 32 + c.lane) * 4,
                (H_GW == 4) ? (ared[c.lane] + ared[32 + c.lane]) + (ared[64 + c.lane] + ared[96 + c.lane])
                            : (ared[c.lane] + ared[32 + c.lane]) + ared[64 + c.lane]);
}

__global__ void __launch_bounds__(H_THREADS, 1) H_KERNEL(MegaArgs a) {
    extern __shared__ __align__(128) uint8_t sm[];
    HCtx c;
    c.sm = sm;
    c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    {
        uint32_t r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        c.rank = (int)r;
    }
    c.head = c.rank >> 1; c.half = c.rank & 1;
    c.grp = (c.warp < H_CW) ? c.warp / H_GW : 0; c.gwarp = c.warp - H_GW * c.grp; c.gtid = c.gwarp * 32 + c.lane;
    // cluster order = scheduling order: a ticket taken at kernel start decides which rows a cluster owns, so the spin on
    // `row_state` of lower rows (row-rank PE rule) can never wait for a cluster that is not resident yet -- co-residency of
    // all clusters is measured on an idle GPU, but another engine or kernel may hold SMs when this one starts
    // (the ticket sits in the first word of the weight ring: nothing is copied there before the cluster barrier that follows
    //  the mbarrier initialisation below)
    int* s_ticket = reinterpret_cast<int*>(sm + HSmem::ring);
    if (c.tid == 0 && c.rank == 0) *s_ticket = atomicAdd(a.ticket, 1);
    h_cluster_sync_all();
    int cluster;
    {
        uint32_t v;
        asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(h_mapa(smem_u32(s_ticket), 0)));
        cluster = (int)v;
    }
    const int row0 = cluster * a.G;
    c.G = min(a.G, a.B - row0);
    c.tile_seq = 0; c.x_seq = 0; c.kv_seq = 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + HSmem::misc);
    c.full = bars; c.empty = bars + H_RING; c.kvfull = bars + 2 * H_RING; c.kvempty = c.kvfull + 2 * H_NG;
    c.pbar = c.kvempty + 2 * H_NG; c.xbar = c.pbar + 2; c.stepbar = c.xbar + 4;
    c.xbar_base = smem_u32(&c.xbar[0]);
    int* s_tok = reinterpret_cast<int*>(c.stepbar + 1);
    int* s_fin = s_tok + H_GMAX;
    int* s_rank = s_fin + H_GMAX;
    int* s_go = s_rank + H_GMAX;

    if (c.tid == 0) {
        for (int i = 0; i < H_RING; ++i) { mbar_init(&c.full[i], 1); mbar_init(&c.empty[i], 8); }
        for (int i = 0; i < 2 * H_NG; ++i) { mbar_init(&c.kvfull[i], 1); mbar_init(&c.kvempty[i], 1); }
        mbar_init(&c.pbar[0], 1); mbar_init(&c.pbar[1], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&c.xbar[i], 1);
        mbar_init(c.stepbar, 1);
        fence_barrier_init();
        for (int g = 0; g < H_GMAX; ++g) { s_tok[g] = a.g.sos; s_fin[g] = (g < c.G) ? 0 : 1; s_rank[g] = 0; }
        *s_go = 1;
    }
    for (int i = c.tid; i < 768; i += H_THREADS) reinterpret_cast<float*>(sm + HSmem::finalp)[i] = a.finalp[i];
    h_cluster_sync_all();

    const size_t kv_layer = (size_t)a.B * 8 * a.T * 32;
    const float* wbase = a.wpack16 + (size_t)c.rank * (MNX_DEC_L * H_TILES_PER_LAYER + 1) * H_TILE_FLOATS;
    const float* pbase = a.ppack16 + (size_t)c.rank * MNX_DEC_L * H_PARAM_FLOATS;

    if (c.warp == H_CW) {
        // ======================= producer: every global -> shared transfer, in consumption order =======================
        if (c.lane == 0) {
            const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
            uint32_t seq = 0, pseq = 0, step = 0, kvseq[H_NG] = {};
            auto weight_tile = [&](int index) {
                const uint32_t slot = seq % H_RING, ph = (seq / H_RING) & 1u;
                mbar_wait(&c.empty[slot], ph ^ 1u);
                mbar_arrive_expect_tx(&c.full[slot], H_TILE_BYTES);
                bulk_g2s_hint(sm + HSmem::ring + slot * H_TILE_BYTES, wbase + (size_t)index * H_TILE_FLOATS, H_TILE_BYTES,
                              &c.full[slot], keep);
                ++seq;
            };
            // K/V tile sequence K0..K(n-1), V0..V(n-1) of both live groups, interleaved
            auto kv_tiles = [&](const float* const* Kb, const float* const* Vb, const bool* live, int nglobal,
                                bool extra, uint64_t pol) {
                const int nkeys = nglobal + (extra ? 1 : 0);
                const int ntiles = (nkeys + H_TK - 1) / H_TK;
                for (int i = 0; i < 2 * ntiles; ++i) {
                    for (int p = 0; p < H_NG; ++p) {
                        if (!live[p]) continue;
                        const uint32_t s = kvseq[p] + (uint32_t)i, slot = s & 1u;
                        mbar_wait(&c.kvempty[p * 2 + slot], ((s >> 1) & 1u) ^ 1u);
                        const int tile = (i < ntiles) ? i : i - ntiles;
                        const float* src = ((i < ntiles) ? Kb[p] : Vb[p]) + (size_t)tile * H_TK * 32;
                        const int rows = min(H_TK, nglobal - tile * H_TK);
                        uint64_t* bar = &c.kvfull[p * 2 + slot];
                        if (rows > 0) {
                            asm volatile("fence.proxy.async;" ::: "memory");
                            mbar_arrive_expect_tx(bar, (uint32_t)rows * 128u);
                            bulk_g2s_hint(sm + HSmem::kv + (p * 2 + slot) * H_TK * 128, src, (uint32_t)rows * 128u, bar, pol);
                        } else {
                            mbar_arrive(bar);
                        }
                    }
                }
                for (int p = 0; p < H_NG; ++p)
                    if (live[p]) kvseq[p] += (uint32_t)(2 * ntiles);
            };
            for (;;) {
                mbar_wait(c.stepbar, step & 1u);
                if (*reinterpret_cast<volatile int*>(s_go) == 0) break;
                const int t = (int)step;
                bool live[H_NG];
                int grow[H_NG];
                for (int p = 0; p < H_NG; ++p) {
                    grow[p] = c.half + 2 * p;
                    live[p] = grow[p] < c.G && reinterpret_cast<volatile int*>(s_fin)[grow[p]] == 0;
                }
                for (int l = 0; l < MNX_DEC_L; ++l) {
                    {
                        const uint32_t pb = pseq & 1u;
                        mbar_arrive_expect_tx(&c.pbar[pb], H_PARAM_FLOATS * 4);
                        bulk_g2s_hint(sm + HSmem::params + pb * H_PARAM_FLOATS * 4, pbase + (size_t)l * H_PARAM_FLOATS,
                                      H_PARAM_FLOATS * 4, &c.pbar[pb], keep);
                        ++pseq;
                    }
                    const int base = l * H_TILES_PER_LAYER;
                    weight_tile(base + 0); weight_tile(base + 1); weight_tile(base + 2);          // q, k, v
                    {
                        const float* Kb[H_NG], *Vb[H_NG];
                        for (int p = 0; p < H_NG; ++p) {
                            const size_t off = l * kv_layer + ((size_t)(row0 + grow[p]) * 8 + c.head) * a.T * 32;
                            Kb[p] = a.selfK + off; Vb[p] = a.selfV + off;
                        }
                        kv_tiles(Kb, Vb, live, t, true, stream);
                    }
                    weight_tile(base + 3); weight_tile(base + 4);                                  // Wo, Wq_ctx
                    {
                        const float* Kb[H_NG], *Vb[H_NG];
                        for (int p = 0; p < H_NG; ++p) {
                            const size_t off = (((size_t)l * a.B + row0 + grow[p]) * 8 + c.head) * (size_t)a.S * 32;
                            Kb[p] = a.crossK + off; Vb[p] = a.crossV + off;
                        }
                        kv_tiles(Kb, Vb, live, a.S, false, keep);
                    }
                    for (int i = 5; i < H_TILES_PER_LAYER; ++i) weight_tile(base + i);              // Wo_ctx, W1 x4, W2 x4
                }
                weight_tile(MNX_DEC_L * H_TILES_PER_LAYER);                                        // vocabulary slice
                ++step;
            }
        }
        __syncwarp();
    } else {
        // ======================= compute warps =======================
        float* xbuf = reinterpret_cast<float*>(sm + HSmem::xbuf);
        float* nbuf = reinterpret_cast<float*>(sm + HSmem::nbuf);
        float* ctxbuf = reinterpret_cast<float*>(sm + HSmem::ctxbuf);
        float* hbuf = reinterpret_cast<float*>(sm + HSmem::hbuf);
        float* lgbuf = reinterpret_cast<float*>(sm + HSmem::lgbuf);
        const float* fp = reinterpret_cast<const float*>(sm + HSmem::finalp);
        uint32_t pseq = 0;
        int pm = 0;
#define H_MARK() do { if (a.prof && t == 100 && l == 1 && cluster == 0 && c.rank == 0 && c.tid == 0 && pm < 64) a.prof[pm++] = clock64(); } while (0)
        for (int t = 0;; ++t) {
            int n_alive = 0;
            for (int g = 0; g < c.G; ++g) n_alive += (s_fin[g] == 0) ? 1 : 0;
            if (c.tid == 0) {
                *reinterpret_cast<volatile int*>(s_go) = n_alive > 0 ? 1 : 0;
                __threadfence_block();
                mbar_arrive(c.stepbar);
            }
            if (n_alive == 0) break;
            // rows of this CTA's two attention groups and how many q/k/v slices it will receive
            const int my_g = c.half + 2 * c.grp;                       // cluster row handled by this thread's group
            const bool my_row = my_g < c.G && s_fin[my_g] == 0;
            int n_my = 0;
            for (int p = 0; p < H_NG; ++p) n_my += (c.half + 2 * p < c.G && s_fin[c.half + 2 * p] == 0) ? 1 : 0;
            // ---- rank of each alive row among all alive rows of the batch (row-rank PE rule) ----
            if (c.warp == 0) {
                int finished_before = 0;
                for (int r = c.lane; r < row0; r += 32) {
                    unsigned s;
                    do { s = h_ld_acquire(a.row_state + r); } while ((s >> 1) < (unsigned)t && (s & 1u) == 0u);
                    if ((s & 1u) && (s >> 1) <= (unsigned)t) ++finished_before;
                }
                finished_before = (int)warp_sum((float)finished_before);
                if (c.lane == 0) {
                    int alive_lower = row0 - finished_before;
                    for (int g = 0; g < c.G; ++g) {
                        s_rank[g] = alive_lower;
                        if (s_fin[g] == 0) ++alive_lower;
                    }
                }
            }
            h_sync();
            for (int i = c.tid; i < c.G * 256; i += H_CT) {
                const int g = i >> 8, d = i & 255;
                xbuf[i] = (s_fin[g] == 0) ? a.emb[s_tok[g] * 256 + d] * 16.0f + a.pe[(size_t)s_rank[g] * 256 + d] : 0.f;
            }
            h_sync();

            for (int l = 0; l < MNX_DEC_L; ++l) {
                mbar_wait(&c.pbar[pseq & 1u], (pseq >> 1) & 1u);
                const float* P = reinterpret_cast<const float*>(sm + HSmem::params + (pseq & 1u) * H_PARAM_FLOATS * 4);
                ++pseq;
                float* Kc = a.selfK + l * kv_layer;
                float* Vc = a.selfV + l * kv_layer;
                H_MARK();   // 0: layer start (after param wait)
                // ---------- self attention ----------
                h_layer_norm(c, P + HP_LN1W, P + HP_LN1B);
                H_MARK();   // 1: LN1
                {
                    float acc[3][H_GMAX];
                    h_gemm_shared_x<3>(c, nbuf, 256, acc);
                    h_reduce_apply<3>(c, acc, [&](int which, int g, int col, float v) {
                        float o = v + P[HP_BQ + which * 16 + col];
                        if (s_fin[g]) return;
                        if (which == 0) o = o / H_QSCALE;
                        else {
                            float* dst = (which == 1) ? Kc : Vc;
                            dst[(((size_t)(row0 + g) * 8 + c.head) * a.T + t) * 32 + c.half * 16 + col] = o;
                        }
                        // the slice goes to the CTA that runs this (row, head): cluster rank 2*head + (g & 1), group g >> 1
                        h_send(c, HSmem::qkvs + (((g >> 1) * 3 + which) * 32 + c.half * 16 + col) * 4,
                               (uint32_t)(2 * c.head + (g & 1)), o);
                    });
                }
                H_MARK();   // 2: QKV gemm + sends
                h_exchange(c, (uint32_t)n_my * 2u * 3u * 16u * 4u);       // q, k, v of my rows, from both halves of the head
                H_MARK();   // 3: qkv exchange
                if (my_row) h_attend(c, my_g, t, true);
                H_MARK();   // 4: self attention
                h_exchange(c, (uint32_t)n_alive * 1024u);                // ctx complete everywhere
                H_MARK();   // 5: ctx exchange
                {
                    float acc[1][H_GMAX];
                    h_gemm_shared_x<1>(c, ctxbuf, 256, acc);
                    h_reduce_bcast<1, 4>(c, acc,
                        [&](int, int g, int col, float v) { return (v + P[HP_BO + col]) + xbuf[g * 256 + c.rank * 16 + col]; },
                        [&](int, int g) { return HSmem::xbuf + (g * 256 + c.rank * 16) * 4; });
                }
                H_MARK();   // 6: Wo gemm
                h_exchange(c, (uint32_t)c.G * 1024u);                     // x1
                H_MARK();   // 7: x1 exchange
                // ---------- context attention ----------
                h_layer_norm(c, P + HP_LN2W, P + HP_LN2B);
                {
                    float acc[1][H_GMAX];
                    h_gemm_shared_x<1>(c, nbuf, 256, acc);
                    h_reduce_apply<1>(c, acc, [&](int, int g, int col, float v) {
                        if (s_fin[g]) return;
                        h_send(c, HSmem::qkvs + (((g >> 1) * 3 + 0) * 32 + c.half * 16 + col) * 4,
                               (uint32_t)(2 * c.head + (g & 1)), (v + P[HP_BQC + col]) / H_QSCALE);
                    });
                }
                H_MARK();   // 8: LN2 + Wq gemm
                h_exchange(c, (uint32_t)n_my * 2u * 16u * 4u);
                H_MARK();   // 9: q exchange
                if (my_row) h_attend(c, my_g, a.S, false);
                H_MARK();   // 10: cross attention
                h_exchange(c, (uint32_t)n_alive * 1024u);
                H_MARK();   // 11: ctx exchange
                {
                    float acc[1][H_GMAX];
                    h_gemm_shared_x<1>(c, ctxbuf, 256, acc);
                    h_reduce_bcast<1, 4>(c, acc,
                        [&](int, int g, int col, float v) { return (v + P[HP_BOC + col]) + xbuf[g * 256 + c.rank * 16 + col]; },
                        [&](int, int g) { return HSmem::xbuf + (g * 256 + c.rank * 16) * 4; });
                }
                H_MARK();   // 12: Wo_c gemm
                h_exchange(c, (uint32_t)c.G * 1024u);                     // x2
                H_MARK();   // 13: x2 exchange
                // ---------- feed forward ----------
                h_layer_norm(c, P + HP_LNFW, P + HP_LNFB);
                {
                    float acc[4][H_GMAX];
                    h_gemm_shared_x<4>(c, nbuf, 256, acc);
                    h_reduce_bcast<4, 2>(c, acc,
                        [&](int j, int, int col, float v) { return gelu_erf(v + P[HP_B1 + j * 16 + col]); },
                        [&](int j, int g) { return HSmem::hbuf + (g * 1024 + c.rank * 64 + j * 16) * 4; });
                }
                H_MARK();   // 14: LN + W1
                h_exchange(c, (uint32_t)c.G * 4096u);                     // FFN hidden
                H_MARK();   // 15: h exchange
                {
                    float acc[1][H_GMAX];
                    if (c.warp >= 8) {
                        c.tile_seq += 4;
                    } else {
                        float a4[H_GMAX][4] = {};
#pragma unroll 1
                        for (int j = 0; j < 4; ++j) {
                            float xr[4][H_GMAX];
                            h_load_x(c, hbuf, 1024, 256 * j, xr);
                            const float* tile = h_tile_acquire(c);
                            h_tile_fma(c, tile, xr, a4);
                            h_tile_release(c);
                        }
                        h_warp_reduce(c, a4, acc[0]);
                    }
                    h_reduce_bcast<1, 4>(c, acc,
                        [&](int, int g, int col, float v) { return (v + P[HP_B2 + col]) + xbuf[g * 256 + c.rank * 16 + col]; },
                        [&](int, int g) { return HSmem::xbuf + (g * 256 + c.rank * 16) * 4; });
                }
                H_MARK();   // 16: W2
                h_exchange(c, (uint32_t)c.G * 1024u);                     // x3
                H_MARK();   // 17: x3 exchange
            }
            // ---------- final LayerNorm, vocabulary slice, logits all-gather ----------
            h_layer_norm(c, fp, fp + 256);
            {
                float acc[1][H_GMAX];
                h_gemm_shared_x<1>(c, nbuf, 256, acc);
                h_reduce_bcast<1, 4>(c, acc,
                    [&](int, int, int col, float v) { return v + fp[512 + c.rank * 16 + col]; },
                    [&](int, int g) { return HSmem::lgbuf + (g * 256 + c.rank * 16) * 4; });
            }
            h_exchange(c, (uint32_t)c.G * 1024u);
            // ---------- log_softmax, grammar mask, argmax (identically in every CTA) ----------
            if (c.warp < c.G && s_fin[c.warp] == 0) {
                const int g = c.warp, row = row0 + g;
                float lg[8];
                float m = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int v = i * 32 + c.lane;
                    lg[i] = (v < a.g.vocab) ? lgbuf[g * 256 + v] : -INFINITY;
                    m = fmaxf(m, lg[i]);
                }
                m = warp_max(m);
                float se = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) se += (i * 32 + c.lane < a.g.vocab) ? expf(lg[i] - m) : 0.f;
                se = warp_sum(se);
                const float lse = logf(se);
                const int tok_in = s_tok[g];
                const bool in_x = tok_in >= a.g.offset && tok_in < a.g.offset + a.g.maxx;
                const bool in_y = tok_in >= a.g.offset + a.g.maxx;
                float bv = -INFINITY;
                int bi = 1 << 30;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int v = i * 32 + c.lane;
                    float lp = (lg[i] - m) - lse;
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
                const int fin = (bi == a.g.eos) || (t == a.g.max_len - 1);
                if (c.rank == 0) {
                    float* hd = a.hidden + ((size_t)row * a.T + t) * 256;
#pragma unroll
                    for (int i = 0; i < 8; ++i) hd[i * 32 + c.lane] = nbuf[g * 256 + i * 32 + c.lane];
                    if (c.lane == 0) {
                        a.ids[(size_t)row * a.T + t] = bi;
                        a.logp[(size_t)row * a.T + t] = bv;
                        if (fin) { a.lens[row] = t + 1; atomicMax(a.steps_run, t + 1); }
                        h_st_release(a.row_state + row, ((unsigned)(t + 1) << 1) | (fin ? 1u : 0u));
                    }
                }
                __syncwarp();
                if (c.lane == 0) { s_tok[g] = bi; s_fin[g] = fin; }
            }
            h_sync();
        }
    }
    h_cluster_sync_all();
}

}  // anonymous namespace

cudaError_t H_CONFIGURE(int* max_clusters) {
    cudaError_t e = cudaFuncSetAttribute(H_KERNEL, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(H_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, HSmem::total);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(H_CS * 8);
    cfg.blockDim = dim3(H_THREADS);
    cfg.dynamicSmemBytes = HSmem::total;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = H_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, H_KERNEL, &cfg);
    if (e != cudaSuccess) { *max_clusters = 0; cudaGetLastError(); return cudaSuccess; }   // unsupported -> path disabled
    *max_clusters = n;
    return cudaSuccess;
}

cudaError_t H_LAUNCH(const MegaArgs& a, int clusters, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(H_CS * clusters);
    cfg.blockDim = dim3(H_THREADS);
    cfg.dynamicSmemBytes = HSmem::total;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = H_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, H_KERNEL, a);
}

}  // namespace mnx
