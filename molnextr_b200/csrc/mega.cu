// Persistent cluster decode kernel: the WHOLE greedy decode (all <= 480 steps, 6 layers each) in one
// launch, with no grid-wide synchronisation.
//
// A thread-block cluster of 8 CTAs owns G <= 4 rows (images) for the entire decode; CTA h of the
// cluster is attention head h and owns a 1/8 column slice of every linear layer:
//   * every weight the CTA needs is a contiguous 32 KB tile [256 k][32 cols] (repacked once at load
//     time); a producer warp streams the tiles L2 -> shared memory with 1-D bulk async copies (TMA
//     engine) through a 3-deep mbarrier ring, running ahead of the math across phase boundaries;
//   * activations (residual stream, attention context, FFN hidden, logits) are replicated in every
//     CTA's shared memory and re-assembled after each sliced GEMM by writing the slice into all 8
//     CTAs through distributed shared memory (st.shared::cluster) + a cluster-scope mbarrier;
//   * K/V tiles of the row are staged with bulk copies exactly like the multi-kernel path;
//   * greedy bookkeeping (log-softmax, grammar mask, argmax, <eos>, max length) is done redundantly
//     in every CTA; the only inter-cluster traffic is one word per row per step (`row_state`), which
//     later clusters read to reproduce the reference's "positional encoding indexed by the rank of
//     the row among alive rows" rule (models/embedding.py:52-59).
// Arithmetic is fp32 and mirrors decoder.cu operation for operation (same reference citations).
#include "mega.cuh"

namespace mnx {

#define MG_COMPUTE_THREADS 256
#define MG_THREADS 288            // + 1 producer warp
#define MG_GMAX 4
#define MG_TILE_FLOATS (256 * 32)
#define MG_TILE_BYTES (MG_TILE_FLOATS * 4)
#define MG_RING 3
#define MG_TILES_PER_LAYER 14
#define MG_PARAM_FLOATS 1920      // 1888 used, padded to a multiple of 32 floats
#define MG_TK 72                  // keys per staged K/V tile (two tiles cover S = 144 at 384x384)
#define MG_QSCALE 5.656854152679443f

// tile ids inside a layer
enum { TQ = 0, TK_ = 1, TV = 2, TOS = 3, TQC = 4, TOC = 5, TW1 = 6, TW2 = 10 };
// parameter block offsets (floats)
enum { P_LN1W = 0, P_LN1B = 256, P_LN2W = 512, P_LN2B = 768, P_LNFW = 1024, P_LNFB = 1280,
       P_BQ = 1536, P_BK = 1568, P_BV = 1600, P_BO = 1632, P_BQC = 1664, P_BOC = 1696, P_B1 = 1728, P_B2 = 1856 };

// ---- shared memory carve-up (bytes) ----------------------------------------------------------
struct MegaSmem {
    static constexpr int ring = 0;
    static constexpr int kv = ring + MG_RING * MG_TILE_BYTES;                 // 98304
    // K/V staging: 4 row groups x 2 buffers x 72 keys.  The FFN hidden activations alias this area
    // (no attention tile is in flight between the x2 and x3 exchanges of a layer).
    static constexpr int hbuf = kv;
    static constexpr int params = kv + 8 * MG_TK * 128;                        // +73728
    static constexpr int finalp = params + 2 * MG_PARAM_FLOATS * 4;            // +15360
    static constexpr int xbuf = finalp + 768 * 4;
    static constexpr int nbuf = xbuf + MG_GMAX * 256 * 4;
    static constexpr int ctxbuf = nbuf + MG_GMAX * 256 * 4;
    static constexpr int lgbuf = ctxbuf + MG_GMAX * 256 * 4;
    static constexpr int qkv = lgbuf + MG_GMAX * 256 * 4;                      // q,k,v [G][32] each
    // GEMM cross-warp reduction scratch [4 sets][8 warps][G][32]; the attention scores (one 1024-float
    // row per group) alias it -- the two are never live at the same time.
    static constexpr int red = qkv + 3 * MG_GMAX * 32 * 4;
    static constexpr int scores = red;
    static constexpr int ared = red + 4 * 8 * MG_GMAX * 32 * 4;               // attention scratch [4 groups][64]
    static constexpr int misc = ared + 4 * 64 * 4;
    static constexpr int total = misc + 512;
};
static_assert(MegaSmem::total <= 232448, "shared memory budget exceeded");
static_assert(MG_GMAX * 1024 * 4 <= 8 * MG_TK * 128, "FFN hidden does not fit in the K/V staging area");

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_C:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_C;\n"
        "bra WAIT_LOOP_C;\n"
        "DONE_C:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// all threads of all CTAs of the cluster (non-.aligned form: callers need not be warp-converged)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct MegaCtx {
    uint8_t* sm;
    int h;              // cluster rank = head
    int tid, lane, warp;
    int G;
    uint64_t *full, *empty, *kvbar, *pbar, *xbar, *stepbar;
    uint32_t tile_seq;  // consumer-side running tile counter
    uint32_t kv_seq;    // running K/V staging counter of this thread's row group
    uint32_t x_seq;     // running exchange counter
    int grp, gtid, gwarp;   // attention row group (0/1), thread and warp index inside it
    uint32_t xbar_remote[8];      // exchange barrier 0 of every CTA (barrier 1 is 8 bytes further)
};

// ---- weight-tile ring (consumer side) ---------------------------------------------------------
__device__ __forceinline__ const float* tile_acquire(MegaCtx& c) {
    const uint32_t slot = c.tile_seq % MG_RING, ph = (c.tile_seq / MG_RING) & 1u;
    mbar_wait(&c.full[slot], ph);
    return reinterpret_cast<const float*>(c.sm + MegaSmem::ring + slot * MG_TILE_BYTES);
}
// each warp releases the slot as soon as IT has finished reading the tile (empty barrier counts 8)
__device__ __forceinline__ void tile_release(MegaCtx& c) {
    const uint32_t slot = c.tile_seq % MG_RING;
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[slot]);
    ++c.tile_seq;
}

// acc[g] += sum over this warp's 32 k of Xs[g][koff + 32*warp + k] * tile[32*warp + k][lane]
__device__ __forceinline__ void tile_fma(const MegaCtx& c, const float* tile, const float* Xs, int ldx, int koff,
                                         float (&acc)[MG_GMAX]) {
    const int kb = 32 * c.warp;
#pragma unroll
    for (int kk = 0; kk < 32; kk += 4) {
        const float w0 = tile[(kb + kk + 0) * 32 + c.lane];
        const float w1 = tile[(kb + kk + 1) * 32 + c.lane];
        const float w2 = tile[(kb + kk + 2) * 32 + c.lane];
        const float w3 = tile[(kb + kk + 3) * 32 + c.lane];
#pragma unroll
        for (int g = 0; g < MG_GMAX; ++g) {
            if (g < c.G) {
                const float4 x = *reinterpret_cast<const float4*>(Xs + g * ldx + koff + kb + kk);
                acc[g] = fmaf(x.x, w0, acc[g]);
                acc[g] = fmaf(x.y, w1, acc[g]);
                acc[g] = fmaf(x.z, w2, acc[g]);
                acc[g] = fmaf(x.w, w3, acc[g]);
            }
        }
    }
}
// cross-warp reduction of the 8 k-splits of NS accumulator sets at once, then f(set, row, value) is
// called by one warp per (set, row) pair with lane = output column.  Two compute_syncs; the trailing one is
// dropped (TAIL_SYNC = false) when f broadcasts into every CTA and the caller waits on that exchange next:
// the exchange cannot complete before every thread of THIS CTA has issued its stores (each CTA is one of
// its own destinations), i.e. before everybody is done reading `red`.
template <int NS, bool TAIL_SYNC = true, class F>
__device__ __forceinline__ void reduce_apply(const MegaCtx& c, const float (&acc)[NS][MG_GMAX], F f) {
    float* red = reinterpret_cast<float*>(c.sm + MegaSmem::red);
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int g = 0; g < MG_GMAX; ++g)
            if (g < c.G) red[((s * 8 + c.warp) * MG_GMAX + g) * 32 + c.lane] = acc[s][g];
    compute_sync();
    for (int idx = c.warp; idx < NS * c.G; idx += 8) {
        const int s = idx / c.G, g = idx - s * c.G;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[((s * 8 + w) * MG_GMAX + g) * 32 + c.lane];
        f(s, g, v);
    }
    if (TAIL_SYNC) compute_sync();
}

// write one float into the same shared-memory location of all 8 CTAs of the cluster.  The store is
// asynchronous and signals the destination CTA's exchange mbarrier with its byte count
// (st.async ... mbarrier::complete_tx), so no fence / arrive round trip is needed afterwards.
__device__ __forceinline__ void bcast_store(const MegaCtx& c, int byte_off, float v) {
    const uint32_t local = smem_u32(c.sm + byte_off);
#pragma unroll
    for (uint32_t d = 0; d < 8; ++d) {
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(mapa_u32(local, d)),
                     "r"(__float_as_uint(v)), "r"(c.xbar_remote[d] + 8u * (c.x_seq & 1u))
                     : "memory");
    }
}
// complete an all-gather in which every CTA of the cluster sent `bytes_per_cta` bytes to every CTA.
// Two barriers alternate between consecutive exchanges so a fast peer's next exchange can never be
// counted into the current phase.
__device__ __forceinline__ void exchange_sync(MegaCtx& c, uint32_t bytes_per_cta) {
    uint64_t* bar = c.xbar + (c.x_seq & 1u);
    if (c.tid == 0) mbar_arrive_expect_tx(bar, 8u * bytes_per_cta);
    mbar_wait_cluster(bar, (c.x_seq >> 1) & 1u);
    ++c.x_seq;
}

// LayerNorm (eps 1e-6) of rows xbuf[g] -> nbuf[g]; warp g handles row g
__device__ __forceinline__ void layer_norm_rows(const MegaCtx& c, const float* w, const float* b) {
    if (c.warp < c.G) {
        const float* x = reinterpret_cast<const float*>(c.sm + MegaSmem::xbuf) + c.warp * 256;
        float* n = reinterpret_cast<float*>(c.sm + MegaSmem::nbuf) + c.warp * 256;
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
    compute_sync();
}

// ---- attention: row g of the cluster is handled by thread group g (64 threads = 2 warps) ------------
__device__ __forceinline__ void group_sync(const MegaCtx& c) {
    asm volatile("bar.sync %0, 64;" ::"r"(2 + c.grp) : "memory");
}
// stage tile i of the sequence K0..K(n-1), V0..V(n-1) of this group's row into buffer (seq & 1)
__device__ __forceinline__ void kv_issue(const MegaCtx& c, const float* Kb, const float* Vb, int nglobal, int ntiles,
                                         int i, uint32_t seq) {
    const int tile = (i < ntiles) ? i : i - ntiles;
    const float* src = ((i < ntiles) ? Kb : Vb) + (size_t)tile * MG_TK * 32;
    const int rows = min(MG_TK, nglobal - tile * MG_TK);
    uint64_t* bar = &c.kvbar[c.grp * 2 + (seq & 1u)];
    if (rows > 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_arrive_expect_tx(bar, (uint32_t)rows * 128u);
        bulk_g2s(c.sm + MegaSmem::kv + (c.grp * 2 + (seq & 1u)) * MG_TK * 128, src, (uint32_t)rows * 128u, bar);
    } else {
        mbar_arrive(bar);   // nothing to copy: complete the phase by hand
    }
}
// prefetch the first two tiles; call (by every thread of the group) well before attend_run
__device__ __forceinline__ void attend_arm(const MegaCtx& c, const float* Kb, const float* Vb, int nglobal, bool extra) {
    const int nkeys = nglobal + (extra ? 1 : 0);
    const int ntiles = (nkeys + MG_TK - 1) / MG_TK;
    if (c.gtid == 0) {
        kv_issue(c, Kb, Vb, nglobal, ntiles, 0, c.kv_seq);
        kv_issue(c, Kb, Vb, nglobal, ntiles, 1, c.kv_seq + 1);
    }
}
// single-query attention of head c.h for the group's row g (q/k/v of the row in qkv smem); the context
// slice is written into every CTA of the cluster.  nglobal keys come from global memory; with `extra`
// the row's freshly computed key/value is appended as key index nglobal.
__device__ void attend_run(MegaCtx& c, int g, const float* Kb, const float* Vb, int nglobal, bool extra) {
    float* scores = reinterpret_cast<float*>(c.sm + MegaSmem::scores) + c.grp * 1024;
    float* ared = reinterpret_cast<float*>(c.sm + MegaSmem::ared) + c.grp * 64;
    const float* qs = reinterpret_cast<const float*>(c.sm + MegaSmem::qkv) + g * 32;
    const float* ks = reinterpret_cast<const float*>(c.sm + MegaSmem::qkv) + (MG_GMAX + g) * 32;
    const float* vs = reinterpret_cast<const float*>(c.sm + MegaSmem::qkv) + (2 * MG_GMAX + g) * 32;
    const int nkeys = nglobal + (extra ? 1 : 0);
    const int ntiles = (nkeys + MG_TK - 1) / MG_TK;
    float acc = 0.f;
    float mloc = -INFINITY;      // running maximum of this thread's scores (saves a pass; measured on mega16)
    for (int i = 0; i < 2 * ntiles; ++i) {
        const uint32_t seq = c.kv_seq + (uint32_t)i;
        mbar_wait(&c.kvbar[c.grp * 2 + (seq & 1u)], (seq >> 1) & 1u);
        float* tb = reinterpret_cast<float*>(c.sm + MegaSmem::kv + (c.grp * 2 + (seq & 1u)) * MG_TK * 128);
        const int tile = (i < ntiles) ? i : i - ntiles;
        const int nk = min(MG_TK, nkeys - tile * MG_TK);
        if (extra && tile == ntiles - 1) {
            if (c.gtid < 32) tb[(nkeys - 1 - tile * MG_TK) * 32 + c.gtid] = (i < ntiles) ? ks[c.gtid] : vs[c.gtid];
            group_sync(c);
        }
        if (i < ntiles) {
            for (int j = c.gtid; j < nk; j += 64) {
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
                scores[tile * MG_TK + j] = s;
                mloc = fmaxf(mloc, s);
            }
        } else {
            if (i == ntiles) {
                // softmax over all keys by the group's first warp (shuffles only), p = exp(s - max) / sum
                {   // group maximum: the two warps exchange their maxima through the scratch row
                    const float wm = warp_max(mloc);
                    if (c.lane == 0) ared[c.gwarp] = wm;
                }
                group_sync(c);                         // also: every score is in shared memory
                if (c.gwarp == 0) {
                    const float m = fmaxf(ared[0], ared[1]);
                    float sum = 0.f;
                    for (int j = c.lane; j < nkeys; j += 32) {
                        const float e = expf(scores[j] - m);
                        scores[j] = e;
                        sum += e;
                    }
                    sum = warp_sum(sum);
                    for (int j = c.lane; j < nkeys; j += 32) scores[j] = scores[j] / sum;
                }
                group_sync(c);
            }
            const float* ps = scores + tile * MG_TK;
            for (int j = c.gwarp; j < nk; j += 2) acc = fmaf(ps[j], tb[j * 32 + c.lane], acc);
        }
        group_sync(c);   // tile consumed: refill its buffer with tile i + 2
        if (c.gtid == 0 && i + 2 < 2 * ntiles) kv_issue(c, Kb, Vb, nglobal, ntiles, i + 2, seq + 2);
    }
    c.kv_seq += (uint32_t)(2 * ntiles);
    ared[c.gwarp * 32 + c.lane] = acc;
    group_sync(c);
    if (c.gwarp == 0) bcast_store(c, MegaSmem::ctxbuf + (g * 256 + c.h * 32 + c.lane) * 4, ared[c.lane] + ared[32 + c.lane]);
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(MegaArgs a) {
    extern __shared__ __align__(128) uint8_t sm[];
    MegaCtx c;
    c.sm = sm;
    c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    {
        uint32_t r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        c.h = (int)r;
    }
    // cluster order = scheduling order (see mega16_impl.cuh): rows are handed out by a ticket taken at kernel start
    int* s_ticket = reinterpret_cast<int*>(sm + MegaSmem::ring);      // free until the barrier after the mbarrier initialisation
    if (c.tid == 0 && c.h == 0) *s_ticket = atomicAdd(a.ticket, 1);
    cluster_sync_all();
    int cluster;
    {
        uint32_t v;
        asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(mapa_u32(smem_u32(s_ticket), 0)));
        cluster = (int)v;
    }
    const int row0 = cluster * a.G;
    c.G = min(a.G, a.B - row0);
    c.tile_seq = 0; c.kv_seq = 0; c.x_seq = 0;
    c.grp = (c.warp >> 1) & 3; c.gtid = c.tid & 63; c.gwarp = c.warp & 1;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + MegaSmem::misc);
    c.full = bars; c.empty = bars + MG_RING; c.kvbar = bars + 2 * MG_RING; c.pbar = c.kvbar + 8;
    c.xbar = c.pbar + 2; c.stepbar = c.xbar + 2;
    int* s_tok = reinterpret_cast<int*>(c.stepbar + 1);      // [G]
    int* s_fin = s_tok + MG_GMAX;                             // [G]
    int* s_rank = s_fin + MG_GMAX;                            // [G]
    int* s_go = s_rank + MG_GMAX;                             // producer: continue flag

    if (c.tid == 0) {
        for (int i = 0; i < MG_RING; ++i) { mbar_init(&c.full[i], 1); mbar_init(&c.empty[i], 8); }
        for (int i = 0; i < 8; ++i) mbar_init(&c.kvbar[i], 1);
        mbar_init(&c.pbar[0], 1); mbar_init(&c.pbar[1], 1);
        mbar_init(&c.xbar[0], 1);
        mbar_init(&c.xbar[1], 1);
        mbar_init(c.stepbar, 1);
        fence_barrier_init();
        for (int g = 0; g < MG_GMAX; ++g) { s_tok[g] = a.g.sos; s_fin[g] = (g < c.G) ? 0 : 1; s_rank[g] = 0; }
        *s_go = 1;
    }
    for (int i = c.tid; i < 768; i += MG_THREADS) reinterpret_cast<float*>(sm + MegaSmem::finalp)[i] = a.finalp[i];
#pragma unroll
    for (uint32_t d = 0; d < 8; ++d) c.xbar_remote[d] = mapa_u32(smem_u32(&c.xbar[0]), d);
    // every CTA's barriers must be initialised before anyone arrives remotely
    cluster_sync_all();

    const size_t kv_layer = (size_t)a.B * 8 * a.T * 32;
    const float* wbase = a.wpack + (size_t)c.h * (MNX_DEC_L * MG_TILES_PER_LAYER + 1) * MG_TILE_FLOATS;
    const float* pbase = a.ppack + (size_t)c.h * MNX_DEC_L * MG_PARAM_FLOATS;

    if (c.warp == 8) {
        // =========================== producer warp: weight tiles + parameter blocks ===========================
        if (c.lane == 0) {
            uint32_t seq = 0, pseq = 0, step = 0;
            for (;;) {
                // wait until the consumers decided that step `step` runs
                mbar_wait(c.stepbar, step & 1u);
                if (*reinterpret_cast<volatile int*>(s_go) == 0) break;
                for (int l = 0; l < MNX_DEC_L; ++l) {
                    {   // parameter block of layer l (double buffered)
                        const uint32_t pb = pseq & 1u;
                        // slot reuse is safe: block pseq-2 was consumed two layers ago (consumers sync every tile)
                        mbar_arrive_expect_tx(&c.pbar[pb], MG_PARAM_FLOATS * 4);
                        bulk_g2s(sm + MegaSmem::params + pb * MG_PARAM_FLOATS * 4, pbase + (size_t)l * MG_PARAM_FLOATS,
                                 MG_PARAM_FLOATS * 4, &c.pbar[pb]);
                        ++pseq;
                    }
                    for (int tI = 0; tI < MG_TILES_PER_LAYER; ++tI, ++seq) {
                        const uint32_t slot = seq % MG_RING, ph = (seq / MG_RING) & 1u;
                        mbar_wait(&c.empty[slot], ph ^ 1u);
                        mbar_arrive_expect_tx(&c.full[slot], MG_TILE_BYTES);
                        bulk_g2s(sm + MegaSmem::ring + slot * MG_TILE_BYTES,
                                 wbase + (size_t)(l * MG_TILES_PER_LAYER + tI) * MG_TILE_FLOATS, MG_TILE_BYTES, &c.full[slot]);
                    }
                }
                {   // vocabulary tile
                    const uint32_t slot = seq % MG_RING, ph = (seq / MG_RING) & 1u;
                    mbar_wait(&c.empty[slot], ph ^ 1u);
                    mbar_arrive_expect_tx(&c.full[slot], MG_TILE_BYTES);
                    bulk_g2s(sm + MegaSmem::ring + slot * MG_TILE_BYTES,
                             wbase + (size_t)(MNX_DEC_L * MG_TILES_PER_LAYER) * MG_TILE_FLOATS, MG_TILE_BYTES, &c.full[slot]);
                    ++seq;
                }
                ++step;
            }
        }
        __syncwarp();
    } else {
        // =========================== 8 compute warps ===========================
        float* xbuf = reinterpret_cast<float*>(sm + MegaSmem::xbuf);
        float* nbuf = reinterpret_cast<float*>(sm + MegaSmem::nbuf);
        float* ctxbuf = reinterpret_cast<float*>(sm + MegaSmem::ctxbuf);
        float* hbuf = reinterpret_cast<float*>(sm + MegaSmem::hbuf);
        float* lgbuf = reinterpret_cast<float*>(sm + MegaSmem::lgbuf);
        float* qkv = reinterpret_cast<float*>(sm + MegaSmem::qkv);
        const float* fp = reinterpret_cast<const float*>(sm + MegaSmem::finalp);
        uint32_t pseq = 0;
        int t = 0;
        int pm = 0;
#define MG_MARK() do { if (a.prof && t == 100 && cluster == 0 && c.h == 0 && c.tid == 0 && pm < 64) a.prof[pm++] = clock64(); } while (0)
        for (;; ++t) {
            MG_MARK();
            // ---- is there anything left to do in this cluster? ----
            int n_alive = 0;
            for (int g = 0; g < c.G; ++g) n_alive += (s_fin[g] == 0) ? 1 : 0;
            const bool any = n_alive > 0;
            if (c.tid == 0) {
                *reinterpret_cast<volatile int*>(s_go) = any ? 1 : 0;
                __threadfence_block();
                mbar_arrive(c.stepbar);
            }
            if (!any) break;
            // ---- rank of each alive row among all alive rows of the batch (row-rank PE rule) ----
            if (c.warp == 0) {
                int finished_before = 0;
                for (int r = c.lane; r < row0; r += 32) {
                    unsigned s;
                    do { s = ld_acquire_u32(a.row_state + r); } while ((s >> 1) < (unsigned)t && (s & 1u) == 0u);
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
            compute_sync();
            MG_MARK();
            // ---- embedding: x = emb[tok] * 16 + pe[rank]  (Embeddings / PositionalEncoding) ----
            for (int i = c.tid; i < c.G * 256; i += MG_COMPUTE_THREADS) {
                const int g = i >> 8, d = i & 255;
                xbuf[i] = (s_fin[g] == 0) ? a.emb[s_tok[g] * 256 + d] * 16.0f + a.pe[(size_t)s_rank[g] * 256 + d] : 0.f;
            }
            compute_sync();

            MG_MARK();
            for (int l = 0; l < MNX_DEC_L; ++l) {
                if (l == 1) MG_MARK();
                mbar_wait(&c.pbar[pseq & 1u], (pseq >> 1) & 1u);
                const float* P = reinterpret_cast<const float*>(sm + MegaSmem::params + (pseq & 1u) * MG_PARAM_FLOATS * 4);
                ++pseq;
                float* Kc = a.selfK + l * kv_layer;
                float* Vc = a.selfV + l * kv_layer;
                // ---------- self attention ----------
                // K/V tiles of positions < t are final: start staging them now, they land during LN1 + QKV
                const bool my_row = (c.grp < c.G) && (s_fin[c.grp] == 0);
                const float* sKb = Kc + ((size_t)(row0 + c.grp) * 8 + c.h) * a.T * 32;
                const float* sVb = Vc + ((size_t)(row0 + c.grp) * 8 + c.h) * a.T * 32;
                if (my_row) attend_arm(c, sKb, sVb, t, true);
                layer_norm_rows(c, P + P_LN1W, P + P_LN1B);
                if (l == 1) MG_MARK();
                {
                    float acc[3][MG_GMAX] = {};
#pragma unroll
                    for (int which = 0; which < 3; ++which) {
                        const float* tile = tile_acquire(c);
                        tile_fma(c, tile, nbuf, 256, 0, acc[which]);
                        tile_release(c);
                    }
                    reduce_apply<3>(c, acc, [&](int which, int g, float v) {
                        const float o = v + P[P_BQ + which * 32 + c.lane];
                        if (which == 0) {
                            qkv[g * 32 + c.lane] = o / MG_QSCALE;
                        } else {
                            qkv[(which * MG_GMAX + g) * 32 + c.lane] = o;
                            if (s_fin[g] == 0) {
                                float* dst = (which == 1) ? Kc : Vc;
                                dst[(((size_t)(row0 + g) * 8 + c.h) * a.T + t) * 32 + c.lane] = o;
                            }
                        }
                    });
                }
                compute_sync();
                if (l == 1) MG_MARK();
                const size_t coff = (((size_t)l * a.B + row0 + c.grp) * 8 + c.h) * (size_t)a.S * 32;
                if (my_row) {
                    attend_run(c, c.grp, sKb, sVb, t, true);
                    // buffers are free again: stage the memory-bank K/V for the context attention right away
                    group_sync(c);
                    attend_arm(c, a.crossK + coff, a.crossV + coff, a.S, false);
                }
                if (l == 1) MG_MARK();
                exchange_sync(c, (uint32_t)n_alive * 128u);         // ctx complete everywhere
                if (l == 1) MG_MARK();
                {   // final_linear slice + residual -> x1
                    float acc[1][MG_GMAX] = {};
                    const float* tile = tile_acquire(c);
                    tile_fma(c, tile, ctxbuf, 256, 0, acc[0]);
                    tile_release(c);
                    reduce_apply<1, false>(c, acc, [&](int, int g, float v) {
                        const int col = c.h * 32 + c.lane;
                        bcast_store(c, MegaSmem::xbuf + (g * 256 + col) * 4, (v + P[P_BO + c.lane]) + xbuf[g * 256 + col]);
                    });
                }
                if (l == 1) MG_MARK();
                exchange_sync(c, (uint32_t)c.G * 128u);             // x1 complete everywhere
                if (l == 1) MG_MARK();
                // ---------- context attention ----------
                layer_norm_rows(c, P + P_LN2W, P + P_LN2B);
                {
                    float acc[1][MG_GMAX] = {};
                    const float* tile = tile_acquire(c);
                    tile_fma(c, tile, nbuf, 256, 0, acc[0]);
                    tile_release(c);
                    reduce_apply<1>(c, acc, [&](int, int g, float v) { qkv[g * 32 + c.lane] = (v + P[P_BQC + c.lane]) / MG_QSCALE; });
                }
                compute_sync();
                if (l == 1) MG_MARK();
                if (my_row) attend_run(c, c.grp, a.crossK + coff, a.crossV + coff, a.S, false);
                if (l == 1) MG_MARK();
                exchange_sync(c, (uint32_t)n_alive * 128u);
                if (l == 1) MG_MARK();
                {
                    float acc[1][MG_GMAX] = {};
                    const float* tile = tile_acquire(c);
                    tile_fma(c, tile, ctxbuf, 256, 0, acc[0]);
                    tile_release(c);
                    reduce_apply<1, false>(c, acc, [&](int, int g, float v) {
                        const int col = c.h * 32 + c.lane;
                        bcast_store(c, MegaSmem::xbuf + (g * 256 + col) * 4, (v + P[P_BOC + c.lane]) + xbuf[g * 256 + col]);
                    });
                }
                if (l == 1) MG_MARK();
                exchange_sync(c, (uint32_t)c.G * 128u);             // x2
                if (l == 1) MG_MARK();
                // ---------- feed forward ----------
                layer_norm_rows(c, P + P_LNFW, P + P_LNFB);
                {
                    float acc[4][MG_GMAX] = {};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float* tile = tile_acquire(c);
                        tile_fma(c, tile, nbuf, 256, 0, acc[j]);
                        tile_release(c);
                    }
                    reduce_apply<4, false>(c, acc, [&](int j, int g, float v) {
                        bcast_store(c, MegaSmem::hbuf + (g * 1024 + c.h * 128 + j * 32 + c.lane) * 4,
                                    gelu_erf(v + P[P_B1 + j * 32 + c.lane]));
                    });
                }
                if (l == 1) MG_MARK();
                exchange_sync(c, (uint32_t)c.G * 512u);             // FFN hidden complete
                if (l == 1) MG_MARK();
                {
                    float acc[1][MG_GMAX] = {};
#pragma unroll 1
                    for (int j = 0; j < 4; ++j) {
                        const float* tile = tile_acquire(c);
                        tile_fma(c, tile, hbuf, 1024, 256 * j, acc[0]);
                        tile_release(c);
                    }
                    reduce_apply<1, false>(c, acc, [&](int, int g, float v) {
                        const int col = c.h * 32 + c.lane;
                        bcast_store(c, MegaSmem::xbuf + (g * 256 + col) * 4, (v + P[P_B2 + c.lane]) + xbuf[g * 256 + col]);
                    });
                }
                if (l == 1) MG_MARK();
                exchange_sync(c, (uint32_t)c.G * 128u);             // x3 = layer output
                if (l == 1) MG_MARK();
            }
            MG_MARK();
            // ---------- final LayerNorm, vocabulary slice, all-gather of the logits ----------
            layer_norm_rows(c, fp, fp + 256);
            {
                float acc[1][MG_GMAX] = {};
                const float* tile = tile_acquire(c);
                tile_fma(c, tile, nbuf, 256, 0, acc[0]);
                tile_release(c);
                reduce_apply<1, false>(c, acc, [&](int, int g, float v) {
                    bcast_store(c, MegaSmem::lgbuf + (g * 256 + c.h * 32 + c.lane) * 4, v + fp[512 + c.h * 32 + c.lane]);
                });
            }
            MG_MARK();
            exchange_sync(c, (uint32_t)c.G * 128u);
            MG_MARK();
            // ---------- log_softmax, grammar mask, argmax: warp g decides row g (identically in every CTA) ----------
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
                    if (lp > bv) { bv = lp; bi = v; }      // ascending v: first maximum kept
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                const int fin = (bi == a.g.eos) || (t == a.g.max_len - 1);
                if (c.h == 0) {
                    // hidden state of this step = final LayerNorm output (greedy_search.py:93-97)
                    float* hd = a.hidden + ((size_t)row * a.T + t) * 256;
#pragma unroll
                    for (int i = 0; i < 8; ++i) hd[i * 32 + c.lane] = nbuf[g * 256 + i * 32 + c.lane];
                    if (c.lane == 0) {
                        a.ids[(size_t)row * a.T + t] = bi;
                        a.logp[(size_t)row * a.T + t] = bv;
                        if (fin) { a.lens[row] = t + 1; atomicMax(a.steps_run, t + 1); }
                        st_release_u32(a.row_state + row, ((unsigned)(t + 1) << 1) | (fin ? 1u : 0u));
                    }
                }
                __syncwarp();
                if (c.lane == 0) { s_tok[g] = bi; s_fin[g] = fin; }
            }
            compute_sync();
        }
    }
    // nobody may exit while peers can still write into its shared memory or arrive on its barriers
    cluster_sync_all();
}

// ---- host side --------------------------------------------------------------------------------
size_t mega_smem_bytes() { return (size_t)MegaSmem::total; }

cudaError_t mega_configure(int* max_clusters) {
    cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MegaSmem::total);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(8 * 16);
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = MegaSmem::total;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, decode_mega_kernel, &cfg);
    if (e != cudaSuccess) return e;
    *max_clusters = n;
    return cudaSuccess;
}

cudaError_t mega_launch(const MegaArgs& a, int clusters, cudaStream_t s) {
    decode_mega_kernel<<<dim3(8 * clusters), MG_THREADS, MegaSmem::total, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mnx
