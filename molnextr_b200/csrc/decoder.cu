// Autoregressive transformer decoder for MolNexTR on sm_100a: hand-written fp32 kernels.
//
// Replaces, per decode step, TransformerDecoderAR.decode's loop body
// (MolNexTR/components.py:284-319), TransformerDecoderLayer._forward
// (MolNexTR/models/decoder.py:224-279, onmt MultiHeadedAttention / PositionwiseFeedForward) and
// GreedySearch.advance / update_finished (MolNexTR/decoding/greedy_search.py:76-127): about
// 1650 ATen dispatches and three host syncs per step in the reference become 38 kernels that
// are captured in a CUDA graph and never talk to the host.
//
// Arithmetic is fp32 throughout (FMA on the CUDA cores): at 32..256 alive rows every GEMM here
// is a [rows x 256] x [256 x N] skinny product whose cost is reading the weights out of L2, and
// greedy ids must match the fp32 reference bit for bit, so tensor cores are deliberately not
// used on this path ("decode step").
//
// Kernels per step:  embed_compact | 6 x { ln1+qkv | self-attn+Wo | sum+ln2+q | cross-attn+Wo |
//                    sum+lnff+W1+gelu | W2+residual } | final-ln+vocab+logsoftmax+mask+argmax
#include "decoder.cuh"

#include <math.h>
#include <stdlib.h>

namespace mnx {

// =====================================================================================
// embed + compaction  (GreedySearch.update_finished compaction, greedy_search.py:119-127;
// Embeddings / PositionalEncoding with the row-rank rule, models/embedding.py:42-61)
// =====================================================================================
// input token of step t for an original row: the previous pick, unless a label is given for this position
// (tgt = tgt * mask + label * (1 - mask), components.py:286-289)
__device__ __forceinline__ int input_token(const DecBuffers& b, const Grammar& g, int lab_len, int row, int t) {
    int tok = (t == 0) ? g.sos : b.cur_tok[row];
    if (t < lab_len) {
        const int lab = b.labels[(size_t)row * (b.T + 1) + t];
        if (lab != MNX_MASK_ID) tok = lab;
    }
    return tok;
}

__global__ void __launch_bounds__(1024) embed_compact_kernel(DecBuffers b, DecWeights w, Grammar g) {
    __shared__ int warp_cnt[32];
    __shared__ int warp_excl[32];
    __shared__ int chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    DecState* st = b.st;
    const int t = st->next_step;
    const int n_prev = (t == 0) ? b.B : st->n_alive;
    const int* prev = b.alive + ((t + 1) & 1) * b.B;
    int* cur = b.alive + (t & 1) * b.B;
    int n = 0;   // running number of survivors (uniform across the block)
    if (t < g.max_len) {
        for (int base = 0; base < n_prev; base += 1024) {
            const int i = base + tid;
            int row = -1, keep = 0;
            if (i < n_prev) {
                row = (t == 0) ? i : prev[i];
                keep = (t == 0) ? 1 : (b.finished[row] == 0);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int pre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) warp_cnt[wid] = __popc(bal);
            __syncthreads();
            if (wid == 0) {
                const int v = warp_cnt[lane];
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                warp_excl[lane] = incl - v;
                if (lane == 31) chunk_total = incl;
            }
            __syncthreads();
            if (keep) cur[n + warp_excl[wid] + pre] = row;   // order of alive rows is preserved
            n += chunk_total;
            __syncthreads();
        }
    }
    if (tid == 0) {
        st->n_alive = n;
        st->step = t;
        st->next_step = t + 1;
        if (n == 0) st->done = 1; else st->steps_run = t + 1;
    }
    __syncthreads();   // cur[] complete before it is read below
    const int lab_len = st->lab_len;
    // x[rank] = emb[tok] * sqrt(256) + pe[rank]      (row-rank rule)
    for (int idx = tid; idx < n * MNX_DEC_D; idx += 1024) {
        const int rank = idx >> 8, d = idx & 255;
        const int row = cur[rank];
        const int tok = input_token(b, g, lab_len, row, t);
        b.xa[idx] = w.emb[tok * MNX_DEC_D + d] * 16.0f + w.pe[rank * MNX_DEC_D + d];
    }
}

// Beam search: the same compaction over IMAGES (BeamSearch.update_finished removes an image
// when its top beam has finished and it holds n_best hypotheses, beam_search.py:156-190); the
// row list handed to the layer kernels is image-major, beam-minor, and row rank r gets pe[r].
__global__ void __launch_bounds__(1024) embed_compact_beam_kernel(DecBuffers b, DecWeights w, Grammar g, BeamBuffers bm) {
    __shared__ int warp_cnt[32];
    __shared__ int warp_excl[32];
    __shared__ int chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    DecState* st = b.st;
    const int t = st->next_step;
    const int K = bm.beam;
    const int n_prev = (t == 0) ? bm.n_img0 : st->n_img;
    const int* prev = bm.alive_img + ((t + 1) & 1) * bm.n_img0;
    int* cur = bm.alive_img + (t & 1) * bm.n_img0;
    int n = 0;
    if (t < g.max_len) {
        for (int base = 0; base < n_prev; base += 1024) {
            const int i = base + tid;
            int img = -1, keep = 0;
            if (i < n_prev) {
                img = (t == 0) ? i : prev[i];
                keep = (t == 0) ? 1 : (bm.img_done[img] == 0);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int pre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) warp_cnt[wid] = __popc(bal);
            __syncthreads();
            if (wid == 0) {
                const int v = warp_cnt[lane];
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                warp_excl[lane] = incl - v;
                if (lane == 31) chunk_total = incl;
            }
            __syncthreads();
            if (keep) cur[n + warp_excl[wid] + pre] = img;
            n += chunk_total;
            __syncthreads();
        }
    }
    if (tid == 0) {
        st->n_img = n;
        st->n_alive = n * K;
        st->step = t;
        st->next_step = t + 1;
        if (n == 0) st->done = 1; else st->steps_run = t + 1;
    }
    __syncthreads();
    int* rows = b.alive + (t & 1) * b.B;
    for (int r = tid; r < n * K; r += 1024) rows[r] = cur[r / K] * K + (r % K);
    for (int idx = tid; idx < n * K * MNX_DEC_D; idx += 1024) {
        const int rank = idx >> 8, d = idx & 255;
        const int slot = cur[rank / K] * K + (rank % K);
        const int tok = (t == 0) ? g.sos : b.cur_tok[slot];
        b.xa[idx] = w.emb[tok * MNX_DEC_D + d] * 16.0f + w.pe[rank * MNX_DEC_D + d];
    }
}

// =====================================================================================
// skinny GEMM: out[rows<=32][N] = f(LN(x))[rows][K] * Wt[K][N]  (+ fused prologue / epilogue)
// lane = output column, 8 warps split K, cross-warp reduction in shared memory.
// =====================================================================================
enum { PRO_LN = 0, PRO_SUM_LN = 1, PRO_PLAIN = 2 };
enum { EPI_QKV = 0, EPI_Q = 1, EPI_GELU = 2, EPI_PART = 3 };

struct SkinnyArgs {
    const DecState* st;
    const int* alive;       // [2][B]
    int B;
    const float* x_in;      // PRO_LN / PRO_SUM_LN: [B][256] residual stream; PRO_PLAIN: [B][x_stride]
    int x_stride;           // PRO_PLAIN: row stride of x_in (the CTA reads columns blockIdx.z*K .. +K)
    const float* part;      // PRO_SUM_LN: [B][8][256] partial sums of the previous linear layer
    const float* bo;        // PRO_SUM_LN: that layer's bias [256]
    float* x_sum_out;       // PRO_SUM_LN: x_in + bo + sum(part), written by column-block 0
    const float* ln_w;
    const float* ln_b;
    const float* wt;        // [Ktotal][N]
    const float* bias;      // [N]
    int N;
    float* out;             // EPI_Q/EPI_QKV: q [B][256]; EPI_GELU: [B][N]; EPI_PART: part [B][8][256]
    float* kc;              // EPI_QKV: this layer's self K cache [B][8][T][32]
    float* vc;
    int T;
};

#define SKINNY_QSCALE 5.656854152679443f   // float(math.sqrt(32)): `query / math.sqrt(dim_per_head)`

// K = reduction length handled by one CTA (256 for the d_model inputs, 128 = one of 8 slices of
// the FFN hidden layer).  8 warps split K; lane = output column; 32 rows per CTA.
template <int K, int PRO, int EPI>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(SkinnyArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* Xs = smem;                 // [32][K]
    float* red = smem + 32 * K;       // [8][32][32]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_alive = a.st->n_alive;
    const int rank0 = blockIdx.y * 32;
    if (rank0 >= n_alive) return;
    const int nrows = min(32, n_alive - rank0);
    const int n = blockIdx.x * 32 + lane;     // this lane's output column
    constexpr int KW = K / 8;                  // k-range per warp
    constexpr int KB = KW < 32 ? KW : 32;      // k-block held in registers
    const int kslice = blockIdx.z;
    const int k0 = kslice * K + wid * KW;      // row of wt where this warp starts

    // weights for the first k-block: issued before the prologue so the L2 round trip overlaps it
    float wv[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) wv[i] = a.wt[(size_t)(k0 + i) * a.N + n];

    // ---------------- prologue: fill Xs ----------------
    if (PRO == PRO_PLAIN) {
#pragma unroll 4
        for (int idx = tid; idx < 32 * (K / 4); idx += 256) {
            const int r = idx / (K / 4), c4 = idx % (K / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows)
                v = reinterpret_cast<const float4*>(a.x_in + (size_t)(rank0 + r) * a.x_stride + kslice * K)[c4];
            reinterpret_cast<float4*>(Xs + r * K)[c4] = v;
        }
        __syncthreads();
    } else {
        // phase 1 (flat, all loads independent): Xs = x (+ bias + partial sums)
#pragma unroll 2
        for (int idx = tid; idx < 32 * 64; idx += 256) {
            const int r = idx >> 6, c4 = idx & 63;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) {
                v = reinterpret_cast<const float4*>(a.x_in + (size_t)(rank0 + r) * MNX_DEC_D)[c4];
                if (PRO == PRO_SUM_LN) {
                    const float4 bb = reinterpret_cast<const float4*>(a.bo)[c4];
                    const float* pr = a.part + (size_t)(rank0 + r) * 8 * MNX_DEC_D;
                    float4 p[8];
#pragma unroll
                    for (int h = 0; h < 8; ++h) p[h] = reinterpret_cast<const float4*>(pr + h * MNX_DEC_D)[c4];
                    float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int h = 0; h < 8; ++h) { sacc.x += p[h].x; sacc.y += p[h].y; sacc.z += p[h].z; sacc.w += p[h].w; }
                    v.x = (sacc.x + bb.x) + v.x; v.y = (sacc.y + bb.y) + v.y;
                    v.z = (sacc.z + bb.z) + v.z; v.w = (sacc.w + bb.w) + v.w;
                    if (blockIdx.x == 0)
                        reinterpret_cast<float4*>(a.x_sum_out + (size_t)(rank0 + r) * MNX_DEC_D)[c4] = v;
                }
            }
            reinterpret_cast<float4*>(Xs + r * K)[c4] = v;
        }
        __syncthreads();
        // phase 2: LayerNorm (eps 1e-6) in place, one warp per row, data already on chip
        const float4 g0 = reinterpret_cast<const float4*>(a.ln_w)[lane];
        const float4 g1 = reinterpret_cast<const float4*>(a.ln_w)[lane + 32];
        const float4 c0 = reinterpret_cast<const float4*>(a.ln_b)[lane];
        const float4 c1 = reinterpret_cast<const float4*>(a.ln_b)[lane + 32];
        for (int r = wid; r < nrows; r += 8) {
            float4 v0 = reinterpret_cast<float4*>(Xs + r * K)[lane];
            float4 v1 = reinterpret_cast<float4*>(Xs + r * K)[lane + 32];
            const float sum = ((v0.x + v0.y) + (v0.z + v0.w)) + ((v1.x + v1.y) + (v1.z + v1.w));
            const float mean = warp_sum(sum) * (1.0f / MNX_DEC_D);
            v0.x -= mean; v0.y -= mean; v0.z -= mean; v0.w -= mean;
            v1.x -= mean; v1.y -= mean; v1.z -= mean; v1.w -= mean;
            const float sq = ((v0.x * v0.x + v0.y * v0.y) + (v0.z * v0.z + v0.w * v0.w)) +
                             ((v1.x * v1.x + v1.y * v1.y) + (v1.z * v1.z + v1.w * v1.w));
            const float var = warp_sum(sq) * (1.0f / MNX_DEC_D);
            const float rstd = 1.0f / sqrtf(var + 1e-6f);
            v0.x = v0.x * rstd * g0.x + c0.x; v0.y = v0.y * rstd * g0.y + c0.y;
            v0.z = v0.z * rstd * g0.z + c0.z; v0.w = v0.w * rstd * g0.w + c0.w;
            v1.x = v1.x * rstd * g1.x + c1.x; v1.y = v1.y * rstd * g1.y + c1.y;
            v1.z = v1.z * rstd * g1.z + c1.z; v1.w = v1.w * rstd * g1.w + c1.w;
            reinterpret_cast<float4*>(Xs + r * K)[lane] = v0;
            reinterpret_cast<float4*>(Xs + r * K)[lane + 32] = v1;
        }
        __syncthreads();
    }

    // ---------------- main: acc[r] += Xs[r][k] * W[k][n] over this warp's k-range ----------------
    float acc[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) acc[r] = 0.f;
#pragma unroll 1
    for (int kb = 0; kb < KW; kb += KB) {
        float wn[KB];
        if (kb + KB < KW) {
#pragma unroll
            for (int i = 0; i < KB; ++i) wn[i] = a.wt[(size_t)(k0 + kb + KB + i) * a.N + n];
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const float4* xr = reinterpret_cast<const float4*>(Xs + r * K + wid * KW + kb);
#pragma unroll
            for (int i4 = 0; i4 < KB / 4; ++i4) {
                const float4 xv = xr[i4];
                acc[r] = fmaf(xv.x, wv[4 * i4 + 0], acc[r]);
                acc[r] = fmaf(xv.y, wv[4 * i4 + 1], acc[r]);
                acc[r] = fmaf(xv.z, wv[4 * i4 + 2], acc[r]);
                acc[r] = fmaf(xv.w, wv[4 * i4 + 3], acc[r]);
            }
        }
        if (kb + KB < KW) {
#pragma unroll
            for (int i = 0; i < KB; ++i) wv[i] = wn[i];
        }
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) red[(wid * 32 + r) * 32 + lane] = acc[r];
    __syncthreads();

    // ---------------- cross-warp reduction + epilogue ----------------
    const float bias = (EPI == EPI_PART) ? 0.f : a.bias[n];
    const int t = a.st->step;
    const int* alive = a.alive + (t & 1) * a.B;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = wid * 4 + rr;
        if (r >= nrows) break;
        float v = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) v += red[(w8 * 32 + r) * 32 + lane];
        v += bias;
        const int rank = rank0 + r;
        if (EPI == EPI_QKV) {
            const int sec = n >> 8, f = n & 255;
            if (sec == 0) {
                a.out[(size_t)rank * MNX_DEC_D + f] = v / SKINNY_QSCALE;
            } else {
                const int row = alive[rank];
                const int h = f >> 5, d = f & 31;
                float* dst = (sec == 1) ? a.kc : a.vc;
                dst[(((size_t)row * 8 + h) * a.T + t) * 32 + d] = v;
            }
        } else if (EPI == EPI_Q) {
            a.out[(size_t)rank * MNX_DEC_D + n] = v / SKINNY_QSCALE;
        } else if (EPI == EPI_GELU) {
            a.out[(size_t)rank * a.N + n] = gelu_erf(v);
        } else {   // EPI_PART: k-slice partial of the W2 product; bias and residual are added by the consumer
            a.out[((size_t)rank * 8 + kslice) * MNX_DEC_D + n] = v;
        }
    }
}

// =====================================================================================
// The same Linear for MANY rows (beam search: images x beams = up to 1280; greedy shards of > 288 rows): a register-
// tiled fp32 GEMM with the skinny kernel's prologues and epilogues.  One CTA = 64 rows x 64 output columns, 256 threads
// as 16 x 16, each thread 4 rows x 4 columns; activations of the 64 rows sit in shared memory ([64][K + 4], the
// LayerNorm is applied there), weights stream through a double-buffered [32 k][64 cols] tile.  Per 4 k a thread issues
// 4 + 4 LDS.128 for 64 FMA (the skinny kernel: one broadcast LDS.128 per 4 FMA per lane, which is what held it to
// ~8 % of the FMA peak at 1280 rows).  Same arithmetic per output as skinny_gemm_kernel except the order of the k sum
// (sequential here, 8 warp-partials there): results agree to fp32 rounding, ids are checked by the same tests.
// =====================================================================================
#define TG_ROWS 64
#define TG_COLS 64
#define TG_KC 32
template <int K, int PRO, int EPI>
__global__ void __launch_bounds__(256, 2) tile_gemm_kernel(SkinnyArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LDX = K + 4;
    float* Xs = smem;                          // [64][K + 4]
    float* Ws = smem + TG_ROWS * LDX;          // [2][32][64]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int n_alive = a.st->n_alive;
    const int rank0 = blockIdx.y * TG_ROWS;
    if (rank0 >= n_alive) return;
    const int nrows = min(TG_ROWS, n_alive - rank0);
    const int n0 = blockIdx.x * TG_COLS;
    const int kslice = blockIdx.z;
    const float* wbase = a.wt + (size_t)kslice * K * a.N + n0;     // row k of this CTA's weight tile: wbase + k * N

    // first weight tile in flight before the prologue
    float4 wreg[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int idx = tid + 256 * j;                 // 512 float4 = 32 k x 16 column quads
        wreg[j] = *reinterpret_cast<const float4*>(wbase + (size_t)(idx >> 4) * a.N + (idx & 15) * 4);
    }

    // ---------------- prologue: fill Xs (identical arithmetic to skinny_gemm_kernel) ----------------
    if (PRO == PRO_PLAIN) {
        for (int idx = tid; idx < TG_ROWS * (K / 4); idx += 256) {
            const int r = idx / (K / 4), c4 = idx % (K / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows)
                v = reinterpret_cast<const float4*>(a.x_in + (size_t)(rank0 + r) * a.x_stride + kslice * K)[c4];
            *reinterpret_cast<float4*>(Xs + r * LDX + 4 * c4) = v;
        }
        __syncthreads();
    } else {
        for (int idx = tid; idx < TG_ROWS * 64; idx += 256) {
            const int r = idx >> 6, c4 = idx & 63;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) {
                v = reinterpret_cast<const float4*>(a.x_in + (size_t)(rank0 + r) * MNX_DEC_D)[c4];
                if (PRO == PRO_SUM_LN) {
                    const float4 bb = reinterpret_cast<const float4*>(a.bo)[c4];
                    const float* pr = a.part + (size_t)(rank0 + r) * 8 * MNX_DEC_D;
                    float4 p[8];
#pragma unroll
                    for (int h = 0; h < 8; ++h) p[h] = reinterpret_cast<const float4*>(pr + h * MNX_DEC_D)[c4];
                    float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int h = 0; h < 8; ++h) { sacc.x += p[h].x; sacc.y += p[h].y; sacc.z += p[h].z; sacc.w += p[h].w; }
                    v.x = (sacc.x + bb.x) + v.x; v.y = (sacc.y + bb.y) + v.y;
                    v.z = (sacc.z + bb.z) + v.z; v.w = (sacc.w + bb.w) + v.w;
                    if (blockIdx.x == 0)
                        reinterpret_cast<float4*>(a.x_sum_out + (size_t)(rank0 + r) * MNX_DEC_D)[c4] = v;
                }
            }
            *reinterpret_cast<float4*>(Xs + r * LDX + 4 * c4) = v;
        }
        __syncthreads();
        const float4 g0 = reinterpret_cast<const float4*>(a.ln_w)[lane];
        const float4 g1 = reinterpret_cast<const float4*>(a.ln_w)[lane + 32];
        const float4 c0 = reinterpret_cast<const float4*>(a.ln_b)[lane];
        const float4 c1 = reinterpret_cast<const float4*>(a.ln_b)[lane + 32];
        for (int r = wid; r < nrows; r += 8) {
            float4 v0 = *reinterpret_cast<float4*>(Xs + r * LDX + 4 * lane);
            float4 v1 = *reinterpret_cast<float4*>(Xs + r * LDX + 4 * (lane + 32));
            const float sum = ((v0.x + v0.y) + (v0.z + v0.w)) + ((v1.x + v1.y) + (v1.z + v1.w));
            const float mean = warp_sum(sum) * (1.0f / MNX_DEC_D);
            v0.x -= mean; v0.y -= mean; v0.z -= mean; v0.w -= mean;
            v1.x -= mean; v1.y -= mean; v1.z -= mean; v1.w -= mean;
            const float sq = ((v0.x * v0.x + v0.y * v0.y) + (v0.z * v0.z + v0.w * v0.w)) +
                             ((v1.x * v1.x + v1.y * v1.y) + (v1.z * v1.z + v1.w * v1.w));
            const float var = warp_sum(sq) * (1.0f / MNX_DEC_D);
            const float rstd = 1.0f / sqrtf(var + 1e-6f);
            v0.x = v0.x * rstd * g0.x + c0.x; v0.y = v0.y * rstd * g0.y + c0.y;
            v0.z = v0.z * rstd * g0.z + c0.z; v0.w = v0.w * rstd * g0.w + c0.w;
            v1.x = v1.x * rstd * g1.x + c1.x; v1.y = v1.y * rstd * g1.y + c1.y;
            v1.z = v1.z * rstd * g1.z + c1.z; v1.w = v1.w * rstd * g1.w + c1.w;
            *reinterpret_cast<float4*>(Xs + r * LDX + 4 * lane) = v0;
            *reinterpret_cast<float4*>(Xs + r * LDX + 4 * (lane + 32)) = v1;
        }
        // (the barrier of the first k-chunk below orders these writes before the reads)
    }

    // ---------------- main loop ----------------
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    constexpr int NCH = K / TG_KC;
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
        float* wt = Ws + (ch & 1) * (TG_KC * TG_COLS);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int idx = tid + 256 * j;
            *reinterpret_cast<float4*>(wt + (idx >> 4) * TG_COLS + (idx & 15) * 4) = wreg[j];
        }
        __syncthreads();          // tile ch visible; every thread is done with tile ch-1 (the buffer the NEXT store overwrites)
        if (ch + 1 < NCH) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int idx = tid + 256 * j;
                wreg[j] = *reinterpret_cast<const float4*>(wbase + (size_t)((ch + 1) * TG_KC + (idx >> 4)) * a.N + (idx & 15) * 4);
            }
        }
        const float* xs = Xs + (4 * ty) * LDX + ch * TG_KC;
#pragma unroll
        for (int kk = 0; kk < TG_KC; kk += 4) {
            float4 xv[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + i * LDX + kk);
#pragma unroll
            for (int q = 0; q < 4; ++q) wv[q] = *reinterpret_cast<const float4*>(wt + (kk + q) * TG_COLS + 4 * tx);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(xv[i].x, wv[0].x, acc[i][0]); acc[i][1] = fmaf(xv[i].x, wv[0].y, acc[i][1]);
                acc[i][2] = fmaf(xv[i].x, wv[0].z, acc[i][2]); acc[i][3] = fmaf(xv[i].x, wv[0].w, acc[i][3]);
                acc[i][0] = fmaf(xv[i].y, wv[1].x, acc[i][0]); acc[i][1] = fmaf(xv[i].y, wv[1].y, acc[i][1]);
                acc[i][2] = fmaf(xv[i].y, wv[1].z, acc[i][2]); acc[i][3] = fmaf(xv[i].y, wv[1].w, acc[i][3]);
                acc[i][0] = fmaf(xv[i].z, wv[2].x, acc[i][0]); acc[i][1] = fmaf(xv[i].z, wv[2].y, acc[i][1]);
                acc[i][2] = fmaf(xv[i].z, wv[2].z, acc[i][2]); acc[i][3] = fmaf(xv[i].z, wv[2].w, acc[i][3]);
                acc[i][0] = fmaf(xv[i].w, wv[3].x, acc[i][0]); acc[i][1] = fmaf(xv[i].w, wv[3].y, acc[i][1]);
                acc[i][2] = fmaf(xv[i].w, wv[3].z, acc[i][2]); acc[i][3] = fmaf(xv[i].w, wv[3].w, acc[i][3]);
            }
        }
    }

    // ---------------- epilogue ----------------
    const int n = n0 + 4 * tx;                                   // first of this thread's four columns
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI != EPI_PART) bias = *reinterpret_cast<const float4*>(a.bias + n);
    const int t = a.st->step;
    const int* alive = a.alive + (t & 1) * a.B;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 4 * ty + i;
        if (r >= nrows) break;
        const int rank = rank0 + r;
        float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
        if (EPI == EPI_QKV) {
            const int sec = n >> 8, f = n & 255;
            if (sec == 0) {
                v.x = v.x / SKINNY_QSCALE; v.y = v.y / SKINNY_QSCALE; v.z = v.z / SKINNY_QSCALE; v.w = v.w / SKINNY_QSCALE;
                *reinterpret_cast<float4*>(a.out + (size_t)rank * MNX_DEC_D + f) = v;
            } else {
                const int row = alive[rank];
                const int h = f >> 5, d = f & 31;
                float* dst = (sec == 1) ? a.kc : a.vc;
                *reinterpret_cast<float4*>(dst + (((size_t)row * 8 + h) * a.T + t) * 32 + d) = v;
            }
        } else if (EPI == EPI_Q) {
            v.x = v.x / SKINNY_QSCALE; v.y = v.y / SKINNY_QSCALE; v.z = v.z / SKINNY_QSCALE; v.w = v.w / SKINNY_QSCALE;
            *reinterpret_cast<float4*>(a.out + (size_t)rank * MNX_DEC_D + n) = v;
        } else if (EPI == EPI_GELU) {
            v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
            *reinterpret_cast<float4*>(a.out + (size_t)rank * a.N + n) = v;
        } else {
            *reinterpret_cast<float4*>(a.out + ((size_t)rank * 8 + kslice) * MNX_DEC_D + n) = v;
        }
    }
}

// =====================================================================================
// single-query attention for one (alive row, head) + that head's slice of final_linear.
// K/V tiles of the row are staged into shared memory with 1-D bulk async copies (TMA engine)
// double-buffered on mbarriers; scores/softmax/PV in fp32 exactly as onmt MultiHeadedAttention.
// =====================================================================================
struct AttnArgs {
    const DecState* st;
    const int* alive;
    int B;
    const float* q;       // [B][256], already divided by sqrt(32)
    const float* Kc;      // layer base, [B][8][cap][32]
    const float* Vc;
    int cap;              // T (self) or S (cross)
    int nkeys_cross;      // S for cross attention; self uses step+1
    const float* wo_t;    // [256][256] final_linear, K-major
    float* part;          // [B][8][256]
    int kv_div;           // cross attention under beam search: K/V row = row / kv_div (beams share their image's memory)
    const int* anc;       // beam self-attention: [B][cap] ancestry rows of this step (nullptr otherwise)
};

#define ATTN_TK 160       // keys per staged tile (20 KB)
#define ATTN_MAXKEYS 1024

template <bool SELF>
__global__ void __launch_bounds__(128) attn_kernel(AttnArgs a) {
    __shared__ __align__(128) float buf[2][ATTN_TK * 32];
    __shared__ __align__(16) float4 q4s[8];
    __shared__ float scores[ATTN_MAXKEYS];
    __shared__ float ctx_red[4][32];
    __shared__ float red_s[4];
    __shared__ __align__(8) uint64_t bar[2];

    const int rank = blockIdx.x, h = blockIdx.y;
    if (rank >= a.st->n_alive) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t = a.st->step;
    const int row = a.alive[(t & 1) * a.B + rank];
    const int nkeys = SELF ? (t + 1) : a.nkeys_cross;
    const int ntiles = (nkeys + ATTN_TK - 1) / ATTN_TK;
    const int kvrow = SELF ? row : row / a.kv_div;
    const float* Kb = a.Kc + ((size_t)kvrow * 8 + h) * a.cap * 32;
    const float* Vb = a.Vc + ((size_t)kvrow * 8 + h) * a.cap * 32;

    // this head's final_linear slice, two output columns per thread: issue early
    float w0[32], w1[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        w0[d] = a.wo_t[(size_t)(h * 32 + d) * MNX_DEC_D + tid];
        w1[d] = a.wo_t[(size_t)(h * 32 + d) * MNX_DEC_D + tid + 128];
    }

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (tid < 8) q4s[tid] = reinterpret_cast<const float4*>(a.q + (size_t)rank * MNX_DEC_D + h * 32)[tid];
    __syncthreads();

    auto issue = [&](int i) {
        const int tile = (i < ntiles) ? i : i - ntiles;
        const float* src = ((i < ntiles) ? Kb : Vb) + (size_t)tile * ATTN_TK * 32;
        const uint32_t bytes = (uint32_t)min(ATTN_TK, nkeys - tile * ATTN_TK) * 128u;
        fence_proxy_async();
        mbar_arrive_expect_tx(&bar[i & 1], bytes);
        bulk_g2s(buf[i & 1], src, bytes, &bar[i & 1]);
    };
    if (tid == 0) issue(0);

    float acc = 0.f;
    for (int i = 0; i < 2 * ntiles; ++i) {
        if (tid == 0 && i + 1 < 2 * ntiles) issue(i + 1);
        mbar_wait(&bar[i & 1], (uint32_t)((i >> 1) & 1));
        const float* tb = buf[i & 1];
        if (i < ntiles) {
            // scores for this K tile: thread = key, float4 chunks visited in a lane-rotated
            // order so the 128-byte key rows are read without bank conflicts
            const int nk = min(ATTN_TK, nkeys - i * ATTN_TK);
            for (int j = tid; j < nk; j += 128) {
                const float4* kr = reinterpret_cast<const float4*>(tb + j * 32);
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int cc = (c + j) & 7;
                    const float4 kv = kr[cc];
                    const float4 qv = q4s[cc];
                    s = fmaf(qv.x, kv.x, s); s = fmaf(qv.y, kv.y, s);
                    s = fmaf(qv.z, kv.z, s); s = fmaf(qv.w, kv.w, s);
                }
                scores[i * ATTN_TK + j] = s;
            }
        } else {
            if (i == ntiles) {
                // softmax over all keys (fp32), p = exp(s - max) / sum
                __syncthreads();
                float m = -INFINITY;
                for (int j = tid; j < nkeys; j += 128) m = fmaxf(m, scores[j]);
                m = warp_max(m);
                if (lane == 0) red_s[wid] = m;
                __syncthreads();
                m = fmaxf(fmaxf(red_s[0], red_s[1]), fmaxf(red_s[2], red_s[3]));
                __syncthreads();
                float sum = 0.f;
                for (int j = tid; j < nkeys; j += 128) {
                    const float e = expf(scores[j] - m);
                    scores[j] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                if (lane == 0) red_s[wid] = sum;
                __syncthreads();
                sum = (red_s[0] + red_s[1]) + (red_s[2] + red_s[3]);
                for (int j = tid; j < nkeys; j += 128) scores[j] = scores[j] / sum;
                __syncthreads();
            }
            const int tile = i - ntiles;
            const int nk = min(ATTN_TK, nkeys - tile * ATTN_TK);
            const float* ps = scores + tile * ATTN_TK;
            for (int j = wid; j < nk; j += 4) acc = fmaf(ps[j], tb[j * 32 + lane], acc);
        }
        __syncthreads();   // tile consumed: its buffer may be refilled
    }
    ctx_red[wid][lane] = acc;
    __syncthreads();
    __shared__ float ctx[32];
    if (tid < 32) ctx[tid] = (ctx_red[0][tid] + ctx_red[1][tid]) + (ctx_red[2][tid] + ctx_red[3][tid]);
    __syncthreads();
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        const float c = ctx[d];
        o0 = fmaf(c, w0[d], o0);
        o1 = fmaf(c, w1[d], o1);
    }
    float* pr = a.part + ((size_t)rank * 8 + h) * MNX_DEC_D;
    pr[tid] = o0;
    pr[tid + 128] = o1;
}

// =====================================================================================
// beam-search self-attention: the keys of a hypothesis live in the slots of its ancestors
// (BeamBuffers::anc), so K/V rows are gathered instead of bulk-copied.  8 lanes x float4 cover
// one 128-byte key row, i.e. each warp-wide load instruction fetches four whole rows; beams of
// one image share most of their ancestry, so the gathered rows hit in L1/L2.
// =====================================================================================
#define ATTN_MAXSELF 512   // >= max_len
__global__ void __launch_bounds__(128) attn_self_beam_kernel(AttnArgs a) {
    __shared__ int slots[ATTN_MAXSELF];
    __shared__ float scores[ATTN_MAXSELF];
    __shared__ float ctx_red[4][32];
    __shared__ float ctx[32];
    __shared__ float red_s[4];
    const int rank = blockIdx.x, h = blockIdx.y;
    if (rank >= a.st->n_alive) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t = a.st->step;
    const int row = a.alive[(t & 1) * a.B + rank];
    const int nkeys = t + 1;

    float w0[32], w1[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        w0[d] = a.wo_t[(size_t)(h * 32 + d) * MNX_DEC_D + tid];
        w1[d] = a.wo_t[(size_t)(h * 32 + d) * MNX_DEC_D + tid + 128];
    }
    const int* anc_row = a.anc + ((size_t)(t & 1) * a.B + row) * a.cap;   // anc[t & 1][row][:]
    for (int j = tid; j < nkeys; j += 128) slots[j] = anc_row[j];
    const int sub = lane & 7, grp = lane >> 3;
    const float4 qv = reinterpret_cast<const float4*>(a.q + (size_t)rank * MNX_DEC_D + h * 32)[sub];
    __syncthreads();
    // scores
    for (int j0 = wid * 4; j0 < nkeys; j0 += 16) {
        const int j = j0 + grp;
        float sdot = 0.f;
        if (j < nkeys) {
            const float4 kv = reinterpret_cast<const float4*>(a.Kc + (((size_t)slots[j] * 8 + h) * a.cap + j) * 32)[sub];
            sdot = fmaf(qv.x, kv.x, sdot); sdot = fmaf(qv.y, kv.y, sdot);
            sdot = fmaf(qv.z, kv.z, sdot); sdot = fmaf(qv.w, kv.w, sdot);
        }
        sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
        sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
        sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
        if (sub == 0 && j < nkeys) scores[j] = sdot;
    }
    __syncthreads();
    float m = -INFINITY;
    for (int j = tid; j < nkeys; j += 128) m = fmaxf(m, scores[j]);
    m = warp_max(m);
    if (lane == 0) red_s[wid] = m;
    __syncthreads();
    m = fmaxf(fmaxf(red_s[0], red_s[1]), fmaxf(red_s[2], red_s[3]));
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < nkeys; j += 128) {
        const float e = expf(scores[j] - m);
        scores[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red_s[wid] = sum;
    __syncthreads();
    sum = (red_s[0] + red_s[1]) + (red_s[2] + red_s[3]);
    for (int j = tid; j < nkeys; j += 128) scores[j] = scores[j] / sum;
    __syncthreads();
    // P.V : lane = feature, the 4 warps split the keys
    float acc = 0.f;
    for (int j = wid; j < nkeys; j += 4)
        acc = fmaf(scores[j], a.Vc[(((size_t)slots[j] * 8 + h) * a.cap + j) * 32 + lane], acc);
    ctx_red[wid][lane] = acc;
    __syncthreads();
    if (tid < 32) ctx[tid] = (ctx_red[0][tid] + ctx_red[1][tid]) + (ctx_red[2][tid] + ctx_red[3][tid]);
    __syncthreads();
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        const float c = ctx[d];
        o0 = fmaf(c, w0[d], o0);
        o1 = fmaf(c, w1[d], o1);
    }
    float* pr = a.part + ((size_t)rank * 8 + h) * MNX_DEC_D;
    pr[tid] = o0;
    pr[tid + 128] = o1;
}

// =====================================================================================
// final LayerNorm -> vocab projection -> log_softmax -> grammar mask -> argmax -> bookkeeping
// (components.py:293-306, greedy_search.py:76-98, decode_strategy.py:51-57)
// =====================================================================================
#define VPAD 256

// x_in = x2 of the last layer; the W2 product arrives as 8 k-slice partials (+ bias b2)
// BEAM: stop after the masked log-probs and hand them (by rank) to beam_topk_kernel
template <bool BEAM>
__global__ void __launch_bounds__(1024) pick_kernel(DecBuffers b, DecWeights w, Grammar g, const float* x_in,
                                                    const float* part, const float* b2, float* lp_out) {
    __shared__ float hs[MNX_DEC_D];
    __shared__ float redf[8];
    __shared__ int redi[8];
    __shared__ float lsum[4][VPAD];
    const int rank = blockIdx.x;
    if (rank >= b.st->n_alive) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t = b.st->step;
    const int row = b.alive[(t & 1) * b.B + rank];

    // vocab weights of this thread's (d-quarter, vocab id): 64 independent loads in flight
    const int v_id = tid & 255, dq = tid >> 8;
    float wreg[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) wreg[i] = w.wout_t[(dq * 64 + i) * VPAD + v_id];

    float xv = 0.f;
    if (tid < MNX_DEC_D) {
        float p[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) p[h] = part[((size_t)rank * 8 + h) * MNX_DEC_D + tid];
        float sacc = 0.f;
#pragma unroll
        for (int h = 0; h < 8; ++h) sacc += p[h];
        xv = (sacc + b2[tid]) + x_in[(size_t)rank * MNX_DEC_D + tid];
    }
    // final LayerNorm (eps 1e-6) over 256 features (threads 0..255 = warps 0..7)
    float s = warp_sum(xv);
    if (lane == 0 && wid < 8) redf[wid] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += redf[i];
    const float mean = tot * (1.0f / MNX_DEC_D);
    __syncthreads();
    const float dv = (tid < MNX_DEC_D) ? xv - mean : 0.f;
    s = warp_sum(dv * dv);
    if (lane == 0 && wid < 8) redf[wid] = s;
    __syncthreads();
    tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += redf[i];
    const float rstd = 1.0f / sqrtf(tot * (1.0f / MNX_DEC_D) + 1e-6f);
    if (tid < MNX_DEC_D) {
        const float hv = dv * rstd * w.lnF_w[tid] + w.lnF_b[tid];
        hs[tid] = hv;
        b.hidden[((size_t)row * b.T + t) * MNX_DEC_D + tid] = hv;
    }
    __syncthreads();

    // logits: 4 d-quarters x 256 vocab slots, reduced through shared memory
    {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) acc = fmaf(hs[dq * 64 + i], wreg[i], acc);
        lsum[dq][v_id] = acc;
    }
    __syncthreads();
    float logit = -INFINITY;
    if (tid < g.vocab) logit = ((lsum[0][tid] + lsum[1][tid]) + (lsum[2][tid] + lsum[3][tid])) + w.bout[tid];
    if (tid >= 256) return;   // warps 8..31 are done (no further block-wide barriers below use them)
    // NOTE: the barriers below are named barriers over the first 256 threads only
#define BAR256() asm volatile("bar.sync 1, 256;" ::: "memory")
    // log_softmax
    float m = warp_max(logit);
    if (lane == 0) redf[wid] = m;
    BAR256();
    m = redf[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, redf[i]);
    BAR256();
    float e = (tid < g.vocab) ? expf(logit - m) : 0.f;
    e = warp_sum(e);
    if (lane == 0) redf[wid] = e;
    BAR256();
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) se += redf[i];
    float lp = (logit - m) - logf(se);
    // grammar mask keyed on the INPUT token of this step (tokenization.py:383-392)
    const int lab_len = BEAM ? 0 : b.st->lab_len;
    const int tok_in = input_token(b, g, lab_len, row, t);
    const bool in_x = tok_in >= g.offset && tok_in < g.offset + g.maxx;
    const bool in_y = tok_in >= g.offset + g.maxx;
    if (in_x && tid < g.offset + g.maxx) lp = -10000.0f;
    if (in_y && tid >= g.offset) lp = -10000.0f;
    if (t == 0 && tid == g.eos) lp = -1e20f;          // ensure_min_length (min_length = 1)
    if (tid >= g.vocab) lp = -INFINITY;
    if (BEAM) {
        lp_out[(size_t)rank * VPAD + tid] = lp;
        return;
    }
    // argmax, lowest index wins ties (torch.topk(1))
    float bv = lp;
    int bi = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    BAR256();
    if (lane == 0) { redf[wid] = bv; redi[wid] = bi; }
    BAR256();
#undef BAR256
    if (tid == 0) {
        float v = redf[0];
        int ix = redi[0];
#pragma unroll
        for (int i = 1; i < 8; ++i)
            if (redf[i] > v || (redf[i] == v && redi[i] < ix)) { v = redf[i]; ix = redi[i]; }
        b.ids[(size_t)row * b.T + t] = ix;
        b.logp[(size_t)row * b.T + t] = v;
        b.cur_tok[row] = ix;
        // with labels the NEXT given token decides (greedy_search.py:83-85: is_finished = label.eq(eos)); the stored id
        // stays the model's pick (alive_seq, :86) until label_merge_kernel
        const int ends = (t + 1 < lab_len) ? b.labels[(size_t)row * (b.T + 1) + t + 1] : ix;
        const int fin = (ends == g.eos) || (t == g.max_len - 1);
        b.finished[row] = fin;
        if (fin) b.lens[row] = t + 1;
    }
}

// =====================================================================================
// BeamSearch.advance + update_finished for one alive image per CTA (beam_search.py:84-190 as
// repaired in oracle/restate.py beam_decode): cumulative + token log-prob, length-normalised
// score, top-`beam` over beam*V with ties to the lowest flat index, back-pointers applied to
// the ancestry / id / log-prob histories, finished hypotheses filed, end condition.
// =====================================================================================
__global__ void __launch_bounds__(256) beam_topk_kernel(DecBuffers b, Grammar g, BeamBuffers bm) {
    __shared__ float s_val[8];
    __shared__ int s_idx[8];
    __shared__ int sel_flat[MNX_MAX_BEAM];
    __shared__ float sel_score[MNX_MAX_BEAM];
    __shared__ int sel_store[MNX_MAX_BEAM];
    const int r = blockIdx.x;
    if (r >= b.st->n_img) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t = b.st->step, K = bm.beam, V = g.vocab, T = b.T, R = b.B;
    const int cur = t & 1, nxt = cur ^ 1;
    const int img = bm.alive_img[cur * bm.n_img0 + r];
    const float len = (float)(t + 2);             // curr_length = len(self) + 1, beam_search.py:99
    const int BIG = 0x7fffffff;

    float sc[MNX_MAX_BEAM];
#pragma unroll
    for (int k = 0; k < MNX_MAX_BEAM; ++k) {
        sc[k] = -INFINITY;
        if (k < K && tid < V) {
            const float total = __fadd_rn(bm.lp[(size_t)(r * K + k) * VPAD + tid], bm.cum[cur * R + img * K + k]);
            sc[k] = __fdiv_rn(total, len);
        }
    }
    unsigned taken = 0;
    for (int j = 0; j < K; ++j) {
        float bv = -INFINITY;
        int bi = BIG;
        if (tid < V) {
#pragma unroll
            for (int k = 0; k < MNX_MAX_BEAM; ++k) {
                if (k < K && !((taken >> k) & 1u)) {
                    const int fi = k * V + tid;
                    if (sc[k] > bv || (sc[k] == bv && fi < bi)) { bv = sc[k]; bi = fi; }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = bv; s_idx[wid] = bi; }
        __syncthreads();
        bv = s_val[0]; bi = s_idx[0];
#pragma unroll
        for (int i = 1; i < 8; ++i)
            if (s_val[i] > bv || (s_val[i] == bv && s_idx[i] < bi)) { bv = s_val[i]; bi = s_idx[i]; }
        if (tid == 0) {
            sel_flat[j] = bi; sel_score[j] = bv;
            bm.trace[((size_t)t * bm.n_img0 + img) * MNX_MAX_BEAM + j] = bi;
        }
        if (bi % V == tid) taken |= 1u << (bi / V);
        __syncthreads();
    }

    if (tid == 0) {
        int count = bm.hyp_count[img];
        int topf = bm.top_fin[img];
        const int NB = bm.n_best;
        int* order = bm.hyp_order + img * MNX_MAX_BEAM;
        float* hscore = bm.hyp_score + img * MNX_MAX_BEAM;
        for (int j = 0; j < K; ++j) {
            const int wtok = sel_flat[j] % V;
            const float score = sel_score[j];
            const int fin = (wtok == g.eos) || (t == g.max_len - 1);
            int store = -1;
            if (fin) {
                const int kept = count < NB ? count : NB;
                int pos = 0;
                while (pos < kept && hscore[order[pos]] >= score) ++pos;   // stable: first stored wins ties
                if (pos < NB) {
                    store = (kept < NB) ? kept : order[NB - 1];
                    for (int q = (kept < NB ? kept : NB - 1); q > pos; --q) order[q] = order[q - 1];
                    order[pos] = store;
                    hscore[store] = score;
                    bm.hyp_len[img * MNX_MAX_BEAM + store] = t + 1;
                }
                ++count;
                if (j == 0) topf = 1;
            }
            sel_store[j] = store;
            const int slot = img * K + j;
            bm.cum[nxt * R + slot] = fin ? -1e10f : __fmul_rn(score, len);   // beam_search.py:105,136
            b.cur_tok[slot] = wtok;
        }
        bm.hyp_count[img] = count;
        bm.top_fin[img] = topf;
        if (topf && count >= NB) bm.img_done[img] = 1;
    }
    __syncthreads();

    // histories follow their back-pointers (alive_seq.index_select(0, select_indices), :117-119)
    for (int j = 0; j < K; ++j) {
        const int pk = sel_flat[j] / V, wtok = sel_flat[j] % V;
        const size_t src = ((size_t)cur * R + img * K + pk) * T;
        const size_t dst = ((size_t)nxt * R + img * K + j) * T;
        const int store = sel_store[j];
        const size_t hdst = ((size_t)img * MNX_MAX_BEAM + (store < 0 ? 0 : store)) * T;
        const float tok_lp = bm.lp[(size_t)(r * K + pk) * VPAD + wtok];
        for (int i = tid; i <= t; i += 256) {
            const int id = (i < t) ? bm.hist_ids[src + i] : wtok;
            const float l = (i < t) ? bm.hist_logp[src + i] : tok_lp;
            const int an = bm.anc[src + i];
            bm.hist_ids[dst + i] = id;
            bm.hist_logp[dst + i] = l;
            bm.anc[dst + i] = an;
            if (store >= 0) {
                bm.hyp_ids[hdst + i] = id;
                bm.hyp_logp[hdst + i] = l;
                bm.hyp_anc[hdst + i] = an;
            }
        }
        if (tid == 0 && t + 1 < T) bm.anc[dst + t + 1] = img * K + j;   // next step's own K/V position
    }
}

// cum = [0, -inf, ...] per image (beam_search.py:45-47); every slot is its own ancestor at position 0
__global__ void beam_init_kernel(DecBuffers b, BeamBuffers bm) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= b.B) return;
    bm.cum[slot] = (slot % bm.beam == 0) ? 0.f : -INFINITY;
    bm.anc[(size_t)slot * b.T] = slot;
    if (slot < bm.n_img0) {
        bm.img_done[slot] = 0; bm.top_fin[slot] = 0; bm.hyp_count[slot] = 0;
    }
}

// results: n_best hypotheses per image, best first; the best one is also laid out like a greedy
// result (ids/lens/logp per image + its hidden states gathered along the ancestry) so that the
// atom scan and the bond head run on it unchanged (Decoder.decode uses pred[0], components.py:455).
struct BeamOut {
    int* ids;        // [n_img][n_best][T]
    int* lens;       // [n_img][n_best]
    float* scores;   // [n_img][n_best]
    float* logp;     // [n_img][n_best][T]
    int* best_ids;   // [n_img][T]
    int* best_lens;  // [n_img]
    float* best_logp;    // [n_img][T]
    float* best_hidden;  // [n_img][T][256]
};
__global__ void __launch_bounds__(256) beam_finalize_kernel(DecBuffers b, BeamBuffers bm, BeamOut o) {
    const int img = blockIdx.x, tid = threadIdx.x, T = b.T, NB = bm.n_best;
    const int have = min(bm.hyp_count[img], NB);
    for (int n = 0; n < NB; ++n) {
        const int store = (n < have) ? bm.hyp_order[img * MNX_MAX_BEAM + n] : -1;
        const int L = (store >= 0) ? bm.hyp_len[img * MNX_MAX_BEAM + store] : 0;
        const size_t src = ((size_t)img * MNX_MAX_BEAM + (store < 0 ? 0 : store)) * T;
        const size_t dst = ((size_t)img * NB + n) * T;
        for (int i = tid; i < T; i += 256) {
            const int id = (i < L) ? bm.hyp_ids[src + i] : 0;
            const float l = (i < L) ? bm.hyp_logp[src + i] : 0.f;
            if (o.ids) o.ids[dst + i] = id;
            if (o.logp) o.logp[dst + i] = l;
            if (n == 0) { o.best_ids[(size_t)img * T + i] = id; o.best_logp[(size_t)img * T + i] = l; }
        }
        if (tid == 0) {
            if (o.lens) o.lens[img * NB + n] = L;
            if (o.scores) o.scores[img * NB + n] = (store >= 0) ? bm.hyp_score[img * MNX_MAX_BEAM + store] : -INFINITY;
            if (n == 0) o.best_lens[img] = L;
        }
        if (n == 0) {
            for (int i = 0; i < L; ++i) {
                const int slot = bm.hyp_anc[src + i];
                o.best_hidden[((size_t)img * T + i) * MNX_DEC_D + tid] = b.hidden[((size_t)slot * T + i) * MNX_DEC_D + tid];
            }
        }
    }
}

// =====================================================================================
// fp32 tiled GEMM for the once-per-call projections (memory bank, cross K/V of all layers,
// bond-head first layer):  C[M][N] = A[M][K] * Bt[K][N] + bias, 64x64 tile, 4x4 per thread
// =====================================================================================
enum { ST_PLAIN = 0, ST_CROSSKV = 1 };
struct GemmStore {
    float* out;       // ST_PLAIN: [M][N]
    float* k_out;     // ST_CROSSKV: [L][B][8][S][32]
    float* v_out;
    int B, S;
};

template <int STORE>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                       const float* __restrict__ bias, int M, int N, int K,
                                                       GemmStore st) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int kk = 0; kk < K; kk += 16) {
        // A tile 64x16 (row-major, K contiguous): 256 threads x float4
        {
            const int r = tid >> 2, c4 = tid & 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M) v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * K + kk + c4 * 4);
            As[c4 * 4 + 0][r] = v.x; As[c4 * 4 + 1][r] = v.y; As[c4 * 4 + 2][r] = v.z; As[c4 * 4 + 3][r] = v.w;
        }
        {
            const int r = tid >> 4, c4 = tid & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + c4 * 4 < N) v = *reinterpret_cast<const float4*>(Bt + (size_t)(kk + r) * N + n0 + c4 * 4);
            *reinterpret_cast<float4*>(&Bs[r][c4 * 4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w};
            const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            const float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (STORE == ST_PLAIN) {
                st.out[(size_t)m * N + n] = v;
            } else {
                const int l = n >> 9, kv = (n >> 8) & 1, h = (n >> 5) & 7, d = n & 31;
                const int bb = m / st.S, s = m - bb * st.S;
                float* dst = kv ? st.v_out : st.k_out;
                dst[((((size_t)l * st.B + bb) * 8 + h) * st.S + s) * 32 + d] = v;
            }
        }
    }
}

// =====================================================================================
// atom scan (CharTokenizer.sequence_to_smiles, tokenization.py:464-515: `indices` only)
// =====================================================================================
// walks the atom tokens of one sequence: on_atom(k, i, j) for atom k whose symbol tokens are seq[i..j) followed by its X and Y
// tokens (so `indices[k]` = j + 2, tokenization.py:464-515); returns the number of atoms
template <class F>
__device__ __forceinline__ int scan_atoms(const int* __restrict__ seq, int n, const uint8_t* __restrict__ cls, const Grammar& g, F on_atom) {
    int i = 0, k = 0;
    auto is_x = [&](int v) { return v >= g.offset && v < g.offset + g.maxx; };
    auto is_y = [&](int v) { return v >= g.offset + g.maxx; };
    while (i < n) {
        const int tok = seq[i];
        if (tok == g.eos || tok == 0) break;
        if (is_x(tok) || is_y(tok)) { ++i; continue; }
        const uint8_t c = cls[tok];
        if (!(c & 2)) { ++i; continue; }          // not an atom token: plain SMILES character
        int j;
        if (c & 4) {                               // '[' ... ']'
            j = i + 1;
            while (j < n) {
                const int v = seq[j];
                const uint8_t cj = (v < g.offset) ? cls[v] : 0;
                if (!(cj & 1)) break;
                ++j;
                if (cj & 8) break;
            }
        } else {
            j = i + 1;
            if (j < n && seq[j] < g.offset) {
                const uint8_t cn = cls[seq[j]];
                if ((cn & 1) && (((c & 16) && (cn & 32)) || ((c & 64) && (cn & 128)))) j = i + 2;
            }
        }
        if (j + 2 < n && is_x(seq[j]) && is_y(seq[j + 1])) {
            on_atom(k, i, j);
            ++k;
            i = j + 2;
        } else {
            i = j;
        }
    }
    return k;
}

__global__ void atom_scan_kernel(const int* __restrict__ ids, const int* __restrict__ lens, int B, int T,
                                 const uint8_t* __restrict__ cls, Grammar g, int max_atoms,
                                 int* __restrict__ atom_idx, int* __restrict__ n_atoms) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= B) return;
    n_atoms[row] = scan_atoms(ids + (size_t)row * T, lens[row], cls, g, [&](int k, int, int j) {
        if (k < max_atoms) atom_idx[(size_t)row * max_atoms + k] = j + 2;
    });
}

// =====================================================================================
// confidences (Decoder.decode with compute_confidence, components.py:456-469,485-491): per atom the geometric mean of
// the probabilities of its symbol tokens, per image exp(mean(token log-prob)) and
// overall = average_token_score * sqrt(prod(edge_scores[:k, :k])).  One CTA per image; the products run in fp64 like
// numpy's (np.prod of Python floats), including its underflow to 0 for large molecules.
// =====================================================================================
__global__ void __launch_bounds__(128) confidence_kernel(const int* __restrict__ ids, const int* __restrict__ lens,
                                                         const float* __restrict__ logp, int T, const uint8_t* __restrict__ cls,
                                                         Grammar g, int max_atoms, const float* __restrict__ edge_score,
                                                         float* __restrict__ atom_scores, float* __restrict__ seq_score,
                                                         double* __restrict__ overall) {
    __shared__ float redf[4];
    __shared__ int s_k;
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = lens[row];
    const float* lp = logp + (size_t)row * T;
    if (tid == 0) {
        s_k = scan_atoms(ids + (size_t)row * T, n, cls, g, [&](int k, int i, int j) {
            if (k >= max_atoms) return;
            double prod = 1.0;
            for (int p = i; p < j; ++p) prod *= (double)expf(lp[p]);          // token_scores = exp(log-prob) in fp32, then Python floats
            atom_scores[(size_t)row * max_atoms + k] = (float)pow(prod, 1.0 / (double)(j - i));
        });
    }
    float s = 0.f;
    for (int p = tid; p < n; p += 128) s += lp[p];
    s = warp_sum(s);
    if (lane == 0) redf[wid] = s;
    __syncthreads();
    const float avg = expf(((redf[0] + redf[1]) + (redf[2] + redf[3])) / (float)max(n, 1));
    const int k = min(s_k, max_atoms);
    // np.prod multiplies the k * k fp64 factors one after the other, and that ORDER is part of the result: once the running
    // product reaches the denormal range it either sticks at the smallest denormal (every later factor > 0.5 rounds back up
    // to it) or drops to 0 -- the reference's fixtures contain both.  So: stage the scores in shared memory, one thread
    // multiplies them in row-major order (fp64 on the GPU keeps denormals and rounds to nearest even like the host).
    extern __shared__ float es[];
    for (int idx = tid; idx < k * k; idx += 128) es[idx] = edge_score[((size_t)row * max_atoms + idx / k) * max_atoms + idx % k];
    __syncthreads();
    if (tid == 0) {
        double prod = 1.0;
        for (int idx = 0; idx < k * k; ++idx) prod *= (double)es[idx];
        seq_score[row] = avg;
        overall[row] = (double)avg * sqrt(prod);
    }
}

// =====================================================================================
// bond head (GraphPredictor.forward + softmax + get_edge_prediction, components.py:365-400)
// =====================================================================================
// gather the hidden rows of the atoms of every image into a dense [B*max_atoms][256] matrix
__global__ void edge_gather_kernel(const float* __restrict__ hidden, const int* __restrict__ atom_idx,
                                   const int* __restrict__ n_atoms, int T, int max_atoms,
                                   float* __restrict__ hg) {
    const int img = blockIdx.y, a = blockIdx.x;
    float v = 0.f;
    if (a < min(n_atoms[img], max_atoms)) {
        const int pos = atom_idx[(size_t)img * max_atoms + a];
        v = hidden[((size_t)img * T + pos) * MNX_DEC_D + threadIdx.x];
    }
    hg[((size_t)img * max_atoms + a) * MNX_DEC_D + threadIdx.x] = v;
}

// one warp per ordered pair (i, j): z = gelu(A_i + B_j + b0); logits = W2 z + b2; softmax over 7
__global__ void __launch_bounds__(256) edge_pair_kernel(const float* __restrict__ AB, const int* __restrict__ n_atoms,
                                                       int max_atoms, DecWeights w, float* __restrict__ prob) {
    const int img = blockIdx.y;
    const int k = min(n_atoms[img], max_atoms);
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pair >= k * k) return;
    const int i = pair / k, j = pair - i * k;
    const float* Ai = AB + ((size_t)img * max_atoms + i) * 512;          // [.., 0:256] = W[:, :256] h
    const float* Bj = AB + ((size_t)img * max_atoms + j) * 512 + 256;    // [.., 256:512] = W[:, 256:] h
    float lg[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int d = u * 32 + lane;
        const float z = gelu_erf(Ai[d] + Bj[d] + w.be0[d]);
#pragma unroll
        for (int c = 0; c < 7; ++c) lg[c] = fmaf(z, w.we2[c * MNX_DEC_D + d], lg[c]);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) lg[c] = warp_sum(lg[c]) + w.be2[c];
    float m = lg[0];
#pragma unroll
    for (int c = 1; c < 7; ++c) m = fmaxf(m, lg[c]);
    float e[7], se = 0.f;
#pragma unroll
    for (int c = 0; c < 7; ++c) { e[c] = expf(lg[c] - m); se += e[c]; }
    if (lane < 7) {
        float v = e[0];
#pragma unroll
        for (int c = 1; c < 7; ++c) if (lane == c) v = e[c];
        prob[(((size_t)img * max_atoms + i) * max_atoms + j) * 8 + lane] = v / se;
    }
}

// symmetrisation in double precision, exactly the in-place update order of the reference
__global__ void edge_sym_kernel(const float* __restrict__ prob, const int* __restrict__ n_atoms, int max_atoms,
                                uint8_t* __restrict__ edges, float* __restrict__ score) {
    const int img = blockIdx.y;
    const int k = min(n_atoms[img], max_atoms);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    const int i = idx / k, j = idx - i * k;
    const float* pij = prob + (((size_t)img * max_atoms + i) * max_atoms + j) * 8;
    const float* pji = prob + (((size_t)img * max_atoms + j) * max_atoms + i) * 8;
    double p[7];
    if (i == j) {
#pragma unroll
        for (int c = 0; c < 7; ++c) p[c] = (double)pij[c];
    } else {
#pragma unroll
        for (int c = 0; c < 5; ++c) p[c] = ((double)pij[c] + (double)pji[c]) / 2;
        if (i < j) {
            p[5] = ((double)pij[5] + (double)pji[6]) / 2;
            p[6] = ((double)pij[6] + (double)pji[5]) / 2;
        } else {   // (i,j) is the mirrored entry of pair (j,i): gets that pair's 6 and 5
            p[5] = ((double)pji[6] + (double)pij[5]) / 2;
            p[6] = ((double)pji[5] + (double)pij[6]) / 2;
        }
    }
    int best = 0;
    double bv = p[0];
#pragma unroll
    for (int c = 1; c < 7; ++c) if (p[c] > bv) { bv = p[c]; best = c; }
    edges[((size_t)img * max_atoms + i) * max_atoms + j] = (uint8_t)best;
    if (score) score[((size_t)img * max_atoms + i) * max_atoms + j] = (float)bv;
}

// =====================================================================================
// host-side launchers
// =====================================================================================
#define TG_MIN_ROWS 128     // row capacity from which the register-tiled kernel replaces the skinny one
static int g_tile_min_rows = TG_MIN_ROWS;   // MNX_TILE_GEMM_MIN_ROWS overrides it (tests run the small fixtures through the tiled kernel)
template <int K>
constexpr size_t tile_gemm_smem() { return (size_t)(TG_ROWS * (K + 4) + 2 * TG_KC * TG_COLS) * sizeof(float); }

template <int K, int PRO, int EPI>
static cudaError_t launch_skinny(const SkinnyArgs& a, int B, cudaStream_t s) {
    if (B >= g_tile_min_rows) {
        dim3 grid(a.N / TG_COLS, (B + TG_ROWS - 1) / TG_ROWS, EPI == EPI_PART ? 8 : 1);
        tile_gemm_kernel<K, PRO, EPI><<<grid, 256, tile_gemm_smem<K>(), s>>>(a);
        return cudaGetLastError();
    }
    const size_t smem = (size_t)(32 * K + 8 * 32 * 32) * sizeof(float);
    dim3 grid(a.N / 32, (B + 31) / 32, EPI == EPI_PART ? 8 : 1);
    skinny_gemm_kernel<K, PRO, EPI><<<grid, 256, smem, s>>>(a);
    return cudaGetLastError();
}

template <int K, int PRO, int EPI>
static cudaError_t configure_skinny() {
    const size_t smem = (size_t)(32 * K + 8 * 32 * 32) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(tile_gemm_kernel<K, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_gemm_smem<K>());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(skinny_gemm_kernel<K, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// opt in to >48 KB dynamic shared memory on the current device (call once per device)
// =====================================================================================
// partial-label decoding: arm the labels, and the final merge of components.py:326-332
//   label = orig_labels[i][1:len(pred)+1]; pred = pred[:len(label)]; pred = pred*mask + label*(1-mask)
// =====================================================================================
__global__ void set_label_len_kernel(DecState* st, int lab_len) { st->lab_len = lab_len; }

__global__ void __launch_bounds__(256) label_merge_kernel(DecBuffers b, int lab_len) {
    const int row = blockIdx.x;
    const int n = b.lens[row];
    const int keep = min(n, lab_len - 1);
    for (int j = threadIdx.x; j < b.T; j += 256) {
        int v = b.ids[(size_t)row * b.T + j];
        if (j < keep) {
            const int lab = b.labels[(size_t)row * (b.T + 1) + 1 + j];
            if (lab != MNX_MASK_ID) v = lab;
        } else {
            v = 0;
        }
        b.ids[(size_t)row * b.T + j] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) b.lens[row] = keep;
}

cudaError_t dec_set_label_len(const DecBuffers& b, int lab_len, cudaStream_t s) {
    set_label_len_kernel<<<1, 1, 0, s>>>(b.st, lab_len);
    return cudaGetLastError();
}
cudaError_t dec_label_merge(const DecBuffers& b, int lab_len, cudaStream_t s) {
    label_merge_kernel<<<b.B, 256, 0, s>>>(b, lab_len);
    return cudaGetLastError();
}

cudaError_t dec_configure() {
    cudaError_t e;
    const char* env = getenv("MNX_TILE_GEMM_MIN_ROWS");
    g_tile_min_rows = (env && atoi(env) > 0) ? atoi(env) : TG_MIN_ROWS;
    if ((e = configure_skinny<256, PRO_LN, EPI_QKV>()) != cudaSuccess) return e;
    if ((e = configure_skinny<256, PRO_SUM_LN, EPI_QKV>()) != cudaSuccess) return e;
    if ((e = configure_skinny<256, PRO_SUM_LN, EPI_Q>()) != cudaSuccess) return e;
    if ((e = configure_skinny<256, PRO_SUM_LN, EPI_GELU>()) != cudaSuccess) return e;
    if ((e = configure_skinny<128, PRO_PLAIN, EPI_PART>()) != cudaSuccess) return e;
    return cudaSuccess;
}

// one decode step = 38 launches; returns the number of launches issued
// bm != nullptr: beam search (rows = image-major slots; 39 launches)
int dec_launch_step(const DecBuffers& b, const DecWeights& w, const Grammar& g, const BeamBuffers* bm, cudaStream_t s,
                    cudaError_t* err) {
    int n = 0;
    cudaError_t e = cudaSuccess;
#define CK(x) do { e = (x); if (e != cudaSuccess) { *err = e; return n; } } while (0)
    if (bm) embed_compact_beam_kernel<<<1, 1024, 0, s>>>(b, w, g, *bm);
    else embed_compact_kernel<<<1, 1024, 0, s>>>(b, w, g);
    CK(cudaGetLastError()); ++n;
    const size_t kv_layer = (size_t)b.B * 8 * b.T * 32;
    const size_t ckv_layer = (size_t)(bm ? bm->n_img0 : b.B) * 8 * b.S * 32;
    // Residual-stream bookkeeping: every linear layer that ends a sub-block (final_linear of both
    // attentions, W2 of the FFN) leaves 8 partial sums in b.part; the NEXT kernel's prologue adds
    // them, the bias and the residual, and column-block 0 of that kernel stores the new stream.
    float* x_cur = b.xa;     // stream entering the layer (complete for l == 0, else x2 of layer l-1)
    float* x_alt = b.xb;
    for (int l = 0; l < MNX_DEC_L; ++l) {
        const DecLayerW& L = w.layer[l];
        SkinnyArgs a{};
        a.st = b.st; a.alive = b.alive; a.B = b.B;
        // (A) [x3 = x2 + b2 + sum(W2 partials)] -> LN1 -> QKV, append K/V
        a.x_in = x_cur; a.ln_w = L.ln1_w; a.ln_b = L.ln1_b; a.wt = L.wqkv_t; a.bias = L.bqkv; a.N = 768;
        a.out = b.q; a.kc = b.selfK + l * kv_layer; a.vc = b.selfV + l * kv_layer; a.T = b.T;
        if (l == 0) {
            CK((launch_skinny<256, PRO_LN, EPI_QKV>(a, b.B, s))); ++n;
        } else {
            a.part = b.part; a.bo = w.layer[l - 1].b2; a.x_sum_out = x_alt;
            CK((launch_skinny<256, PRO_SUM_LN, EPI_QKV>(a, b.B, s))); ++n;
            float* tmp = x_cur; x_cur = x_alt; x_alt = tmp;      // x_cur now holds this layer's input
        }
        // (B) self attention + per-head final_linear partials
        AttnArgs at{};
        at.st = b.st; at.alive = b.alive; at.B = b.B; at.q = b.q;
        at.Kc = b.selfK + l * kv_layer; at.Vc = b.selfV + l * kv_layer; at.cap = b.T; at.nkeys_cross = 0;
        at.wo_t = L.wo_s_t; at.part = b.part2; at.kv_div = 1;
        if (bm) {
            // ancestry rows of step t live in anc[t & 1]; both parities are passed and the kernel picks
            at.anc = bm->anc;
            attn_self_beam_kernel<<<dim3(b.B, 8), 128, 0, s>>>(at);
        } else {
            attn_kernel<true><<<dim3(b.B, 8), 128, 0, s>>>(at);
        }
        CK(cudaGetLastError()); ++n;
        // (D) x1 = x + bo + sum(part); LN2 -> context query
        a = SkinnyArgs{};
        a.st = b.st; a.alive = b.alive; a.B = b.B;
        a.x_in = x_cur; a.part = b.part2; a.bo = L.bo_s; a.x_sum_out = x_alt;
        a.ln_w = L.ln2_w; a.ln_b = L.ln2_b; a.wt = L.wq_c_t; a.bias = L.bq_c; a.N = 256; a.out = b.q;
        CK((launch_skinny<256, PRO_SUM_LN, EPI_Q>(a, b.B, s))); ++n;
        // (E) cross attention over the S memory positions + partials
        at.Kc = b.crossK + l * ckv_layer; at.Vc = b.crossV + l * ckv_layer; at.cap = b.S; at.nkeys_cross = b.S;
        at.wo_t = L.wo_c_t; at.part = b.part; at.kv_div = bm ? bm->beam : 1; at.anc = nullptr;
        attn_kernel<false><<<dim3(b.B, 8), 128, 0, s>>>(at);
        CK(cudaGetLastError()); ++n;
        // (G) x2 = x1 + bo + sum(part); LN_ff -> W1 -> GELU
        a = SkinnyArgs{};
        a.st = b.st; a.alive = b.alive; a.B = b.B;
        a.x_in = x_alt; a.part = b.part; a.bo = L.bo_c; a.x_sum_out = x_cur;
        a.ln_w = L.lnf_w; a.ln_b = L.lnf_b; a.wt = L.w1_t; a.bias = L.b1; a.N = 1024; a.out = b.hbuf;
        CK((launch_skinny<256, PRO_SUM_LN, EPI_GELU>(a, b.B, s))); ++n;
        // (H) 8 k-slice partials of W2 h (bias + residual x2 are added by the next consumer)
        a = SkinnyArgs{};
        a.st = b.st; a.alive = b.alive; a.B = b.B;
        a.x_in = b.hbuf; a.x_stride = MNX_DEC_FF; a.wt = L.w2_t; a.N = 256; a.out = b.part;
        CK((launch_skinny<128, PRO_PLAIN, EPI_PART>(a, b.B, s))); ++n;
        // x_cur holds x2; the W2 partials are in b.part
    }
    if (bm) {
        pick_kernel<true><<<b.B, 1024, 0, s>>>(b, w, g, x_cur, b.part, w.layer[MNX_DEC_L - 1].b2, bm->lp);
        CK(cudaGetLastError()); ++n;
        beam_topk_kernel<<<bm->n_img0, 256, 0, s>>>(b, g, *bm);
        CK(cudaGetLastError()); ++n;
    } else {
        pick_kernel<false><<<b.B, 1024, 0, s>>>(b, w, g, x_cur, b.part, w.layer[MNX_DEC_L - 1].b2, nullptr);
        CK(cudaGetLastError()); ++n;
    }
#undef CK
    *err = cudaSuccess;
    return n;
}

cudaError_t dec_precompute(const DecBuffers& b, const DecWeights& w, const float* features, int enc_dim,
                           cudaStream_t s, int* launches) {
    const int M = b.B * b.S;
    GemmStore st{};
    st.out = b.membank;
    gemm_f32_kernel<ST_PLAIN><<<dim3(MNX_DEC_D / 64, (M + 63) / 64), 256, 0, s>>>(features, w.wenc_t, w.benc, M,
                                                                                MNX_DEC_D, enc_dim, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    st = GemmStore{};
    st.k_out = b.crossK; st.v_out = b.crossV; st.B = b.B; st.S = b.S;
    const int N = MNX_DEC_L * 512;
    gemm_f32_kernel<ST_CROSSKV><<<dim3(N / 64, (M + 63) / 64), 256, 0, s>>>(b.membank, w.wkv_c_t, w.bkv_c, M, N,
                                                                          MNX_DEC_D, st);
    *launches += 2;
    return cudaGetLastError();
}

cudaError_t dec_atom_scan(const int* ids, const int* lens, int B, int T, const uint8_t* cls, const Grammar& g,
                          int max_atoms, int* atom_idx, int* n_atoms, cudaStream_t s) {
    atom_scan_kernel<<<(B + 63) / 64, 64, 0, s>>>(ids, lens, B, T, cls, g, max_atoms, atom_idx, n_atoms);
    return cudaGetLastError();
}

cudaError_t dec_confidence(const int* ids, const int* lens, const float* logp, int B, int T, const uint8_t* cls, const Grammar& g,
                           int max_atoms, const float* edge_score, float* atom_scores, float* seq_score, double* overall, cudaStream_t s) {
    const size_t smem = (size_t)max_atoms * max_atoms * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(confidence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    confidence_kernel<<<B, 128, smem, s>>>(ids, lens, logp, T, cls, g, max_atoms, edge_score, atom_scores, seq_score, overall);
    return cudaGetLastError();
}

cudaError_t dec_edges(const float* hidden, const int* atom_idx, const int* n_atoms, int B, int T, int max_atoms,
                      const DecWeights& w, float* hg, float* AB, float* prob, uint8_t* edges, float* score,
                      cudaStream_t s, int* launches) {
    edge_gather_kernel<<<dim3(max_atoms, B), MNX_DEC_D, 0, s>>>(hidden, atom_idx, n_atoms, T, max_atoms, hg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int M = B * max_atoms;
    GemmStore st{};
    st.out = AB;
    // AB[m][0:256] = W0[:, :256] h_m ; AB[m][256:512] = W0[:, 256:] h_m   (we_a_t is stored as one [256][512])
    gemm_f32_kernel<ST_PLAIN><<<dim3(512 / 64, (M + 63) / 64), 256, 0, s>>>(hg, w.we_a_t, nullptr, M, 512, MNX_DEC_D, st);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int pairs = max_atoms * max_atoms;
    edge_pair_kernel<<<dim3((pairs + 7) / 8, B), 256, 0, s>>>(AB, n_atoms, max_atoms, w, prob);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    edge_sym_kernel<<<dim3((pairs + 255) / 256, B), 256, 0, s>>>(prob, n_atoms, max_atoms, edges, score);
    *launches += 4;
    return cudaGetLastError();
}

cudaError_t dec_beam_init(const DecBuffers& b, const BeamBuffers& bm, cudaStream_t s) {
    beam_init_kernel<<<(b.B + 255) / 256, 256, 0, s>>>(b, bm);
    return cudaGetLastError();
}

cudaError_t dec_beam_finalize(const DecBuffers& b, const BeamBuffers& bm, int* ids, int* lens, float* scores, float* logp,
                              int* best_ids, int* best_lens, float* best_logp, float* best_hidden, cudaStream_t s) {
    BeamOut o{ids, lens, scores, logp, best_ids, best_lens, best_logp, best_hidden};
    beam_finalize_kernel<<<bm.n_img0, 256, 0, s>>>(b, bm, o);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------
// isolated timing of one decode kernel on the shapes of the last call (roofline evidence):
// which = 1 cross-attention, 2 self-attention at t=step, 3 ln1+qkv, 4 sum+lnff+W1, 5 W2+res, 6 pick
// -------------------------------------------------------------------------------------
__global__ void timing_setup_kernel(DecBuffers b, int step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.B) { b.alive[i] = i; b.alive[b.B + i] = i; }
    if (i == 0) { b.st->n_alive = b.B; b.st->step = step; }
}

cudaError_t dec_time_kernel(int which, int iters, const DecBuffers& b, const DecWeights& w, const Grammar& g,
                            int step, float* ms, cudaStream_t s) {
    if (which < 1 || which > 6) return cudaErrorInvalidValue;
    DecState saved{};
    cudaError_t e = cudaMemcpyAsync(&saved, b.st, sizeof(DecState), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    timing_setup_kernel<<<(b.B + 255) / 256, 256, 0, s>>>(b, step);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const DecLayerW& L = w.layer[0];
    const size_t kv_layer = (size_t)b.B * 8 * b.T * 32;
    (void)kv_layer;
    int call = 0;
    auto once = [&]() -> cudaError_t {
        SkinnyArgs a{};
        a.st = b.st; a.alive = b.alive; a.B = b.B;
        AttnArgs at{};
        at.st = b.st; at.alive = b.alive; at.B = b.B; at.q = b.q; at.part = b.part; at.kv_div = 1;
        switch (which) {
            case 1: {
                // walk the six layers' memory-bank K/V like a decode step does (56 MB at bs = 32: L2-resident, as in the real
                // loop; 453 MB at bs = 256: every launch streams from HBM)
                const size_t cross_layer = (size_t)b.B * 8 * b.S * 32;
                const int l = call++ % MNX_DEC_L;
                at.Kc = b.crossK + l * cross_layer; at.Vc = b.crossV + l * cross_layer; at.cap = b.S; at.nkeys_cross = b.S; at.wo_t = L.wo_c_t;
                attn_kernel<false><<<dim3(b.B, 8), 128, 0, s>>>(at);
                return cudaGetLastError();
            }
            case 2:
                at.Kc = b.selfK; at.Vc = b.selfV; at.cap = b.T; at.wo_t = L.wo_s_t;
                attn_kernel<true><<<dim3(b.B, 8), 128, 0, s>>>(at);
                return cudaGetLastError();
            case 3:
                a.x_in = b.xa; a.ln_w = L.ln1_w; a.ln_b = L.ln1_b; a.wt = L.wqkv_t; a.bias = L.bqkv; a.N = 768;
                a.out = b.q; a.kc = b.selfK; a.vc = b.selfV; a.T = b.T;
                return launch_skinny<256, PRO_LN, EPI_QKV>(a, b.B, s);
            case 4:
                a.x_in = b.xb; a.part = b.part; a.bo = L.bo_c; a.x_sum_out = b.xa;
                a.ln_w = L.lnf_w; a.ln_b = L.lnf_b; a.wt = L.w1_t; a.bias = L.b1; a.N = 1024; a.out = b.hbuf;
                return launch_skinny<256, PRO_SUM_LN, EPI_GELU>(a, b.B, s);
            case 5:
                a.x_in = b.hbuf; a.x_stride = MNX_DEC_FF; a.wt = L.w2_t; a.N = 256; a.out = b.part;
                return launch_skinny<128, PRO_PLAIN, EPI_PART>(a, b.B, s);
            default:
                pick_kernel<false><<<b.B, 1024, 0, s>>>(b, w, g, b.xa, b.part, L.b2, nullptr);
                return cudaGetLastError();
        }
    };
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = once();
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters && e == cudaSuccess; ++i) e = once();
    cudaEventRecord(e1, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    *ms = t / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaMemcpyAsync(b.st, &saved, sizeof(DecState), cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    return e;
}

}  // namespace mnx
