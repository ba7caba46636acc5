// Swin-B encoder (patch4 / window12 / dims 128-256-512-1024 / depths 2-2-18-2 / heads 4-8-16-32)
// on sm_100a -- the encoder the reference actually executes (MolNexTR/models/transformers.py:
// PatchEmbed :405-419, SwinTransformerBlock :245-292, WindowAttention :147-178, PatchMerging
// :310-336, Vision_Transformer.forward :504-515).
//
// Data layout: the residual stream x is fp32 [B*H*W][C] (token-major, NHWC); every Linear runs on
// the tcgen05 GEMM (gemm_tc.cu) with bf16 operands that the LayerNorm kernels emit directly in the
// order the GEMM wants (window-partitioned and cyclically shifted for attention), so the
// reference's pad / roll / window_partition / permute / contiguous copies never materialise;
// window_reverse + un-roll + crop + residual add happen in the proj GEMM's epilogue via a row map.
#include <cuda_bf16.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "encoder.cuh"

#include <cstring>
#include <cuda_fp16.h>
#include "gemm_tc.cuh"

namespace mnx {

static constexpr int WS = 12, WIN = 144;
static const int SW_DEPTH[4] = {2, 2, 18, 2};
static const int SW_HEADS[4] = {4, 8, 16, 32};

struct SwinBlockW {
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
    const __nv_bfloat16 *qkv_w, *proj_w, *fc1_w, *fc2_w;
    const float *qkv_b, *proj_b, *fc1_b, *fc2_b;
    const float* rpb_t;   // [nH][529] relative position bias table, head-major
    const uint2* rpb_frag;   // [nH][9 warps][18 key tiles][32 lanes]: bias / scale of the two query rows x two keys a lane
                            // owns in the QK^T accumulator, as packed fp16 pairs (window_attn_kernel)
};
struct SwinMergeW {
    const float *ln_w, *ln_b;
    const __nv_bfloat16* red_w;   // [2C][4C]
};

struct SwinState {
    const float *pe_w, *pe_b, *pe_ln_w, *pe_ln_b;   // patch embed: [48][128] (k-major: coalesced per-channel loads), [128]
    std::vector<SwinBlockW> blocks[4];
    SwinMergeW merge[3];
    const float *norm_w, *norm_b;
    // workspaces
    float *x0 = nullptr, *x1 = nullptr;             // residual stream ping-pong (merging switches)
    __nv_bfloat16 *abuf = nullptr, *qkv = nullptr, *attn = nullptr, *hbuf = nullptr;
    int* row_map = nullptr;                          // per (stage, shift): window row -> token row, kept across calls of one shape
    size_t map_off[4][2] = {};                       // offsets into row_map
    int map_B = 0, map_H = 0, map_W = 0;             // the request the cached maps were built for ...
    cudaStream_t map_stream = nullptr;               // ... and the stream that built them (another stream rebuilds)
    size_t max_tokens = 0;
    int last_B = 0, last_H = 0, last_W = 0;
    int cta_limit = 0;   // cap of the persistent GEMM grids for the forward in progress (EncoderState::cta_limit)
    int num_sms = 148;
};

// ------------------------------------------------------------------------------------------
// patch embed: conv 4x4 stride 4 (3 -> 128) + LayerNorm(128, eps 1e-5); 128 tokens per CTA; weights k-major [48][128]
// ------------------------------------------------------------------------------------------
#define PE_TOK_PER_CTA 128
#define PE_TB 16          // tokens per inner iteration
__global__ void __launch_bounds__(128) patch_embed_kernel(const float* __restrict__ img, int B, int H, int W, int Hp,
                                                          int Wp, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                          const float* __restrict__ ln_b, float eps,
                                                          float* __restrict__ x) {
    __shared__ __align__(16) float patch[PE_TB][48];
    __shared__ float red[PE_TB][4][2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned ntok = (unsigned)B * Hp * Wp;        // < 2^31: checked by the launcher (32-bit index arithmetic below)
    const unsigned hw = (unsigned)Hp * Wp;
    float wr[48];
#pragma unroll
    for (int k = 0; k < 48; ++k) wr[k] = w[k * 128 + tid];      // w is [48][128]
    const float bi = bias[tid], g = ln_w[tid], be = ln_b[tid];
    // thread = output channel: its 48 weights stay in registers while the CTA walks over 64 tokens, 16 at a time; the patch
    // pixels are broadcast out of shared memory four at a time (one LDS.128 per 4 FMA: the scalar version was bound by
    // one shared-memory load per FMA and ran 10x off its HBM roofline)
    for (int it = 0; it < PE_TOK_PER_CTA / PE_TB; ++it) {
        const unsigned tok0 = blockIdx.x * PE_TOK_PER_CTA + it * PE_TB;
        if (tok0 >= ntok) break;
        for (int i = tid; i < PE_TB * 12; i += 128) {       // one (token, channel, row) = 4 contiguous pixels per thread
            const int t = i / 12, cr = i % 12, c = cr >> 2, dy = cr & 3;
            const unsigned tok = tok0 + t;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tok < ntok) {
                const unsigned b = tok / hw, r = tok - b * hw;
                const int py = (int)(r / (unsigned)Wp), px = (int)(r - (unsigned)py * Wp);
                const int yy = py * 4 + dy, xx = px * 4;
                if (yy < H) {
                    const float* src = img + (((size_t)b * 3 + c) * H + yy) * W + xx;
                    if (xx + 3 < W && (W & 3) == 0) {
                        v = *reinterpret_cast<const float4*>(src);
                    } else {
                        if (xx < W) v.x = src[0];
                        if (xx + 1 < W) v.y = src[1];
                        if (xx + 2 < W) v.z = src[2];
                        if (xx + 3 < W) v.w = src[3];
                    }
                }
            }
            *reinterpret_cast<float4*>(&patch[t][c * 16 + dy * 4]) = v;
        }
        __syncthreads();
        float acc[PE_TB];
#pragma unroll
        for (int t = 0; t < PE_TB; ++t) {
            float a = bi;
#pragma unroll
            for (int k4 = 0; k4 < 12; ++k4) {
                const float4 pv = *reinterpret_cast<const float4*>(&patch[t][4 * k4]);
                a = fmaf(pv.x, wr[4 * k4], a); a = fmaf(pv.y, wr[4 * k4 + 1], a);
                a = fmaf(pv.z, wr[4 * k4 + 2], a); a = fmaf(pv.w, wr[4 * k4 + 3], a);
            }
            acc[t] = a;
        }
        // LayerNorm over the 128 channels of each token (two-pass)
#pragma unroll
        for (int t = 0; t < PE_TB; ++t) {
            const float s = warp_sum(acc[t]);
            if (lane == 0) red[t][wid][0] = s;
        }
        __syncthreads();
        float mean[PE_TB];
#pragma unroll
        for (int t = 0; t < PE_TB; ++t) mean[t] = ((red[t][0][0] + red[t][1][0]) + (red[t][2][0] + red[t][3][0])) * (1.0f / 128.0f);
#pragma unroll
        for (int t = 0; t < PE_TB; ++t) {
            const float d = acc[t] - mean[t];
            const float s = warp_sum(d * d);
            if (lane == 0) red[t][wid][1] = s;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < PE_TB; ++t) {
            const float var = ((red[t][0][1] + red[t][1][1]) + (red[t][2][1] + red[t][3][1])) * (1.0f / 128.0f);
            const float rstd = 1.0f / sqrtf(var + eps);
            if (tok0 + t < ntok) x[(size_t)(tok0 + t) * 128 + tid] = (acc[t] - mean[t]) * rstd * g + be;
        }
        __syncthreads();   // patch / red are rewritten by the next iteration
    }
}

// shared with convnext.cu (its stem is the same 4x4/4 patchify + channel LayerNorm, eps 1e-6)
cudaError_t launch_patch_embed(const float* img, int B, int H, int W, const float* w, const float* bias,
                               const float* ln_w, const float* ln_b, float eps, float* x, cudaStream_t s) {
    const int Hc = (H + 3) / 4, Wc = (W + 3) / 4;
    const long long ntok = (long long)B * Hc * Wc;
    if (ntok >= (1ll << 31)) return cudaErrorInvalidValue;
    patch_embed_kernel<<<(unsigned)((ntok + PE_TOK_PER_CTA - 1) / PE_TOK_PER_CTA), 128, 0, s>>>(img, B, H, W, Hc, Wc, w, bias, ln_w, ln_b, eps, x);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// row map of one attention block: window-order row -> token row of x (or -1 for a padded position)
// window order = (b, wh, ww, ph, pw) over the padded, cyclically shifted map (transformers.py:252-266)
// ------------------------------------------------------------------------------------------
__global__ void swin_row_map_kernel(int B, int H, int W, int Hp, int Wp, int shift, int* __restrict__ map) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Hp * Wp;
    if (idx >= total) return;
    const int nww = Wp / WS;
    const int pos = (int)(idx % WIN);
    const long long win = idx / WIN;
    const int nwin = (Hp / WS) * nww;
    const int b = (int)(win / nwin), wi = (int)(win % nwin);
    const int wh = wi / nww, ww = wi % nww;
    const int hs = wh * WS + pos / WS, wsx = ww * WS + pos % WS;
    const int hp = (hs + shift) % Hp, wp = (wsx + shift) % Wp;
    map[idx] = (hp < H && wp < W) ? (int)(((long long)b * H + hp) * W + wp) : -1;
}

// ------------------------------------------------------------------------------------------
// LayerNorm of C-wide rows -> bf16 (GEMM A operand) or fp32; one warp per output row, optional
// gather through a row map (-1 -> zeros: the reference pads AFTER norm1).
// ------------------------------------------------------------------------------------------
// NV = C / 128 float4 per lane: the row lives in registers (one global read), and a warp handles LN_RPW rows whose loads
// are all issued before the first reduction (the kernel is a latency chain per row: map -> row -> two shuffle
// reductions -> store; at one row per warp the 64 resident warps of an SM kept only 32 KB in flight).
#define LN_RPW 4
template <bool OUT_BF16, int NV>
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ x, const int* __restrict__ map,
                                                      long long rows, const float* __restrict__ w,
                                                      const float* __restrict__ b, float eps, void* __restrict__ out) {
    constexpr int C = NV * 128;
    const int lane = threadIdx.x & 31;
    const long long r0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * LN_RPW;
    if (r0 >= rows) return;
    long long src[LN_RPW];
#pragma unroll
    for (int j = 0; j < LN_RPW; ++j) {
        const long long r = r0 + j;
        src[j] = r < rows ? (map ? (long long)map[r] : r) : -2;
    }
    float4 v[LN_RPW][NV];
#pragma unroll
    for (int j = 0; j < LN_RPW; ++j) {
        if (src[j] >= 0) {
            const float4* xr = reinterpret_cast<const float4*>(x + (size_t)src[j] * C);
#pragma unroll
            for (int i = 0; i < NV; ++i) v[j][i] = xr[lane + 32 * i];
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) v[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int j = 0; j < LN_RPW; ++j) {
        if (src[j] == -2) break;                         // past the last row (warp-uniform)
        const long long r = r0 + j;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
        const float mean = warp_sum(s) / (float)C;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float a = v[j][i].x - mean, bb = v[j][i].y - mean, c = v[j][i].z - mean, d = v[j][i].w - mean;
            sq += (a * a + bb * bb) + (c * c + d * d);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)C + eps);
        const bool pad = src[j] < 0;                     // -1: zero row (the reference pads AFTER norm1)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 g = w4[lane + 32 * i], be = b4[lane + 32 * i];
            float y0 = (v[j][i].x - mean) * rstd * g.x + be.x, y1 = (v[j][i].y - mean) * rstd * g.y + be.y;
            float y2 = (v[j][i].z - mean) * rstd * g.z + be.z, y3 = (v[j][i].w - mean) * rstd * g.w + be.w;
            if (pad) y0 = y1 = y2 = y3 = 0.f;
            if (OUT_BF16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(y0, y1), hi = __floats2bfloat162_rn(y2, y3);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)r * C)[lane + 32 * i] = pk;
            } else {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)r * C)[lane + 32 * i] = make_float4(y0, y1, y2, y3);
            }
        }
    }
}
template <bool OUT_BF16>
static cudaError_t launch_ln_rows(const float* x, const int* map, long long rows, int C, const float* w, const float* b, float eps,
                                  void* out, cudaStream_t s) {
    const unsigned grid = (unsigned)((rows + 8 * LN_RPW - 1) / (8 * LN_RPW));
    switch (C) {
        case 128: ln_rows_kernel<OUT_BF16, 1><<<grid, 256, 0, s>>>(x, map, rows, w, b, eps, out); break;
        case 256: ln_rows_kernel<OUT_BF16, 2><<<grid, 256, 0, s>>>(x, map, rows, w, b, eps, out); break;
        case 512: ln_rows_kernel<OUT_BF16, 4><<<grid, 256, 0, s>>>(x, map, rows, w, b, eps, out); break;
        case 1024: ln_rows_kernel<OUT_BF16, 8><<<grid, 256, 0, s>>>(x, map, rows, w, b, eps, out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// PatchMerging gather + LayerNorm(4C): out row (b,h2,w2) = LN(cat[x(2h2,2w2), x(2h2+1,2w2), x(2h2,2w2+1),
// x(2h2+1,2w2+1)]) with zero padding of odd maps (transformers.py:318-333); one warp per output row.
__global__ void __launch_bounds__(256) merge_ln_kernel(const float* __restrict__ x, int B, int H, int W, int C,
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       float eps, __nv_bfloat16* __restrict__ out) {
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= (long long)B * H2 * W2) return;
    const int lane = threadIdx.x & 31;
    const int bb = (int)(r / ((long long)H2 * W2));
    const int rem = (int)(r % ((long long)H2 * W2));
    const int h2 = rem / W2, w2 = rem % W2;
    const int n4 = C >> 2;
    const float4* src[4];
    const int dh[4] = {0, 1, 0, 1}, dw[4] = {0, 0, 1, 1};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int hh = 2 * h2 + dh[p], ww = 2 * w2 + dw[p];
        src[p] = (hh < H && ww < W) ? reinterpret_cast<const float4*>(x + (((size_t)bb * H + hh) * W + ww) * C) : nullptr;
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p)
        if (src[p]) for (int i = lane; i < n4; i += 32) { const float4 v = src[p][i]; s += (v.x + v.y) + (v.z + v.w); }
    const float inv = 1.0f / (float)(4 * C);
    const float mean = warp_sum(s) * inv;
    float sq = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        for (int i = lane; i < n4; i += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src[p]) v = src[p][i];
            const float a = v.x - mean, b2 = v.y - mean, c = v.z - mean, d = v.w - mean;
            sq += (a * a + b2 * b2) + (c * c + d * d);
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * inv + eps);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float4* w4 = reinterpret_cast<const float4*>(w + p * C);
        const float4* b4 = reinterpret_cast<const float4*>(b + p * C);
        uint2* o = reinterpret_cast<uint2*>(out + (size_t)r * 4 * C + p * C);
        for (int i = lane; i < n4; i += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src[p]) v = src[p][i];
            const float4 g = w4[i], be = b4[i];
            __nv_bfloat162 lo = __floats2bfloat162_rn((v.x - mean) * rstd * g.x + be.x, (v.y - mean) * rstd * g.y + be.y);
            __nv_bfloat162 hi = __floats2bfloat162_rn((v.z - mean) * rstd * g.z + be.z, (v.w - mean) * rstd * g.w + be.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            o[i] = pk;
        }
    }
}

// ------------------------------------------------------------------------------------------
// window attention: one CTA per (window, head); 9 warps x 16 query rows; QK^T and PV on
// mma.sync m16n8k16 (bf16 in, fp32 accumulate); bias-table gather, shift mask (-100) and softmax
// in fp32 registers.  144 x 144 x 32 per tile is too small to amortise a tcgen05/TMEM round trip.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_row)));
}

#define WA_SCALE 0.17677669529663687f                       // 32 ** -0.5
#define WA_SCALE_LOG2E (0.17677669529663687f * 1.4426950408889634f)
#define WA_KV_ELEMS (WIN * 40)                              // one [key][32 dims] tile with an 80-byte row stride
#define WA_BIAS_BYTES (9 * 18 * 32 * 8)
#define WA_SMEM (WA_BIAS_BYTES + 6 * WA_KV_ELEMS * 2 + 2 * WIN)

__device__ __forceinline__ float ex2_approx(float x) {      // MUFU.EX2 alone (exp2f adds a denormal-range fix-up per element)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// grid (chunks, heads); 9 warps x 16 query rows; a CTA walks a contiguous run of windows for ONE head.
// The relative-position bias of the head sits in shared memory in accumulator-fragment order for the whole run
// (QK^T accumulators START from bias / scale, so that softmax((qk + b/s) * s) = softmax(qk * s + b)); the K / V rows
// of window i+1 stream into the second buffer with cp.async while window i is computed, its Q fragments are
// prefetched into registers; one block barrier per window.  The shift mask (-100) is only evaluated for windows that
// touch the rolled edge; exp2 with the scale folded into the exponent; probabilities are packed to bf16 as they are
// produced and normalised AFTER the PV product (1 / l on the 16 outputs instead of the 72 probabilities); V^T
// fragments come from ldmatrix.trans.  (The one-window-per-CTA version spent 45 % of its samples waiting on the
// K / V / bias / Q loads: profiles/r2b ncu capture.)
__global__ void __launch_bounds__(288, 2) window_attn_kernel(const __nv_bfloat16* __restrict__ qkv, int C, int nH, int n_win,
                                                          const uint2* __restrict__ rpb_frag, int Hp, int Wp, int shift,
                                                          const int* __restrict__ row_map, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t wa_smem[];
    uint2* bias_s = reinterpret_cast<uint2*>(wa_smem);                                   // [9][18][32]
    __nv_bfloat16* kv_s = reinterpret_cast<__nv_bfloat16*>(wa_smem + WA_BIAS_BYTES);      // [2 buffers][Q | K | V][WIN][40]
    uint8_t* region_s = wa_smem + WA_BIAS_BYTES + 6 * WA_KV_ELEMS * 2;                    // [2][WIN]
    const int h = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = 3 * C;
    const int qr = lane >> 2, qc = (lane & 3) * 2;
    const int i0 = warp * 16 + qr, i1 = i0 + 8;             // the two query rows this thread owns
    const int nww = Wp / WS, nwh = Hp / WS;
    const int w_begin = (int)(((long long)blockIdx.x * n_win) / gridDim.x);
    const int w_end = (int)(((long long)(blockIdx.x + 1) * n_win) / gridDim.x);
    if (w_begin >= w_end) return;

    auto stage_kv = [&](int win, int buf) {
        const size_t row0 = (size_t)win * WIN;
        __nv_bfloat16* Qd = kv_s + (size_t)buf * 3 * WA_KV_ELEMS;
        __nv_bfloat16* Kd = Qd + WA_KV_ELEMS;
        __nv_bfloat16* Vd = Kd + WA_KV_ELEMS;
        for (int i = tid; i < WIN * 4; i += 288) {
            const int key = i >> 2, c8 = i & 3;
            cp_async16(Qd + key * 40 + c8 * 8, qkv + (row0 + key) * ld + h * 32 + c8 * 8);
            cp_async16(Kd + key * 40 + c8 * 8, qkv + (row0 + key) * ld + C + h * 32 + c8 * 8);
            cp_async16(Vd + key * 40 + c8 * 8, qkv + (row0 + key) * ld + 2 * C + h * 32 + c8 * 8);
        }
    };
    // prologue: bias fragments of the head + K / V of the first window (one cp.async group)
    {
        const uint4* src = reinterpret_cast<const uint4*>(rpb_frag + (size_t)h * 9 * 18 * 32);
        uint4* dst = reinterpret_cast<uint4*>(bias_s);
        for (int i = tid; i < WA_BIAS_BYTES / 16; i += 288) cp_async16(dst + i, src + i);
    }
    stage_kv(w_begin, 0);
    cp_async_commit();
    const uint2* bsrc = bias_s + (warp * 18) * 32 + lane;

#pragma unroll 1
    for (int win = w_begin; win < w_end; ++win) {
        const int buf = (win - w_begin) & 1;
        // only windows in the last row / column of the rolled map mix regions (transformers.py:220-243)
        const int wimg = win % (nwh * nww);
        const bool edge = shift > 0 && (wimg / nww == nwh - 1 || wimg % nww == nww - 1);
        uint8_t* region = region_s + buf * WIN;
        if (edge && tid < WIN) {
            const int hs = (wimg / nww) * WS + tid / WS, wsx = (wimg % nww) * WS + tid % WS;
            const int hid = hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2);
            const int wid = wsx < Wp - WS ? 0 : (wsx < Wp - shift ? 1 : 2);
            region[tid] = (uint8_t)(hid * 3 + wid);
        }
        cp_async_wait_all();
        __syncthreads();            // K / V (and region) of this window visible; everybody is done with the other buffer
        if (win + 1 < w_end) {
            stage_kv(win + 1, buf ^ 1);
            cp_async_commit();
        }
        const __nv_bfloat16* Qs = kv_s + (size_t)buf * 3 * WA_KV_ELEMS;
        const __nv_bfloat16* Ks = Qs + WA_KV_ELEMS;
        const __nv_bfloat16* Vs = Ks + WA_KV_ELEMS;
        // Q fragments (A operand, 2 k-steps of 16 dims): lane L supplies the address of row (L & 15) at dims 8 * (L >> 4)
        uint32_t qa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ldmatrix_x4(qa[ks], Qs + (warp * 16 + (lane & 15)) * 40 + ks * 16 + 8 * (lane >> 4));
        float s[18][4];
#pragma unroll
        for (int np = 0; np < 9; ++np) {
            // K fragments of key tiles 2np and 2np+1 (B operand, [key][dim] rows = k-contiguous columns): lane L supplies the
            // address of key 16 np + (L & 7) + 8 (L >> 4) at dims 8 ((L >> 3) & 1) (+16 for the second k-step)
            uint32_t kb[2][4];      // kb[ks] = {b0, b1 of tile 2np, b0, b1 of tile 2np+1}
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
                ldmatrix_x4(kb[ks], Ks + (np * 16 + (lane & 7) + 8 * (lane >> 4)) * 40 + ks * 16 + 8 * ((lane >> 3) & 1));
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int nt = 2 * np + half;
                const uint2 bf = bsrc[nt * 32];
                const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&bf.x));
                const float2 b1 = __half22float2(*reinterpret_cast<const __half2*>(&bf.y));
                s[nt][0] = b0.x; s[nt][1] = b0.y; s[nt][2] = b1.x; s[nt][3] = b1.y;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) mma_bf16_16816(s[nt], qa[ks], kb[ks][2 * half], kb[ks][2 * half + 1]);
            }
        }
        if (edge) {                                          // CTA-uniform
            const int r0 = region[i0], r1 = region[i1];
            const float neg = -100.0f / WA_SCALE;
#pragma unroll
            for (int nt = 0; nt < 18; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int rj = region[nt * 8 + qc + e];
                    if (rj != r0) s[nt][e] += neg;
                    if (rj != r1) s[nt][2 + e] += neg;
                }
            }
        }
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 18; ++nt) {
            m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
            m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        const float mb0 = m0 * WA_SCALE_LOG2E, mb1 = m1 * WA_SCALE_LOG2E;
        float l0 = 0.f, l1 = 0.f;
        uint32_t pk[18][2];                                  // un-normalised probabilities, bf16 pairs (rows i0 / i1)
#pragma unroll
        for (int nt = 0; nt < 18; ++nt) {
            const float p00 = ex2_approx(fmaf(s[nt][0], WA_SCALE_LOG2E, -mb0)), p01 = ex2_approx(fmaf(s[nt][1], WA_SCALE_LOG2E, -mb0));
            const float p10 = ex2_approx(fmaf(s[nt][2], WA_SCALE_LOG2E, -mb1)), p11 = ex2_approx(fmaf(s[nt][3], WA_SCALE_LOG2E, -mb1));
            const __nv_bfloat162 t0 = __floats2bfloat162_rn(p00, p01), t1 = __floats2bfloat162_rn(p10, p11);
            l0 += p00 + p01; l1 += p10 + p11;
            pk[nt][0] = *reinterpret_cast<const uint32_t*>(&t0);
            pk[nt][1] = *reinterpret_cast<const uint32_t*>(&t1);
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        // O = P V : 9 k-steps of 16 keys, 4 n-tiles of 8 dims; V^T fragments via ldmatrix.trans:
        // lane L supplies the row address of key (L & 15) at dims 8 * (L >> 4) (+16 for the second instruction)
        float o[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) {
            const uint32_t pa[4] = {pk[2 * kk][0], pk[2 * kk][1], pk[2 * kk + 1][0], pk[2 * kk + 1][1]};
            uint32_t vb[2][4];      // vb[p] = {b0, b1 of dims 16p..16p+7, b0, b1 of dims 16p+8..16p+15}
#pragma unroll
            for (int p2 = 0; p2 < 2; ++p2) ldmatrix_x4_trans(vb[p2], Vs + (kk * 16 + (lane & 15)) * 40 + 16 * p2 + 8 * (lane >> 4));
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(o[nt], pa, vb[nt >> 1][(nt & 1) * 2], vb[nt >> 1][(nt & 1) * 2 + 1]);
        }
        // window_reverse + un-roll + crop happen HERE: the context of window row r goes to token row_map[r] (-1 = padding),
        // so that the proj GEMM runs in token order and adds into the residual stream with TMA tile reductions
        const size_t row0 = (size_t)win * WIN;
        const int d0 = row_map[row0 + i0], d1 = row_map[row0 + i1];
        __nv_bfloat16* o0 = out + (size_t)(d0 < 0 ? 0 : d0) * C + h * 32;
        __nv_bfloat16* o1 = out + (size_t)(d1 < 0 ? 0 : d1) * C + h * 32;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            __nv_bfloat162 a2 = __floats2bfloat162_rn(o[nt][0] * inv0, o[nt][1] * inv0);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(o[nt][2] * inv1, o[nt][3] * inv1);
            if (d0 >= 0) *reinterpret_cast<__nv_bfloat162*>(o0 + nt * 8 + qc) = a2;
            if (d1 >= 0) *reinterpret_cast<__nv_bfloat162*>(o1 + nt * 8 + qc) = b2;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static std::vector<__nv_bfloat16> to_bf16(const std::vector<float>& v) {
    std::vector<__nv_bfloat16> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16_rn(v[i]);
    return o;
}

#define SW_CUDA(e, x)                                                                                     \
    do {                                                                                                  \
        cudaError_t _c = (x);                                                                             \
        if (_c != cudaSuccess) {                                                                          \
            std::string m = std::string(#x) + " failed: " + cudaGetErrorString(_c);                       \
            mnx_set_error(e, m.c_str());                                                                  \
            return MNX_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

static int up_f32(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape, const float** out) {
    const std::vector<float>* v = mnx_need(e, key, shape);
    if (!v) return MNX_ERR_WEIGHTS;
    SW_CUDA(e, mnx_upload(e, *v, out));
    return MNX_OK;
}
static int up_bf16(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape, const __nv_bfloat16** out) {
    const std::vector<float>* v = mnx_need(e, key, shape);
    if (!v) return MNX_ERR_WEIGHTS;
    std::vector<__nv_bfloat16> h = to_bf16(*v);
    void* d = nullptr;
    SW_CUDA(e, mnx_upload_raw(e, h.data(), h.size() * sizeof(__nv_bfloat16), &d));
    *out = reinterpret_cast<const __nv_bfloat16*>(d);
    return MNX_OK;
}
#define SW_TRY(x) do { int _r = (x); if (_r != MNX_OK) return _r; } while (0)

int swin_finalize(mnx_engine* e, SwinState** out, const mnx_config& cfg) {
    SW_CUDA(e, gemm_tc_configure());
    SW_CUDA(e, cudaFuncSetAttribute(window_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM));
    SwinState* st = new SwinState();
    *out = st;
    SW_CUDA(e, cudaDeviceGetAttribute(&st->num_sms, cudaDevAttrMultiProcessorCount, cfg.device));
    const std::string P = "encoder.transformer.";
    {
        const std::vector<float>* pw = mnx_need(e, P + "patch_embed.proj.weight", {128, 3, 4, 4});
        if (!pw) return MNX_ERR_WEIGHTS;
        std::vector<float> t(48 * 128);
        for (int c = 0; c < 128; ++c)
            for (int k = 0; k < 48; ++k) t[k * 128 + c] = (*pw)[c * 48 + k];
        SW_CUDA(e, mnx_upload(e, t, &st->pe_w));
    }
    SW_TRY(up_f32(e, P + "patch_embed.proj.bias", {128}, &st->pe_b));
    SW_TRY(up_f32(e, P + "patch_embed.norm.weight", {128}, &st->pe_ln_w));
    SW_TRY(up_f32(e, P + "patch_embed.norm.bias", {128}, &st->pe_ln_b));
    // the index buffer is recomputed on the fly in the attention kernel; verify the checkpoint agrees
    std::vector<int64_t> ref_idx(WIN * WIN);
    for (int i = 0; i < WIN; ++i)
        for (int j = 0; j < WIN; ++j)
            ref_idx[i * WIN + j] = (i / WS - j / WS + WS - 1) * (2 * WS - 1) + (i % WS - j % WS + WS - 1);
    for (int s = 0; s < 4; ++s) {
        const int64_t C = 128 << s, nH = SW_HEADS[s];
        st->blocks[s].resize(SW_DEPTH[s]);
        for (int j = 0; j < SW_DEPTH[s]; ++j) {
            const std::string B = P + "layers." + std::to_string(s) + ".blocks." + std::to_string(j) + ".";
            SwinBlockW& w = st->blocks[s][j];
            SW_TRY(up_f32(e, B + "norm1.weight", {C}, &w.ln1_w));
            SW_TRY(up_f32(e, B + "norm1.bias", {C}, &w.ln1_b));
            SW_TRY(up_f32(e, B + "norm2.weight", {C}, &w.ln2_w));
            SW_TRY(up_f32(e, B + "norm2.bias", {C}, &w.ln2_b));
            SW_TRY(up_bf16(e, B + "attn.qkv.weight", {3 * C, C}, &w.qkv_w));
            SW_TRY(up_f32(e, B + "attn.qkv.bias", {3 * C}, &w.qkv_b));
            SW_TRY(up_bf16(e, B + "attn.proj.weight", {C, C}, &w.proj_w));
            SW_TRY(up_f32(e, B + "attn.proj.bias", {C}, &w.proj_b));
            SW_TRY(up_bf16(e, B + "mlp.fc1.weight", {4 * C, C}, &w.fc1_w));
            SW_TRY(up_f32(e, B + "mlp.fc1.bias", {4 * C}, &w.fc1_b));
            SW_TRY(up_bf16(e, B + "mlp.fc2.weight", {C, 4 * C}, &w.fc2_w));
            SW_TRY(up_f32(e, B + "mlp.fc2.bias", {C}, &w.fc2_b));
            const std::vector<float>* tab = mnx_need(e, B + "attn.relative_position_bias_table", {529, nH});
            if (!tab) return MNX_ERR_WEIGHTS;
            std::vector<float> tt((size_t)nH * 529);
            for (int i = 0; i < 529; ++i)
                for (int hh = 0; hh < nH; ++hh) tt[(size_t)hh * 529 + i] = (*tab)[(size_t)i * nH + hh];
            SW_CUDA(e, mnx_upload(e, tt, &w.rpb_t));
            {   // accumulator-fragment order of window_attn_kernel, bias / scale, packed bf16 pairs
                std::vector<uint32_t> fr((size_t)nH * 9 * 18 * 32 * 2);
                // fp16, not bf16: |bias / scale| stays far below 65504 and keeps 11 significant bits (the bias is added to an
                // fp32 QK^T accumulator; bf16 here cost as much accuracy as the bf16 q / k operands themselves)
                auto bf16_bits = [](float v) { __half b = __float2half_rn(v); uint16_t u; memcpy(&u, &b, 2); return (uint32_t)u; };
                const float inv_scale = 1.0f / 0.17677669529663687f;
                for (int hh = 0; hh < nH; ++hh)
                    for (int wp = 0; wp < 9; ++wp)
                        for (int nt = 0; nt < 18; ++nt)
                            for (int ln = 0; ln < 32; ++ln) {
                                const int qi0 = wp * 16 + (ln >> 2), qi1 = qi0 + 8, jj = nt * 8 + (ln & 3) * 2;
                                auto bias = [&](int i, int j) { return tt[(size_t)hh * 529 + ref_idx[i * WIN + j]] * inv_scale; };
                                const size_t o = ((((size_t)hh * 9 + wp) * 18 + nt) * 32 + ln) * 2;
                                fr[o] = bf16_bits(bias(qi0, jj)) | (bf16_bits(bias(qi0, jj + 1)) << 16);
                                fr[o + 1] = bf16_bits(bias(qi1, jj)) | (bf16_bits(bias(qi1, jj + 1)) << 16);
                            }
                void* dptr = nullptr;
                SW_CUDA(e, mnx_upload_raw(e, fr.data(), fr.size() * 4, &dptr));
                w.rpb_frag = reinterpret_cast<const uint2*>(dptr);
            }
            const std::vector<int64_t>* idx = mnx_need_i64(e, B + "attn.relative_position_index", {WIN, WIN});
            if (!idx) return MNX_ERR_WEIGHTS;
            if (*idx != ref_idx) {
                mnx_set_error(e, (B + "attn.relative_position_index differs from the window-12 table the kernels assume").c_str());
                return MNX_ERR_WEIGHTS;
            }
        }
        if (s < 3) {
            const std::string D = P + "layers." + std::to_string(s) + ".downsample.";
            SW_TRY(up_f32(e, D + "norm.weight", {4 * C}, &st->merge[s].ln_w));
            SW_TRY(up_f32(e, D + "norm.bias", {4 * C}, &st->merge[s].ln_b));
            SW_TRY(up_bf16(e, D + "reduction.weight", {2 * C, 4 * C}, &st->merge[s].red_w));
        }
    }
    SW_TRY(up_f32(e, P + "norm.weight", {1024}, &st->norm_w));
    SW_TRY(up_f32(e, P + "norm.bias", {1024}, &st->norm_b));
    // workspaces for the largest request: stage-0 map padded to window multiples
    const size_t Hq = (cfg.max_height + 3) / 4, Wq = (cfg.max_width + 3) / 4;
    const size_t Hp = (Hq + WS - 1) / WS * WS, Wp = (Wq + WS - 1) / WS * WS;
    const size_t tok = (size_t)cfg.max_batch * Hp * Wp;      // >= tokens (and padded tokens) of every stage
    st->max_tokens = tok;
    void* p = nullptr;
    SW_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 128 * sizeof(float))); st->x0 = (float*)p;
    SW_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 128 * sizeof(float) / 2 + 1024)); st->x1 = (float*)p;
    SW_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 128 * 2)); st->abuf = (__nv_bfloat16*)p;
    SW_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 384 * 2)); st->qkv = (__nv_bfloat16*)p;
    SW_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 128 * 2)); st->attn = (__nv_bfloat16*)p;
    SW_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 512 * 2)); st->hbuf = (__nv_bfloat16*)p;
    {
        size_t total = 0, hq = Hq, wq = Wq;
        for (int sidx = 0; sidx < 4; ++sidx) {
            const size_t mw = (size_t)cfg.max_batch * ((hq + WS - 1) / WS * WS) * ((wq + WS - 1) / WS * WS);
            st->map_off[sidx][0] = total; st->map_off[sidx][1] = total + mw;
            total += 2 * mw;
            hq = (hq + 1) / 2; wq = (wq + 1) / 2;
        }
        SW_CUDA(e, mnx_dev_alloc_bytes(e, &p, total * sizeof(int))); st->row_map = (int*)p;
    }
    return MNX_OK;
}

void swin_destroy(SwinState* st) { delete st; }

static cudaError_t gemm(const __nv_bfloat16* A, const __nv_bfloat16* W, long long M, int N, int K, int epi,
                        const float* bias, const int* row_map, void* out, cudaStream_t s, int cta_limit) {
    GemmParams p{};
    p.cta_limit = cta_limit;
    p.A = A; p.W = W; p.M = (int)M; p.N = N; p.K = K; p.epilogue = epi; p.bias = bias; p.row_map = row_map; p.out = out;
    return gemm_tc_launch(p, s);
}

int swin_forward(mnx_engine* e, SwinState* st, const float* images, int B, int H, int W, float* features,
                 cudaStream_t s, int* launches, int cta_limit) {
    int nl = 0;
    st->cta_limit = cta_limit;
    int Hc = (H + 3) / 4, Wc = (W + 3) / 4;
    {
        const long long ntok = (long long)B * Hc * Wc;
        patch_embed_kernel<<<(unsigned)((ntok + PE_TOK_PER_CTA - 1) / PE_TOK_PER_CTA), 128, 0, s>>>(
            images, B, H, W, Hc, Wc, st->pe_w, st->pe_b, st->pe_ln_w, st->pe_ln_b, 1e-5f, st->x0);
        SW_CUDA(e, cudaGetLastError()); ++nl;
    }
    float* x = st->x0;
    float* x_other = st->x1;
    const bool maps_cached = st->map_B == B && st->map_H == H && st->map_W == W && st->map_stream == s;
    st->map_B = 0;      // (a failed forward leaves no half-built cache behind)
    for (int stage = 0; stage < 4; ++stage) {
        const int C = 128 << stage, nH = SW_HEADS[stage];
        const int Hp = (Hc + WS - 1) / WS * WS, Wp = (Wc + WS - 1) / WS * WS;
        const long long M = (long long)B * Hc * Wc, Mw = (long long)B * Hp * Wp;
        if ((size_t)Mw * C > st->max_tokens * 128) {
            mnx_set_error(e, "swin workspace too small for this request");
            return MNX_ERR_CAPACITY;
        }
        // the two maps of the stage (plain / shifted windows) depend on the shapes only: built once per request shape
        if (!maps_cached) {
            for (int z = 0; z < 2; ++z) {
                swin_row_map_kernel<<<(unsigned)((Mw + 255) / 256), 256, 0, s>>>(B, Hc, Wc, Hp, Wp, z * (WS / 2), st->row_map + st->map_off[stage][z]);
                SW_CUDA(e, cudaGetLastError()); ++nl;
            }
        }
        for (int j = 0; j < SW_DEPTH[stage]; ++j) {
            const SwinBlockW& w = st->blocks[stage][j];
            const int shift = (j % 2 == 0) ? 0 : WS / 2;
            const int* map0 = st->row_map + st->map_off[stage][j % 2];
            // norm1 -> (pad, roll, window partition) -> bf16
            SW_CUDA(e, launch_ln_rows<true>(x, map0, Mw, C, w.ln1_w, w.ln1_b, 1e-5f, st->abuf, s)); ++nl;
            SW_CUDA(e, gemm(st->abuf, w.qkv_w, Mw, 3 * C, C, GEMM_EPI_BF16, w.qkv_b, nullptr, st->qkv, s, st->cta_limit)); ++nl;
            {
                // two resident CTAs per SM, each walking a contiguous run of windows of one head
                const int n_win = (int)(Mw / WIN);
                int chunks = (2 * st->num_sms) / nH;
                if (chunks < 1) chunks = 1;
                if (chunks > n_win) chunks = n_win;
                window_attn_kernel<<<dim3((unsigned)chunks, nH), 288, WA_SMEM, s>>>(st->qkv, C, nH, n_win, w.rpb_frag, Hp, Wp, shift, map0, st->attn);
            }
            SW_CUDA(e, cudaGetLastError()); ++nl;
            // proj + residual, in token order (the attention kernel already undid the window partition / roll / padding)
            SW_CUDA(e, gemm(st->attn, w.proj_w, M, C, C, GEMM_EPI_RESADD_F32, w.proj_b, nullptr, x, s, st->cta_limit)); ++nl;
            // norm2 -> fc1 (GELU) -> fc2 + residual
            SW_CUDA(e, launch_ln_rows<true>(x, nullptr, M, C, w.ln2_w, w.ln2_b, 1e-5f, st->abuf, s)); ++nl;
            SW_CUDA(e, gemm(st->abuf, w.fc1_w, M, 4 * C, C, GEMM_EPI_GELU_BF16, w.fc1_b, nullptr, st->hbuf, s, st->cta_limit)); ++nl;
            SW_CUDA(e, gemm(st->hbuf, w.fc2_w, M, C, 4 * C, GEMM_EPI_RESADD_F32, w.fc2_b, nullptr, x, s, st->cta_limit)); ++nl;
        }
        if (stage < 3) {
            const int H2 = (Hc + 1) / 2, W2 = (Wc + 1) / 2;
            const long long M2 = (long long)B * H2 * W2;
            merge_ln_kernel<<<(unsigned)((M2 + 7) / 8), 256, 0, s>>>(x, B, Hc, Wc, C, st->merge[stage].ln_w,
                                                                    st->merge[stage].ln_b, 1e-5f, st->abuf);
            SW_CUDA(e, cudaGetLastError()); ++nl;
            SW_CUDA(e, gemm(st->abuf, st->merge[stage].red_w, M2, 2 * C, 4 * C, GEMM_EPI_F32, nullptr, nullptr, x_other, s, st->cta_limit)); ++nl;
            float* t = x; x = x_other; x_other = t;
            Hc = H2; Wc = W2;
        }
    }
    const long long M = (long long)B * Hc * Wc;
    SW_CUDA(e, launch_ln_rows<false>(x, nullptr, M, 1024, st->norm_w, st->norm_b, 1e-5f, features, s)); ++nl;
    st->last_B = B; st->last_H = H; st->last_W = W;
    st->map_B = B; st->map_H = H; st->map_W = W; st->map_stream = s;
    *launches += nl;
    return MNX_OK;
}

int swin_time_kernel(mnx_engine* e, SwinState*, int, int, float*, cudaStream_t) {
    mnx_set_error(e, "swin kernel timing ids are not wired up yet");
    return MNX_ERR_INVALID;
}

}  // namespace mnx
