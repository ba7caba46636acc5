// ConvNeXt-B encoder (dims 128/256/512/1024, depths 3/3/27/3) on sm_100a -- the encoder `north_star`
// names.  The reference's branch for it (MolNexTR/components.py:121-126,163-166) is dead code
// (SURVEY.md F2), so the semantics implemented are timm's ConvNeXt `forward_features`:
//   stem conv4x4/4 + LayerNorm2d | per stage: [LayerNorm2d + conv2x2/2] then blocks of
//   dwconv7x7 -> LayerNorm(C, eps 1e-6) -> Linear C->4C -> GELU -> Linear 4C->C -> *gamma -> +x.
//
// Kernels: dwconv_stats_kernel (7x7 depthwise conv, NHWC fp32 in, bf16 GEMM operand out + the statistics of
// the channel LayerNorm, which the fc1 GEMM applies in its epilogue; input halo tiles arrive by 4-D TMA with
// out-of-bounds zero fill = the conv's zero padding), downsample_ln_kernel (per-pixel LN + 2x2 patch gather -> bf16), and the shared
// tcgen05 GEMM (gemm_tc.cu) for fc1 (+GELU), fc2 (+gamma, +residual, in place) and the 2x2 conv.
#include <cuda.h>
#include <cuda_bf16.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "encoder.cuh"
#include "gemm_tc.cuh"

namespace mnx {

cudaError_t launch_patch_embed(const float* img, int B, int H, int W, const float* w, const float* bias,
                               const float* ln_w, const float* ln_b, float eps, float* x, cudaStream_t s);

static const int CN_DEPTH[4] = {3, 3, 27, 3};

struct CnBlockW {
    const float *dw_w, *dw_b;        // [49][C] tap-major, [C]
    const __nv_bfloat16 *fc1_w, *fc2_w;   // fc1_w = mlp.fc1.weight * diag(norm.weight)  (LayerNorm folded, see gemm_tc.cuh)
    const float *fc1_b, *fc2_b, *gamma;   // fc1_b = mlp.fc1.weight @ norm.bias + mlp.fc1.bias
    const float* fc1_colsum;              // [4C] row sums of the bf16 fc1_w
};
struct CnDownW {
    const float *ln_w, *ln_b;        // LayerNorm2d over C_in
    const __nv_bfloat16* w;          // [C_out][4*C_in], k = (kh*2 + kw)*C_in + c
    const float* b;
};
struct ConvNextState {
    const float *stem_w, *stem_b, *stem_ln_w, *stem_ln_b;
    std::vector<CnBlockW> blocks[4];
    CnDownW down[4];
    float *x0 = nullptr, *x1 = nullptr;
    __nv_bfloat16 *abuf = nullptr, *hbuf = nullptr;
    float* stats = nullptr;          // [splits][tokens][2] LayerNorm statistics of the block in flight
    int tile[4] = {8, 8, 8, 12};     // tile side of dwconv_stats_kernel per stage (MNX_DW_TILE="a,b,c,d", values 8 or 12)
    int split[4] = {1, 1, 1, 8};     // channel split (gridDim.z) of dwconv_stats_kernel per stage; MNX_DW_SPLIT="a,b,c,d" overrides
    size_t max_tokens = 0;
    int last_B = 0, last_H = 0, last_W = 0;
    int cta_limit = 0;   // cap of the persistent GEMM grids for the forward in progress (EncoderState::cta_limit)
};

// ------------------------------------------------------------------------------------------
// depthwise 7x7 (pad 3) + bias -> bf16, plus the per-pixel statistics of the channel LayerNorm that follows
// (timm ConvNeXtBlock: conv_dw -> norm -> mlp.fc1).  The LayerNorm itself is applied inside the fc1 GEMM's epilogue
// (GEMM_EPI_LNFOLD_GELU_BF16, gemm_tc.cuh):  fc1(LN(x)) = rstd * (W' x - mean * colsum(W')) + (W beta + b) with
// W' = W diag(gamma) -- so this kernel never revisits its output.  (The previous version normalised in place once all
// channels of a pixel were known: re-reading its own bf16 output from L2 was 45 % of its samples in the ncu capture
// profiles/r1c_summary.md, a latency chain of 128 dependent load -> store pairs per lane.)
// One CTA = 12 x 12 output pixels x (C / gridDim.z) channels, 6 warps; warp w owns output rows 2w and 2w+1.  Channels
// are processed 64 at a time (lane = a PAIR of channels, arithmetic on packed fp32x2 FFMA2): the 18 x 18 x 64 input
// halo tile and the 49 x 64 weight tile of a chunk are fetched by TMA (4-D / 2-D tensor maps; out-of-bounds zero fill
// is the conv's zero padding) into a single 93 KB stage, two CTAs per SM.  Tile size: the kernel's real bound is the
// L2 -> shared-memory fill -- with 8 x 8 tiles ((14 x 14 x 256 + 12.5 K) bytes per 64 x 64 outputs = 15.3 B per output)
// every stage of the network ran at the same ~6 TB/s of fill whatever the grid / occupancy / pipelining (measured: one
// to three resident CTAs, 1-16 channel splits, a persistent two-stage ring), i.e. the FMA roofline of stage 2 would
// have needed 11.6 TB/s of the chip's ~12.4 TB/s L2 cap; 12 x 12 tiles fetch 10.4 B per output and divide every map of
// a 384 x 384 image exactly.  Every input row that is read (18 x 8 bytes per lane) feeds BOTH output rows of the warp
// (kernel rows ky and ky-1; the previous weight row stays in registers): 25 shared loads per 168 FFMA2.  Statistics:
// fp32 sum / sum of squares of the UNROUNDED conv outputs per pixel, reduced over the warp's lanes after every chunk
// (47 shuffles: the 48 per-lane partials would not fit in registers next to the accumulators) and carried in two
// registers -> stats[blockIdx.z][pixel] (the GEMM epilogue adds the gridDim.z partials in a fixed order).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}

// tile side TS = 8 (4 warps, three CTAs per SM) or 12 (6 warps, two CTAs per SM; 10.4 instead of 15.3 fetched bytes per output)
template <int TS> struct DwCfg {
    static constexpr int WARPS = TS / 2, IN = TS + 6, IN_FLOATS = IN * IN * 64;
    static constexpr int SMEM = (IN_FLOATS + 49 * 64) * 4 + 16 + 128;
    static constexpr int CTAS = TS == 8 ? 3 : 2;
};
#define DW_W_FLOATS (49 * 64)

template <int C, int TS>
__global__ void __launch_bounds__(32 * DwCfg<TS>::WARPS, DwCfg<TS>::CTAS) dwconv_stats_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                              const __grid_constant__ CUtensorMap tmap_w, int H, int W,
                                                              const float* __restrict__ dw_b,
                                                              __nv_bfloat16* __restrict__ out, float* __restrict__ stats) {
    constexpr int DW_TILE = TS, DW_IN = DwCfg<TS>::IN, DW_IN_FLOATS = DwCfg<TS>::IN_FLOATS;
    extern __shared__ __align__(128) uint8_t smem_raw[];           // (no integer round trip: keeps LDS, not generic LD)
    float* in_buf = reinterpret_cast<float*>(smem_raw);            // [TS + 6][TS + 6][64]
    float* w_buf = in_buf + DW_IN_FLOATS;                          // [49][64]
    uint64_t* bar = reinterpret_cast<uint64_t*>(w_buf + DW_W_FLOATS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_x = (W + DW_TILE - 1) / DW_TILE;
    const int b = blockIdx.y;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
    const int y0 = ty * DW_TILE, x0 = tx * DW_TILE;
    const int yA = y0 + 2 * warp;                                  // this warp's first output row
    const int nchunk = (C / 64) / (int)gridDim.z;                  // chunks of this CTA
    const int chunk0 = (int)blockIdx.z * nchunk;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    }
    __syncthreads();
    auto issue = [&](int chunk) {
        mbar_arrive_expect_tx(bar, (DW_IN_FLOATS + DW_W_FLOATS) * 4);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                smem_u32(in_buf)),
            "l"(reinterpret_cast<uint64_t>(&tmap_x)), "r"(smem_u32(bar)), "r"(chunk * 64), "r"(x0 - 3), "r"(y0 - 3), "r"(b)
            : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                smem_u32(w_buf)),
            "l"(reinterpret_cast<uint64_t>(&tmap_w)), "r"(smem_u32(bar)), "r"(chunk * 64), "r"(0)
            : "memory");
    };
    if (threadIdx.x == 0) issue(chunk0);

    // running warp totals of the 48 statistics value v = (r * 12 + ox) * 2 + {0: sum, 1: squares}: lane L carries value L in
    // tot_a and value 32 + (L >> 1) in tot_b
    float tot_a = 0.f, tot_b = 0.f;
    float st8[TS == 8 ? 32 : 1];       // TS == 8: per-lane partials of value (r * 8 + ox) * 2 + k, reduced once at the end
#pragma unroll
    for (int i = 0; i < (TS == 8 ? 32 : 1); ++i) st8[i] = 0.f;

    const float2* tin = reinterpret_cast<const float2*>(in_buf);
    const float2* tw = reinterpret_cast<const float2*>(w_buf);
#pragma unroll 1
    for (int ci = 0; ci < nchunk; ++ci) {
        const int c = (chunk0 + ci) * 64 + 2 * lane;
        const float2 bias = *reinterpret_cast<const float2*>(dw_b + c);
        mbar_wait(bar, (uint32_t)ci & 1u);
        float2 acc0[DW_TILE], acc1[DW_TILE];
#pragma unroll
        for (int ox = 0; ox < DW_TILE; ++ox) acc0[ox] = acc1[ox] = bias;
        float2 wprev[7];
#pragma unroll
        for (int ir = 0; ir < 8; ++ir) {       // input row 2*warp + ir of the halo tile: ky = ir for row A, ir - 1 for row B
            float2 row[DW_IN];
#pragma unroll
            for (int ix = 0; ix < DW_IN; ++ix) row[ix] = tin[((2 * warp + ir) * DW_IN + ix) * 32 + lane];
            float2 wcur[7];
            if (ir < 7) {
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) wcur[kx] = tw[(ir * 7 + kx) * 32 + lane];
            }
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                if (ir < 7) {
#pragma unroll
                    for (int ox = 0; ox < DW_TILE; ++ox) acc0[ox] = ffma2(row[ox + kx], wcur[kx], acc0[ox]);
                }
                if (ir > 0) {
#pragma unroll
                    for (int ox = 0; ox < DW_TILE; ++ox) acc1[ox] = ffma2(row[ox + kx], wprev[kx], acc1[ox]);
                }
            }
            if (ir < 7) {
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) wprev[kx] = wcur[kx];
            }
        }
        __syncthreads();   // everyone is done with the stage: refill it while the results are written out
        if (threadIdx.x == 0 && ci + 1 < nchunk) issue(chunk0 + ci + 1);
        if constexpr (TS == 8) {
            // 32 per-lane partials fit in registers next to the 8-wide accumulators: one reduction per CTA (below)
#pragma unroll
            for (int ox = 0; ox < DW_TILE; ++ox) {
                st8[ox * 2] += acc0[ox].x + acc0[ox].y;
                st8[ox * 2 + 1] = fmaf(acc0[ox].x, acc0[ox].x, fmaf(acc0[ox].y, acc0[ox].y, st8[ox * 2 + 1]));
                st8[16 + ox * 2] += acc1[ox].x + acc1[ox].y;
                st8[16 + ox * 2 + 1] = fmaf(acc1[ox].x, acc1[ox].x, fmaf(acc1[ox].y, acc1[ox].y, st8[16 + ox * 2 + 1]));
            }
        }
        float st[TS == 12 ? 48 : 1];
#pragma unroll
        for (int ox = 0; ox < DW_TILE; ++ox) {
            if constexpr (TS == 12) {
                st[ox * 2] = acc0[ox].x + acc0[ox].y;
                st[ox * 2 + 1] = fmaf(acc0[ox].x, acc0[ox].x, acc0[ox].y * acc0[ox].y);
                st[2 * DW_TILE + ox * 2] = acc1[ox].x + acc1[ox].y;
                st[2 * DW_TILE + ox * 2 + 1] = fmaf(acc1[ox].x, acc1[ox].x, acc1[ox].y * acc1[ox].y);
            }
            const int x = x0 + ox;
            if (x < W) {
                if (yA < H)
                    *reinterpret_cast<__nv_bfloat162*>(out + (((size_t)b * H + yA) * W + x) * C + c) = __floats2bfloat162_rn(acc0[ox].x, acc0[ox].y);
                if (yA + 1 < H)
                    *reinterpret_cast<__nv_bfloat162*>(out + (((size_t)b * H + yA + 1) * W + x) * C + c) = __floats2bfloat162_rn(acc1[ox].x, acc1[ox].y);
            }
        }
        if constexpr (TS == 12) {
            // transposing warp reductions after every chunk (the 48 per-lane partials would not fit in registers next to the
            // 12-wide accumulators): values 0..31 -> lane L ends with the warp total of value L (31 shuffles); values 32..47 ->
            // lanes 2i and 2i+1 end with the total of value 32 + i (16 shuffles)
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                    const float keep = hi ? st[i + off] : st[i];
                    const float send = hi ? st[i] : st[i + off];
                    st[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            tot_a += st[0];
#pragma unroll
            for (int off = 16, n = 8; off >= 2; off >>= 1, n >>= 1) {
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const float keep = hi ? st[32 + i + n] : st[32 + i];
                    const float send = hi ? st[32 + i] : st[32 + i + n];
                    st[32 + i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            st[32] += __shfl_xor_sync(0xffffffffu, st[32], 1);
            tot_b += st[32];
        }
    }
    if constexpr (TS == 8) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
                const float keep = hi ? st8[i + off] : st8[i];
                const float send = hi ? st8[i] : st8[i + off];
                st8[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        tot_a = st8[0];
    }
    {
        const size_t M = (size_t)gridDim.y * H * W;
        float* dst = stats + (size_t)blockIdx.z * M * 2;
        {   // value L = (r * 12 + ox) * 2 + k
            const int pix = lane >> 1, r = pix / DW_TILE, ox = pix % DW_TILE;
            const int y = yA + r, x = x0 + ox;
            if (y < H && x < W) dst[(((size_t)b * H + y) * W + x) * 2 + (lane & 1)] = tot_a;
        }
        if (TS == 12 && (lane & 1) == 0) {   // value 32 + (L >> 1): after the four halving steps lane L holds index 8 b16 + 4 b8 + 2 b4 + b2
            const int v = 32 + (lane >> 1);
            const int pix = v >> 1, r = pix / DW_TILE, ox = pix % DW_TILE;
            const int y = yA + r, x = x0 + ox;
            if (y < H && x < W) dst[(((size_t)b * H + y) * W + x) * 2 + (v & 1)] = tot_b;
        }
    }
}

// LayerNorm2d (per pixel over C_in) then 2x2 / stride-2 patch gather: out row (b, h2, w2) holds the four
// normalised pixels in (kh, kw, c) order.  One warp per output row.
__global__ void __launch_bounds__(256) downsample_ln_kernel(const float* __restrict__ x, int B, int H, int W, int C,
                                                            const float* __restrict__ w, const float* __restrict__ bvec,
                                                            float eps, __nv_bfloat16* __restrict__ out) {
    const int H2 = H / 2, W2 = W / 2;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= (long long)B * H2 * W2) return;
    const int lane = threadIdx.x & 31;
    const int bb = (int)(r / ((long long)H2 * W2));
    const int rem = (int)(r % ((long long)H2 * W2));
    const int h2 = rem / W2, w2 = rem % W2;
    const int n4 = C >> 2;
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(bvec);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int hh = 2 * h2 + (p >> 1), ww = 2 * w2 + (p & 1);
        const float4* src = reinterpret_cast<const float4*>(x + (((size_t)bb * H + hh) * W + ww) * C);
        float s = 0.f;
        for (int i = lane; i < n4; i += 32) { const float4 v = src[i]; s += (v.x + v.y) + (v.z + v.w); }
        const float mean = warp_sum(s) / (float)C;
        float sq = 0.f;
        for (int i = lane; i < n4; i += 32) {
            const float4 v = src[i];
            const float a = v.x - mean, b2 = v.y - mean, c = v.z - mean, d = v.w - mean;
            sq += (a * a + b2 * b2) + (c * c + d * d);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)C + eps);
        uint2* o = reinterpret_cast<uint2*>(out + (size_t)r * 4 * C + (size_t)p * C);
        for (int i = lane; i < n4; i += 32) {
            const float4 v = src[i], g = w4[i], be = b4[i];
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
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_cn_encode = nullptr;

#define CN_CUDA(e, x)                                                                                     \
    do {                                                                                                  \
        cudaError_t _c = (x);                                                                             \
        if (_c != cudaSuccess) {                                                                          \
            std::string m = std::string(#x) + " failed: " + cudaGetErrorString(_c);                       \
            mnx_set_error(e, m.c_str());                                                                  \
            return MNX_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)
#define CN_TRY(x) do { int _r = (x); if (_r != MNX_OK) return _r; } while (0)

static int cn_f32(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape, const float** out) {
    const std::vector<float>* v = mnx_need(e, key, shape);
    if (!v) return MNX_ERR_WEIGHTS;
    CN_CUDA(e, mnx_upload(e, *v, out));
    return MNX_OK;
}
static int cn_bf16_vec(mnx_engine* e, const std::vector<float>& v, const __nv_bfloat16** out) {
    std::vector<__nv_bfloat16> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16_rn(v[i]);
    void* d = nullptr;
    CN_CUDA(e, mnx_upload_raw(e, h.data(), h.size() * sizeof(__nv_bfloat16), &d));
    *out = reinterpret_cast<const __nv_bfloat16*>(d);
    return MNX_OK;
}
static int cn_bf16(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape, const __nv_bfloat16** out) {
    const std::vector<float>* v = mnx_need(e, key, shape);
    if (!v) return MNX_ERR_WEIGHTS;
    return cn_bf16_vec(e, *v, out);
}

int convnext_finalize(mnx_engine* e, ConvNextState** out, const mnx_config& cfg) {
    CN_CUDA(e, gemm_tc_configure());
    if (!g_cn_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CN_CUDA(e, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) { mnx_set_error(e, "cuTensorMapEncodeTiled unavailable"); return MNX_ERR_CUDA; }
        g_cn_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    CN_CUDA(e, (cudaFuncSetAttribute(dwconv_stats_kernel<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<8>::SMEM)));
    CN_CUDA(e, (cudaFuncSetAttribute(dwconv_stats_kernel<256, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<8>::SMEM)));
    CN_CUDA(e, (cudaFuncSetAttribute(dwconv_stats_kernel<512, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<8>::SMEM)));
    CN_CUDA(e, (cudaFuncSetAttribute(dwconv_stats_kernel<1024, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<8>::SMEM)));
    CN_CUDA(e, (cudaFuncSetAttribute(dwconv_stats_kernel<128, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<12>::SMEM)));
    CN_CUDA(e, (cudaFuncSetAttribute(dwconv_stats_kernel<256, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<12>::SMEM)));
    CN_CUDA(e, (cudaFuncSetAttribute(dwconv_stats_kernel<512, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<12>::SMEM)));
    CN_CUDA(e, (cudaFuncSetAttribute(dwconv_stats_kernel<1024, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<12>::SMEM)));
    if (cfg.max_height % 32 != 0 || cfg.max_width % 32 != 0) {
        mnx_set_error(e, "ConvNeXt-B needs image bounds that are multiples of 32");
        return MNX_ERR_INVALID;
    }
    ConvNextState* st = new ConvNextState();
    *out = st;
    const std::string P = "encoder.cnn.";
    {   // k-major [48][128] for launch_patch_embed (coalesced per-channel loads)
        const std::vector<float>* pw = mnx_need(e, P + "stem.0.weight", {128, 3, 4, 4});
        if (!pw) return MNX_ERR_WEIGHTS;
        std::vector<float> t(48 * 128);
        for (int c = 0; c < 128; ++c)
            for (int k = 0; k < 48; ++k) t[k * 128 + c] = (*pw)[c * 48 + k];
        CN_CUDA(e, mnx_upload(e, t, &st->stem_w));
    }
    CN_TRY(cn_f32(e, P + "stem.0.bias", {128}, &st->stem_b));
    CN_TRY(cn_f32(e, P + "stem.1.weight", {128}, &st->stem_ln_w));
    CN_TRY(cn_f32(e, P + "stem.1.bias", {128}, &st->stem_ln_b));
    for (int s = 0; s < 4; ++s) {
        const int64_t C = 128 << s;
        if (s > 0) {
            const int64_t Ci = C / 2;
            const std::string D = P + "stages." + std::to_string(s) + ".downsample.";
            CN_TRY(cn_f32(e, D + "0.weight", {Ci}, &st->down[s].ln_w));
            CN_TRY(cn_f32(e, D + "0.bias", {Ci}, &st->down[s].ln_b));
            const std::vector<float>* w = mnx_need(e, D + "1.weight", {C, Ci, 2, 2});
            if (!w) return MNX_ERR_WEIGHTS;
            std::vector<float> r((size_t)C * 4 * Ci);
            for (int64_t n = 0; n < C; ++n)
                for (int64_t c = 0; c < Ci; ++c)
                    for (int kh = 0; kh < 2; ++kh)
                        for (int kw = 0; kw < 2; ++kw)
                            r[(size_t)n * 4 * Ci + (size_t)(kh * 2 + kw) * Ci + c] = (*w)[(((size_t)n * Ci + c) * 2 + kh) * 2 + kw];
            CN_TRY(cn_bf16_vec(e, r, &st->down[s].w));
            CN_TRY(cn_f32(e, D + "1.bias", {C}, &st->down[s].b));
        }
        st->blocks[s].resize(CN_DEPTH[s]);
        for (int j = 0; j < CN_DEPTH[s]; ++j) {
            const std::string B = P + "stages." + std::to_string(s) + ".blocks." + std::to_string(j) + ".";
            CnBlockW& w = st->blocks[s][j];
            const std::vector<float>* dw = mnx_need(e, B + "conv_dw.weight", {C, 1, 7, 7});
            if (!dw) return MNX_ERR_WEIGHTS;
            std::vector<float> t((size_t)49 * C);
            for (int64_t c = 0; c < C; ++c)
                for (int i = 0; i < 49; ++i) t[(size_t)i * C + c] = (*dw)[(size_t)c * 49 + i];
            CN_CUDA(e, mnx_upload(e, t, &w.dw_w));
            CN_TRY(cn_f32(e, B + "conv_dw.bias", {C}, &w.dw_b));
            {   // fold the LayerNorm's affine part into fc1: W' = W diag(g), b' = W beta + b, colsum = W' 1 (of the bf16 W')
                const std::vector<float>* lw = mnx_need(e, B + "norm.weight", {C});
                const std::vector<float>* lb = mnx_need(e, B + "norm.bias", {C});
                const std::vector<float>* fw = mnx_need(e, B + "mlp.fc1.weight", {4 * C, C});
                const std::vector<float>* fb = mnx_need(e, B + "mlp.fc1.bias", {4 * C});
                if (!lw || !lb || !fw || !fb) return MNX_ERR_WEIGHTS;
                std::vector<float> wf((size_t)4 * C * C), bf((size_t)4 * C), cs((size_t)4 * C);
                for (int64_t n = 0; n < 4 * C; ++n) {
                    double bacc = (*fb)[n], sacc = 0.0;
                    for (int64_t k = 0; k < C; ++k) {
                        const float wv = (*fw)[(size_t)n * C + k];
                        const float wg = wv * (*lw)[k];
                        wf[(size_t)n * C + k] = wg;
                        sacc += (double)__bfloat162float(__float2bfloat16_rn(wg));
                        bacc += (double)wv * (double)(*lb)[k];
                    }
                    bf[n] = (float)bacc;
                    cs[n] = (float)sacc;
                }
                CN_TRY(cn_bf16_vec(e, wf, &w.fc1_w));
                CN_CUDA(e, mnx_upload(e, bf, &w.fc1_b));
                CN_CUDA(e, mnx_upload(e, cs, &w.fc1_colsum));
            }
            CN_TRY(cn_bf16(e, B + "mlp.fc2.weight", {C, 4 * C}, &w.fc2_w));
            CN_TRY(cn_f32(e, B + "mlp.fc2.bias", {C}, &w.fc2_b));
            CN_TRY(cn_f32(e, B + "gamma", {C}, &w.gamma));
        }
    }
    const size_t tok = (size_t)cfg.max_batch * (cfg.max_height / 4) * (cfg.max_width / 4);
    st->max_tokens = tok;
    void* p = nullptr;
    CN_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 128 * sizeof(float))); st->x0 = (float*)p;
    CN_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 128 * sizeof(float) / 2 + 1024)); st->x1 = (float*)p;
    CN_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 128 * 2)); st->abuf = (__nv_bfloat16*)p;
    CN_CUDA(e, mnx_dev_alloc_bytes(e, &p, tok * 512 * 2)); st->hbuf = (__nv_bfloat16*)p;
    if (const char* env = getenv("MNX_DW_TILE")) {
        int v[4];
        if (sscanf(env, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4)
            for (int i = 0; i < 4; ++i)
                if (v[i] == 8 || v[i] == 12) st->tile[i] = v[i];
    }
    if (const char* env = getenv("MNX_DW_SPLIT")) {     // kernel tuning only (A/B timing on the GPU box)
        int v[4];
        if (sscanf(env, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4)
            for (int i = 0; i < 4; ++i)
                if (v[i] >= 1 && v[i] <= (2 << i) && ((2 << i) % v[i]) == 0) st->split[i] = v[i];
    }
    size_t stat_floats = 0;
    for (int i = 0; i < 4; ++i) {
        const size_t need = (size_t)st->split[i] * (tok >> (2 * i)) * 2;
        if (need > stat_floats) stat_floats = need;
    }
    CN_CUDA(e, mnx_dev_alloc_bytes(e, &p, stat_floats * sizeof(float))); st->stats = (float*)p;
    return MNX_OK;
}

void convnext_destroy(ConvNextState* st) { delete st; }

static cudaError_t cn_gemm(const __nv_bfloat16* A, const __nv_bfloat16* W, long long M, int N, int K, int epi,
                           const float* bias, const float* gamma, void* out, cudaStream_t s, int cta_limit,
                           const float* colsum = nullptr, const float* ln_stats = nullptr, int ln_splits = 0) {
    GemmParams p{};
    p.cta_limit = cta_limit;
    p.A = A; p.W = W; p.M = (int)M; p.N = N; p.K = K; p.epilogue = epi; p.bias = bias; p.gamma = gamma; p.out = out;
    p.colsum = colsum; p.ln_stats = ln_stats; p.ln_splits = ln_splits; p.ln_eps = 1e-6f;
    return gemm_tc_launch(p, s);
}

template <int C, int TS>
static int launch_dwconv_t(mnx_engine* e, const float* x, int B, int H, int W, const CnBlockW& w, __nv_bfloat16* out,
                           float* stats, int split, cudaStream_t s) {
    constexpr int DW_TILE = TS, DW_IN = DwCfg<TS>::IN;
    CUtensorMap map, wmap;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        const cuuint32_t box[4] = {64, DW_IN, DW_IN, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_cn_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mnx_set_error(e, "cuTensorMapEncodeTiled failed for the dwconv input"); return MNX_ERR_CUDA; }
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)C, 49};
        const cuuint64_t strides[1] = {(cuuint64_t)C * 4};
        const cuuint32_t box[2] = {64, 49};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = g_cn_encode(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w.dw_w), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mnx_set_error(e, "cuTensorMapEncodeTiled failed for the dwconv weights"); return MNX_ERR_CUDA; }
    }
    const int tiles = ((H + DW_TILE - 1) / DW_TILE) * ((W + DW_TILE - 1) / DW_TILE);
    dwconv_stats_kernel<C, TS><<<dim3(tiles, B, split), 32 * DwCfg<TS>::WARPS, DwCfg<TS>::SMEM, s>>>(map, wmap, H, W, w.dw_b, out, stats);
    CN_CUDA(e, cudaGetLastError());
    return MNX_OK;
}
static int launch_dwconv(mnx_engine* e, const float* x, int B, int H, int W, int C, const CnBlockW& w, __nv_bfloat16* out,
                         float* stats, int split, int tile, cudaStream_t s) {
    if (tile == 12) {
        switch (C) {
            case 128: return launch_dwconv_t<128, 12>(e, x, B, H, W, w, out, stats, split, s);
            case 256: return launch_dwconv_t<256, 12>(e, x, B, H, W, w, out, stats, split, s);
            case 512: return launch_dwconv_t<512, 12>(e, x, B, H, W, w, out, stats, split, s);
            default: return launch_dwconv_t<1024, 12>(e, x, B, H, W, w, out, stats, split, s);
        }
    }
    switch (C) {
        case 128: return launch_dwconv_t<128, 8>(e, x, B, H, W, w, out, stats, split, s);
        case 256: return launch_dwconv_t<256, 8>(e, x, B, H, W, w, out, stats, split, s);
        case 512: return launch_dwconv_t<512, 8>(e, x, B, H, W, w, out, stats, split, s);
        default: return launch_dwconv_t<1024, 8>(e, x, B, H, W, w, out, stats, split, s);
    }
}

int convnext_forward(mnx_engine* e, ConvNextState* st, const float* images, int B, int H, int W, float* features,
                     cudaStream_t s, int* launches, int cta_limit) {
    st->cta_limit = cta_limit;
    if (H % 32 != 0 || W % 32 != 0) {
        mnx_set_error(e, "ConvNeXt-B path needs H and W to be multiples of 32");
        return MNX_ERR_INVALID;
    }
    int nl = 0;
    int Hc = H / 4, Wc = W / 4;
    if ((size_t)B * Hc * Wc > st->max_tokens) { mnx_set_error(e, "convnext workspace too small for this request"); return MNX_ERR_CAPACITY; }
    CN_CUDA(e, launch_patch_embed(images, B, H, W, st->stem_w, st->stem_b, st->stem_ln_w, st->stem_ln_b, 1e-6f, st->x0, s));
    ++nl;
    float* x = st->x0;
    float* x_other = st->x1;
    for (int stage = 0; stage < 4; ++stage) {
        const int C = 128 << stage;
        if (stage > 0) {
            const int Ci = C / 2, H2 = Hc / 2, W2 = Wc / 2;
            const long long M2 = (long long)B * H2 * W2;
            downsample_ln_kernel<<<(unsigned)((M2 + 7) / 8), 256, 0, s>>>(x, B, Hc, Wc, Ci, st->down[stage].ln_w,
                                                                         st->down[stage].ln_b, 1e-6f, st->abuf);
            CN_CUDA(e, cudaGetLastError()); ++nl;
            float* dst = (stage == 3) ? features : x_other;     // the last stage lives in the caller's buffer
            CN_CUDA(e, cn_gemm(st->abuf, st->down[stage].w, M2, C, 4 * Ci, GEMM_EPI_F32, st->down[stage].b, nullptr, dst, s, st->cta_limit)); ++nl;
            if (stage == 3) { x = features; } else { float* t = x; x = x_other; x_other = t; }
            Hc = H2; Wc = W2;
        }
        const long long M = (long long)B * Hc * Wc;
        for (int j = 0; j < CN_DEPTH[stage]; ++j) {
            const CnBlockW& w = st->blocks[stage][j];
            CN_TRY(launch_dwconv(e, x, B, Hc, Wc, C, w, st->abuf, st->stats, st->split[stage], st->tile[stage], s));
            ++nl;
            CN_CUDA(e, cn_gemm(st->abuf, w.fc1_w, M, 4 * C, C, GEMM_EPI_LNFOLD_GELU_BF16, w.fc1_b, nullptr, st->hbuf, s, st->cta_limit,
                               w.fc1_colsum, st->stats, st->split[stage])); ++nl;
            CN_CUDA(e, cn_gemm(st->hbuf, w.fc2_w, M, C, 4 * C, GEMM_EPI_RESADD_F32, w.fc2_b, w.gamma, x, s, st->cta_limit)); ++nl;
        }
    }
    st->last_B = B; st->last_H = H; st->last_W = W;
    *launches += nl;
    return MNX_OK;
}

// isolated timing of the dwconv (+ LayerNorm statistics) kernel at the shapes of stage (which - 101) of the last call
int convnext_time_kernel(mnx_engine* e, ConvNextState* st, int which, int iters, float* ms, cudaStream_t s) {
    const int stage = which - 101;
    if (stage < 0 || stage > 3 || st->last_B == 0) { mnx_set_error(e, "convnext timing: ids 101..104 after an encode"); return MNX_ERR_INVALID; }
    const int C = 128 << stage, B = st->last_B;
    const int Hc = st->last_H / (4 << stage), Wc = st->last_W / (4 << stage);
    const CnBlockW& w = st->blocks[stage][0];
    const float* x = (stage == 3) ? st->x0 : ((stage & 1) ? st->x1 : st->x0);   // any resident map of the right size
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = MNX_OK;
    for (int i = 0; i < 3 + iters && rc == MNX_OK; ++i) {
        if (i == 3) cudaEventRecord(e0, s);
        rc = launch_dwconv(e, x, B, Hc, Wc, C, w, st->abuf, st->stats, st->split[stage], st->tile[stage], s);
    }
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    *ms = t / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

}  // namespace mnx
