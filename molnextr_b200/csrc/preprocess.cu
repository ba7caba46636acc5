// Device-side image preprocessing: the reference's inference transform
//   CropWhite(pad=50) -> Resize(384, 384, INTER_LINEAR) -> ToGray -> Normalize(ImageNet) -> ToTensorV2
// (MolNexTR/dataset.py:158-185 `get_transforms`, MolNexTR/data_aug.py:98-143 `CropWhite`; albumentations 1.1.0
// calls cv2.resize / cv2.cvtColor underneath) on raw RGB uint8 images of arbitrary size, bit-exact with
// OpenCV's 8-bit code paths:
//   * crop = tight bounding box of the pixels that differ from (255,255,255) (whole image if there are none),
//     then a constant white border of `pad` pixels -- never materialised: the resize reads a virtual image;
//   * cv2.resize INTER_LINEAR, CV_8UC3: coefficients in 11-bit fixed point (cvRound(coef * 2048)), source
//     coordinate fx = float((dx + 0.5) * scale - 0.5) computed in double, horizontal pass in int32, vertical
//     pass (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2; exact 2x decimation takes OpenCV's
//     area path ((a + b + c + d + 2) >> 2);
//   * RGB2GRAY in 15-bit fixed point: (9798 R + 19235 G + 3735 B + 16384) >> 15, replicated to 3 channels;
//   * (x - mean*255) * (1 / (std*255)) in fp32, one rounding per operation, constants supplied by the host.
// HBM-bound byte work: one read of the image for the bounding box, then 4 source pixels per output pixel.
#include <cuda_runtime.h>
#include <stdint.h>

namespace mnx {

#define PP_MAX_IMAGES 96

struct PpBatch {
    unsigned long long offset[PP_MAX_IMAGES];   // byte offset of image i in the packed buffer
    int h[PP_MAX_IMAGES], w[PP_MAX_IMAGES];
    int n;
};

// bbox[i] = {top, bottom (exclusive), left, right (exclusive)}; initialised to {INT_MAX, 0, INT_MAX, 0}
__global__ void pp_bbox_init_kernel(int* bbox, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { bbox[4 * i + 0] = 0x7fffffff; bbox[4 * i + 1] = 0; bbox[4 * i + 2] = 0x7fffffff; bbox[4 * i + 3] = 0; }
}

__global__ void __launch_bounds__(256) pp_bbox_kernel(const uint8_t* __restrict__ rgb, PpBatch b, int* __restrict__ bbox) {
    const int img = blockIdx.y;
    const int H = b.h[img], W = b.w[img];
    const uint8_t* src = rgb + b.offset[img];
    int top = 0x7fffffff, bot = 0, left = 0x7fffffff, right = 0;
    const long long npix = (long long)H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const uint8_t* px = src + 3 * p;
        if (px[0] != 255 || px[1] != 255 || px[2] != 255) {
            const int y = (int)(p / W), x = (int)(p - (long long)y * W);
            top = min(top, y); bot = max(bot, y + 1); left = min(left, x); right = max(right, x + 1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        top = min(top, __shfl_xor_sync(0xffffffffu, top, o)); bot = max(bot, __shfl_xor_sync(0xffffffffu, bot, o));
        left = min(left, __shfl_xor_sync(0xffffffffu, left, o)); right = max(right, __shfl_xor_sync(0xffffffffu, right, o));
    }
    if ((threadIdx.x & 31) == 0 && bot > 0) {
        atomicMin(&bbox[4 * img + 0], top); atomicMax(&bbox[4 * img + 1], bot);
        atomicMin(&bbox[4 * img + 2], left); atomicMax(&bbox[4 * img + 3], right);
    }
}

struct PpNorm { float mean255[3], inv_std255[3]; };

// source coordinate and 11-bit coefficients of cv2.resize INTER_LINEAR for destination index d
__device__ __forceinline__ void pp_coef(int d, int dn, int sn, bool clamp_edges, int* s, int* c0, int* c1) {
    const double inv = __ddiv_rn((double)dn, (double)sn);
    const double scale = __ddiv_rn(1.0, inv);
    float f = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
    int si = (int)floorf(f);
    f = __fsub_rn(f, (float)si);
    if (clamp_edges) {      // x only: resize.cpp zeroes the fraction at the borders; rows are clipped instead
        if (si < 0) { f = 0.f; si = 0; }
        if (si >= sn - 1) { f = 0.f; si = sn - 1; }
    }
    *s = si;
    *c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));    // cvRound: ties to even
    *c1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

__global__ void __launch_bounds__(256) pp_resize_kernel(const uint8_t* __restrict__ rgb, PpBatch b, const int* __restrict__ bbox,
                                                        int pad, int S, PpNorm nm, float* __restrict__ out) {
    const int img = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * S) return;
    const int dy = idx / S, dx = idx - dy * S;
    const int H = b.h[img], W = b.w[img];
    const uint8_t* src = rgb + b.offset[img];
    int top = bbox[4 * img + 0], bot = bbox[4 * img + 1], left = bbox[4 * img + 2], right = bbox[4 * img + 3];
    if (bot == 0) { top = 0; bot = H; left = 0; right = W; }        // all white: no crop (data_aug.py:113-114)
    const int ch = bot - top, cw = right - left;
    const int ph = ch + 2 * pad, pw = cw + 2 * pad;                 // the padded crop the reference resizes
    // pixel (y, x) of the padded crop; white outside the crop
    auto load = [&](int y, int x, int* r, int* g, int* bl) {
        y -= pad; x -= pad;
        if (y < 0 || y >= ch || x < 0 || x >= cw) { *r = *g = *bl = 255; return; }
        const uint8_t* p = src + 3 * ((size_t)(top + y) * W + left + x);
        *r = p[0]; *g = p[1]; *bl = p[2];
    };
    int R, G, B;
    if (pw == 2 * S && ph == 2 * S) {
        int r[4], g[4], bl[4];
        load(2 * dy, 2 * dx, &r[0], &g[0], &bl[0]); load(2 * dy, 2 * dx + 1, &r[1], &g[1], &bl[1]);
        load(2 * dy + 1, 2 * dx, &r[2], &g[2], &bl[2]); load(2 * dy + 1, 2 * dx + 1, &r[3], &g[3], &bl[3]);
        R = (r[0] + r[1] + r[2] + r[3] + 2) >> 2; G = (g[0] + g[1] + g[2] + g[3] + 2) >> 2; B = (bl[0] + bl[1] + bl[2] + bl[3] + 2) >> 2;
    } else {
        int sx, a0, a1, sy, b0, b1;
        pp_coef(dx, S, pw, true, &sx, &a0, &a1);
        pp_coef(dy, S, ph, false, &sy, &b0, &b1);
        const int x1 = min(sx + 1, pw - 1);
        const int y0 = min(max(sy, 0), ph - 1), y1 = min(max(sy + 1, 0), ph - 1);
        int r00, g00, b00, r01, g01, b01, r10, g10, b10, r11, g11, b11;
        load(y0, sx, &r00, &g00, &b00); load(y0, x1, &r01, &g01, &b01);
        load(y1, sx, &r10, &g10, &b10); load(y1, x1, &r11, &g11, &b11);
        auto lerp = [&](int p00, int p01, int p10, int p11) {
            const int h0 = p00 * a0 + p01 * a1, h1 = p10 * a0 + p11 * a1;
            return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        };
        R = lerp(r00, r01, r10, r11); G = lerp(g00, g01, g10, g11); B = lerp(b00, b01, b10, b11);
        R = min(max(R, 0), 255); G = min(max(G, 0), 255); B = min(max(B, 0), 255);
    }
    const int gray = (R * 9798 + G * 19235 + B * 3735 + (1 << 14)) >> 15;
    const float gf = (float)gray;
    float* o = out + ((size_t)img * 3) * S * S + (size_t)dy * S + dx;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[(size_t)c * S * S] = __fmul_rn(__fsub_rn(gf, nm.mean255[c]), nm.inv_std255[c]);
}

cudaError_t pp_run(const uint8_t* rgb, const unsigned long long* offsets, const int* hs, const int* ws, int n, int pad, int S,
                   const float* mean255, const float* inv_std255, int* bbox, float* out, cudaStream_t s, int* launches) {
    PpNorm nm;
    for (int c = 0; c < 3; ++c) { nm.mean255[c] = mean255[c]; nm.inv_std255[c] = inv_std255[c]; }
    for (int i0 = 0; i0 < n; i0 += PP_MAX_IMAGES) {
        PpBatch b{};
        b.n = n - i0 < PP_MAX_IMAGES ? n - i0 : PP_MAX_IMAGES;
        for (int i = 0; i < b.n; ++i) { b.offset[i] = offsets[i0 + i]; b.h[i] = hs[i0 + i]; b.w[i] = ws[i0 + i]; }
        pp_bbox_init_kernel<<<(b.n + 127) / 128, 128, 0, s>>>(bbox + 4 * i0, b.n);
        pp_bbox_kernel<<<dim3(64, b.n), 256, 0, s>>>(rgb, b, bbox + 4 * i0);
        pp_resize_kernel<<<dim3((S * S + 255) / 256, b.n), 256, 0, s>>>(rgb, b, bbox + 4 * i0, pad, S, nm,
                                                                       out + (size_t)i0 * 3 * S * S);
        *launches += 3;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace mnx
