// tcgen05 / TMEM / TMA GEMM used by both encoders (every Linear of Swin-B and every 1x1 /
// patchify convolution of ConvNeXt-B):   D[M][N] = A[M][K](bf16) * W[N][K]^T(bf16), fp32 accumulate,
// with the consumer's elementwise work fused into the epilogue.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace mnx {

enum GemmEpilogue {
    GEMM_EPI_BF16 = 0,        // out_bf16[m][n] = acc + bias[n]
    GEMM_EPI_GELU_BF16 = 1,   // out_bf16[m][n] = gelu_erf(acc + bias[n])
    GEMM_EPI_RESADD_F32 = 2,  // resid[row_map ? row_map[m] : m][n] += (gamma ? gamma[n] : 1) * (acc + bias[n])
    GEMM_EPI_F32 = 3,         // out_f32[m][n] = acc + (bias ? bias[n] : 0)
    GEMM_EPI_LNFOLD_GELU_BF16 = 4,  // A holds UN-normalised rows x; the LayerNorm over K that precedes this Linear is applied
                              // here:  out_bf16[m][n] = gelu_erf(rstd[m] * (acc - mean[m] * colsum[n]) + bias[n]),  with
                              // W = weight * diag(ln_weight) (bf16), colsum[n] = sum_k W[n][k], bias = weight @ ln_bias + b,
                              // mean / rstd from ln_stats (per-row sum and sum of squares of x, `ln_splits` partials)
};

struct GemmParams {
    const __nv_bfloat16* A;   // [M][K] row-major (K contiguous)
    const __nv_bfloat16* W;   // [N][K] row-major (torch Linear layout)
    int M, N, K;              // N % 128 == 0, K % 64 == 0
    int epilogue;
    const float* bias;        // [N] or nullptr (GEMM_EPI_F32 only)
    const float* gamma;       // [N] or nullptr
    const int* row_map;       // [M] destination row (or -1 = drop) or nullptr
    void* out;                // bf16 [M][N] / fp32 [M][N] / fp32 residual stream [*][N]
    const float* colsum;      // [N]               (GEMM_EPI_LNFOLD_GELU_BF16)
    const float* ln_stats;    // [ln_splits][M][2] {sum, sum of squares} over a K / ln_splits channel range each
    int ln_splits;
    float ln_eps;
    int cta_limit;            // > 0: cap the persistent grid at this many CTAs (0 = one per SM): lets the encoder of the
                              // next batch run on the SMs that concurrently running persistent decode kernels leave free
};

// Launches the kernel on `s`.  Returns cudaErrorInvalidValue for unsupported shapes.
cudaError_t gemm_tc_launch(const GemmParams& p, cudaStream_t s);
// one-time per-device setup (driver entry point for tensor maps, smem opt-in)
cudaError_t gemm_tc_configure();

}  // namespace mnx
