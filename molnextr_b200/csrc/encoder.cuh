// Encoder side of the engine: Swin-B (the encoder the reference executes) and ConvNeXt-B (the
// encoder `north_star` names).  engine.cu sees only this interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <initializer_list>
#include <string>
#include <vector>

#include "../../include/molnextr_b200.h"

#define ATTN_MAXKEYS_HOST 1024   // cross-attention keys the decode kernel can score (S at 1024x1024)

struct mnx_engine;

// helpers exported by engine.cu to the encoder translation units
const std::vector<float>* mnx_need(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape);
const std::vector<int64_t>* mnx_need_i64(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape);
cudaError_t mnx_upload(mnx_engine* e, const std::vector<float>& h, const float** out);
cudaError_t mnx_upload_raw(mnx_engine* e, const void* h, size_t bytes, void** out);
cudaError_t mnx_dev_alloc_bytes(mnx_engine* e, void** p, size_t bytes);
void mnx_set_error(mnx_engine* e, const char* msg);

namespace mnx {

struct SwinState;
struct ConvNextState;

struct EncoderState {
    int kind = MNX_ENCODER_NONE;
    SwinState* swin = nullptr;
    ConvNextState* cnx = nullptr;
    int cta_limit = 0;      // cap of the persistent GEMM grids of this engine's encoder (0 = one CTA per SM)
};

int encoder_seq_len(int kind, int H, int W);
int encoder_finalize(mnx_engine* e, EncoderState& st, const mnx_config& cfg);
int encoder_forward(mnx_engine* e, EncoderState& st, const float* images, int B, int H, int W, float* features,
                    cudaStream_t s, int* launches);
int encoder_time_kernel(mnx_engine* e, EncoderState& st, int which, int iters, float* ms, cudaStream_t s);
void encoder_destroy(EncoderState& st);

}  // namespace mnx
