// Encoder dispatch (by state-dict prefix: `transformer.` -> Swin-B, `cnn.` -> ConvNeXt-B).
#include "encoder.cuh"

namespace mnx {

int swin_finalize(mnx_engine* e, SwinState** st, const mnx_config& cfg);
int swin_forward(mnx_engine* e, SwinState* st, const float* images, int B, int H, int W, float* features,
                 cudaStream_t s, int* launches, int cta_limit);
int swin_time_kernel(mnx_engine* e, SwinState* st, int which, int iters, float* ms, cudaStream_t s);
void swin_destroy(SwinState* st);
int convnext_finalize(mnx_engine* e, ConvNextState** st, const mnx_config& cfg);
int convnext_forward(mnx_engine* e, ConvNextState* st, const float* images, int B, int H, int W, float* features,
                     cudaStream_t s, int* launches, int cta_limit);
void convnext_destroy(ConvNextState* st);
int convnext_time_kernel(mnx_engine* e, ConvNextState* st, int which, int iters, float* ms, cudaStream_t s);

int encoder_seq_len(int kind, int H, int W) {
    if (kind == MNX_ENCODER_SWIN_B) {
        // patch embed pads to a multiple of 4, every PatchMerging pads odd maps (transformers.py:408-411,320-322)
        int h = (H + 3) / 4, w = (W + 3) / 4;
        for (int i = 0; i < 3; ++i) { h = (h + 1) / 2; w = (w + 1) / 2; }
        return h * w;
    }
    return (H / 32) * (W / 32);
}

int encoder_finalize(mnx_engine* e, EncoderState& st, const mnx_config& cfg) {
    st.kind = cfg.encoder_kind;
    if (st.kind == MNX_ENCODER_SWIN_B) return swin_finalize(e, &st.swin, cfg);
    if (st.kind == MNX_ENCODER_CONVNEXT_B) return convnext_finalize(e, &st.cnx, cfg);
    mnx_set_error(e, "unknown encoder kind");
    return MNX_ERR_INVALID;
}

int encoder_forward(mnx_engine* e, EncoderState& st, const float* images, int B, int H, int W, float* features,
                    cudaStream_t s, int* launches) {
    if (st.kind == MNX_ENCODER_SWIN_B) return swin_forward(e, st.swin, images, B, H, W, features, s, launches, st.cta_limit);
    if (st.kind == MNX_ENCODER_CONVNEXT_B) return convnext_forward(e, st.cnx, images, B, H, W, features, s, launches, st.cta_limit);
    mnx_set_error(e, "unknown encoder kind");
    return MNX_ERR_INVALID;
}

int encoder_time_kernel(mnx_engine* e, EncoderState& st, int which, int iters, float* ms, cudaStream_t s) {
    if (st.kind == MNX_ENCODER_SWIN_B) return swin_time_kernel(e, st.swin, which, iters, ms, s);
    if (st.kind == MNX_ENCODER_CONVNEXT_B) return convnext_time_kernel(e, st.cnx, which, iters, ms, s);
    mnx_set_error(e, "no encoder on this handle");
    return MNX_ERR_INVALID;
}

void encoder_destroy(EncoderState& st) {
    if (st.swin) swin_destroy(st.swin);
    if (st.cnx) convnext_destroy(st.cnx);
    st.swin = nullptr;
    st.cnx = nullptr;
}

}  // namespace mnx
