// Device-side data structures of the autoregressive decoder (one engine handle owns one set).
#pragma once
#include "common.cuh"

namespace mnx {

// Repacked decoder weights (all fp32, GEMM operands stored K-major-transposed [K][N] so that a
// warp whose lanes are output columns reads them with unit stride).
struct DecLayerW {
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *lnf_w, *lnf_b;
    const float *wqkv_t, *bqkv;    // [256][768]  columns: q | k | v   (self-attention)
    const float *wo_s_t, *bo_s;    // [256][256]  self-attention final_linear
    const float *wq_c_t, *bq_c;    // [256][256]  context-attention linear_query
    const float *wo_c_t, *bo_c;    // [256][256]  context-attention final_linear
    const float *w1_t, *b1;        // [256][1024]
    const float *w2_t, *b2;        // [1024][256]
};

struct DecWeights {
    DecLayerW layer[MNX_DEC_L];
    const float *lnF_w, *lnF_b;    // final LayerNorm (eps 1e-6)
    const float *wout_t, *bout;    // [256][VPAD] output_layer, VPAD = 256 (zero padded)
    const float *emb;              // [V][256]
    const float *pe;               // [5000][256]
    const float *wenc_t, *benc;    // [1024][256] enc_trans_layer.0
    const float *wkv_c_t, *bkv_c;  // [256][L*512] context linear_keys | linear_values of every layer
    const float *we_a_t, *we_b_t, *be0;  // bond head first layer split: [256][256] x2, bias
    const float *we2, *be2;        // [7][256], [7]
};

// Greedy-search bookkeeping that lives on the device (no host round trips inside a step).
struct DecState {
    int n_alive;     // rows alive in the step being executed
    int step;        // index t of the step being executed
    int next_step;   // t of the next step (advanced by the embed kernel)
    int done;        // 1 once every row has finished
    int steps_run;   // number of steps that had at least one alive row
};

struct DecBuffers {
    DecState* st;
    int* alive;        // [2][B]   ordered list of alive original rows, ping-pong on step parity
    int* cur_tok;      // [B]      last chosen id per original row
    int* finished;     // [B]      set by the pick kernel
    float* xa;         // [B][256] residual stream (ping)
    float* xb;         // [B][256] residual stream (pong)
    float* q;          // [B][256] scaled query of the current attention
    float* part;       // [B][8][256] partial sums (context final_linear per head / W2 per k-slice)
    float* part2;      // [B][8][256] partial sums of the self-attention final_linear
    float* hbuf;       // [B][1024] FFN hidden
    float* selfK;      // [L][B][8][T][32]
    float* selfV;      // [L][B][8][T][32]
    float* crossK;     // [L][B][8][S][32]
    float* crossV;     // [L][B][8][S][32]
    float* membank;    // [B*S][256]
    int B, S, T;       // capacities of this call
    // outputs (caller-owned or engine-owned)
    int* ids;          // [B][T]
    int* lens;         // [B]
    float* logp;       // [B][T]
    float* hidden;     // [B][T][256]
};

struct Grammar {
    int vocab, offset, maxx, maxy, eos, sos, max_len;
};

}  // namespace mnx
