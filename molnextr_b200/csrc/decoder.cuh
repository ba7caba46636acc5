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
    int n_img;       // beam search: images alive in the step being executed (n_alive = n_img * beam)
    int lab_len;     // partial-label decoding (components.py:286-289): columns of DecBuffers::labels in use, 0 = none
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
    // partial-label decoding (TransformerDecoderAR.decode(labels=...), components.py:286-289,305,326-332): given
    // tokens by ORIGINAL row (so the reference's labels.index_select on compaction, :317-318, is implicit);
    // MNX_MASK_ID = position left to the model.  Read only while st->lab_len > 0.
    const int* labels; // [B][T+1]
};
#define MNX_MASK_ID 4   // tokenization.py:13

// Beam-search state (decoding/beam_search.py, repaired as described in oracle/restate.py
// beam_decode).  A *slot* is a physical decoder row: slot = image * beam + k.  Hypotheses move
// between slots every step; instead of permuting the self-attention K/V cache (what
// `map_state(index_select)` does in OpenNMT) every slot keeps an ancestry row anc[slot][t'] =
// the slot whose K/V (and hidden state) at position t' belongs to this hypothesis.
#define MNX_MAX_BEAM 8
struct BeamBuffers {
    int beam, n_best, n_img0;   // beam width, hypotheses returned per image, images of this call
    int* alive_img;     // [2][n_img0]  ordered list of alive images, ping-pong on step parity
    int* img_done;      // [n_img0]
    int* top_fin;       // [n_img0]     top_beam_finished
    float* lp;          // [R][256]     masked log-probs of the current step, by rank
    float* cum;         // [2][R]       topk_log_probs, by slot, ping-pong
    int* anc;           // [2][R][T]    ancestry rows, ping-pong
    int* hist_ids;      // [2][R][T]    alive_seq without <sos>, ping-pong
    float* hist_logp;   // [2][R][T]    masked log-prob of each chosen token, ping-pong
    // finished hypotheses kept per image: the best n_best so far, stable in insertion order
    int* hyp_count;     // [n_img0]     hypotheses stored so far (all of them, not only the kept ones)
    int* hyp_order;     // [n_img0][MNX_MAX_BEAM]  rank -> storage index
    float* hyp_score;   // [n_img0][MNX_MAX_BEAM]  by storage index
    int* hyp_len;       // [n_img0][MNX_MAX_BEAM]
    int* hyp_ids;       // [n_img0][MNX_MAX_BEAM][T]
    float* hyp_logp;    // [n_img0][MNX_MAX_BEAM][T]
    int* hyp_anc;       // [n_img0][MNX_MAX_BEAM][T]
    int* trace;         // [T][n_img0][MNX_MAX_BEAM] flat index (beam * V + token) selected per step, -1 = not run
};

struct Grammar {
    int vocab, offset, maxx, maxy, eos, sos, max_len;
};

}  // namespace mnx
