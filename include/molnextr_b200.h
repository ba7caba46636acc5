/*
 * molnextr_b200 -- C ABI of the B200-native MolNexTR inference engine.
 *
 * One engine handle per GPU.  The handle owns the repacked weights and every workspace
 * (KV caches, memory K/V, activations, CUDA graphs); the caller owns all input / output
 * buffers and passes plain device pointers plus the CUDA stream to run on.  No torch types
 * cross this boundary.  One call in flight per handle AND context (the reference decoder is not
 * re-entrant either: its KV cache is a module attribute, MolNexTR/models/decoder.py:287); see
 * mnx_reserve_contexts for keeping several batches in flight.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference repository root):
 *
 *   mnx_create / mnx_load_tensor / mnx_finalize_weights
 *        <- molnextr.__init__/_get_model + loading()          MolNexTR/model.py:40-48,83-95,17-28
 *           (state-dict keys are the reference's own; unlike `strict=False` there, a
 *            missing or unexpected key is an error here)
 *   mnx_preprocess        <- get_transforms(...)(image=...) + CropWhite
 *                            MolNexTR/dataset.py:158-185, MolNexTR/data_aug.py:98-143, MolNexTR/model.py:104
 *   mnx_encode            <- Encoder.forward                  MolNexTR/components.py:162-174
 *                            (Swin-B: MolNexTR/models/transformers.py:504-515;
 *                             ConvNeXt-B: timm forward_features, components.py:121-126)
 *   mnx_decode_greedy     <- TransformerDecoderAR.decode + GreedySearch
 *                            MolNexTR/components.py:253-334, MolNexTR/decoding/greedy_search.py:33-128
 *   mnx_decode_greedy_labels <- TransformerDecoderAR.decode(..., labels=...) "partial prediction"
 *                            MolNexTR/components.py:286-289,305,317-318,326-332,
 *                            MolNexTR/decoding/greedy_search.py:83-85
 *   mnx_decode_beam       <- TransformerDecoderAR.decode + BeamSearch (repaired; see below)
 *                            MolNexTR/components.py:253-334, MolNexTR/decoding/beam_search.py:84-190
 *   mnx_atom_indices      <- CharTokenizer.sequence_to_smiles (the `indices` output only)
 *                            MolNexTR/tokenization.py:464-515
 *   mnx_edges             <- GraphPredictor.forward + get_edge_prediction
 *                            MolNexTR/components.py:365-400, driver :470-484
 *   mnx_confidence        <- Decoder.decode with compute_confidence (atom_scores, average token score,
 *                            overall_score)             MolNexTR/components.py:456-469,485-491
 *   mnx_predict           <- `features, hiddens = self.encoder(images)` followed by
 *                            `self.decoder.decode(features, hiddens)`  MolNexTR/model.py:106-108
 *
 * All functions return MNX_OK (0) or a negative status; mnx_last_error() gives the message.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef MOLNEXTR_B200_H
#define MOLNEXTR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNX_OK 0
#define MNX_ERR_INVALID (-1)   /* bad argument / shape / state            */
#define MNX_ERR_CUDA (-2)      /* a CUDA runtime or driver call failed    */
#define MNX_ERR_WEIGHTS (-3)   /* missing / unexpected / mis-shaped tensor */
#define MNX_ERR_CAPACITY (-4)  /* request exceeds the sizes given at create */

#define MNX_ENCODER_NONE 0      /* decoder-only handle: mnx_encode unavailable */
#define MNX_ENCODER_SWIN_B 1    /* state-dict prefix `transformer.`            */
#define MNX_ENCODER_CONVNEXT_B 2 /* state-dict prefix `cnn.`                   */

#define MNX_DEC_DIM 256
#define MNX_EDGE_CLASSES 7

typedef struct mnx_engine mnx_engine;

typedef struct mnx_config {
    int32_t device;        /* CUDA device ordinal                                          */
    int32_t encoder_kind;  /* MNX_ENCODER_*                                                */
    int32_t max_batch;     /* images (= decoder rows) per call                             */
    int32_t max_height;    /* input image height bound (pixels)                            */
    int32_t max_width;     /* input image width bound                                      */
    int32_t max_len;       /* decode cap, 480 for chartok_coords (MolNexTR/utils.py:25)    */
    int32_t vocab;         /* 229 = 101 symbols + 64 x-bins + 64 y-bins                    */
    int32_t tok_offset;    /* first x-bin id (101)                                         */
    int32_t max_x;         /* number of x bins (64)                                        */
    int32_t max_y;         /* number of y bins (64)                                        */
    int32_t max_atoms;     /* bond-head capacity per image (<= max_len/3 = 160)            */
    int32_t encoder_dim;   /* 1024                                                         */
    int32_t max_beam;      /* largest beam width mnx_decode_beam may be called with (0 or 1:
                            * greedy only); decoder rows per call = max_batch * max_beam <= 5000 */
    /* per-id class bits for the device-side atom scan (bit0 symbol, bit1 atom, bit2 '[',
     * bit3 ']', bit4 'C', bit5 'l', bit6 'B', bit7 'r'); `vocab` entries, host pointer.   */
    const uint8_t* token_class;
} mnx_config;

/* lifecycle ------------------------------------------------------------------------- */
int mnx_create(const mnx_config* cfg, mnx_engine** out);
int mnx_destroy(mnx_engine* e);
/* message of the last failure on this handle (or of the last failed mnx_create if e==NULL) */
const char* mnx_last_error(const mnx_engine* e);

/* weights: one call per state-dict entry, `name` = "encoder.<key>" or "decoder.<key>" with the
 * reference's own keys; data is host fp32 (or int64 for index buffers, which are verified
 * against the recomputed table and otherwise ignored). */
int mnx_load_tensor(mnx_engine* e, const char* name, const void* host_data,
                    const int64_t* shape, int32_t ndim, int32_t is_int64);
int mnx_finalize_weights(mnx_engine* e);

/* upstream of the hot path ------------------------------ */
/* The reference's inference transform on the device, bit-exact with its OpenCV 8-bit code paths:
 * CropWhite(pad) -> Resize(out_size, out_size, INTER_LINEAR) -> ToGray -> Normalize -> CHW fp32
 * (MolNexTR/dataset.py:158-185 get_transforms, MolNexTR/data_aug.py:98-143 CropWhite, applied per
 * image at MolNexTR/model.py:104).
 * rgb        device uint8: B packed HxWx3 RGB images, image i at byte offset offsets[i]
 * offsets / heights / widths   HOST arrays of length B
 * mean255, inv_std255          HOST float[3]: mean*255 and 1/(std*255) exactly as the caller's numpy
 *                              float32 arithmetic produced them (the kernel does (g - m) * d in fp32)
 * images     device fp32 (B, 3, out_size, out_size): the tensor mnx_encode / mnx_predict take */
int mnx_preprocess(mnx_engine* e, const uint8_t* rgb, const int64_t* offsets, const int32_t* heights,
                   const int32_t* widths, int32_t B, int32_t pad, int32_t out_size, const float* mean255,
                   const float* inv_std255, float* images, void* cuda_stream);

/* hot path --------------------------------------------------------------------------- */
/* images: device fp32 NCHW (B,3,H,W), already normalised.  features: device fp32
 * (B, S, encoder_dim) with S = ceil(H/32)*ceil(W/32) (Swin) or (H/32)*(W/32) (ConvNeXt). */
int mnx_encode(mnx_engine* e, const float* images, int32_t B, int32_t H, int32_t W,
               float* features, void* cuda_stream);

/* Greedy decode of B rows against their (B,S,encoder_dim) feature maps.
 * ids        int32 (B, max_len)  chosen ids without <sos>, including <eos>, 0-padded
 * lens       int32 (B)           number of valid ids per row
 * token_logp fp32  (B, max_len)  masked log-prob of each chosen id
 * hidden     fp32  (B, max_len, 256) final-LayerNorm output per step (may be NULL: the
 *                                engine then keeps it internally for mnx_edges)
 * Row r of the alive batch receives positional encoding pe[rank of r among alive rows],
 * exactly as the reference does. */
int mnx_decode_greedy(mnx_engine* e, const float* features, int32_t B, int32_t S,
                      int32_t* ids, int32_t* lens, float* token_logp, float* hidden,
                      void* cuda_stream);

/* Greedy decode with part of every sequence GIVEN ("partial prediction", components.py:256-257).
 * labels int32 (B, n_labels), device memory: column 0 is <sos>; a position holding MASK_ID (4,
 * tokenization.py:13) is left to the model, any other value is teacher-forced as the INPUT of
 * that step (components.py:286-289) and keys the grammar mask (:301-303); a row finishes when
 * its NEXT label is <eos> (greedy_search.py:83-85; the model's own <eos> ends it only beyond the
 * labels) or at max_len.  Rows leave the batch as in mnx_decode_greedy (row-rank PE rule); labels
 * are addressed by original row, which is what labels.index_select (:317-318) keeps true in the
 * reference.  On return ids/lens hold the merged result of :326-332:
 *   lens[i] = min(len(pred_i), n_labels - 1), ids[i][j] = labels[i][1+j] unless that is MASK_ID;
 * token_logp / hidden keep the model's own picks over the full decoded length, as in the reference.
 * MNX_ERR_INVALID if the decode runs more steps than labels has columns (the reference raises
 * IndexError at components.py:287).  Runs on the multi-kernel path, context 0. */
int mnx_decode_greedy_labels(mnx_engine* e, const float* features, int32_t B, int32_t S,
                             const int32_t* labels, int32_t n_labels,
                             int32_t* ids, int32_t* lens, float* token_logp, float* hidden,
                             void* cuda_stream);

/* Beam-search decode of B images with `beam` hypotheses each (n_best <= beam returned, best first).
 * Replaces TransformerDecoderAR.decode with BeamSearch (MolNexTR/components.py:253-334,
 * MolNexTR/decoding/beam_search.py:84-190).  That branch cannot execute in the reference
 * (SURVEY.md F4); the semantics implemented are the repaired ones spelled out in
 * oracle/restate.py beam_decode ("parity unpinned": checked against that oracle only).
 * ids        int32 (B, n_best, max_len)  ids without <sos>, including <eos>, 0-padded
 * lens       int32 (B, n_best)
 * scores     fp32  (B, n_best)           length-normalised log score BeamSearch files a hypothesis under
 * token_logp fp32  (B, n_best, max_len)  masked log-prob of each chosen id
 * hidden     fp32  (B, max_len, 256)     final-LayerNorm outputs of each image's BEST hypothesis
 *                                        (may be NULL: kept internally for mnx_edges)
 * Any output pointer may be NULL.  Afterwards mnx_atom_indices(ids = lens = NULL) and
 * mnx_edges(hidden = NULL) operate on each image's best hypothesis (what Decoder.decode uses,
 * MolNexTR/components.py:455,477).
 * Rows are image-major, beam-minor; row r of the alive batch receives pe[r]. */
int mnx_decode_beam(mnx_engine* e, const float* features, int32_t B, int32_t S, int32_t beam, int32_t n_best,
                    int32_t* ids, int32_t* lens, float* scores, float* token_logp, float* hidden,
                    void* cuda_stream);

/* atom_idx int32 (B, max_atoms) positions (into ids) right after each atom's Y token,
 * n_atoms int32 (B).  ids and lens may both be NULL: the ids / lengths of the last decode on
 * this handle (greedy result, or best beam hypothesis) are scanned. */
int mnx_atom_indices(mnx_engine* e, const int32_t* ids, const int32_t* lens, int32_t B,
                     int32_t* atom_idx, int32_t* n_atoms, void* cuda_stream);

/* edges uint8 (B, max_atoms, max_atoms): argmax class after the reference's symmetrisation;
 * edge_score fp32 (B, max_atoms, max_atoms) max symmetrised probability (may be NULL).
 * hidden may be NULL to use the engine-internal copy from the last mnx_decode_greedy. */
int mnx_edges(mnx_engine* e, const float* hidden, const int32_t* atom_idx, const int32_t* n_atoms,
              int32_t B, uint8_t* edges, float* edge_score, void* cuda_stream);

/* Confidence outputs of Decoder.decode(compute_confidence=True) (MolNexTR/components.py:456-469,485-491) from the results
 * of the calls above: ids / lens / token_logp of a greedy (or best-beam) decode, edge_score of mnx_edges.
 * atom_scores   fp32 (B, max_atoms)  geometric mean of the probabilities of the atom's symbol tokens
 * seq_score     fp32 (B)             exp(mean(token log-prob)) = the reference's `scores` / average_token_score
 * overall_score fp64 (B)             seq_score * sqrt(prod(edge_score[:k, :k])) -- fp64 like the reference's numpy product,
 *                                    including its underflow to 0 for large molecules */
int mnx_confidence(mnx_engine* e, const int32_t* ids, const int32_t* lens, const float* token_logp, int32_t B,
                   const float* edge_score, float* atom_scores, float* seq_score, double* overall_score,
                   void* cuda_stream);

/* encoder -> decode -> atom scan -> bond head in one call, device pointers in and out. */
int mnx_predict(mnx_engine* e, const float* images, int32_t B, int32_t H, int32_t W,
                int32_t* ids, int32_t* lens, float* token_logp,
                int32_t* atom_idx, int32_t* n_atoms, uint8_t* edges, void* cuda_stream);

/* Same, with HOST buffers (pinned or pageable): copies in, runs, copies out, synchronises. */
int mnx_predict_host(mnx_engine* e, const float* images_host, int32_t B, int32_t H, int32_t W,
                     int32_t* ids_host, int32_t* lens_host, float* token_logp_host,
                     int32_t* atom_idx_host, int32_t* n_atoms_host, uint8_t* edges_host);

/* pipelining across batches -------------------------------------------------------------- */
/* The reference processes one mini-batch at a time (`for idx in range(0, len(input_images), batch_size)`,
 * MolNexTR/model.py:102-109; eval loop main.py:273-293).  The cluster decode kernels are single persistent
 * launches that never synchronise the host, so a caller with more batches ready can keep several in flight:
 *
 *   mnx_reserve_contexts(e, n)  allocate n complete sets of per-call device buffers (KV caches, memory-bank
 *                               K/V, ids / hidden / bond-head scratch).  Context 0 exists after finalize.
 *   mnx_set_context(e, i)       every later call on this handle enqueues its work on context i's buffers.
 *                               Calls that use different contexts may run concurrently on different streams;
 *                               two calls on the SAME context must be ordered by the caller (same stream).
 *                               The encoder has one activation workspace: mnx_encode calls must be ordered
 *                               among themselves (one encoder stream).
 *   mnx_set_decode_path(e, p)   0 = automatic (lowest single-batch latency: 16-CTA clusters of <= 5 rows on
 *                               ~112 SMs up to 35 rows, 8-CTA clusters up to 60 rows, the throughput kernel above --
 *                               as consecutive launches once a batch needs more clusters than are co-resident, 240
 *                               rows per launch on B200 -- and the multi-kernel graph path for S > 512 memory
 *                               positions), 1 = multi-kernel graph path, 2 = 8-CTA clusters of <= 4 rows,
 *                               3 = 16-CTA clusters, 6 = throughput kernel (8-CTA clusters of <= 16 rows: a
 *                               batch of 32 occupies 16 SMs, so ~4-8 batches decode side by side with the
 *                               encoder of the next ones).  Results are identical on every path.
 *   mnx_set_wide_rows(e, r)     throughput kernel only: rows per 8-CTA cluster (0 = 16).  Fewer rows per cluster spread
 *                               a batch over more SMs (bs 32: r = 16 -> 16 SMs, 8 -> 32 SMs, 4 -> 64 SMs) and shorten
 *                               its decode: used for the last batches of a run, when SMs would otherwise idle.
 *   mnx_set_encoder_cta_limit(e, n)  cap the persistent grid of THIS handle's encoder GEMMs at n CTAs
 *                               (0 = one per SM) so that they fit on the SMs running decode kernels leave free
 *                               instead of queueing behind them.
 * See Engine.predict_pipelined for the host-side schedule. */
int mnx_reserve_contexts(mnx_engine* e, int32_t n);
int mnx_set_context(mnx_engine* e, int32_t i);
int mnx_set_decode_path(mnx_engine* e, int32_t path);
int mnx_set_wide_rows(mnx_engine* e, int32_t rows);
int mnx_set_encoder_cta_limit(mnx_engine* e, int32_t n);

/* introspection ------------------------------------------------------------------------ */
/* number of kernel launches issued by this handle since creation (graph nodes counted
 * per replay); used by bench.py for `gpu_launches`. */
int64_t mnx_launch_count(const mnx_engine* e);
/* selections of the last mnx_decode_beam, for parity tests: trace_host int32 (max_len, B, 8),
 * entry [t][b][j] = flat index (parent beam * vocab + token) chosen as new beam j of image b at
 * step t, -1 where image b was no longer decoded.  Synchronises the device. */
#define MNX_MAX_BEAM 8
int mnx_beam_trace(mnx_engine* e, int32_t* trace_host, int32_t B);
/* steps executed by the last decode (<= max_len); waits for it if it is still running */
int32_t mnx_last_decode_steps(const mnx_engine* e);
/* time one internal phase in isolation for the roofline report: fills ms with the mean
 * device time of `iters` launches of kernel `which` on the shapes of the last call
 * (see DESIGN.md for the ids).  Returns MNX_ERR_INVALID for unknown ids. */
int mnx_time_kernel(mnx_engine* e, int32_t which, int32_t iters, float* ms, void* cuda_stream);

/* stand-alone GEMM entry used by the parity tests of the tensor-core kernel:
 * C[M,N] (fp32) = A[M,K] (fp32, rounded to bf16) * W[N,K]^T (fp32, rounded to bf16) + bias */
int mnx_test_gemm_bf16(const float* A, const float* W, const float* bias, float* C,
                       int32_t M, int32_t N, int32_t K, int32_t epilogue, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MOLNEXTR_B200_H */
