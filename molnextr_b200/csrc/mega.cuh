// Persistent cluster decode kernel (mega.cu): host-visible interface.
#pragma once
#include "decoder.cuh"

namespace mnx {

#define MG_TILE_FLOATS_H (256 * 32)
#define MG_TILES_PER_LAYER_H 14
#define MG_PARAM_FLOATS_H 1920
#define MG_GMAX_H 4      // rows per 8-CTA cluster (mega.cu)
#define MG16_GMAX_H 5    // rows per 16-CTA cluster (mega16.cu)
#define MG16S_GMAX_H 4   // rows per 16-CTA cluster, small-batch configuration (mega16s.cu)

struct MegaArgs {
    const float* wpack;
    const float* ppack;
    const float* wpack16;   // 16-CTA-cluster variant: [16][L*14 + 1][256*16]
    const float* ppack16;   // [16][L][1728]
    const float* wpackW;    // wide kernel (wide.cu): [8][L*14 + 1] slots of 8192 floats, see finalize_decoder
    int* ticket;            // cluster-order ticket counter (every cluster kernel), zeroed before every launch
    const float* finalp;
    const float* emb;
    const float* pe;
    float* selfK;
    float* selfV;
    const float* crossK;
    const float* crossV;
    int B, S, T, G;
    int row_base;           // wide kernel: first row of this launch (batches above 16 x the resident clusters run as several launches)
    int* ids;
    float* logp;
    float* hidden;
    int* lens;
    unsigned int* row_state;
    int* steps_run;
    long long* prof;        // optional [64] cycle stamps of cluster 0 / CTA 0 at step 100 (nullptr = off)
    Grammar g;
};

cudaError_t mega_configure(int* max_clusters);
cudaError_t mega_launch(const MegaArgs& a, int clusters, cudaStream_t s);
cudaError_t mega16_configure(int* max_clusters);
cudaError_t mega16_launch(const MegaArgs& a, int clusters, cudaStream_t s);
cudaError_t mega16s_configure(int* max_clusters);     // <= 28 rows: two 4-warp attention groups, G <= 4 (mega16s.cu)
cudaError_t mega16s_launch(const MegaArgs& a, int clusters, cudaStream_t s);
#define MGW_GMAX_H 16    // rows per 8-CTA cluster, throughput kernel (wide.cu)
#define MGW_MAX_KEYS_H 512   // keys one attention of wide.cu can score (16 sub-tiles of 32): needs T <= 513 and S <= 512
cudaError_t wide_configure(int* max_clusters);
cudaError_t wide_launch(const MegaArgs& a, int clusters, cudaStream_t s);

}  // namespace mnx
