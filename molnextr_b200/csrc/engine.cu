// C ABI of the engine (include/molnextr_b200.h): handle lifetime, weight repacking, workspaces,
// CUDA-graph management of the decode loop, and the calls that chain encoder -> decoder -> bond
// head.  Host logic only; every arithmetic kernel lives in decoder.cu / gemm_tc.cu / swin.cu /
// convnext.cu.
#include "../../include/molnextr_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "decoder.cuh"
#include "encoder.cuh"
#include "mega.cuh"

namespace mnx {
// decoder.cu
cudaError_t dec_configure();
cudaError_t dec_confidence(const int* ids, const int* lens, const float* logp, int B, int T, const uint8_t* cls, const Grammar& g,
                           int max_atoms, const float* edge_score, float* atom_scores, float* seq_score, double* overall, cudaStream_t s);
cudaError_t dec_set_label_len(const DecBuffers& b, int lab_len, cudaStream_t s);
cudaError_t dec_label_merge(const DecBuffers& b, int lab_len, cudaStream_t s);
int dec_launch_step(const DecBuffers& b, const DecWeights& w, const Grammar& g, const BeamBuffers* bm, cudaStream_t s,
                    cudaError_t* err);
cudaError_t dec_beam_init(const DecBuffers& b, const BeamBuffers& bm, cudaStream_t s);
cudaError_t dec_beam_finalize(const DecBuffers& b, const BeamBuffers& bm, int* ids, int* lens, float* scores, float* logp,
                              int* best_ids, int* best_lens, float* best_logp, float* best_hidden, cudaStream_t s);
cudaError_t dec_precompute(const DecBuffers& b, const DecWeights& w, const float* features, int enc_dim,
                           cudaStream_t s, int* launches);
cudaError_t dec_atom_scan(const int* ids, const int* lens, int B, int T, const uint8_t* cls, const Grammar& g,
                          int max_atoms, int* atom_idx, int* n_atoms, cudaStream_t s);
cudaError_t dec_edges(const float* hidden, const int* atom_idx, const int* n_atoms, int B, int T, int max_atoms,
                      const DecWeights& w, float* hg, float* AB, float* prob, uint8_t* edges, float* score,
                      cudaStream_t s, int* launches);
cudaError_t dec_time_kernel(int which, int iters, const DecBuffers& b, const DecWeights& w, const Grammar& g,
                            int step, float* ms, cudaStream_t s);
// preprocess.cu
cudaError_t pp_run(const uint8_t* rgb, const unsigned long long* offsets, const int* hs, const int* ws, int n, int pad, int S,
                   const float* mean255, const float* inv_std255, int* bbox, float* out, cudaStream_t s, int* launches);
}  // namespace mnx

using namespace mnx;

static thread_local std::string g_create_error;

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> f;
    std::vector<int64_t> i;
    bool used = false;
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : shape) n *= d;
        return n;
    }
};

struct CtxPtrs {
    DecState* st;
    int *alive, *cur_tok, *finished;
    float *xa, *xb, *q, *part, *part2, *hbuf;
    float *selfK, *selfV, *crossK, *crossV, *membank;
    int *ids, *lens;
    float *logp, *hidden;
    unsigned int* row_state;
    int *steps_run_dev, *ticket;
    float *hg, *AB, *prob, *features;
    int *atom_idx, *n_atoms;
    uint8_t* edges;
};

struct mnx_engine {
    mnx_config cfg{};
    std::string err;
    std::map<std::string, HostTensor> host_w;
    bool finalized = false;
    std::vector<void*> allocs;
    std::vector<uint8_t> cls_host;
    uint8_t* d_cls = nullptr;
    DecWeights dw{};
    Grammar g{};
    int S_max = 0;
    // decoder workspaces (sized for max_batch / S_max / max_len)
    DecState* st = nullptr;
    int *alive = nullptr, *cur_tok = nullptr, *finished = nullptr;
    int* labels = nullptr;   // [max_batch][max_len + 1] given tokens of a partial-label decode (context 0, graph path)
    float *xa = nullptr, *xb = nullptr, *q = nullptr, *part = nullptr, *part2 = nullptr, *hbuf = nullptr;
    float *selfK = nullptr, *selfV = nullptr, *crossK = nullptr, *crossV = nullptr, *membank = nullptr;
    int *ids = nullptr, *lens = nullptr;
    float *logp = nullptr, *hidden = nullptr;
    // persistent cluster decode kernel (mega.cu)
    const float *wpack = nullptr, *ppack = nullptr, *finalp = nullptr, *wpack16 = nullptr, *ppack16 = nullptr;
    int max_clusters16 = 0, max_clusters16s = 0;
    const float* wpackW = nullptr;       // wide.cu weight slots
    int max_clusters_w = 0;
    int* ticket = nullptr;               // wide.cu cluster-order ticket
    unsigned int* row_state = nullptr;
    int* steps_run_dev = nullptr;
    long long* prof_dev = nullptr;
    int max_clusters = 0;
    int decode_path = 0;   // 0 auto, 1 force multi-kernel graph path, 2 force 8-CTA cluster kernel, 3 force 16-CTA cluster kernel,
                           // 6 force the throughput kernel (wide.cu: 8-CTA clusters of <= 16 rows)
    bool time_launches = false;    // bench instrumentation: CUDA events around every cluster decode launch (mnx_time_kernel 1005 / 1006)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> launch_events;
    int wide_rows = 0;             // throughput kernel: rows per cluster (0 = MGW_GMAX_H); fewer rows = more SMs per batch, lower latency
    bool decode_profile = false;   // MNX_DECODE_PROFILE, read once at create
    // decode contexts: complete sets of per-call device buffers, so that several batches can be in flight on
    // different streams (Engine.predict_pipelined).  Context 0 is allocated at finalize; the flat pointer fields
    // of this struct always hold the CURRENT context (mnx_set_context copies a saved set over them).
    std::vector<struct CtxPtrs> ctxs;
    int cur_ctx = 0;
    int num_sms = 0;
    // bond head
    float *hg = nullptr, *AB = nullptr, *prob = nullptr;
    // predict-path staging
    float* features = nullptr;
    float* images = nullptr;
    int *atom_idx = nullptr, *n_atoms = nullptr;
    uint8_t* edges = nullptr;
    int* pp_bbox = nullptr;   // [max_batch][4] crop boxes of mnx_preprocess
    // graph of STEPS_PER_GRAPH decode steps for one (B, S)
    cudaGraphExec_t graph = nullptr;
    int graph_B = -1, graph_S = -1, graph_nodes = 0;
    // beam search (graph path only): second graph keyed on (images, S, beam)
    BeamBuffers bm{};
    float* hid_best = nullptr;           // [max_batch][T][256] hidden states of each image's best hypothesis
    const float* edge_hidden = nullptr;  // what mnx_edges reads when the caller passes hidden == NULL
    cudaGraphExec_t graph_beam = nullptr;
    int gb_B = -1, gb_S = -1, gb_K = -1, gb_NB = -1, gb_nodes = 0;
    int last_beam_B = 0;
    cudaEvent_t ev_last = nullptr;       // recorded behind the work of every entry point on the caller's stream: mnx_predict_host
                                         // (which runs on cap_stream) orders itself after it, so that calls on one handle from
                                         // different streams never overlap on the shared workspaces
    cudaStream_t cap_stream = nullptr;   // engine-owned non-blocking stream: graph capture and mnx_predict_host
    int* h_done = nullptr;   // pinned
    int64_t launches = 0;
    int last_steps = 0;
    bool steps_pending = false;   // cluster paths return without synchronising: steps are read back on demand
    int last_B = 0, last_S = 0, last_path = 0;
    EncoderState enc{};
};

static const int STEPS_PER_GRAPH = 16;

static int fail(mnx_engine* e, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (e) e->err = buf; else g_create_error = buf;
    return code;
}
#define CUDA_TRY(e, x)                                                                         \
    do {                                                                                       \
        cudaError_t _c = (x);                                                                  \
        if (_c != cudaSuccess)                                                                 \
            return fail(e, MNX_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(_c), __FILE__, __LINE__); \
    } while (0)

// Entry points run on the engine's device and put the caller's current device back on return (a process that drives
// several engines, or torch on another GPU, keeps its own current device).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); else if (err == cudaSuccess) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
static void mark_work(mnx_engine* e, cudaStream_t s) {
    if (e->ev_last) cudaEventRecord(e->ev_last, s);
}
struct WorkMark {
    mnx_engine* e;
    cudaStream_t s;
    WorkMark(mnx_engine* e_, cudaStream_t s_) : e(e_), s(s_) {}
    ~WorkMark() { mark_work(e, s); }
};
#define ON_ENGINE_DEVICE(e)                 \
    DeviceGuard _dev_guard((e)->cfg.device); \
    CUDA_TRY(e, _dev_guard.err)

template <typename T>
static cudaError_t dev_alloc(mnx_engine* e, T** p, size_t count) {
    void* v = nullptr;
    cudaError_t c = cudaMalloc(&v, count * sizeof(T) + 256);
    if (c != cudaSuccess) return c;
    e->allocs.push_back(v);
    *p = reinterpret_cast<T*>(v);
    return cudaSuccess;
}

// upload a host fp32 array to a fresh device buffer
cudaError_t mnx_upload(mnx_engine* e, const std::vector<float>& h, const float** out) {
    float* d = nullptr;
    cudaError_t c = dev_alloc(e, &d, h.size());
    if (c != cudaSuccess) return c;
    c = cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
    *out = d;
    return c;
}
cudaError_t mnx_upload_raw(mnx_engine* e, const void* h, size_t bytes, void** out) {
    uint8_t* d = nullptr;
    cudaError_t c = dev_alloc(e, &d, bytes);
    if (c != cudaSuccess) return c;
    c = cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
    *out = d;
    return c;
}
cudaError_t mnx_dev_alloc_bytes(mnx_engine* e, void** p, size_t bytes) {
    uint8_t* d = nullptr;
    cudaError_t c = dev_alloc(e, &d, bytes);
    *p = d;
    return c;
}

extern "C" const char* mnx_last_error(const mnx_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" int mnx_create(const mnx_config* cfg, mnx_engine** out) {
    if (!cfg || !out) return fail(nullptr, MNX_ERR_INVALID, "mnx_create: null argument");
    if (cfg->max_batch < 1 || cfg->max_batch > 5000)
        return fail(nullptr, MNX_ERR_INVALID, "max_batch must be in [1,5000] (positional-encoding table has 5000 rows)");
    if (cfg->vocab < 1 || cfg->vocab > 256 || cfg->max_len < 1 || cfg->max_len % STEPS_PER_GRAPH != 0)
        return fail(nullptr, MNX_ERR_INVALID, "vocab must be <=256 and max_len a multiple of %d", STEPS_PER_GRAPH);
    if (cfg->tok_offset + cfg->max_x + cfg->max_y != cfg->vocab)
        return fail(nullptr, MNX_ERR_INVALID, "vocab != tok_offset + max_x + max_y");
    if (!cfg->token_class) return fail(nullptr, MNX_ERR_INVALID, "token_class table is required");
    if (cfg->max_atoms < 1 || cfg->max_atoms * 3 > cfg->max_len)
        return fail(nullptr, MNX_ERR_INVALID, "max_atoms must be in [1, max_len/3]");
    if (cfg->encoder_dim % 16 != 0) return fail(nullptr, MNX_ERR_INVALID, "encoder_dim must be a multiple of 16");
    if (cfg->max_beam < 0 || cfg->max_beam > MNX_MAX_BEAM)
        return fail(nullptr, MNX_ERR_INVALID, "max_beam must be in [0,%d]", MNX_MAX_BEAM);
    if ((int64_t)cfg->max_batch * (cfg->max_beam > 1 ? cfg->max_beam : 1) > 5000)
        return fail(nullptr, MNX_ERR_INVALID, "max_batch * max_beam must be <= 5000 (positional-encoding table has 5000 rows)");
    if (cfg->max_beam > 1 && cfg->max_len > 512) return fail(nullptr, MNX_ERR_INVALID, "beam search supports max_len <= 512");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, MNX_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MNX_ERR_INVALID, "bad device ordinal %d", cfg->device);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, MNX_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", cfg->device,
                    prop.major, prop.minor);
    mnx_engine* e = new mnx_engine();
    e->cfg = *cfg;
    e->cls_host.assign(cfg->token_class, cfg->token_class + cfg->vocab);
    e->cfg.token_class = nullptr;
    e->g = Grammar{cfg->vocab, cfg->tok_offset, cfg->max_x, cfg->max_y, /*eos*/ 2, /*sos*/ 1, cfg->max_len};
    const int hs = (cfg->max_height + 31) / 32, ws = (cfg->max_width + 31) / 32;
    e->S_max = hs * ws;
    if (e->S_max < 1 || e->S_max > ATTN_MAXKEYS_HOST) {
        delete e;
        return fail(nullptr, MNX_ERR_INVALID, "image bound gives %d memory positions; supported range is [1,%d]", hs * ws,
                    ATTN_MAXKEYS_HOST);
    }
    DeviceGuard guard(cfg->device);
    cudaError_t c = guard.err;
    if (c == cudaSuccess) c = dec_configure();
    if (c == cudaSuccess) c = mega_configure(&e->max_clusters);
    if (c == cudaSuccess) c = mega16_configure(&e->max_clusters16);
    if (c == cudaSuccess) c = mega16s_configure(&e->max_clusters16s);
    if (c == cudaSuccess) c = wide_configure(&e->max_clusters_w);
    e->num_sms = prop.multiProcessorCount;
    e->decode_profile = getenv("MNX_DECODE_PROFILE") != nullptr;
    if (const char* env = getenv("MNX_DECODE_PATH")) {
        if (!strcmp(env, "graph")) e->decode_path = 1;
        else if (!strcmp(env, "cluster")) e->decode_path = 2;
        else if (!strcmp(env, "cluster16")) e->decode_path = 3;
        else if (!strcmp(env, "wide")) e->decode_path = 6;
    }
    if (c == cudaSuccess) c = cudaMallocHost(&e->h_done, sizeof(int));
    if (c == cudaSuccess) c = cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking);
    if (c == cudaSuccess) c = cudaEventCreateWithFlags(&e->ev_last, cudaEventDisableTiming);
    if (c != cudaSuccess) {
        std::string m = cudaGetErrorString(c);
        delete e;
        return fail(nullptr, MNX_ERR_CUDA, "device setup failed: %s", m.c_str());
    }
    *out = e;
    return MNX_OK;
}

extern "C" int mnx_destroy(mnx_engine* e) {
    if (!e) return MNX_OK;
    DeviceGuard guard(e->cfg.device);
    cudaDeviceSynchronize();
    if (e->graph) cudaGraphExecDestroy(e->graph);
    if (e->graph_beam) cudaGraphExecDestroy(e->graph_beam);
    encoder_destroy(e->enc);
    for (void* p : e->allocs) cudaFree(p);
    if (e->h_done) cudaFreeHost(e->h_done);
    if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
    if (e->ev_last) cudaEventDestroy(e->ev_last);
    delete e;
    return MNX_OK;
}

extern "C" int mnx_load_tensor(mnx_engine* e, const char* name, const void* host_data, const int64_t* shape,
                               int32_t ndim, int32_t is_int64) {
    if (!e || !name || !host_data || (!shape && ndim > 0)) return fail(e, MNX_ERR_INVALID, "mnx_load_tensor: null argument");
    if (e->finalized) return fail(e, MNX_ERR_INVALID, "weights already finalized");
    HostTensor t;
    t.shape.assign(shape, shape + ndim);
    const int64_t n = t.numel();
    if (is_int64) t.i.assign((const int64_t*)host_data, (const int64_t*)host_data + n);
    else t.f.assign((const float*)host_data, (const float*)host_data + n);
    std::string key(name);
    size_t pos;
    while ((pos = key.find("module.")) != std::string::npos) key.erase(pos, 7);   // DDP prefix, model.py:25-26
    if (e->host_w.count(key)) return fail(e, MNX_ERR_WEIGHTS, "tensor %s loaded twice", key.c_str());
    e->host_w.emplace(key, std::move(t));
    return MNX_OK;
}

// fetch a required fp32 tensor with an exact shape
static const HostTensor* need(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape) {
    auto it = e->host_w.find(key);
    if (it == e->host_w.end()) {
        fail(e, MNX_ERR_WEIGHTS, "missing tensor %s", key.c_str());
        return nullptr;
    }
    HostTensor& t = it->second;
    if (t.shape != std::vector<int64_t>(shape) || t.f.empty()) {
        std::string got;
        for (auto d : t.shape) got += std::to_string(d) + ",";
        fail(e, MNX_ERR_WEIGHTS, "tensor %s has shape (%s), not the expected one", key.c_str(), got.c_str());
        return nullptr;
    }
    t.used = true;
    return &t;
}
const std::vector<float>* mnx_need(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape) {
    const HostTensor* t = need(e, key, shape);
    return t ? &t->f : nullptr;
}
const std::vector<int64_t>* mnx_need_i64(mnx_engine* e, const std::string& key, std::initializer_list<int64_t> shape) {
    auto it = e->host_w.find(key);
    if (it == e->host_w.end() || it->second.shape != std::vector<int64_t>(shape) || it->second.i.empty()) {
        fail(e, MNX_ERR_WEIGHTS, "missing or mis-shaped index tensor %s", key.c_str());
        return nullptr;
    }
    it->second.used = true;
    return &it->second.i;
}
void mnx_set_error(mnx_engine* e, const char* msg) { e->err = msg; }

// transpose [N][K] (torch Linear weight) into K-major [K][Npad] at column offset c0
static void put_transposed(std::vector<float>& dst, int Npad, int c0, const std::vector<float>& W, int N, int K) {
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) dst[(size_t)k * Npad + c0 + n] = W[(size_t)n * K + k];
}

static int finalize_decoder(mnx_engine* e) {
    const int D = MNX_DEC_D, V = e->cfg.vocab, ED = e->cfg.encoder_dim;
    const std::string P = "decoder.decoder.chartok_coords.";
#define NEED(var, key, ...)                          \
    const HostTensor* var = need(e, key, {__VA_ARGS__}); \
    if (!var) return MNX_ERR_WEIGHTS
#define UP(dst, vec) CUDA_TRY(e, mnx_upload(e, vec, &(dst)))
    std::vector<float> wkv((size_t)D * MNX_DEC_L * 512), bkv((size_t)MNX_DEC_L * 512);
    // per-head tile stream of the persistent cluster kernel: [8][L*14 + 1][256 k][32 cols]
    const size_t TILE = MG_TILE_FLOATS_H, TPL = MG_TILES_PER_LAYER_H, NTILE = MNX_DEC_L * TPL + 1;
    std::vector<float> wpack((size_t)8 * NTILE * TILE, 0.f), ppack((size_t)8 * MNX_DEC_L * MG_PARAM_FLOATS_H, 0.f);
    // tile[k][c] = W[row0 + c][k0 + k] for a torch Linear weight W [out][in]
    auto put_tile = [&](int h, int l, int t, const std::vector<float>& Wm, int in_dim, int row0, int k0) {
        float* dst = wpack.data() + ((size_t)h * NTILE + (size_t)l * TPL + t) * TILE;
        for (int k = 0; k < 256; ++k)
            for (int c = 0; c < 32; ++c) dst[k * 32 + c] = Wm[(size_t)(row0 + c) * in_dim + k0 + k];
    };
    // 16-CTA-cluster variant: [16][L*14 + 1][256 k][16 cols] and [16][L][1728]
    const size_t TILE16 = 256 * 16, PF16 = 1728;
    std::vector<float> wpack16((size_t)16 * NTILE * TILE16, 0.f), ppack16((size_t)16 * MNX_DEC_L * PF16, 0.f);
    auto put_tile16 = [&](int i, int l, int t, const std::vector<float>& Wm, int in_dim, int row0, int k0) {
        float* dst = wpack16.data() + ((size_t)i * NTILE + (size_t)l * TPL + t) * TILE16;
        for (int k = 0; k < 256; ++k)
            for (int c = 0; c < 16; ++c) dst[k * 16 + c] = Wm[(size_t)(row0 + c) * in_dim + k0 + k];
    };
    // throughput kernel (wide.cu): [8 heads][L*14 + 1] slots of 8192 floats, k-pair interleaved for packed f32x2 FMAs.
    // Column-parallel slot (q, k, v, Wq_ctx, W1 x4, vocabulary): [128 kp][32 cols][2].  Row-parallel slot (Wo, Wo_ctx,
    // W2 x4 k-chunks): [2 halves][16 kp][128 cols][2], k = this head's context features / this CTA's FFN columns.
    const size_t SLOTW = 8192;
    std::vector<float> wpackW((size_t)8 * NTILE * SLOTW, 0.f);
    auto put_col = [&](int h, int l, int t, const std::vector<float>& Wm, int in_dim, int row0) {
        float* dst = wpackW.data() + ((size_t)h * NTILE + (size_t)l * TPL + t) * SLOTW;
        for (int k = 0; k < 256; ++k)
            for (int c = 0; c < 32; ++c) dst[((k >> 1) * 32 + c) * 2 + (k & 1)] = Wm[(size_t)(row0 + c) * in_dim + k];
    };
    auto put_row = [&](int h, int l, int t, const std::vector<float>& Wm, int in_dim, int k0) {
        float* dst = wpackW.data() + ((size_t)h * NTILE + (size_t)l * TPL + t) * SLOTW;
        for (int half = 0; half < 2; ++half)
            for (int k = 0; k < 32; ++k)
                for (int c = 0; c < 128; ++c)
                    dst[half * 4096 + ((k >> 1) * 128 + c) * 2 + (k & 1)] = Wm[(size_t)(half * 128 + c) * in_dim + k0 + k];
    };
    for (int l = 0; l < MNX_DEC_L; ++l) {
        const std::string L = P + "decoder.transformer_layers." + std::to_string(l) + ".";
        DecLayerW& w = e->dw.layer[l];
        NEED(ln1w, L + "layer_norm_1.weight", D); NEED(ln1b, L + "layer_norm_1.bias", D);
        NEED(ln2w, L + "layer_norm_2.weight", D); NEED(ln2b, L + "layer_norm_2.bias", D);
        NEED(lnfw, L + "feed_forward.layer_norm.weight", D); NEED(lnfb, L + "feed_forward.layer_norm.bias", D);
        UP(w.ln1_w, ln1w->f); UP(w.ln1_b, ln1b->f); UP(w.ln2_w, ln2w->f); UP(w.ln2_b, ln2b->f);
        UP(w.lnf_w, lnfw->f); UP(w.lnf_b, lnfb->f);
        NEED(sq, L + "self_attn.linear_query.weight", D, D); NEED(sqb, L + "self_attn.linear_query.bias", D);
        NEED(sk, L + "self_attn.linear_keys.weight", D, D); NEED(skb, L + "self_attn.linear_keys.bias", D);
        NEED(sv, L + "self_attn.linear_values.weight", D, D); NEED(svb, L + "self_attn.linear_values.bias", D);
        NEED(so, L + "self_attn.final_linear.weight", D, D); NEED(sob, L + "self_attn.final_linear.bias", D);
        std::vector<float> qkv((size_t)D * 768), bq(768);
        put_transposed(qkv, 768, 0, sq->f, D, D);
        put_transposed(qkv, 768, 256, sk->f, D, D);
        put_transposed(qkv, 768, 512, sv->f, D, D);
        std::copy(sqb->f.begin(), sqb->f.end(), bq.begin());
        std::copy(skb->f.begin(), skb->f.end(), bq.begin() + 256);
        std::copy(svb->f.begin(), svb->f.end(), bq.begin() + 512);
        UP(w.wqkv_t, qkv); UP(w.bqkv, bq);
        std::vector<float> tmp((size_t)D * D);
        put_transposed(tmp, D, 0, so->f, D, D);
        UP(w.wo_s_t, tmp); UP(w.bo_s, sob->f);
        NEED(cq, L + "context_attn.linear_query.weight", D, D); NEED(cqb, L + "context_attn.linear_query.bias", D);
        NEED(ck, L + "context_attn.linear_keys.weight", D, D); NEED(ckb, L + "context_attn.linear_keys.bias", D);
        NEED(cv, L + "context_attn.linear_values.weight", D, D); NEED(cvb, L + "context_attn.linear_values.bias", D);
        NEED(co, L + "context_attn.final_linear.weight", D, D); NEED(cob, L + "context_attn.final_linear.bias", D);
        put_transposed(tmp, D, 0, cq->f, D, D);
        UP(w.wq_c_t, tmp); UP(w.bq_c, cqb->f);
        put_transposed(tmp, D, 0, co->f, D, D);
        UP(w.wo_c_t, tmp); UP(w.bo_c, cob->f);
        put_transposed(wkv, MNX_DEC_L * 512, l * 512, ck->f, D, D);
        put_transposed(wkv, MNX_DEC_L * 512, l * 512 + 256, cv->f, D, D);
        std::copy(ckb->f.begin(), ckb->f.end(), bkv.begin() + l * 512);
        std::copy(cvb->f.begin(), cvb->f.end(), bkv.begin() + l * 512 + 256);
        NEED(w1, L + "feed_forward.w_1.weight", MNX_DEC_FF, D); NEED(b1, L + "feed_forward.w_1.bias", MNX_DEC_FF);
        NEED(w2, L + "feed_forward.w_2.weight", D, MNX_DEC_FF); NEED(b2, L + "feed_forward.w_2.bias", D);
        std::vector<float> w1t((size_t)D * MNX_DEC_FF), w2t((size_t)MNX_DEC_FF * D);
        put_transposed(w1t, MNX_DEC_FF, 0, w1->f, MNX_DEC_FF, D);
        put_transposed(w2t, D, 0, w2->f, D, MNX_DEC_FF);
        UP(w.w1_t, w1t); UP(w.b1, b1->f); UP(w.w2_t, w2t); UP(w.b2, b2->f);
        for (int h = 0; h < 8; ++h) {
            put_tile(h, l, 0, sq->f, D, h * 32, 0);
            put_tile(h, l, 1, sk->f, D, h * 32, 0);
            put_tile(h, l, 2, sv->f, D, h * 32, 0);
            put_tile(h, l, 3, so->f, D, h * 32, 0);
            put_tile(h, l, 4, cq->f, D, h * 32, 0);
            put_tile(h, l, 5, co->f, D, h * 32, 0);
            for (int j = 0; j < 4; ++j) put_tile(h, l, 6 + j, w1->f, D, h * 128 + j * 32, 0);
            for (int j = 0; j < 4; ++j) put_tile(h, l, 10 + j, w2->f, MNX_DEC_FF, h * 32, 256 * j);
            float* pp = ppack.data() + ((size_t)h * MNX_DEC_L + l) * MG_PARAM_FLOATS_H;
            std::copy(ln1w->f.begin(), ln1w->f.end(), pp + 0);    std::copy(ln1b->f.begin(), ln1b->f.end(), pp + 256);
            std::copy(ln2w->f.begin(), ln2w->f.end(), pp + 512);  std::copy(ln2b->f.begin(), ln2b->f.end(), pp + 768);
            std::copy(lnfw->f.begin(), lnfw->f.end(), pp + 1024); std::copy(lnfb->f.begin(), lnfb->f.end(), pp + 1280);
            for (int c = 0; c < 32; ++c) {
                pp[1536 + c] = sqb->f[h * 32 + c]; pp[1568 + c] = skb->f[h * 32 + c]; pp[1600 + c] = svb->f[h * 32 + c];
                pp[1632 + c] = sob->f[h * 32 + c]; pp[1664 + c] = cqb->f[h * 32 + c]; pp[1696 + c] = cob->f[h * 32 + c];
                pp[1856 + c] = b2->f[h * 32 + c];
            }
            for (int c = 0; c < 128; ++c) pp[1728 + c] = b1->f[h * 128 + c];
            put_col(h, l, 0, sq->f, D, h * 32); put_col(h, l, 1, sk->f, D, h * 32); put_col(h, l, 2, sv->f, D, h * 32);
            put_row(h, l, 3, so->f, D, h * 32);
            put_col(h, l, 4, cq->f, D, h * 32);
            put_row(h, l, 5, co->f, D, h * 32);
            for (int j = 0; j < 4; ++j) put_col(h, l, 6 + j, w1->f, D, h * 128 + j * 32);
            for (int j = 0; j < 4; ++j) put_row(h, l, 10 + j, w2->f, MNX_DEC_FF, h * 128 + j * 32);
        }
        for (int i = 0; i < 16; ++i) {
            put_tile16(i, l, 0, sq->f, D, i * 16, 0);
            put_tile16(i, l, 1, sk->f, D, i * 16, 0);
            put_tile16(i, l, 2, sv->f, D, i * 16, 0);
            put_tile16(i, l, 3, so->f, D, i * 16, 0);
            put_tile16(i, l, 4, cq->f, D, i * 16, 0);
            put_tile16(i, l, 5, co->f, D, i * 16, 0);
            for (int j = 0; j < 4; ++j) put_tile16(i, l, 6 + j, w1->f, D, i * 64 + j * 16, 0);
            for (int j = 0; j < 4; ++j) put_tile16(i, l, 10 + j, w2->f, MNX_DEC_FF, i * 16, 256 * j);
            float* pp = ppack16.data() + ((size_t)i * MNX_DEC_L + l) * PF16;
            std::copy(ln1w->f.begin(), ln1w->f.end(), pp + 0);    std::copy(ln1b->f.begin(), ln1b->f.end(), pp + 256);
            std::copy(ln2w->f.begin(), ln2w->f.end(), pp + 512);  std::copy(ln2b->f.begin(), ln2b->f.end(), pp + 768);
            std::copy(lnfw->f.begin(), lnfw->f.end(), pp + 1024); std::copy(lnfb->f.begin(), lnfb->f.end(), pp + 1280);
            for (int c = 0; c < 16; ++c) {
                pp[1536 + c] = sqb->f[i * 16 + c]; pp[1552 + c] = skb->f[i * 16 + c]; pp[1568 + c] = svb->f[i * 16 + c];
                pp[1584 + c] = sob->f[i * 16 + c]; pp[1600 + c] = cqb->f[i * 16 + c]; pp[1616 + c] = cob->f[i * 16 + c];
                pp[1632 + c] = b2->f[i * 16 + c];
            }
            for (int c = 0; c < 64; ++c) pp[1648 + c] = b1->f[i * 64 + c];
        }
    }
    UP(e->dw.wkv_c_t, wkv); UP(e->dw.bkv_c, bkv);
    NEED(lnw, P + "decoder.layer_norm.weight", D); NEED(lnb, P + "decoder.layer_norm.bias", D);
    UP(e->dw.lnF_w, lnw->f); UP(e->dw.lnF_b, lnb->f);
    NEED(ow, P + "output_layer.weight", V, D); NEED(ob, P + "output_layer.bias", V);
    std::vector<float> wout((size_t)D * 256, 0.f), bout(256, 0.f);
    put_transposed(wout, 256, 0, ow->f, V, D);
    std::copy(ob->f.begin(), ob->f.end(), bout.begin());
    UP(e->dw.wout_t, wout); UP(e->dw.bout, bout);
    {
        std::vector<float> fin(768, 0.f);
        std::copy(lnw->f.begin(), lnw->f.end(), fin.begin());
        std::copy(lnb->f.begin(), lnb->f.end(), fin.begin() + 256);
        std::copy(ob->f.begin(), ob->f.end(), fin.begin() + 512);
        for (int h = 0; h < 8; ++h) {
            float* dst = wpack.data() + ((size_t)h * NTILE + (size_t)MNX_DEC_L * TPL) * TILE;
            for (int k = 0; k < 256; ++k)
                for (int c = 0; c < 32; ++c) dst[k * 32 + c] = (h * 32 + c < V) ? ow->f[(size_t)(h * 32 + c) * D + k] : 0.f;
        }
        for (int i = 0; i < 16; ++i) {
            float* dst = wpack16.data() + ((size_t)i * NTILE + (size_t)MNX_DEC_L * TPL) * TILE16;
            for (int k = 0; k < 256; ++k)
                for (int c = 0; c < 16; ++c) dst[k * 16 + c] = (i * 16 + c < V) ? ow->f[(size_t)(i * 16 + c) * D + k] : 0.f;
        }
        for (int h = 0; h < 8; ++h) {
            float* dst = wpackW.data() + ((size_t)h * NTILE + (size_t)MNX_DEC_L * TPL) * SLOTW;
            for (int k = 0; k < 256; ++k)
                for (int c = 0; c < 32; ++c)
                    dst[((k >> 1) * 32 + c) * 2 + (k & 1)] = (h * 32 + c < V) ? ow->f[(size_t)(h * 32 + c) * D + k] : 0.f;
        }
        UP(e->wpackW, wpackW);
        UP(e->wpack, wpack); UP(e->ppack, ppack); UP(e->finalp, fin);
        UP(e->wpack16, wpack16); UP(e->ppack16, ppack16);
    }
    NEED(emb, P + "embeddings.make_embedding.emb_luts.0.weight", V, D);
    UP(e->dw.emb, emb->f);
    NEED(pe, P + "embeddings.make_embedding.pe.pe", 5000, 1, D);
    UP(e->dw.pe, pe->f);
    NEED(ew, P + "enc_trans_layer.0.weight", D, ED); NEED(eb, P + "enc_trans_layer.0.bias", D);
    std::vector<float> wenc((size_t)ED * D);
    put_transposed(wenc, D, 0, ew->f, D, ED);
    UP(e->dw.wenc_t, wenc); UP(e->dw.benc, eb->f);
    NEED(g0, "decoder.decoder.edges.mlp.0.weight", D, 2 * D); NEED(g0b, "decoder.decoder.edges.mlp.0.bias", D);
    NEED(g2, "decoder.decoder.edges.mlp.2.weight", MNX_EDGE_CLASSES, D); NEED(g2b, "decoder.decoder.edges.mlp.2.bias", MNX_EDGE_CLASSES);
    std::vector<float> wab((size_t)D * 512);
    for (int n = 0; n < D; ++n)
        for (int k = 0; k < D; ++k) {
            wab[(size_t)k * 512 + n] = g0->f[(size_t)n * 512 + k];
            wab[(size_t)k * 512 + 256 + n] = g0->f[(size_t)n * 512 + 256 + k];
        }
    UP(e->dw.we_a_t, wab); e->dw.we_b_t = nullptr;
    UP(e->dw.be0, g0b->f); UP(e->dw.we2, g2->f); UP(e->dw.be2, g2b->f);
#undef NEED
#undef UP
    return MNX_OK;
}

// per-context device buffers (everything one predict / decode call writes), allocated into the flat fields of `e`;
// R = decoder rows (images x beams for context 0 of a beam-search handle, images otherwise)
static int alloc_context(mnx_engine* e, size_t R) {
    const size_t B = e->cfg.max_batch, T = e->cfg.max_len, S = e->S_max, KA = e->cfg.max_atoms;
    CUDA_TRY(e, dev_alloc(e, &e->st, 1));
    CUDA_TRY(e, dev_alloc(e, &e->alive, 2 * R));
    CUDA_TRY(e, dev_alloc(e, &e->cur_tok, R));
    CUDA_TRY(e, dev_alloc(e, &e->finished, R));
    CUDA_TRY(e, dev_alloc(e, &e->xa, R * 256));
    CUDA_TRY(e, dev_alloc(e, &e->xb, R * 256));
    CUDA_TRY(e, dev_alloc(e, &e->q, R * 256));
    CUDA_TRY(e, dev_alloc(e, &e->part, R * 8 * 256));
    CUDA_TRY(e, dev_alloc(e, &e->part2, R * 8 * 256));
    CUDA_TRY(e, dev_alloc(e, &e->hbuf, R * 1024));
    CUDA_TRY(e, dev_alloc(e, &e->selfK, MNX_DEC_L * R * T * 256));
    CUDA_TRY(e, dev_alloc(e, &e->selfV, MNX_DEC_L * R * T * 256));
    CUDA_TRY(e, dev_alloc(e, &e->crossK, MNX_DEC_L * B * S * 256));
    CUDA_TRY(e, dev_alloc(e, &e->crossV, MNX_DEC_L * B * S * 256));
    CUDA_TRY(e, dev_alloc(e, &e->membank, B * S * 256));
    CUDA_TRY(e, dev_alloc(e, &e->ids, B * T));
    CUDA_TRY(e, dev_alloc(e, &e->lens, B));
    CUDA_TRY(e, dev_alloc(e, &e->logp, B * T));
    CUDA_TRY(e, dev_alloc(e, &e->hidden, R * T * 256));
    CUDA_TRY(e, dev_alloc(e, &e->row_state, B));
    CUDA_TRY(e, dev_alloc(e, &e->steps_run_dev, 1));
    CUDA_TRY(e, dev_alloc(e, &e->ticket, 16));     // one ticket counter per wide-kernel launch of a call
    CUDA_TRY(e, dev_alloc(e, &e->hg, B * KA * 256));
    CUDA_TRY(e, dev_alloc(e, &e->AB, B * KA * 512));
    CUDA_TRY(e, dev_alloc(e, &e->prob, B * KA * KA * 8));
    CUDA_TRY(e, dev_alloc(e, &e->features, B * S * (size_t)e->cfg.encoder_dim));
    CUDA_TRY(e, dev_alloc(e, &e->atom_idx, B * KA));
    CUDA_TRY(e, dev_alloc(e, &e->n_atoms, B));
    CUDA_TRY(e, dev_alloc(e, &e->edges, B * KA * KA));
    return MNX_OK;
}
static CtxPtrs capture_context(const mnx_engine* e) {
    return CtxPtrs{e->st, e->alive, e->cur_tok, e->finished, e->xa, e->xb, e->q, e->part, e->part2, e->hbuf,
                   e->selfK, e->selfV, e->crossK, e->crossV, e->membank, e->ids, e->lens, e->logp, e->hidden,
                   e->row_state, e->steps_run_dev, e->ticket, e->hg, e->AB, e->prob, e->features,
                   e->atom_idx, e->n_atoms, e->edges};
}
static void apply_context(mnx_engine* e, const CtxPtrs& c) {
    e->st = c.st; e->alive = c.alive; e->cur_tok = c.cur_tok; e->finished = c.finished;
    e->xa = c.xa; e->xb = c.xb; e->q = c.q; e->part = c.part; e->part2 = c.part2; e->hbuf = c.hbuf;
    e->selfK = c.selfK; e->selfV = c.selfV; e->crossK = c.crossK; e->crossV = c.crossV; e->membank = c.membank;
    e->ids = c.ids; e->lens = c.lens; e->logp = c.logp; e->hidden = c.hidden;
    e->row_state = c.row_state; e->steps_run_dev = c.steps_run_dev; e->ticket = c.ticket;
    e->hg = c.hg; e->AB = c.AB; e->prob = c.prob; e->features = c.features;
    e->atom_idx = c.atom_idx; e->n_atoms = c.n_atoms; e->edges = c.edges;
    e->edge_hidden = e->hidden;
}

static int alloc_workspaces(mnx_engine* e) {
    // B = images per call; R = decoder rows per call (images x beams under beam search)
    const size_t B = e->cfg.max_batch, T = e->cfg.max_len;
    const size_t R = B * (e->cfg.max_beam > 1 ? e->cfg.max_beam : 1);
    int rc = alloc_context(e, R);
    if (rc != MNX_OK) return rc;
    e->ctxs.assign(1, capture_context(e));
    e->cur_ctx = 0;
    e->edge_hidden = e->hidden;
    CUDA_TRY(e, dev_alloc(e, &e->labels, B * (T + 1)));
    if (e->cfg.max_beam > 1) {
        BeamBuffers& m = e->bm;
        CUDA_TRY(e, dev_alloc(e, &m.alive_img, 2 * B));
        CUDA_TRY(e, dev_alloc(e, &m.img_done, B));
        CUDA_TRY(e, dev_alloc(e, &m.top_fin, B));
        CUDA_TRY(e, dev_alloc(e, &m.lp, R * 256));
        CUDA_TRY(e, dev_alloc(e, &m.cum, 2 * R));
        CUDA_TRY(e, dev_alloc(e, &m.anc, 2 * R * T));
        CUDA_TRY(e, dev_alloc(e, &m.hist_ids, 2 * R * T));
        CUDA_TRY(e, dev_alloc(e, &m.hist_logp, 2 * R * T));
        CUDA_TRY(e, dev_alloc(e, &m.hyp_count, B));
        CUDA_TRY(e, dev_alloc(e, &m.hyp_order, B * MNX_MAX_BEAM));
        CUDA_TRY(e, dev_alloc(e, &m.hyp_score, B * MNX_MAX_BEAM));
        CUDA_TRY(e, dev_alloc(e, &m.hyp_len, B * MNX_MAX_BEAM));
        CUDA_TRY(e, dev_alloc(e, &m.hyp_ids, B * MNX_MAX_BEAM * T));
        CUDA_TRY(e, dev_alloc(e, &m.hyp_logp, B * MNX_MAX_BEAM * T));
        CUDA_TRY(e, dev_alloc(e, &m.hyp_anc, B * MNX_MAX_BEAM * T));
        CUDA_TRY(e, dev_alloc(e, &m.trace, T * B * MNX_MAX_BEAM));
        CUDA_TRY(e, dev_alloc(e, &e->hid_best, B * T * 256));
    }
    CUDA_TRY(e, dev_alloc(e, &e->prof_dev, 64));
    CUDA_TRY(e, cudaMemset(e->prof_dev, 0, 64 * sizeof(long long)));
    CUDA_TRY(e, dev_alloc(e, &e->pp_bbox, B * 4));
    if (e->cfg.encoder_kind != MNX_ENCODER_NONE)
        CUDA_TRY(e, dev_alloc(e, &e->images, B * 3 * (size_t)e->cfg.max_height * e->cfg.max_width));
    void* cls = nullptr;
    CUDA_TRY(e, mnx_upload_raw(e, e->cls_host.data(), e->cls_host.size(), &cls));
    e->d_cls = (uint8_t*)cls;
    return MNX_OK;
}

extern "C" int mnx_reserve_contexts(mnx_engine* e, int32_t n) {
    if (!e || n < 1 || n > 64) return fail(e, MNX_ERR_INVALID, "mnx_reserve_contexts: n must be in [1,64]");
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    ON_ENGINE_DEVICE(e);
    const CtxPtrs cur = capture_context(e);
    while ((int)e->ctxs.size() < n) {
        int rc = alloc_context(e, e->cfg.max_batch);
        if (rc != MNX_OK) { apply_context(e, cur); return rc; }
        e->ctxs.push_back(capture_context(e));
    }
    apply_context(e, cur);
    return MNX_OK;
}

extern "C" int mnx_set_context(mnx_engine* e, int32_t i) {
    if (!e || i < 0 || i >= (int)e->ctxs.size()) return fail(e, MNX_ERR_INVALID, "mnx_set_context: context %d not reserved", i);
    apply_context(e, e->ctxs[i]);
    e->cur_ctx = i;
    return MNX_OK;
}

extern "C" int mnx_set_wide_rows(mnx_engine* e, int32_t rows) {
    if (!e || rows < 0 || rows > MGW_GMAX_H) return fail(e, MNX_ERR_INVALID, "mnx_set_wide_rows: rows must be in [0,%d]", MGW_GMAX_H);
    e->wide_rows = rows;
    return MNX_OK;
}

extern "C" int mnx_set_decode_path(mnx_engine* e, int32_t path) {
    if (!e || (path != 0 && path != 1 && path != 2 && path != 3 && path != 6))
        return fail(e, MNX_ERR_INVALID, "mnx_set_decode_path: path must be 0 (auto), 1 (graph), 2 (cluster8), 3 (cluster16) or 6 (wide)");
    e->decode_path = path;
    return MNX_OK;
}

extern "C" int mnx_finalize_weights(mnx_engine* e) {
    if (!e) return MNX_ERR_INVALID;
    if (e->finalized) return fail(e, MNX_ERR_INVALID, "weights already finalized");
    ON_ENGINE_DEVICE(e);
    int rc = finalize_decoder(e);
    if (rc != MNX_OK) return rc;
    if (e->cfg.encoder_kind != MNX_ENCODER_NONE) {
        rc = encoder_finalize(e, e->enc, e->cfg);
        if (rc != MNX_OK) return rc;
    }
    for (auto& kv : e->host_w)
        if (!kv.second.used) return fail(e, MNX_ERR_WEIGHTS, "unexpected tensor %s (not part of the model)", kv.first.c_str());
    rc = alloc_workspaces(e);
    if (rc != MNX_OK) return rc;
    e->host_w.clear();
    e->finalized = true;
    CUDA_TRY(e, cudaDeviceSynchronize());
    return MNX_OK;
}

static DecBuffers make_buffers(mnx_engine* e, int B, int S) {
    DecBuffers b{};
    b.st = e->st; b.alive = e->alive; b.cur_tok = e->cur_tok; b.finished = e->finished;
    b.xa = e->xa; b.xb = e->xb; b.q = e->q; b.part = e->part; b.part2 = e->part2; b.hbuf = e->hbuf;
    b.selfK = e->selfK; b.selfV = e->selfV; b.crossK = e->crossK; b.crossV = e->crossV; b.membank = e->membank;
    b.B = B; b.S = S; b.T = e->cfg.max_len;
    b.ids = e->ids; b.lens = e->lens; b.logp = e->logp; b.hidden = e->hidden;
    b.labels = e->labels;
    return b;
}

static int ensure_graph(mnx_engine* e, const DecBuffers& b) {
    if (e->graph && e->graph_B == b.B && e->graph_S == b.S) return MNX_OK;
    cudaStream_t s = e->cap_stream;
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
    cudaGraph_t graph = nullptr;
    CUDA_TRY(e, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int nodes = 0;
    cudaError_t kerr = cudaSuccess;
    for (int i = 0; i < STEPS_PER_GRAPH && kerr == cudaSuccess; ++i) nodes += dec_launch_step(b, e->dw, e->g, nullptr, s, &kerr);
    cudaError_t cerr = cudaStreamEndCapture(s, &graph);
    if (kerr != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        return fail(e, MNX_ERR_CUDA, "decode-step launch failed during capture: %s", cudaGetErrorString(kerr));
    }
    CUDA_TRY(e, cerr);
    cudaError_t ierr = cudaGraphInstantiate(&e->graph, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_TRY(e, ierr);
    e->graph_B = b.B; e->graph_S = b.S; e->graph_nodes = nodes;
    return MNX_OK;
}

static int decode_internal(mnx_engine* e, const float* features, int B, int S, cudaStream_t s, const int32_t* labels = nullptr,
                           int n_labels = 0) {
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    if (B < 1 || B > e->cfg.max_batch) return fail(e, MNX_ERR_CAPACITY, "batch %d exceeds max_batch %d", B, e->cfg.max_batch);
    if (S < 1 || S > e->S_max) return fail(e, MNX_ERR_CAPACITY, "memory length %d exceeds capacity %d", S, e->S_max);
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, s);
    DecBuffers b = make_buffers(e, B, S);
    const int T = e->cfg.max_len;
    e->edge_hidden = e->hidden;
    CUDA_TRY(e, cudaMemsetAsync(e->st, 0, sizeof(DecState), s));
    CUDA_TRY(e, cudaMemsetAsync(e->finished, 0, sizeof(int) * B, s));
    CUDA_TRY(e, cudaMemsetAsync(e->lens, 0, sizeof(int) * B, s));
    CUDA_TRY(e, cudaMemsetAsync(e->ids, 0, sizeof(int) * (size_t)B * T, s));
    CUDA_TRY(e, cudaMemsetAsync(e->logp, 0, sizeof(float) * (size_t)B * T, s));
    int nl = 0;
    CUDA_TRY(e, dec_precompute(b, e->dw, features, e->cfg.encoder_dim, s, &nl));
    e->launches += nl;
    // persistent cluster kernels when every cluster (<= 4 rows each) can be co-resident: prefer 16-CTA clusters
    // (half the per-SM byte stream), then 8-CTA clusters, else the multi-kernel graph path
    const int usable16 = e->max_clusters16 < 8 ? e->max_clusters16 : 8;
    const int usable16s = e->max_clusters16s < 8 ? e->max_clusters16s : 8;
    const int usable8 = e->max_clusters < 16 ? e->max_clusters : 16;
    const bool fits16s = usable16s > 0 && B <= usable16s * MG16S_GMAX_H;   // small-batch configuration (4-warp groups)
    const bool fits16 = fits16s || (usable16 > 0 && B <= usable16 * MG16_GMAX_H);
    const bool fits8 = usable8 > 0 && B <= usable8 * MG_GMAX_H;
    if (e->decode_path == 3 && !fits16) return fail(e, MNX_ERR_CAPACITY, "16-CTA cluster path forced but %d rows do not fit %d clusters", B, usable16);
    if (e->decode_path == 2 && !fits8) return fail(e, MNX_ERR_CAPACITY, "8-CTA cluster path forced but %d rows do not fit %d clusters", B, usable8);
    const int growsw = (e->wide_rows > 0 && e->wide_rows < MGW_GMAX_H) ? e->wide_rows : MGW_GMAX_H;
    const int nclw = (B + growsw - 1) / growsw;
    const bool fitsw = nclw <= e->max_clusters_w && T <= MGW_MAX_KEYS_H + 1 && S <= MGW_MAX_KEYS_H;
    if (e->decode_path == 6 && !fitsw)
        return fail(e, MNX_ERR_CAPACITY, "throughput decode kernel forced but B=%d S=%d T=%d does not fit (%d clusters resident, <= %d keys)",
                    B, S, T, e->max_clusters_w, MGW_MAX_KEYS_H);
    // partial-label decoding lives on the multi-kernel graph path only
    // the throughput kernel: forced (predict_pipelined / mnx_set_decode_path), or auto for batches too large for the latency
    // kernels.  Its step time barely depends on the number of clusters (bs 32 / 64 / 128 / 192: 143 / 145 / 148 / 150 ms for
    // 480 steps; the multi-kernel graph path: 303 ms at 128 rows, 375 ms at 256), so a batch above 16 rows x the resident
    // clusters (15 on this part = 240 rows) runs as consecutive launches: the row-rank rule of a later launch reads the
    // final row_state words of the earlier rows (they carry the step at which each row finished)
    const bool wide_shapes = e->max_clusters_w > 0 && T <= MGW_MAX_KEYS_H + 1 && S <= MGW_MAX_KEYS_H;
    const int wide_launches = wide_shapes ? (nclw + e->max_clusters_w - 1) / e->max_clusters_w : 0;
    const bool usew = !labels && (e->decode_path == 6 ||
                                  (e->decode_path == 0 && !fits16 && !fits8 && wide_shapes && wide_launches <= 16 && e->wide_rows == 0));
    const bool use16 = !labels && !usew && ((e->decode_path == 3) || (e->decode_path == 0 && fits16));
    const bool use8 = !labels && !usew && !use16 && ((e->decode_path == 2) || (e->decode_path == 0 && fits8));
    const bool use16s = use16 && fits16s;
    if (e->cur_ctx != 0 && !(usew || use16 || use8))
        return fail(e, MNX_ERR_INVALID, "the multi-kernel graph path runs in context 0 only");
    if (usew || use16 || use8) {
        const int usable = usew ? nclw : use16s ? usable16s : use16 ? usable16 : usable8;
        const int G = (B + usable - 1) / usable;
        const int clusters = (B + G - 1) / G;
        CUDA_TRY(e, cudaMemsetAsync(e->row_state, 0, sizeof(unsigned) * B, s));
        CUDA_TRY(e, cudaMemsetAsync(e->steps_run_dev, 0, sizeof(int), s));
        CUDA_TRY(e, cudaMemsetAsync(e->ticket, 0, sizeof(int) * 16, s));
        MegaArgs a{};
        a.wpack = e->wpack; a.ppack = e->ppack; a.wpack16 = e->wpack16; a.ppack16 = e->ppack16;
        a.wpackW = e->wpackW; a.ticket = e->ticket;
        a.finalp = e->finalp; a.emb = e->dw.emb; a.pe = e->dw.pe;
        a.selfK = e->selfK; a.selfV = e->selfV; a.crossK = e->crossK; a.crossV = e->crossV;
        a.B = B; a.S = S; a.T = T; a.G = G;
        a.ids = e->ids; a.logp = e->logp; a.hidden = e->hidden; a.lens = e->lens;
        a.row_state = e->row_state; a.steps_run = e->steps_run_dev; a.g = e->g;
        a.prof = e->decode_profile ? e->prof_dev : nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (e->time_launches) {
            CUDA_TRY(e, cudaEventCreate(&ev0)); CUDA_TRY(e, cudaEventCreate(&ev1));
            CUDA_TRY(e, cudaEventRecord(ev0, s));
        }
        if (usew) {
            // as many launches as the resident clusters require (one for every batch up to 16 x max_clusters_w rows)
            int nlaunch = 0;
            for (int c0 = 0; c0 < clusters; c0 += e->max_clusters_w, ++nlaunch) {
                a.row_base = c0 * G;
                a.ticket = e->ticket + nlaunch;
                const int ncl = clusters - c0 < e->max_clusters_w ? clusters - c0 : e->max_clusters_w;
                CUDA_TRY(e, wide_launch(a, ncl, s));
            }
            e->launches += nlaunch - 1;
        } else {
            CUDA_TRY(e, use16s ? mega16s_launch(a, clusters, s) : use16 ? mega16_launch(a, clusters, s) : mega_launch(a, clusters, s));
        }
        if (e->time_launches) {
            CUDA_TRY(e, cudaEventRecord(ev1, s));
            e->launch_events.emplace_back(ev0, ev1);
        }
        e->launches += 1;
        e->last_path = usew ? 6 : use16s ? 5 : use16 ? 3 : 2;
        // no host synchronisation: the whole decode is one kernel, so the call is asynchronous like any other
        // launch on `s` (lets the caller overlap the next batch's encoder with it); mnx_last_decode_steps
        // reads the step count back on demand
        e->steps_pending = true;
        e->last_B = B; e->last_S = S;
        return MNX_OK;
    }
    e->last_path = 1;
    int rc = ensure_graph(e, b);
    if (rc != MNX_OK) return rc;
    const int lab_len = labels ? (n_labels < T + 1 ? n_labels : T + 1) : 0;
    if (labels) {
        // the columns a decode of <= T steps can read (t and t + 1), by original row
        CUDA_TRY(e, cudaMemcpy2DAsync(e->labels, sizeof(int) * (size_t)(T + 1), labels, sizeof(int) * (size_t)n_labels, sizeof(int) * (size_t)lab_len,
                                      (size_t)B, cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(e, dec_set_label_len(b, lab_len, s));
        e->launches += 1;
    }
    for (int chunk = 0; chunk < T / STEPS_PER_GRAPH; ++chunk) {
        CUDA_TRY(e, cudaGraphLaunch(e->graph, s));
        e->launches += e->graph_nodes;
        CUDA_TRY(e, cudaMemcpyAsync(e->h_done, &e->st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(e, cudaStreamSynchronize(s));
        if (*e->h_done) break;
    }
    DecState hs{};
    CUDA_TRY(e, cudaMemcpyAsync(&hs, e->st, sizeof(DecState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(e, cudaStreamSynchronize(s));
    e->last_steps = hs.steps_run;
    e->steps_pending = false;
    e->last_B = B; e->last_S = S;
    if (labels) {
        // the reference indexes labels[:, step] at every step it runs (components.py:287): a decode that outlives the
        // labels is its IndexError
        if (hs.steps_run > n_labels)
            return fail(e, MNX_ERR_INVALID, "labels have %d columns but the decode ran %d steps (IndexError at components.py:287 in the reference)",
                        n_labels, hs.steps_run);
        CUDA_TRY(e, dec_label_merge(b, lab_len, s));
        e->launches += 1;
    }
    return MNX_OK;
}

extern "C" int mnx_decode_greedy_labels(mnx_engine* e, const float* features, int32_t B, int32_t S, const int32_t* labels, int32_t n_labels,
                                        int32_t* ids, int32_t* lens, float* token_logp, float* hidden, void* cuda_stream) {
    if (!e || !features || !labels) return fail(e, MNX_ERR_INVALID, "mnx_decode_greedy_labels: null argument");
    if (n_labels < 1) return fail(e, MNX_ERR_INVALID, "mnx_decode_greedy_labels: labels need at least one column, got %d", n_labels);
    if (e->cur_ctx != 0) return fail(e, MNX_ERR_INVALID, "partial-label decoding runs in context 0 only");
    cudaStream_t s = (cudaStream_t)cuda_stream;
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, s);
    int rc = decode_internal(e, features, B, S, s, labels, n_labels);
    if (rc != MNX_OK) return rc;
    const size_t T = e->cfg.max_len;
    if (ids) CUDA_TRY(e, cudaMemcpyAsync(ids, e->ids, sizeof(int) * B * T, cudaMemcpyDeviceToDevice, s));
    if (lens) CUDA_TRY(e, cudaMemcpyAsync(lens, e->lens, sizeof(int) * B, cudaMemcpyDeviceToDevice, s));
    if (token_logp) CUDA_TRY(e, cudaMemcpyAsync(token_logp, e->logp, sizeof(float) * B * T, cudaMemcpyDeviceToDevice, s));
    if (hidden) CUDA_TRY(e, cudaMemcpyAsync(hidden, e->hidden, sizeof(float) * B * T * 256, cudaMemcpyDeviceToDevice, s));
    return MNX_OK;
}

extern "C" int mnx_decode_greedy(mnx_engine* e, const float* features, int32_t B, int32_t S, int32_t* ids,
                                 int32_t* lens, float* token_logp, float* hidden, void* cuda_stream) {
    if (!e || !features) return fail(e, MNX_ERR_INVALID, "mnx_decode_greedy: null argument");
    cudaStream_t s = (cudaStream_t)cuda_stream;
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, s);
    int rc = decode_internal(e, features, B, S, s);
    if (rc != MNX_OK) return rc;
    const size_t T = e->cfg.max_len;
    if (ids) CUDA_TRY(e, cudaMemcpyAsync(ids, e->ids, sizeof(int) * B * T, cudaMemcpyDeviceToDevice, s));
    if (lens) CUDA_TRY(e, cudaMemcpyAsync(lens, e->lens, sizeof(int) * B, cudaMemcpyDeviceToDevice, s));
    if (token_logp) CUDA_TRY(e, cudaMemcpyAsync(token_logp, e->logp, sizeof(float) * B * T, cudaMemcpyDeviceToDevice, s));
    if (hidden) CUDA_TRY(e, cudaMemcpyAsync(hidden, e->hidden, sizeof(float) * B * T * 256, cudaMemcpyDeviceToDevice, s));
    return MNX_OK;
}

// ---------------------------------------------------------------------------------------
// beam search: multi-kernel graph path with rows = images x beams
// ---------------------------------------------------------------------------------------
static int ensure_graph_beam(mnx_engine* e, const DecBuffers& b, const BeamBuffers& bm) {
    if (e->graph_beam && e->gb_B == bm.n_img0 && e->gb_S == b.S && e->gb_K == bm.beam && e->gb_NB == bm.n_best) return MNX_OK;
    cudaStream_t s = e->cap_stream;
    if (e->graph_beam) { cudaGraphExecDestroy(e->graph_beam); e->graph_beam = nullptr; }
    cudaGraph_t graph = nullptr;
    CUDA_TRY(e, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int nodes = 0;
    cudaError_t kerr = cudaSuccess;
    for (int i = 0; i < STEPS_PER_GRAPH && kerr == cudaSuccess; ++i) nodes += dec_launch_step(b, e->dw, e->g, &bm, s, &kerr);
    cudaError_t cerr = cudaStreamEndCapture(s, &graph);
    if (kerr != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        return fail(e, MNX_ERR_CUDA, "beam decode-step launch failed during capture: %s", cudaGetErrorString(kerr));
    }
    CUDA_TRY(e, cerr);
    cudaError_t ierr = cudaGraphInstantiate(&e->graph_beam, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_TRY(e, ierr);
    e->gb_B = bm.n_img0; e->gb_S = b.S; e->gb_K = bm.beam; e->gb_NB = bm.n_best; e->gb_nodes = nodes;
    return MNX_OK;
}

extern "C" int mnx_decode_beam(mnx_engine* e, const float* features, int32_t B, int32_t S, int32_t beam, int32_t n_best,
                               int32_t* ids, int32_t* lens, float* scores, float* token_logp, float* hidden,
                               void* cuda_stream) {
    if (!e || !features) return fail(e, MNX_ERR_INVALID, "mnx_decode_beam: null argument");
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    if (beam < 1 || beam > e->cfg.max_beam || e->cfg.max_beam < 2)
        return fail(e, MNX_ERR_CAPACITY, "beam %d outside [1, max_beam=%d] (create the handle with max_beam >= 2)", beam, e->cfg.max_beam);
    if (n_best < 1 || n_best > beam) return fail(e, MNX_ERR_INVALID, "n_best must be in [1, beam]");
    if (B < 1 || B > e->cfg.max_batch) return fail(e, MNX_ERR_CAPACITY, "batch %d exceeds max_batch %d", B, e->cfg.max_batch);
    if (S < 1 || S > e->S_max) return fail(e, MNX_ERR_CAPACITY, "memory length %d exceeds capacity %d", S, e->S_max);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, (cudaStream_t)cuda_stream);
    const int T = e->cfg.max_len, R = B * beam;
    DecBuffers bi = make_buffers(e, B, S);     // image-shaped view for the once-per-call projections
    DecBuffers b = make_buffers(e, R, S);      // row-shaped view for the step kernels
    BeamBuffers bm = e->bm;
    bm.beam = beam; bm.n_best = n_best; bm.n_img0 = B;
    CUDA_TRY(e, cudaMemsetAsync(e->st, 0, sizeof(DecState), s));
    CUDA_TRY(e, dec_beam_init(b, bm, s));
    CUDA_TRY(e, cudaMemsetAsync(bm.trace, 0xff, sizeof(int) * (size_t)T * B * MNX_MAX_BEAM, s));
    e->last_beam_B = B;
    int nl = 1;
    CUDA_TRY(e, dec_precompute(bi, e->dw, features, e->cfg.encoder_dim, s, &nl));
    e->launches += nl;
    e->last_path = 4;
    int rc = ensure_graph_beam(e, b, bm);
    if (rc != MNX_OK) return rc;
    for (int chunk = 0; chunk < T / STEPS_PER_GRAPH; ++chunk) {
        CUDA_TRY(e, cudaGraphLaunch(e->graph_beam, s));
        e->launches += e->gb_nodes;
        CUDA_TRY(e, cudaMemcpyAsync(e->h_done, &e->st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(e, cudaStreamSynchronize(s));
        if (*e->h_done) break;
    }
    CUDA_TRY(e, dec_beam_finalize(b, bm, ids, lens, scores, token_logp, e->ids, e->lens, e->logp, e->hid_best, s));
    e->launches += 1;
    e->edge_hidden = e->hid_best;
    if (hidden) CUDA_TRY(e, cudaMemcpyAsync(hidden, e->hid_best, sizeof(float) * (size_t)B * T * 256, cudaMemcpyDeviceToDevice, s));
    DecState hs{};
    CUDA_TRY(e, cudaMemcpyAsync(&hs, e->st, sizeof(DecState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(e, cudaStreamSynchronize(s));
    e->last_steps = hs.steps_run;
    e->steps_pending = false;
    e->last_B = 0; e->last_S = S;   // isolated kernel timing (mnx_time_kernel) is defined on greedy shapes only
    return MNX_OK;
}

extern "C" int mnx_atom_indices(mnx_engine* e, const int32_t* ids, const int32_t* lens, int32_t B, int32_t* atom_idx,
                                int32_t* n_atoms, void* cuda_stream) {
    if (!e || !atom_idx || !n_atoms || (!ids != !lens)) return fail(e, MNX_ERR_INVALID, "mnx_atom_indices: null argument");
    if (!ids) { ids = e->ids; lens = e->lens; }
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    if (B < 1 || B > e->cfg.max_batch) return fail(e, MNX_ERR_CAPACITY, "batch %d exceeds max_batch %d", B, e->cfg.max_batch);
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, (cudaStream_t)cuda_stream);
    CUDA_TRY(e, dec_atom_scan(ids, lens, B, e->cfg.max_len, e->d_cls, e->g, e->cfg.max_atoms, atom_idx, n_atoms,
                              (cudaStream_t)cuda_stream));
    e->launches += 1;
    return MNX_OK;
}

extern "C" int mnx_edges(mnx_engine* e, const float* hidden, const int32_t* atom_idx, const int32_t* n_atoms, int32_t B,
                         uint8_t* edges, float* edge_score, void* cuda_stream) {
    if (!e || !atom_idx || !n_atoms || !edges) return fail(e, MNX_ERR_INVALID, "mnx_edges: null argument");
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    if (B < 1 || B > e->cfg.max_batch) return fail(e, MNX_ERR_CAPACITY, "batch %d exceeds max_batch %d", B, e->cfg.max_batch);
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, (cudaStream_t)cuda_stream);
    int nl = 0;
    CUDA_TRY(e, dec_edges(hidden ? hidden : e->edge_hidden, atom_idx, n_atoms, B, e->cfg.max_len, e->cfg.max_atoms, e->dw,
                          e->hg, e->AB, e->prob, edges, edge_score, (cudaStream_t)cuda_stream, &nl));
    e->launches += nl;
    return MNX_OK;
}

extern "C" int mnx_confidence(mnx_engine* e, const int32_t* ids, const int32_t* lens, const float* token_logp, int32_t B,
                              const float* edge_score, float* atom_scores, float* seq_score, double* overall_score, void* cuda_stream) {
    if (!e || !ids || !lens || !token_logp || !edge_score || !atom_scores || !seq_score || !overall_score)
        return fail(e, MNX_ERR_INVALID, "mnx_confidence: null argument");
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    if (B < 1 || B > e->cfg.max_batch) return fail(e, MNX_ERR_CAPACITY, "batch %d exceeds max_batch %d", B, e->cfg.max_batch);
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, (cudaStream_t)cuda_stream);
    CUDA_TRY(e, dec_confidence(ids, lens, token_logp, B, e->cfg.max_len, e->d_cls, e->g, e->cfg.max_atoms, edge_score, atom_scores,
                               seq_score, overall_score, (cudaStream_t)cuda_stream));
    e->launches += 1;
    return MNX_OK;
}

extern "C" int mnx_encode(mnx_engine* e, const float* images, int32_t B, int32_t H, int32_t W, float* features,
                          void* cuda_stream) {
    if (!e || !images || !features) return fail(e, MNX_ERR_INVALID, "mnx_encode: null argument");
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    if (e->cfg.encoder_kind == MNX_ENCODER_NONE) return fail(e, MNX_ERR_INVALID, "this handle was created without an encoder");
    if (B < 1 || B > e->cfg.max_batch) return fail(e, MNX_ERR_CAPACITY, "batch %d exceeds max_batch %d", B, e->cfg.max_batch);
    if (H < 32 || W < 32 || H > e->cfg.max_height || W > e->cfg.max_width)
        return fail(e, MNX_ERR_CAPACITY, "image %dx%d outside [32, %dx%d]", H, W, e->cfg.max_height, e->cfg.max_width);
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, (cudaStream_t)cuda_stream);
    int nl = 0;
    int rc = encoder_forward(e, e->enc, images, B, H, W, features, (cudaStream_t)cuda_stream, &nl);
    e->launches += nl;
    return rc;
}

extern "C" int mnx_preprocess(mnx_engine* e, const uint8_t* rgb, const int64_t* offsets, const int32_t* heights,
                              const int32_t* widths, int32_t B, int32_t pad, int32_t out_size, const float* mean255,
                              const float* inv_std255, float* images, void* cuda_stream) {
    if (!e || !rgb || !offsets || !heights || !widths || !mean255 || !inv_std255 || !images)
        return fail(e, MNX_ERR_INVALID, "mnx_preprocess: null argument");
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    if (B < 1 || B > e->cfg.max_batch) return fail(e, MNX_ERR_CAPACITY, "batch %d exceeds max_batch %d", B, e->cfg.max_batch);
    if (pad < 0 || out_size < 1) return fail(e, MNX_ERR_INVALID, "mnx_preprocess: bad pad / out_size");
    std::vector<unsigned long long> off(B);
    for (int i = 0; i < B; ++i) {
        if (heights[i] < 1 || widths[i] < 1 || offsets[i] < 0) return fail(e, MNX_ERR_INVALID, "image %d has a bad size or offset", i);
        off[i] = (unsigned long long)offsets[i];
    }
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, (cudaStream_t)cuda_stream);
    int nl = 0;
    CUDA_TRY(e, pp_run(rgb, off.data(), heights, widths, B, pad, out_size, mean255, inv_std255, e->pp_bbox, images,
                       (cudaStream_t)cuda_stream, &nl));
    e->launches += nl;
    return MNX_OK;
}

extern "C" int mnx_predict(mnx_engine* e, const float* images, int32_t B, int32_t H, int32_t W, int32_t* ids,
                           int32_t* lens, float* token_logp, int32_t* atom_idx, int32_t* n_atoms, uint8_t* edges,
                           void* cuda_stream) {
    if (!e) return MNX_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    ON_ENGINE_DEVICE(e);
    WorkMark work_mark(e, s);   // recorded behind the result copies below
    int rc = mnx_encode(e, images, B, H, W, e->features, s);
    if (rc != MNX_OK) return rc;
    const int S = encoder_seq_len(e->cfg.encoder_kind, H, W);
    rc = decode_internal(e, e->features, B, S, s);
    if (rc != MNX_OK) return rc;
    rc = mnx_atom_indices(e, e->ids, e->lens, B, e->atom_idx, e->n_atoms, s);
    if (rc != MNX_OK) return rc;
    rc = mnx_edges(e, nullptr, e->atom_idx, e->n_atoms, B, e->edges, nullptr, s);
    if (rc != MNX_OK) return rc;
    const size_t T = e->cfg.max_len, KA = e->cfg.max_atoms;
    if (ids) CUDA_TRY(e, cudaMemcpyAsync(ids, e->ids, sizeof(int) * B * T, cudaMemcpyDefault, s));
    if (lens) CUDA_TRY(e, cudaMemcpyAsync(lens, e->lens, sizeof(int) * B, cudaMemcpyDefault, s));
    if (token_logp) CUDA_TRY(e, cudaMemcpyAsync(token_logp, e->logp, sizeof(float) * B * T, cudaMemcpyDefault, s));
    if (atom_idx) CUDA_TRY(e, cudaMemcpyAsync(atom_idx, e->atom_idx, sizeof(int) * B * KA, cudaMemcpyDefault, s));
    if (n_atoms) CUDA_TRY(e, cudaMemcpyAsync(n_atoms, e->n_atoms, sizeof(int) * B, cudaMemcpyDefault, s));
    if (edges) CUDA_TRY(e, cudaMemcpyAsync(edges, e->edges, B * KA * KA, cudaMemcpyDefault, s));
    return MNX_OK;
}

extern "C" int mnx_predict_host(mnx_engine* e, const float* images_host, int32_t B, int32_t H, int32_t W,
                                int32_t* ids_host, int32_t* lens_host, float* token_logp_host, int32_t* atom_idx_host,
                                int32_t* n_atoms_host, uint8_t* edges_host) {
    if (!e || !images_host) return fail(e, MNX_ERR_INVALID, "mnx_predict_host: null argument");
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    if (e->cfg.encoder_kind == MNX_ENCODER_NONE) return fail(e, MNX_ERR_INVALID, "this handle was created without an encoder");
    if (B < 1 || B > e->cfg.max_batch || H > e->cfg.max_height || W > e->cfg.max_width)
        return fail(e, MNX_ERR_CAPACITY, "request (%d,%d,%d) exceeds the sizes given at create", B, H, W);
    ON_ENGINE_DEVICE(e);
    cudaStream_t s = e->cap_stream;   // engine-owned non-blocking stream (never the legacy default stream)
    CUDA_TRY(e, cudaStreamWaitEvent(s, e->ev_last, 0));   // earlier calls on this handle, whatever stream they used
    CUDA_TRY(e, cudaMemcpyAsync(e->images, images_host, sizeof(float) * (size_t)B * 3 * H * W, cudaMemcpyHostToDevice, s));
    int rc = mnx_predict(e, e->images, B, H, W, ids_host, lens_host, token_logp_host, atom_idx_host, n_atoms_host,
                         edges_host, s);
    if (rc != MNX_OK) return rc;
    CUDA_TRY(e, cudaStreamSynchronize(s));
    return MNX_OK;
}

extern "C" int mnx_beam_trace(mnx_engine* e, int32_t* trace_host, int32_t B) {
    if (!e || !trace_host) return fail(e, MNX_ERR_INVALID, "mnx_beam_trace: null argument");
    if (e->cfg.max_beam < 2 || e->last_beam_B == 0 || B != e->last_beam_B)
        return fail(e, MNX_ERR_INVALID, "mnx_beam_trace: B must equal the batch of the last mnx_decode_beam (%d)", e->last_beam_B);
    ON_ENGINE_DEVICE(e);
    CUDA_TRY(e, cudaMemcpy(trace_host, e->bm.trace, sizeof(int) * (size_t)e->cfg.max_len * B * MNX_MAX_BEAM, cudaMemcpyDeviceToHost));
    return MNX_OK;
}

extern "C" int64_t mnx_launch_count(const mnx_engine* e) { return e ? e->launches : 0; }
extern "C" int32_t mnx_last_decode_steps(const mnx_engine* ce) {
    mnx_engine* e = const_cast<mnx_engine*>(ce);
    if (!e) return 0;
    if (e->steps_pending) {      // cluster paths: wait for the decode kernel and fetch its step count
        int steps = 0;
        DeviceGuard guard(e->cfg.device);
        if (cudaMemcpy(&steps, e->steps_run_dev, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess) e->last_steps = steps;
        e->steps_pending = false;
    }
    return e->last_steps;
}

extern "C" int mnx_set_encoder_cta_limit(mnx_engine* e, int32_t n) {
    if (!e || n < 0) return fail(e, MNX_ERR_INVALID, "mnx_set_encoder_cta_limit: bad argument");
    e->enc.cta_limit = n;      // per handle: picked up by the next mnx_encode / mnx_predict of THIS engine
    return MNX_OK;
}

extern "C" int mnx_time_kernel(mnx_engine* e, int32_t which, int32_t iters, float* ms, void* cuda_stream) {
    if (!e || !ms || iters < 1) return fail(e, MNX_ERR_INVALID, "mnx_time_kernel: bad argument");
    if (which == 1002) { *ms = (float)e->max_clusters16; return MNX_OK; }
    if (which == 1004) { *ms = (float)e->max_clusters_w; return MNX_OK; }
    if (which == 1006) {   // start (iters = 1) / stop (iters = 2) recording CUDA events around every cluster decode launch
        e->time_launches = (iters == 1);
        *ms = 0.f;
        return MNX_OK;
    }
    if (which == 1005) {   // mean device time of the cluster decode launches recorded since the last call (waits for them)
        double total = 0.0;
        int n = 0;
        for (auto& p : e->launch_events) {
            float t = 0.f;
            if (cudaEventSynchronize(p.second) == cudaSuccess && cudaEventElapsedTime(&t, p.first, p.second) == cudaSuccess) { total += t; ++n; }
            cudaEventDestroy(p.first); cudaEventDestroy(p.second);
        }
        e->launch_events.clear();
        *ms = n ? (float)(total / n) : 0.f;
        return MNX_OK;
    }
    if (which == 1003) { *ms = (float)e->last_path; return MNX_OK; }
    if (which == 1000) { *ms = (float)e->max_clusters; return MNX_OK; }   // introspection: co-resident 8-CTA clusters
    if (which == 1001) {   // dump the cycle stamps recorded by the last profiled cluster decode (MNX_DECODE_PROFILE=1)
        long long h[64];
        CUDA_TRY(e, cudaMemcpy(h, e->prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
        for (int i = 1; i < 64 && h[i] != 0; ++i) fprintf(stderr, "mark %2d: +%lld cycles (total %lld)\n", i, h[i] - h[i - 1], h[i] - h[0]);
        *ms = 0.f;
        return MNX_OK;
    }
    if (which == 1008) { e->decode_profile = iters == 1; *ms = 0.f; return MNX_OK; }   // cycle stamps: iters 1 = on, 2 = off (as MNX_DECODE_PROFILE)
    if (which == 1009) {   // cycles between stamp iters-1 and stamp iters of the last profiled cluster decode (cluster 0, CTA 0, step 100, layer 1)
        long long h[64];
        if (iters < 1 || iters > 63) return fail(e, MNX_ERR_INVALID, "stamp index out of range");
        CUDA_TRY(e, cudaMemcpy(h, e->prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
        *ms = (h[iters] != 0 && h[iters - 1] != 0) ? (float)(h[iters] - h[iters - 1]) : 0.f;
        return MNX_OK;
    }
    if (!e->finalized) return fail(e, MNX_ERR_INVALID, "weights not finalized");
    ON_ENGINE_DEVICE(e);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (which >= 100) return encoder_time_kernel(e, e->enc, which, iters, ms, s);
    if (e->last_B == 0) return fail(e, MNX_ERR_INVALID, "run a decode first: timing uses its shapes and caches");
    if (which == 7) {   // the persistent cluster decode kernel alone, on the K/V of the last call
        if (e->last_path != 2 && e->last_path != 3 && e->last_path != 5 && e->last_path != 6) return fail(e, MNX_ERR_INVALID, "last decode did not use a cluster kernel");
        const bool use16 = e->last_path == 3, use16s = e->last_path == 5, usew = e->last_path == 6;
        const int B = e->last_B, S = e->last_S, T = e->cfg.max_len;
        const int growsw = (e->wide_rows > 0 && e->wide_rows < MGW_GMAX_H) ? e->wide_rows : MGW_GMAX_H;
        const int usable = usew ? (B + growsw - 1) / growsw : use16s ? (e->max_clusters16s < 8 ? e->max_clusters16s : 8)
                                  : use16 ? (e->max_clusters16 < 8 ? e->max_clusters16 : 8) : (e->max_clusters < 16 ? e->max_clusters : 16);
        const int G = (B + usable - 1) / usable, clusters = (B + G - 1) / G;
        if (usew && clusters > e->max_clusters_w)
            return fail(e, MNX_ERR_INVALID, "isolated timing covers single-launch decodes only (%d clusters, %d resident)", clusters, e->max_clusters_w);
        MegaArgs a{};
        a.wpack = e->wpack; a.ppack = e->ppack; a.wpack16 = e->wpack16; a.ppack16 = e->ppack16;
        a.wpackW = e->wpackW; a.ticket = e->ticket;
        a.finalp = e->finalp; a.emb = e->dw.emb; a.pe = e->dw.pe;
        a.selfK = e->selfK; a.selfV = e->selfV; a.crossK = e->crossK; a.crossV = e->crossV;
        a.B = B; a.S = S; a.T = T; a.G = G;
        a.ids = e->ids; a.logp = e->logp; a.hidden = e->hidden; a.lens = e->lens;
        a.row_state = e->row_state; a.steps_run = e->steps_run_dev; a.g = e->g; a.prof = nullptr;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float total = 0.f;
        for (int i = 0; i < iters; ++i) {
            CUDA_TRY(e, cudaMemsetAsync(e->row_state, 0, sizeof(unsigned) * B, s));
            CUDA_TRY(e, cudaMemsetAsync(e->steps_run_dev, 0, sizeof(int), s));
            CUDA_TRY(e, cudaMemsetAsync(e->ticket, 0, sizeof(int), s));
            cudaEventRecord(e0, s);
            CUDA_TRY(e, usew ? wide_launch(a, clusters, s) : use16s ? mega16s_launch(a, clusters, s) : use16 ? mega16_launch(a, clusters, s) : mega_launch(a, clusters, s));
            cudaEventRecord(e1, s);
            CUDA_TRY(e, cudaStreamSynchronize(s));
            float t = 0.f;
            cudaEventElapsedTime(&t, e0, e1);
            total += t;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        *ms = total / iters;
        return MNX_OK;
    }
    DecBuffers b = make_buffers(e, e->last_B, e->last_S);
    cudaError_t c = dec_time_kernel(which, iters, b, e->dw, e->g, e->cfg.max_len / 2, ms, s);
    if (c == cudaErrorInvalidValue) return fail(e, MNX_ERR_INVALID, "unknown kernel id %d", which);
    CUDA_TRY(e, c);
    return MNX_OK;
}
