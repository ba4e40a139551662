// engine.h -- the engine object behind the C-ABI (include/parakeet_b200.h), shared by engine.cu (offline path,
// jobs) and stream_engine.cu (streaming eou path).  See engine.cu for the reference call stack being replaced.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/parakeet_b200.h"
#include "kernels.h"
#include "nccl_dl.h"
#include "safetensors.h"

using namespace pk;


namespace pk_detail {

std::string &create_err();   // last error of a failed pk_engine_create / engine-less call (thread-local, engine.cu)

#define PK_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            return fail(PK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
        }                                                                                     \
    } while (0)

inline int conv_len(int L) { return (L - 1) / 2 + 1; }  // k3 s2 p1 (operations.cpp:3191-3196)

// A linear layer's parameters on the device: fp32 master [N][K] + bias, and (tcgen05
// modes) the bf16 hi/lo split planes of the weight.
struct GemmWeight {
    float *w = nullptr;
    bf16 *hi = nullptr, *lo = nullptr;
    float *bias = nullptr;
    int N = 0, K = 0;
    TcOperand tc;   // TMA tensor maps of hi/lo (tcgen05 modes)
};

// An activation buffer that feeds GEMMs, with its TMA tensor maps (tcgen05 modes).
struct Act : ActBuf {
    TcOperand tc;
    TcOperand tc32, tc64;   // the same planes with a 32- / 64-row box (A tiles fetched in slices and TMA-multicast across a 4- / 2-CTA cluster)
};

struct LayerW {
    float *ffn_ln_w[2], *ffn_ln_b[2];
    GemmWeight fc1[2], fc2[2];
    float *att_ln_w, *att_ln_b;
    GemmWeight qkv, out;
    float *pos_u, *pos_v;
    float *pp;  // [(2*Tmax-1)][d] projected relative-position table
    bf16 *pp_hi = nullptr, *pp_lo = nullptr;   // its bf16 split planes (tensor-core attention)
    TcOperand pp_tc;                           // their TMA maps, box 64 x 256 (tcgen05 attention: one box = every relative position of a 128 x 128 tile)
    bool pp_tc_ok = false;
    float *conv_ln_w, *conv_ln_b;
    GemmWeight pw1, pw2;
    float *dw_w, *dw_b;  // BatchNorm folded; [d][k]
    float *dw_wt;        // the same weights tap-major [k][d] (offline kernel: one float4 per tap and 4 channels)
    float *fin_ln_w, *fin_ln_b;
};

}  // namespace pk_detail
using namespace pk_detail;

struct pk_engine {
    pk_config cfg;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_h2d = nullptr;              // recorded after the staging copies of a batch
    // Host PCM arrives in H2D_CHUNKS utterance groups on `copy_stream`; the front end (mel, conv1+dw1)
    // of group i runs on `stream` as soon as its samples have landed, i.e. under the DMA of group i+1.
    static constexpr int H2D_CHUNKS = 8;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_chunk[H2D_CHUNKS] = {}, ev_front = nullptr;
    bool front_done = false;                   // front end of the staged batch already launched (chunked path)
    std::string err;
    int64_t launches = 0;
    std::vector<void *> allocs;
    void *l2_scratch = nullptr;
    size_t l2_scratch_bytes = 0;
    int64_t l2_flushes = 0;

    // ---- capacity
    int Bmax = 0, Fmax = 0, Tmax = 0;          // per-utterance max mel frames / encoder frames
    int f1n = 0, f2n = 0, f3n = 0;
    int cap = 0;                               // token capacity per utterance

    // ---- weights
    MelTables mel_tb{};
    float *c1_w, *c1_b, *dw1_w, *dw1_b, *dw2_w, *dw2_b;
    float *dw2_wt = nullptr;                   // dw2_ weights tap-major [9][C]: one float4 per tap and 4 channels
    GemmWeight conv2, conv3, proj;
    std::vector<LayerW> layers;
    GemmWeight ctc_head;
    GemmWeight enc_proj;                       // joint enc_proj_ [J][d] + bias
    float *G0 = nullptr;                       // [V][4P]
    float *Whh[PK_MAX_LSTM] = {}, *Wih[PK_MAX_LSTM] = {}, *bih[PK_MAX_LSTM] = {};
    float *Whh_um[PK_MAX_LSTM] = {}, *Wih_um[PK_MAX_LSTM] = {};   // unit-major copies (decode kernel)
    float *Wp = nullptr, *Wout = nullptr, *bout = nullptr;
    bf16 *Whh_s[PK_MAX_LSTM] = {}, *Wih_s[PK_MAX_LSTM] = {}, *Wp_s = nullptr, *Wout_s = nullptr;   // pre-split rows (tdt.cu)

    // ---- workspace
    float *d_pcm = nullptr;
    float *d_pcm_alt = nullptr;               // second PCM buffer (pk_prefetch_pcm); swapped with d_pcm on adoption
    cudaEvent_t ev_pcm_free[2] = {}, ev_prefetch = nullptr;   // [k]: last front end reading buffer k has run; prefetch copy done
    int pcm_cur = 0;                           // which physical buffer d_pcm currently is
    struct { const float *pcm = nullptr; int32_t n = 0; std::vector<int64_t> off; bool valid = false; } pref;
    int64_t *d_pcm_off = nullptr;
    int32_t *d_frame_off = nullptr, *d_s2_off = nullptr, *d_row_off = nullptr, *d_t2_rows = nullptr;
    float *logmel = nullptr, *feats = nullptr;
    float *mel_part = nullptr;                 // per-chunk statistics of the mel normalisation (mel.cu K2), one slice per utterance
    Act sub1, sub3, sub4, ln, ffh, ctx, cv;
    float *sub2 = nullptr, *x = nullptr, *qkv = nullptr, *glu = nullptr, *logits = nullptr, *EP = nullptr;
    bf16 *qkvp_hi = nullptr, *qkvp_lo = nullptr;   // [Mx, 2 d] planes [k | v] for the tensor-core attention (q stays fp32 in `qkv`)
    int32_t *best = nullptr;
    float *bconf = nullptr;
    int32_t *tok = nullptr, *t_start = nullptr, *t_end = nullptr;
    float *t_conf = nullptr;
    // TDT state
    int Bpad = 0;
    float *hbuf = nullptr, *cbuf = nullptr, *zbuf = nullptr, *pl_max = nullptr, *pl_sum = nullptr;
    int32_t *tdt_ints = nullptr;               // overflow[Bpad] | barrier counter
    unsigned long long *tdt_keys = nullptr;    // arg-max keys: label[3][Bpad] | duration[3][Bpad]
    // pinned host staging
    float *h_pcm = nullptr;
    int32_t *h_meta = nullptr;                 // offsets staging
    int32_t *h_tok = nullptr, *h_ts = nullptr, *h_te = nullptr;
    float *h_tc = nullptr;

    // ---- CUDA graphs of the staged pipeline, keyed by (decoder, utterance lengths)
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; int seen = 0; };
    std::map<std::string, GraphEntry> graphs;
    bool use_graphs = true;
    bool attn_umma = true;                     // tcgen05 attention (attention_umma.cu) for head_dim 64 and batches of <= 128-frame utterances (PK_ATTN_UMMA=0: mma.sync kernel)
    std::map<int, TcOperand> kv_maps;          // TMA maps of the k | v planes, keyed by the number of rows of the batch
    bool attn_tc = true;                       // mma.sync attention for head_dim 64 / 128 (PK_ATTN_TC=0: fp32 kernel)

    // ---- the staged batch
    int n_utt = 0;
    std::vector<int64_t> pcm_off;
    std::vector<int32_t> frame_off, s2_off, row_off, t2_rows;
    int maxF = 0, maxT2 = 0, maxT = 0, M = 0, M2 = 0;

    // ---- a JOB: many micro-batches on this GPU, ONE exchange at the end (BASELINE configs[4])
    int32_t *job_tok = nullptr, *job_all = nullptr;   // [job_cap_rows][1 + cap] local rows; [job_world * job_cap_rows][1 + cap] gathered
    int64_t job_cap_rows = 0, job_rows = 0, job_alloc_rows = 0;   // rows per rank of this job / appended so far / allocated (x world)
    int job_world = 1;
    int32_t *h_job = nullptr;                          // pinned staging of the gathered rows
    size_t h_job_ints = 0;
    float *job_pcm = nullptr;                          // device-resident PCM of a whole job (pk_job_stage_pcm)
    size_t job_pcm_cap = 0;
    std::vector<int64_t> job_off;
    const float *pcm_src = nullptr;                    // front end reads this instead of d_pcm (a slice of job_pcm)
    float *d_raw = nullptr;                            // input at its own sample rate (pk_stage_pcm_rate / pk_resample_batch)
    size_t d_raw_cap = 0;
    int64_t *d_raw_off = nullptr;                      // [Bmax + 1] offsets of the raw utterances
    void *nccl_comm = nullptr;                         // ncclComm_t (pk_comm_init_rank) -- owned
    int nccl_rank = 0, nccl_world = 1;
    bool last_tdt = false;                             // the token buffer holds a TDT decode (overflow flags are valid)
    int32_t truncated = 0;                             // utterances of the last fetch whose TDT hypothesis hit the token capacity

    // ---- phrase boosting (pk_set_boost): ContextTrie on the device + per-utterance trie state of the TDT kernel
    DeviceTrie trie{};
    bool boost_on = false;
    float boost = 0.f;
    int boost_gen = 0;                                 // bumps on every pk_set_boost (part of the CUDA-graph key)
    uint32_t *boost_bits = nullptr;                    // [Bpad][(V+31)/32]
    int32_t *trie_active = nullptr, *trie_nact = nullptr;

    // ---- optional per-kernel-class timing (CUDA events on the engine stream)
    enum { CAT_MEL, CAT_SUBSAMPLE, CAT_GEMM, CAT_LAYERNORM, CAT_ATTENTION, CAT_DWCONV, CAT_CTC, CAT_TDT, CAT_N };
    struct ProfRec { int cat; cudaEvent_t a, b; double flops; };
    bool prof_on = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t prof_event() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    struct Scope {
        pk_engine *e; int idx = -1;
        Scope(pk_engine *e_, int cat, double flops = 0.0) : e(e_) {
            if (!e->prof_on) return;
            ProfRec r{cat, e->prof_event(), e->prof_event(), flops};
            cudaEventRecord(r.a, e->stream);
            idx = (int)e->prof.size();
            e->prof.push_back(r);
        }
        ~Scope() { if (idx >= 0) cudaEventRecord(e->prof[idx].b, e->stream); }
    };

    pk_status fail(pk_status s, const std::string &m) {
        err = m;
        return s;
    }
    template <typename T>
    T *dalloc(size_t n) {
        void *p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
        allocs.push_back(p);
        return static_cast<T *>(p);
    }
    template <typename T>
    T *upload(const std::vector<T> &h) {
        T *d = dalloc<T>(h.size());
        // On the engine's own (non-blocking) stream, then wait: a legacy-stream cudaMemcpy from
        // pageable memory may still be in flight when a kernel on `stream` starts.
        if (d && !h.empty()) {
            cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream);
            cudaStreamSynchronize(stream);
        }
        return d;
    }
    // [rows][K] activation that feeds a GEMM as the A operand
    Act act_alloc(size_t rows, size_t K) {
        Act a;
        const size_t n = rows * K;
        if (cfg.math == PK_MATH_FP32) {
            a.f32 = dalloc<float>(n);
        } else {
            a.hi = dalloc<bf16>(n);
            if (cfg.math == PK_MATH_BF16X3) a.lo = dalloc<bf16>(n);
            if (a.hi && (!make_tc_operand(&a.tc, a.hi, a.lo, rows, K, 128) || !make_tc_operand(&a.tc32, a.hi, a.lo, rows, K, 32) || !make_tc_operand(&a.tc64, a.hi, a.lo, rows, K, 64))) a.hi = nullptr;   // reported by the caller
        }
        return a;
    }

    pk_status load(const char *path);
    pk_status make_weight(const SafeTensors &st, const std::string &wname, const std::string &bname, int N, int K,
                          GemmWeight &out, const std::vector<int> *row_perm = nullptr,
                          const std::vector<int> *col_perm = nullptr);
    pk_status finish_weight(std::vector<float> &w, std::vector<float> *b, int N, int K, GemmWeight &out);
    pk_status get_vec(const SafeTensors &st, const std::string &name, int n, float **out);
    pk_status alloc_workspace();
    pk_status set_batch_shapes(const int32_t *n_frames_or_null, const int64_t *offsets_or_null, int n);
    pk_status upload_shapes();
    void gemm(const Act &A, int lda, const GemmWeight &W, int M_, EpiParams epi);
    // x = resid + alpha * (A . W^T + b) followed by LayerNorm(s): ONE kernel (gemm_tc_ln.cu) when fuse_ln applies,
    // else the residual GEMM and layernorm_kernel.  resid_in_x: the residual is x itself (false: x = A . W^T + b).
    // out_ln1: x receives LayerNorm_1 of the sum (block end) instead of the sum; planes = split of the last LayerNorm.
    pk_status gemm_ln(const Act &A, int lda, const GemmWeight &W, int M_, bool resid_in_x, float alpha, const float *ln1_w, const float *ln1_b,
                      bool out_ln1, const float *ln2_w, const float *ln2_b, ActBuf planes);
    int gemm_cluster = 0;                      // PK_GEMM_CLUSTER=2|4: wide GEMMs (fc1, q/k/v, pw1) run as clusters of 2 | 4 CTAs along N with the A tile multicast
    bool ln_mcast = false;                     // PK_LN_MCAST=1: the A tile is fetched in quarters and TMA-multicast across the cluster (measured: no gain)
    int fuse_ln_min_k = 0;                     // PK_FUSE_LN_MINK: fuse only GEMMs with K >= this (short-K launches are epilogue-bound either way)
    bool fuse_ln = false;                      // PK_FUSE_LN=1: LayerNorm in the epilogue of the GEMM that produces its input
    // output-side tensor maps of the TMA-store epilogue, keyed by (buffer, leading dimension, rows)
    std::map<std::tuple<const void *, int, int, int>, CUtensorMap> out_maps;
    const CUtensorMap *out_map(const void *ptr, bool is_f32, int rows, int ld);
    bool tma_out = true;                       // PK_GEMM_TMA_OUT=0: results leave through st.global instead
    // few-row GEMMs (M <= 128: streaming steps, short utterances) go to gemm_skinny.cu (PK_GEMM_SKINNY=0: never)
    bool skinny = true;
    float *skinny_ws = nullptr;
    size_t skinny_ws_floats = 0;
    unsigned int *skinny_tickets = nullptr;
    static constexpr int SKINNY_TICKETS = 1024;
    pk_status gemm_err = PK_OK;
    pk_status run_mel(int u0 = 0, int u1 = -1);
    pk_status run_conv1(int u0 = 0, int u1 = -1);
    pk_status run_graphed(const std::string &key, const std::function<pk_status()> &body);
    pk_status run_subsample_tail(bool with_first_ln = false);
    pk_status run_encoder(float *sub_out_host, float *layers_out_host);
    // streaming eou path (stream_engine.cu)
    struct StreamSet *ss = nullptr;
    pk_status run_stream_layers();
    pk_status run_stream_decode();
    pk_status run_ctc(float *logprobs_dev_or_null);
    pk_status run_tdt();
    pk_status fetch(pk_tokens *out);
};

void pk_stream_free(pk_engine *e);   // stream_engine.cu
