// oracle/ref_harness_stream.cpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// C-ABI around the UNMODIFIED reference's STREAMING path (eou-120m), compiled into
// oracle/_ref/libpkref.so with the rest of the harness.  One handle = one stream; a chunk call does exactly
// what StreamingTranscriber::transcribe_chunk does (src/eou.cpp:112-145), with the intermediate tensors
// copied out:
//   StreamingAudioPreprocessor::process_chunk          src/audio.cpp:195-259
//   StreamingFastConformerEncoder::forward_chunk       src/streaming_encoder.cpp:425-472
//     (CausalConvSubsampling::forward_cached :339-378, StreamingConformerBlock::forward_cached :289-301,
//      StreamingConformerAttention::forward_cached :160-272, CausalConformerConvModule::forward_cached :41-80)
//   rnnt_streaming_decode_chunk                        src/eou.cpp:17-98
// Two facts about the reference's CPU streaming path that this harness has to live with (found while pinning
// the oracle; both are documented in oracle/oracle.py and DESIGN.md):
//   * under axiom's default LAZY evaluation forward_cached crashes in ops::masked_fill (Tensor::copy of an
//     unmaterialised score tensor -> CPUStorage::copy_from(null)); the calls below therefore run inside
//     axiom::graph::EagerModeScope, the reference's own switch (graph_registry.hpp:144);
//   * the bounded-context attention mask has no effect on CPU (float mask read bytewise by masked_fill).
// Only tests/ and golden generators may load this library.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <axiom/axiom.hpp>
#include <axiom/graph/graph_registry.hpp>
#include <axiom/io/safetensors.hpp>

#include "parakeet/audio.hpp"
#include "parakeet/eou.hpp"

using namespace parakeet;
using axiom::Shape;
using axiom::Tensor;

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
namespace {
void segv_handler(int) {
    void *bt[64];
    int n = backtrace(bt, 64);
    backtrace_symbols_fd(bt, n, 2);
    _exit(139);
}
struct RefStream {
    EOUConfig cfg;
    std::unique_ptr<ParakeetEOU> model;
    std::map<std::string, Tensor> weights;
    StreamingAudioPreprocessor pre;
    EncoderCache cache;
    StreamingDecodeState st;
};
thread_local std::string g_serr;
}  // namespace

extern "C" {

const char *pkref_stream_last_error() { return g_serr.c_str(); }

// dims <= 0 keep make_eou_120m_config()'s values
void *pkref_stream_new(const char *weights_path, int mel, int sub_ch, int d, int layers, int heads, int ff, int vocab,
                       int pred_hidden, int lstm_layers, int joint_hidden, int ctx_left, int ctx_right) {
    try {
        auto s = std::make_unique<RefStream>();
        auto &c = s->cfg;
        c = make_eou_120m_config();
        if (mel > 0) c.encoder.mel_bins = mel;
        if (sub_ch > 0) c.encoder.subsampling_channels = sub_ch;
        if (d > 0) { c.encoder.hidden_size = d; c.joint.encoder_hidden = d; }
        if (layers > 0) c.encoder.num_layers = layers;
        if (heads > 0) c.encoder.num_heads = heads;
        if (ff > 0) c.encoder.ffn_intermediate = ff;
        if (vocab > 0) { c.prediction.vocab_size = vocab; c.joint.vocab_size = vocab; }
        if (pred_hidden > 0) { c.prediction.pred_hidden = pred_hidden; c.joint.pred_hidden = pred_hidden; }
        if (lstm_layers > 0) c.prediction.num_lstm_layers = lstm_layers;
        if (joint_hidden > 0) c.joint.joint_hidden = joint_hidden;
        if (ctx_left >= 0) c.encoder.att_context_left = ctx_left;
        if (ctx_right >= 0) c.encoder.att_context_right = ctx_right;
        AudioConfig ac;
        ac.n_mels = c.encoder.mel_bins;
        s->pre = StreamingAudioPreprocessor(ac);
        s->weights = axiom::io::safetensors::load(weights_path);
        s->model = std::make_unique<ParakeetEOU>(c);
        s->model->load_state_dict(s->weights, "", false);
        return s.release();
    } catch (const std::exception &e) {
        g_serr = e.what();
        return nullptr;
    }
}

void pkref_stream_free(void *h) { delete static_cast<RefStream *>(h); }

// One chunk of PCM.  Outputs (each may be empty): log-mel frames [n_frames][mel], encoder frames
// [n_enc][d], new tokens (id, start, end) + confidences.  Returns 0, or -1 (see pkref_stream_last_error).
int pkref_stream_chunk(void *h, const float *pcm, int n, float *feats_out, int feats_cap_frames, int *n_frames,
                       float *enc_out, int enc_cap_frames, int *n_enc, int32_t *tok_out, float *conf_out, int tok_cap,
                       int *n_tok) {
    auto *s = static_cast<RefStream *>(h);
    *n_frames = *n_enc = *n_tok = 0;
    try {
        axiom::graph::EagerModeScope eager;
        const bool dbg = getenv("PKREF_STREAM_DEBUG") != nullptr;
        if (dbg) signal(SIGSEGV, segv_handler);
        auto samples = Tensor::from_data(pcm, Shape{(size_t)n}, true);
        if (dbg) fprintf(stderr, "[pkref_stream] process_chunk n=%d\n", n);
        auto feats = s->pre.process_chunk(samples);
        if (dbg) fprintf(stderr, "[pkref_stream] feats ok storage=%d\n", (int)(bool)feats.storage());
        if (!feats.storage()) return 0;
        const int nf = (int)feats.shape()[1], nm = (int)feats.shape()[2];
        if (nf > feats_cap_frames) { g_serr = "feats capacity"; return -1; }
        *n_frames = nf;
        auto fc = feats.cpu().ascontiguousarray();
        std::memcpy(feats_out, fc.typed_data<float>(), (size_t)nf * nm * sizeof(float));
        if (dbg) fprintf(stderr, "[pkref_stream] forward_chunk\n");
        auto enc = s->model->encoder().forward_chunk(feats, s->cache);
        if (dbg) fprintf(stderr, "[pkref_stream] enc ok\n");
        if (!enc.storage() || enc.shape().size() == 0) return 0;
        const int ne = (int)enc.shape()[1], d = (int)enc.shape()[2];
        if (ne > enc_cap_frames) { g_serr = "enc capacity"; return -1; }
        *n_enc = ne;
        auto ec = enc.cpu().ascontiguousarray();
        std::memcpy(enc_out, ec.typed_data<float>(), (size_t)ne * d * sizeof(float));
        const size_t before = s->st.timestamped_tokens.size();
        // blank = vocab - 1, passed explicitly: the default argument (1024, eou.hpp:94) is only right for the
        // 1025-label preset and indexes past the embedding table of a smaller test vocabulary
        rnnt_streaming_decode_chunk(s->model->prediction(), s->model->joint(), enc, s->cfg.durations, s->st,
                                    s->cfg.joint.vocab_size - 1);
        const size_t after = s->st.timestamped_tokens.size();
        if ((int)(after - before) > tok_cap) { g_serr = "token capacity"; return -1; }
        for (size_t i = before; i < after; ++i) {
            const auto &t = s->st.timestamped_tokens[i];
            tok_out[3 * (i - before) + 0] = t.token_id;
            tok_out[3 * (i - before) + 1] = t.start_frame;
            tok_out[3 * (i - before) + 2] = t.end_frame;
            conf_out[i - before] = t.confidence;
        }
        *n_tok = (int)(after - before);
        return 0;
    } catch (const std::exception &e) {
        g_serr = e.what();
        return -1;
    }
}

}  // extern "C"
