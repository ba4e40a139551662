// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// A thin C-ABI around the UNMODIFIED reference (compiled from /root/reference
// by oracle/Makefile into oracle/_ref/libpkref.so) so Python tests / bench.py
// can run the reference's own CPU implementation of the hot path and read its
// intermediate tensors.  Every entry point only *calls* reference functions:
//   preprocess_audio                     src/audio.cpp:100-158
//   ConvSubsampling / ConformerBlock     src/encoder.cpp:196-241
//   FastConformerEncoder::forward        src/encoder.cpp:253-271
//   sinusoidal_position_embedding        src/encoder.cpp:9-30
//   CTCDecoder::forward                  src/ctc.cpp:12-25
//   ctc_greedy_decode(_with_timestamps)  src/ctc.cpp:40-127
//   tdt_greedy_decode(_with_timestamps)  src/tdt.cpp:36-201
//   Tokenizer / group_timestamps         src/vocab.cpp, src/timestamp.cpp
//   ContextTrie, *_decode_boosted        src/phrase_boost.cpp (next-row groundwork)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <axiom/axiom.hpp>
#include <axiom/io/safetensors.hpp>

#include "parakeet/audio.hpp"
#include "parakeet/audio_io.hpp"
#include "parakeet/config.hpp"
#include "parakeet/ctc.hpp"
#include "parakeet/encoder.hpp"
#include "parakeet/phrase_boost.hpp"
#include "parakeet/tdt.hpp"
#include "parakeet/tdt_ctc.hpp"
#include "parakeet/timestamp.hpp"
#include "parakeet/vocab.hpp"

using namespace parakeet;
using axiom::Shape;
using axiom::Tensor;

namespace {

struct RefModel {
    int preset = 0; // 0 = tdt-ctc-110m, 1 = tdt-600m
    TDTCTCConfig cfg110;
    TDTConfig cfg600;
    std::unique_ptr<ParakeetTDTCTC> m110;
    std::unique_ptr<ParakeetTDT> m600;
    std::map<std::string, Tensor> weights;
    Tokenizer tok;

    const EncoderConfig &enc_cfg() const { return preset == 0 ? cfg110.encoder : cfg600.encoder; }
    FastConformerEncoder &encoder() { return preset == 0 ? m110->encoder() : m600->encoder(); }
    RNNTPrediction &prediction() { return preset == 0 ? m110->prediction() : m600->prediction(); }
    TDTJoint &joint() { return preset == 0 ? m110->tdt_joint() : m600->joint(); }
    const std::vector<int> &durations() const { return preset == 0 ? cfg110.durations : cfg600.durations; }
    int blank() const { return (preset == 0 ? cfg110.joint.vocab_size : cfg600.joint.vocab_size) - 1; }
};

thread_local std::string g_err;

void copy_out(const Tensor &t, float *dst) {
    auto c = t.cpu().ascontiguousarray();
    std::memcpy(dst, c.typed_data<float>(), c.size() * sizeof(float));
}

Tensor feats_tensor(const float *f, int n_frames, int n_mels) {
    return Tensor::from_data(f, Shape{1, (size_t)n_frames, (size_t)n_mels}, true);
}

} // namespace

extern "C" {

const char *pkref_last_error() { return g_err.c_str(); }

void *pkref_load(const char *weights_path, const char *vocab_path, int preset) {
    try {
        auto m = std::make_unique<RefModel>();
        m->preset = preset;
        m->weights = axiom::io::safetensors::load(weights_path);
        if (preset == 0) {
            m->cfg110 = make_110m_config();
            m->m110 = std::make_unique<ParakeetTDTCTC>(m->cfg110);
            m->m110->load_state_dict(m->weights, "", false);
        } else {
            m->cfg600 = make_tdt_600m_config();
            m->m600 = std::make_unique<ParakeetTDT>(m->cfg600);
            m->m600->load_state_dict(m->weights, "", false);
        }
        if (vocab_path && vocab_path[0]) m->tok.load(vocab_path);
        return m.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

// Same as preset 0 (ParakeetTDTCTC) but with explicit dimensions (test-only tiny shapes).
void *pkref_load_custom(const char *weights_path, const char *vocab_path, int mel, int sub_ch, int d,
                        int layers, int heads, int ff, int vocab, int pred_hidden, int lstm_layers,
                        int joint_hidden) {
    try {
        auto m = std::make_unique<RefModel>();
        m->preset = 0;
        m->weights = axiom::io::safetensors::load(weights_path);
        auto &c = m->cfg110;
        c = make_110m_config();
        c.encoder.mel_bins = mel;
        c.encoder.subsampling_channels = sub_ch;
        c.encoder.hidden_size = d;
        c.encoder.num_layers = layers;
        c.encoder.num_heads = heads;
        c.encoder.ffn_intermediate = ff;
        c.prediction.vocab_size = vocab;
        c.prediction.pred_hidden = pred_hidden;
        c.prediction.num_lstm_layers = lstm_layers;
        c.joint.encoder_hidden = d;
        c.joint.pred_hidden = pred_hidden;
        c.joint.joint_hidden = joint_hidden;
        c.joint.vocab_size = vocab;
        c.ctc_vocab_size = vocab;
        m->m110 = std::make_unique<ParakeetTDTCTC>(c);
        m->m110->load_state_dict(m->weights, "", false);
        if (vocab_path && vocab_path[0]) m->tok.load(vocab_path);
        return m.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void pkref_free(void *h) { delete static_cast<RefModel *>(h); }

// PCM -> normalised log-mel.  out must hold (1 + n/160) * n_mels floats.
// Returns n_frames or -1.
int pkref_mel(const float *pcm, int64_t n, int n_mels, float *out) {
    try {
        AudioConfig cfg;
        cfg.n_mels = n_mels;
        auto wav = Tensor::from_data(pcm, Shape{(size_t)n}, true);
        auto f = preprocess_audio(wav, cfg); // (1, n_frames, n_mels)
        copy_out(f, out);
        return (int)f.shape()[1];
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

int pkref_posemb(int seq_len, int d_model, float *out) {
    try {
        copy_out(sinusoidal_position_embedding(seq_len, d_model), out);
        return 2 * seq_len - 1;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// feats (n_frames, n_mels) -> encoder output (T', d).  Returns T' or -1.
int pkref_encode(void *h, const float *feats, int n_frames, int n_mels, float *out) {
    try {
        auto *m = static_cast<RefModel *>(h);
        auto y = m->encoder()(feats_tensor(feats, n_frames, n_mels));
        copy_out(y, out);
        return (int)y.shape()[1];
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// Layer-by-layer: standalone reference modules loaded from the same state dict
// under the encoder's key prefixes.  sub_out (T', d); layer_out (L, T', d).
int pkref_encode_layers(void *h, const float *feats, int n_frames, int n_mels,
                        float *sub_out, float *layer_out) {
    try {
        auto *m = static_cast<RefModel *>(h);
        const auto &ec = m->enc_cfg();
        ConvSubsampling sub(ec.subsampling_channels);
        sub.load_state_dict(m->weights, "encoder_.subsampling_.", true);
        auto x = sub(feats_tensor(feats, n_frames, n_mels));
        int T = (int)x.shape()[1], d = (int)x.shape()[2];
        if (sub_out) copy_out(x, sub_out);
        auto pos = sinusoidal_position_embedding(T, d);
        for (int i = 0; i < ec.num_layers; ++i) {
            ConformerBlock blk(ec);
            blk.load_state_dict(m->weights, "encoder_.layers_." + std::to_string(i) + ".", false);
            x = blk(x, pos);
            if (layer_out) copy_out(x, layer_out + (size_t)i * T * d);
        }
        return T;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// enc (T, d) -> CTC log-probs (T, V) via the reference CTC head (110m only).
int pkref_ctc_logprobs(void *h, const float *enc, int T, int d, float *out) {
    try {
        auto *m = static_cast<RefModel *>(h);
        if (m->preset != 0) throw std::runtime_error("no CTC head on this preset");
        auto e = Tensor::from_data(enc, Shape{1, (size_t)T, (size_t)d}, true);
        auto lp = m->m110->ctc_decoder()(e);
        copy_out(lp, out);
        return (int)lp.shape()[2];
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// Pure decode-loop entry points (no weights): log-probs (B,T,V) -> tokens.
// ids/start/end/conf hold T entries per batch row; lens[b] = count.
int pkref_ctc_greedy(const float *lp, int B, int T, int V, int blank, int with_ts,
                     int *ids, int *start, int *end, float *conf, int *lens) {
    try {
        auto t = Tensor::from_data(lp, Shape{(size_t)B, (size_t)T, (size_t)V}, true);
        if (with_ts) {
            auto r = ctc_greedy_decode_with_timestamps(t, blank);
            for (int b = 0; b < B; ++b) {
                lens[b] = (int)r[b].size();
                for (size_t i = 0; i < r[b].size(); ++i) {
                    ids[b * T + i] = r[b][i].token_id;
                    start[b * T + i] = r[b][i].start_frame;
                    end[b * T + i] = r[b][i].end_frame;
                    conf[b * T + i] = r[b][i].confidence;
                }
            }
        } else {
            auto r = ctc_greedy_decode(t, blank);
            for (int b = 0; b < B; ++b) {
                lens[b] = (int)r[b].size();
                for (size_t i = 0; i < r[b].size(); ++i) ids[b * T + i] = r[b][i];
            }
        }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// enc (T, d) -> TDT greedy tokens.  cap = capacity of ids/start/end/conf.
// Returns token count (may exceed cap; only cap are written) or -1.
int pkref_tdt_greedy(void *h, const float *enc, int T, int d, int with_ts, int cap,
                     int *ids, int *start, int *end, float *conf) {
    try {
        auto *m = static_cast<RefModel *>(h);
        auto e = Tensor::from_data(enc, Shape{1, (size_t)T, (size_t)d}, true);
        if (with_ts) {
            auto r = tdt_greedy_decode_with_timestamps(m->prediction(), m->joint(), e,
                                                       m->durations(), m->blank());
            int n = (int)r[0].size();
            for (int i = 0; i < n && i < cap; ++i) {
                ids[i] = r[0][i].token_id;
                start[i] = r[0][i].start_frame;
                end[i] = r[0][i].end_frame;
                conf[i] = r[0][i].confidence;
            }
            return n;
        }
        auto r = tdt_greedy_decode(m->prediction(), m->joint(), e, m->durations(), m->blank());
        int n = (int)r[0].size();
        for (int i = 0; i < n && i < cap; ++i) ids[i] = r[0][i];
        return n;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// Tokenizer::decode -> UTF-8 text in buf (NUL-terminated); returns length.
int pkref_detok(void *h, const int *ids, int n, char *buf, int cap) {
    try {
        auto *m = static_cast<RefModel *>(h);
        auto s = m->tok.decode(std::vector<int>(ids, ids + n));
        int k = (int)std::min<size_t>(s.size(), (size_t)cap - 1);
        std::memcpy(buf, s.data(), k);
        buf[k] = 0;
        return (int)s.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// group_timestamps(Words): words are written '\n'-separated into buf.
int pkref_group_words(void *h, const int *ids, const int *start, const int *end,
                      const float *conf, int n, char *buf, int cap, float *w_start,
                      float *w_end, float *w_conf) {
    try {
        auto *m = static_cast<RefModel *>(h);
        std::vector<TimestampedToken> toks(n);
        for (int i = 0; i < n; ++i) toks[i] = {ids[i], start[i], end[i], conf[i]};
        auto words = group_timestamps(toks, m->tok.pieces());
        std::string s;
        for (size_t i = 0; i < words.size(); ++i) {
            s += words[i].word;
            s += '\n';
            w_start[i] = words[i].start;
            w_end[i] = words[i].end;
            w_conf[i] = words[i].confidence;
        }
        int k = (int)std::min<size_t>(s.size(), (size_t)cap - 1);
        std::memcpy(buf, s.data(), k);
        buf[k] = 0;
        return (int)words.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// The reference Transcriber::transcribe path (transcribe.hpp:99-179) for one
// utterance, stage-timed: mel -> encoder -> (CTC head + greedy | TDT greedy).
// decoder: 0 = CTC, 1 = TDT.  ms[3] = {preprocess, encoder, decode}.
int pkref_transcribe(void *h, const float *pcm, int64_t n, int decoder, int cap, int *ids,
                     double *ms) {
    try {
        auto *m = static_cast<RefModel *>(h);
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto dt = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        AudioConfig cfg;
        cfg.n_mels = m->enc_cfg().mel_bins;
        auto t0 = now();
        auto wav = Tensor::from_data(pcm, Shape{(size_t)n}, true);
        auto f = preprocess_audio(wav, cfg).ascontiguousarray();
        auto t1 = now();
        auto enc = m->encoder()(f).ascontiguousarray();
        auto t2 = now();
        std::vector<std::vector<int>> toks;
        if (decoder == 0) {
            auto lp = m->m110->ctc_decoder()(enc).cpu();
            toks = ctc_greedy_decode(lp, m->blank());
        } else {
            toks = tdt_greedy_decode(m->prediction(), m->joint(), enc, m->durations(), m->blank());
        }
        auto t3 = now();
        if (ms) { ms[0] = dt(t0, t1); ms[1] = dt(t1, t2); ms[2] = dt(t2, t3); }
        int k = (int)toks[0].size();
        for (int i = 0; i < k && i < cap; ++i) ids[i] = toks[0][i];
        return k;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// ---- phrase-boosted decode (src/phrase_boost.cpp).  Phrases arrive as token-id sequences:
// ids = concatenation, phrase p = ids[off[p] .. off[p+1]); ContextTrie::insert is the reference's own.
static ContextTrie make_trie(const int *ids, const int *off, int n_phrases) {
    ContextTrie trie;
    for (int p = 0; p < n_phrases; ++p) trie.insert(std::vector<int>(ids + off[p], ids + off[p + 1]));
    return trie;
}

// lp (T, V) -> boosted CTC greedy with timestamps; returns the token count (<= T).
int pkref_ctc_greedy_boosted(const float *lp, int T, int V, int blank, const int *ph_ids, const int *ph_off, int n_phrases,
                             float boost, int *ids, int *start, int *end, float *conf) {
    try {
        auto t = Tensor::from_data(lp, Shape{1, (size_t)T, (size_t)V}, true);
        auto trie = make_trie(ph_ids, ph_off, n_phrases);
        auto r = ctc_greedy_decode_with_timestamps_boosted(t, trie, boost, blank);
        auto r2 = ctc_greedy_decode_boosted(t, trie, boost, blank);
        if (r2[0].size() != r[0].size()) { g_err = "boosted CTC: id-only and timestamped variants disagree"; return -1; }
        for (size_t i = 0; i < r[0].size(); ++i) {
            if (r2[0][i] != r[0][i].token_id) { g_err = "boosted CTC: id-only and timestamped variants disagree"; return -1; }
            ids[i] = r[0][i].token_id;
            start[i] = r[0][i].start_frame;
            end[i] = r[0][i].end_frame;
            conf[i] = r[0][i].confidence;
        }
        return (int)r[0].size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// enc (T, d) -> boosted TDT greedy with timestamps; returns the token count (only cap are written) or -1.
int pkref_tdt_greedy_boosted(void *h, const float *enc, int T, int d, const int *ph_ids, const int *ph_off, int n_phrases,
                             float boost, int cap, int *ids, int *start, int *end, float *conf) {
    try {
        auto *m = static_cast<RefModel *>(h);
        auto e = Tensor::from_data(enc, Shape{1, (size_t)T, (size_t)d}, true);
        auto trie = make_trie(ph_ids, ph_off, n_phrases);
        auto r = tdt_greedy_decode_with_timestamps_boosted(m->prediction(), m->joint(), e, m->durations(), trie, boost,
                                                           m->blank());
        int n = (int)r[0].size();
        for (int i = 0; i < n && i < cap; ++i) {
            ids[i] = r[0][i].token_id;
            start[i] = r[0][i].start_frame;
            end[i] = r[0][i].end_frame;
            conf[i] = r[0][i].confidence;
        }
        return n;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// parakeet::resample (src/audio_io.cpp:238-251 -> sinc_resample :123-195); returns the output length.
int pkref_resample(const float *in, int n, int src_rate, int dst_rate, float *out, int cap) {
    try {
        auto t = Tensor::from_data(in, Shape{(size_t)n}, true);
        auto r = resample(t, src_rate, dst_rate).ascontiguousarray();
        const int m = (int)r.shape()[0];
        std::memcpy(out, r.typed_data<float>(), sizeof(float) * (size_t)std::min(m, cap));
        return m;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// Tokenizer::encode (src/vocab.cpp:76-117) with the model's vocabulary; returns the id count.
int pkref_tok_encode(void *h, const char *text, int cap, int *ids) {
    try {
        auto *m = static_cast<RefModel *>(h);
        auto r = m->tok.encode(text);
        for (size_t i = 0; i < r.size() && (int)i < cap; ++i) ids[i] = r[i];
        return (int)r.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

} // extern "C"
