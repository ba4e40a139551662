// text.cpp -- host-side text helpers behind the C-ABI (no device work).
//
// Behavioural parity targets (reference):
//   Tokenizer::load   src/vocab.cpp:10-27   "piece<TAB>score" lines; a line without a tab is
//                                            taken whole; empty lines are skipped.
//   Tokenizer::decode src/vocab.cpp:29-64   concatenate pieces, out-of-range id -> "[id]",
//                                            U+2581 -> ' ', strip exactly one leading space.
//   group_timestamps  src/timestamp.cpp:24-75 (Words mode) a piece starting with U+2581
//                                            opens a new word; word conf = min token conf;
//                                            seconds = frame * 0.08f (timestamp.hpp:31-35).
#include <algorithm>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/parakeet_b200.h"

struct pk_vocab {
    std::vector<std::string> pieces;
};

namespace {
const char kMark[] = "\xe2\x96\x81";  // U+2581
bool has_mark(const std::string &s) { return s.size() >= 3 && s.compare(0, 3, kMark) == 0; }

int32_t emit(const std::string &s, char *buf, int32_t cap) {
    if (buf && cap > 0) {
        size_t k = std::min(s.size(), (size_t)cap - 1);
        memcpy(buf, s.data(), k);
        buf[k] = 0;
    }
    return (int32_t)s.size();
}
}  // namespace

extern "C" {

pk_status pk_vocab_load(const char *vocab_path, pk_vocab **out) {
    if (!vocab_path || !out) return PK_ERR_INVALID;
    std::ifstream f(vocab_path);
    if (!f) return PK_ERR_IO;
    auto *v = new pk_vocab();
    std::string line;
    while (std::getline(f, line)) {
        size_t tab = line.find('\t');
        if (tab != std::string::npos)
            v->pieces.emplace_back(line, 0, tab);
        else if (!line.empty())
            v->pieces.push_back(line);
    }
    *out = v;
    return PK_OK;
}

void pk_vocab_free(pk_vocab *v) { delete v; }

int32_t pk_vocab_size(const pk_vocab *v) { return v ? (int32_t)v->pieces.size() : 0; }

int32_t pk_detokenize(const pk_vocab *v, const int32_t *ids, int32_t n, char *buf, int32_t cap) {
    if (!v) return -1;
    std::string joined;
    for (int32_t i = 0; i < n; ++i) {
        const int32_t id = ids[i];
        if (id < 0 || id >= (int32_t)v->pieces.size())
            joined += "[" + std::to_string(id) + "]";
        else
            joined += v->pieces[id];
    }
    std::string text;
    text.reserve(joined.size());
    for (size_t pos = 0; pos < joined.size();) {
        if (joined.compare(pos, 3, kMark) == 0 && pos + 3 <= joined.size()) {
            text += ' ';
            pos += 3;
        } else {
            text += joined[pos++];
        }
    }
    if (!text.empty() && text[0] == ' ') text.erase(0, 1);
    return emit(text, buf, cap);
}

int32_t pk_group_words(const pk_vocab *v, const int32_t *ids, const int32_t *start, const int32_t *end,
                       const float *conf, int32_t n, char *buf, int32_t cap, float *w_start, float *w_end,
                       float *w_conf) {
    if (!v) return -1;
    if (n <= 0) {
        emit("", buf, cap);
        return 0;
    }
    const float kFrame = 0.08f;
    std::string all, word;
    int32_t n_words = 0;
    int ws = start[0], we = end[0];
    float wc = 1.0f;
    auto flush = [&]() {
        all += word;
        all += '\n';
        w_start[n_words] = (float)ws * kFrame;
        w_end[n_words] = (float)we * kFrame;
        w_conf[n_words] = wc;
        ++n_words;
        word.clear();
    };
    for (int32_t i = 0; i < n; ++i) {
        if (ids[i] < 0 || ids[i] >= (int32_t)v->pieces.size()) continue;
        const std::string &piece = v->pieces[ids[i]];
        const bool opens = has_mark(piece);
        if (opens && !word.empty()) {
            flush();
            ws = start[i];
            wc = 1.0f;
        }
        word += opens ? piece.substr(3) : piece;
        we = end[i];
        wc = std::min(wc, conf ? conf[i] : 1.0f);
    }
    if (!word.empty()) flush();
    emit(all, buf, cap);
    return n_words;
}

}  // extern "C"
