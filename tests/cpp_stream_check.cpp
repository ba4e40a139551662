// tests/cpp_stream_check.cpp -- the C++ drop-in's StreamingTranscriber (reference eou.hpp:101-141 usage): feed a raw fp32
// PCM file chunk by chunk, print the tokens every chunk produced.  Built and run by tests/test_gpu_parity.py.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "parakeet/transcribe.hpp"

int main(int argc, char **argv) {
    if (argc < 5) return 2;      // weights vocab pcm.f32 chunk,chunk,...
    try {
        parakeet::EOUConfig cfg = parakeet::make_eou_120m_config();
        // the tiny streaming test shape (oracle.make_tiny_stream_config)
        cfg.encoder.subsampling_channels = 64; cfg.encoder.hidden_size = 128; cfg.encoder.num_layers = 2; cfg.encoder.num_heads = 2;
        cfg.encoder.ffn_intermediate = 256; cfg.encoder.att_context_left = 12; cfg.encoder.att_context_right = 1;
        cfg.prediction.vocab_size = 33; cfg.prediction.pred_hidden = 64; cfg.prediction.num_lstm_layers = 1;
        cfg.joint.encoder_hidden = 128; cfg.joint.pred_hidden = 64; cfg.joint.joint_hidden = 64; cfg.joint.vocab_size = 33;
        parakeet::StreamingTranscriber t(argv[1], argv[2], cfg);
        t.to_gpu();
        int n_cb = 0;
        t.set_partial_callback([&](const std::string &) { ++n_cb; });
        std::ifstream f(argv[3], std::ios::binary);
        std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        const float *pcm = reinterpret_cast<const float *>(raw.data());
        size_t pos = 0, emitted = 0;
        for (char *p = argv[4]; *p;) {
            const long n = std::strtol(p, &p, 10);
            if (*p == ',') ++p;
            t.transcribe_chunk(pcm + pos, (size_t)n);
            pos += (size_t)n;
            const auto &all = t.get_timestamped_tokens();
            std::cout << "CHUNK";
            for (; emitted < all.size(); ++emitted) std::cout << " " << all[emitted].token_id << ":" << all[emitted].start_frame << ":" << all[emitted].end_frame;
            std::cout << "\n";
        }
        std::cout << "TEXT " << t.get_text() << "\nCALLBACKS " << n_cb << "\n";
        t.reset();
        std::cout << "AFTER_RESET " << t.get_timestamped_tokens().size() << "\n";
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
