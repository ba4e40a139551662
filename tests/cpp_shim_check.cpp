// tests/cpp_shim_check.cpp -- exercises the header-only C++ drop-in (include/parakeet/transcribe.hpp)
// exactly the way the reference's README uses parakeet::Transcriber.  Built and run by
// tests/test_gpu_parity.py::test_cpp_shim; prints token ids / text for comparison.
#include <cstdio>
#include <iostream>

#include "parakeet/transcribe.hpp"

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    try {
        parakeet::TDTCTCConfig cfg = parakeet::make_110m_config();
        if (argc > 4 && std::string(argv[4]) == "tiny") {
            cfg.encoder.subsampling_channels = 64; cfg.encoder.hidden_size = 128; cfg.encoder.num_layers = 2;
            cfg.encoder.num_heads = 2; cfg.encoder.ffn_intermediate = 256;
            cfg.prediction.vocab_size = 33; cfg.prediction.pred_hidden = 64; cfg.prediction.num_lstm_layers = 1;
            cfg.joint.encoder_hidden = 128; cfg.joint.pred_hidden = 64; cfg.joint.joint_hidden = 64; cfg.joint.vocab_size = 33;
            cfg.ctc_vocab_size = 33;
        }
        parakeet::Transcriber t(argv[1], argv[2], cfg, 0, 4, 64000);
        t.to_gpu();
        for (auto dec : {parakeet::Decoder::TDT, parakeet::Decoder::CTC}) {
            auto r = t.transcribe(std::string(argv[3]), dec, true);
            std::cout << (dec == parakeet::Decoder::TDT ? "TDT" : "CTC");
            for (auto &tk : r.timestamped_tokens) std::cout << " " << tk.token_id << ":" << tk.start_frame << ":" << tk.end_frame;
            std::cout << "\nTEXT " << r.text << "\nWORDS";
            for (auto &w : r.word_timestamps) std::cout << " " << w.word;
            std::cout << "\n";
        }
        if (argc > 5) {      // TranscribeOptions::boost_phrases (transcribe.hpp:38-43): boosted decode on the device, then plain again
            for (auto dec : {parakeet::Decoder::CTC, parakeet::Decoder::TDT}) {
                parakeet::TranscribeOptions o;
                o.decoder = dec;
                o.timestamps = true;
                o.boost_phrases = {argv[5]};
                o.boost_score = 6.0f;
                auto r = t.transcribe(std::string(argv[3]), o);
                std::cout << "BOOST";
                for (auto &tk : r.timestamped_tokens) std::cout << " " << tk.token_id << ":" << tk.start_frame << ":" << tk.end_frame;
                std::cout << "\n";
            }
            auto r = t.transcribe(std::string(argv[3]), parakeet::Decoder::CTC, true);
            std::cout << "PLAIN";
            for (auto &tk : r.timestamped_tokens) std::cout << " " << tk.token_id << ":" << tk.start_frame << ":" << tk.end_frame;
            std::cout << "\n";
        }
        if (argc > 6) {      // a WAV that is not 16 kHz: converted on the device (pk_stage_pcm_rate)
            auto r = t.transcribe(std::string(argv[6]), parakeet::Decoder::TDT, true);
            std::cout << "RATE";
            for (auto &tk : r.timestamped_tokens) std::cout << " " << tk.token_id << ":" << tk.start_frame << ":" << tk.end_frame;
            std::cout << "\n";
        }
        try {
            t.transcribe(std::string("/nonexistent.wav"));
            return 3;
        } catch (const std::runtime_error &e) {
            std::cout << "ERR " << e.what() << "\n";
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
