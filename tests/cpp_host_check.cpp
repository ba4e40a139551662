// Host-only part of the C++ shim (include/parakeet/transcribe.hpp): Tokenizer::encode / decode, ContextTrie and the
// boosted CTC greedy decode, exercised WITHOUT a device (tests/test_abi.py builds and runs this on the CPU).
// Known answers: the reference's BoostedCTCDecode.* tests (tests/test_all.cpp:1369-1452).
#include <cstdio>
#include <string>
#include <vector>

#include <parakeet/transcribe.hpp>

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    parakeet::Tokenizer tok;
    tok.load(argv[1]);
    const std::string text = argv[2];
    auto ids = tok.encode(text);
    std::printf("encode");
    for (int i : ids) std::printf(" %d", i);
    std::printf("\ndecode %s\n", tok.decode(ids).c_str());

    const int V = 1025;
    std::vector<float> lp(3 * V, -10.0f);
    lp[0 * V + 42] = -0.1f; lp[0 * V + 43] = -0.2f; lp[0 * V + 1024] = -5.0f;
    lp[1 * V + 1024] = 0.0f; lp[2 * V + 1024] = 0.0f;
    parakeet::ContextTrie empty, trie;
    trie.insert({43});
    auto a = parakeet::ctc_greedy_decode_boosted(lp.data(), 3, V, empty, 5.0f, 1024);
    auto b = parakeet::ctc_greedy_decode_with_timestamps_boosted(lp.data(), 3, V, trie, 5.0f, 1024);
    std::printf("plain %d n=%zu\nboosted %d start=%d end=%d n=%zu empty=%d/%d\n", a.empty() ? -1 : a[0], a.size(),
                b.empty() ? -1 : b[0].token_id, b.empty() ? -1 : b[0].start_frame, b.empty() ? -1 : b[0].end_frame, b.size(),
                (int)empty.empty(), (int)trie.empty());
    parakeet::ContextTrie built;
    built.build({text}, tok);
    std::printf("built %zu\n", built.ids().size());
    if (argc > 3) {   // read_audio on a WAV that is not 16 kHz: resampled like the reference's read_audio
        auto pcm = parakeet::read_audio(argv[3]);
        double acc = 0.0;
        for (size_t i = 0; i < pcm.size(); ++i) acc += (double)pcm[i] * (double)((i % 7) + 1);
        std::printf("wav %zu %.9e %.9e\n", pcm.size(), pcm.empty() ? 0.0 : (double)pcm[pcm.size() / 2], acc);
    }
    return 0;
}
