// examples/sharded_transcribe.cpp -- a C++ host sharding a job of clips over the GPUs of one box through the C-ABI only
// (SURVEY.md section 8e): one thread per GPU, each with its own pk_engine; contiguous blocks of clips; every micro-batch's
// token rows are appended to the engine's device job buffer; ONE ncclAllGather (pk_allgather_tokens) assembles all rows on
// every GPU.  No Python, no torch: NCCL is resolved by the library at run time (dlopen libnccl.so.2).
//
//   g++ -std=c++17 -O2 -Iinclude examples/sharded_transcribe.cpp -Lparakeet.cpp_b200 -lparakeet_b200 \
//       -Wl,-rpath,$PWD/parakeet.cpp_b200 -lpthread -o sharded_transcribe
//   ./sharded_transcribe model.safetensors <n_gpus> <clips_per_gpu> [tiny]
//
// The audio is synthetic (seeded tones + noise, 4 s clips): the point is the control flow a reference maintainer would
// write around Transcriber::transcribe for a multi-GPU server.  Prints a checksum of every rank's gathered rows (they
// must agree) and the gathered row count.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "parakeet_b200.h"

static std::vector<float> make_clip(int n, unsigned seed) {
    std::vector<float> x(n);
    unsigned s = seed * 2654435761u + 1u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f; };
    const float f1 = 200.f + 1800.f * rnd(), f2 = 300.f + 3000.f * rnd();
    for (int i = 0; i < n; ++i) {
        const float t = i / 16000.0f;
        x[i] = 0.2f * std::sin(6.2831853f * f1 * t) * (0.5f + 0.5f * std::sin(6.2831853f * 3.f * t)) + 0.15f * std::sin(6.2831853f * f2 * t) + 0.02f * (rnd() - 0.5f);
    }
    return x;
}

int main(int argc, char **argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s weights.safetensors n_gpus clips_per_gpu [tiny]\n", argv[0]); return 2; }
    const std::string weights = argv[1];
    const int world = std::atoi(argv[2]), per_gpu = std::atoi(argv[3]);
    const bool tiny = argc > 4 && std::string(argv[4]) == "tiny";
    const int clip = 64000, micro = 8;

    char id[PK_NCCL_UNIQUE_ID_BYTES];
    if (pk_nccl_unique_id(id) != PK_OK) { std::fprintf(stderr, "NCCL: %s\n", pk_last_error(nullptr)); return 1; }

    std::vector<unsigned long long> checksum(world, 0);
    std::vector<int> status(world, 0);
    std::vector<std::thread> threads;
    for (int rank = 0; rank < world; ++rank)
        threads.emplace_back([&, rank]() {
            pk_config c;
            pk_config_110m(&c);
            if (tiny) {
                c.sub_channels = 64; c.d_model = 128; c.n_layers = 2; c.n_heads = 2; c.ff = 256; c.vocab = 33; c.pred_hidden = 64; c.joint_hidden = 64;
            }
            c.max_batch = micro; c.max_samples = clip;
            pk_engine *e = nullptr;
            if (pk_engine_create(&c, weights.c_str(), rank, &e) != PK_OK) { std::fprintf(stderr, "rank %d: %s\n", rank, pk_last_error(nullptr)); status[rank] = 1; return; }
            auto fail = [&](const char *what) { std::fprintf(stderr, "rank %d: %s: %s\n", rank, what, pk_last_error(e)); status[rank] = 1; };
            if (pk_comm_init_rank(e, id, rank, world) != PK_OK) { fail("pk_comm_init_rank"); return; }
            if (pk_job_begin(e, per_gpu, world) != PK_OK) { fail("pk_job_begin"); return; }
            for (int first = 0; first < per_gpu && !status[rank]; first += micro) {
                const int n = std::min(micro, per_gpu - first);
                std::vector<float> pcm;
                std::vector<int64_t> off(n + 1, 0);
                for (int i = 0; i < n; ++i) {                       // this rank's block: global clip index rank * per_gpu + first + i
                    auto x = make_clip(clip - 1000 * ((first + i) % 5), 1000u + (unsigned)(rank * per_gpu + first + i));
                    off[i + 1] = off[i] + (int64_t)x.size();
                    pcm.insert(pcm.end(), x.begin(), x.end());
                }
                if (pk_stage_pcm(e, pcm.data(), off.data(), n) != PK_OK || pk_run_staged(e, PK_DECODER_TDT) != PK_OK || pk_job_append(e) != PK_OK) fail("micro-batch");
            }
            if (!status[rank] && pk_allgather_tokens(e, nullptr) != PK_OK) fail("pk_allgather_tokens");
            int32_t w = 0;
            pk_job_fetch(e, 1, nullptr, 0, &w);
            std::vector<int32_t> rows((size_t)world * per_gpu * w);
            if (!status[rank] && pk_job_fetch(e, 1, rows.data(), (int64_t)world * per_gpu, &w) != PK_OK) fail("pk_job_fetch");
            unsigned long long h = 1469598103934665603ull;
            for (int r = 0; r < world * per_gpu && !status[rank]; ++r)
                for (int i = 0; i <= rows[(size_t)r * w]; ++i) h = (h ^ (unsigned long long)(unsigned)rows[(size_t)r * w + i]) * 1099511628211ull;
            checksum[rank] = h;
            if (rank == 0 && !status[rank]) {
                long tok = 0;
                for (int r = 0; r < world * per_gpu; ++r) tok += rows[(size_t)r * w];
                std::printf("gathered %d rows, %ld tokens\n", world * per_gpu, tok);
            }
            pk_engine_destroy(e);
        });
    for (auto &t : threads) t.join();
    int bad = 0;
    for (int r = 0; r < world; ++r) bad |= status[r] | (checksum[r] != checksum[0]);
    std::printf("checksum %016llx %s\n", checksum[0], bad ? "MISMATCH" : "identical on every rank");
    return bad ? 1 : 0;
}
