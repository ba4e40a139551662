// ctc.cu -- K9: CTC head reductions and greedy collapse.
//
// The head GEMM (CTCDecoder::forward's k=1 Conv1d, reference src/ctc.cpp:12-18) runs in the
// GEMM kernel; this file replaces log_softmax (:24) + the argmax scan + collapse of
// ctc_greedy_decode / ctc_greedy_decode_with_timestamps (src/ctc.cpp:40-127):
//   ctc_frame_argmax_kernel : per frame, first maximum (strict '>' scan order, :59-66) and
//                             exp(max log-prob) = 1 / sum exp(l - max)  (:110)
//   ctc_collapse_kernel     : per utterance, drop blanks/repeats (prev updates on every
//                             frame, :69-72), token spans: start = first frame of the run,
//                             end closed at t-1 when the argmax changes after a non-blank
//                             run, last token's end forced to T-1 (:104-121).
#include "kernels.h"

namespace pk {
namespace {

__global__ void __launch_bounds__(256)
ctc_frame_argmax_kernel(const float *__restrict__ logits, int M, int V, int ld, int32_t *__restrict__ best,
                        float *__restrict__ conf, float *__restrict__ logprobs) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *l = logits + (size_t)row * ld;
    float mx = -INFINITY;
    int idx = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
        const float x = l[v];
        if (x > mx) {
            mx = x;
            idx = v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (om > mx || (om == mx && oi < idx)) {
            mx = om;
            idx = oi;
        }
    }
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(l[v] - mx);
    s = warp_sum(s);
    if (lane == 0) {
        best[row] = idx;
        conf[row] = 1.0f / s;
    }
    if (logprobs) {
        const float lse = logf(s);
        for (int v = lane; v < V; v += 32) logprobs[(size_t)row * V + v] = (l[v] - mx) - lse;
    }
}

// one warp per utterance; lane 0 scans (T' <= a few hundred frames)
__global__ void ctc_collapse_kernel(const int32_t *__restrict__ best, const float *__restrict__ conf,
                                    const int32_t *__restrict__ row_off, int n_utt, int blank, int cap,
                                    int32_t *__restrict__ tok /* [n_utt][1+cap] */,
                                    int32_t *__restrict__ t_start, int32_t *__restrict__ t_end,
                                    float *__restrict__ t_conf) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_utt) return;
    const int r0 = row_off[b], T = row_off[b + 1] - r0;
    int32_t *ids = tok + (size_t)b * (1 + cap) + 1;
    int32_t *st = t_start + (size_t)b * cap, *en = t_end + (size_t)b * cap;
    float *cf = t_conf + (size_t)b * cap;
    int prev = -1, n = 0;
    for (int t = 0; t < T; ++t) {
        const int cur = best[r0 + t];
        if (cur != prev) {
            if (prev != -1 && prev != blank && n > 0 && n <= cap) en[n - 1] = t - 1;
            if (cur != blank) {
                if (n < cap) {
                    ids[n] = cur;
                    st[n] = t;
                    en[n] = t;
                    cf[n] = conf[r0 + t];
                }
                ++n;
            }
        }
        prev = cur;
    }
    if (n > 0 && n <= cap) en[n - 1] = T - 1;
    tok[(size_t)b * (1 + cap)] = n < cap ? n : cap;
}

}  // namespace

void launch_ctc_frame_argmax(const float *logits, int M, int V, int ld, int32_t *best, float *conf,
                             float *logprobs, cudaStream_t st) {
    if (M <= 0) return;
    ctc_frame_argmax_kernel<<<(M + 7) / 8, 256, 0, st>>>(logits, M, V, ld, best, conf, logprobs);
}

void launch_ctc_collapse(const int32_t *best, const float *conf, const int32_t *row_off, int n_utt, int blank,
                         int cap, int32_t *tok, int32_t *t_start, int32_t *t_end, float *t_conf,
                         cudaStream_t st) {
    ctc_collapse_kernel<<<(n_utt + 63) / 64, 64, 0, st>>>(best, conf, row_off, n_utt, blank, cap, tok, t_start,
                                                        t_end, t_conf);
}

}  // namespace pk
