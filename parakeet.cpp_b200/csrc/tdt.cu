// tdt.cu -- K10: batched TDT greedy decode as ONE persistent cooperative kernel.
//
// Replaces tdt_greedy_decode / tdt_greedy_decode_with_timestamps (reference
// src/tdt.cpp:36-110, :122-201) and what they call per step: RNNTPrediction::step
// (src/rnnt.cpp:22-28) -> Embedding -> LSTM::step (src/lstm.cpp:40-49) -> LSTMCell::forward
// (:11-29, gate order i,f,g,o, one merged bias) and TDTJoint::forward (src/tdt.cpp:15-24,
// pred_proj_ without bias, two log_softmax heads), argmax label (first maximum) / duration.
//
// The reference decodes utterances one after another with a host round trip per symbol.
// Here all utterances of the batch advance in lock step inside one kernel; the serial
// chain per utterance is unchanged:
//     saved = state; pred = LSTM(embed(token)); (label, dur) = joint(enc[t], pred)
//     blank  -> state = saved, t += max(skip, 1)
//     symbol -> emit (start = t, end = min(t + max(skip,1) - 1, T-1), conf = exp(lp)),
//               token = symbol, t += skip (skip = 0 stays on the frame)
// max_symbols_per_step has no observable effect in the reference (after 10 zero-duration
// symbols the inner loop is simply re-entered on the same frame with the same state), so it
// is not modelled; a token capacity bounds the loop instead (reference would livelock).
//
// Design (weights-stationary): the grid is one CTA per SM; every CTA keeps its slice of
// W_hh / W_ih / pred_proj / label+duration rows in shared memory for the whole decode and
// only the per-utterance vectors (h, z) travel through L2 between the phases of a step:
//   P1 LSTM gates + cell (per layer) | P2 joint hidden | P3 logits -> per-CTA (max, sum-exp)
//   partials + atomicMax of packed (value, index) keys | P4 state update, replicated in every
//   CTA from the keys (no barrier before the next P1; confidences are finalised one phase later
//   from the partials, in a fixed order, by the CTA that owns the utterance, while it waits at a
//   grid barrier).
// Three monotonic-counter grid barriers per step (cooperative launch guarantees co-residency).
// enc_proj(enc)+bias for all frames and the layer-0 input table W_ih.E[token]+b for all
// tokens are precomputed by GEMMs (engine.cu).
//
// Every product  out[r][b] = sum_k W[r][k] x_b[k]  (weight rows r of this CTA, all utterances b) runs
// on mma.sync.m16n8k16 with the bf16 hi/lo operand split of the encoder GEMMs
// (x_hi.W_hi + x_hi.W_lo + x_lo.W_hi, fp32 accumulate: ~16 mantissa bits):
//   * weights are split once at load (engine.cu): row r = [hi: K+4 bf16][lo: K+4 bf16];
//   * the PRODUCER of a vector (LSTM cell, joint hidden) stores it already split, as bf16 hi / lo
//     planes [utterance][k], so consumers only copy: 64 utterances x 64 k per cp.async chunk,
//     A fragments by ldmatrix, no conversion in the inner loop;
//   * operands of the phase epilogues (G0[token], EP[t], biases) are fetched BEFORE the product
//     so their L2 latency overlaps the streaming; the LSTM cell state never leaves shared memory.
#include "kernels.h"

namespace pk {
namespace {

constexpr int RMAX = 20;        // weight rows per pass (5 LSTM units x 4 gates)
constexpr int RPAD = 24;        // RMAX rounded up to whole 8-row MMA n-blocks
constexpr int NWARP = 8;
constexpr int BCH = 64;         // utterances per pass (4 MMA m-blocks)
constexpr int KC = 64;          // k-values staged per chunk (2 k-steps per warp half)
constexpr int NST = 4;          // cp.async ring depth
constexpr int XLD = KC + 8;     // staged row stride (bf16): 144 B, conflict-free ldmatrix
constexpr int RLD = BCH + 4;    // row stride of the partial-sum buffer (floats)
constexpr int STAGE_ELEMS = 2 * BCH * XLD;   // one ring stage: hi plane + lo plane (bf16 elements)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// 32-bit weight fragment word (two consecutive k of one row); weights never change during the kernel
template <bool WS>
__device__ __forceinline__ uint32_t ldw32(const bf16 *p) {
    uint32_t v;
    if (WS) asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)));
    else asm("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void store_split(bf16 *hi, bf16 *lo, size_t idx, float v) {
    const bf16 h = __float2bfloat16_rn(v);
    hi[idx] = h;
    lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// out(r, b) = sum_k W[r][k] * x_b[k] for r < R (R <= NB*8), b in [bc, bc+64).
//   W     : pre-split rows, row r at W + r * 2*(K+4): [hi K+4][lo K+4] (shared memory when WS, else global/L2)
//   xsrc  : b -> pointer to the K bf16 "hi" values of utterance b (global; lives in L2); lo = hi + lo_off
//   pre   : (r, b) -> float, called BEFORE the product for the outputs this thread will finalise
//   fin   : (r, b, sum, pre value)
// M = utterances (4 blocks of 16), N = weight rows (NB blocks of 8).  Warp w owns utterance block
// (w & 3) and k-half (w >> 2) of every 64-wide chunk; the two halves are added in `red` in a fixed order.
template <int NB, bool WS, typename XSrc, typename Pre, typename Fin>
__device__ __forceinline__ void rows_times_batch_mma(const bf16 *W, int R, int K, int Bpad, int bc, XSrc xsrc,
                                                     size_t lo_off, bf16 *xs, float *red, Pre pre, Fin fin) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, cq = lane & 3;
    const int mb = warp & 3, kh = warp >> 2;
    const int KP = K + 4, RS = 2 * KP;
    const int nchunks = K / KC;
    // this thread copies 16-byte pieces q and q+4 (hi and lo) of staged row `row`
    const int row = threadIdx.x >> 2, q = threadIdx.x & 3;
    const bool vrow = (bc + row) < Bpad;
    const bf16 *xsrc_hi = xsrc(vrow ? bc + row : 0) + q * 8;
    const uint32_t xs_s = smem_addr(xs);
    const uint32_t dst0 = xs_s + (uint32_t)(row * XLD + q * 8) * 2u;
    auto issue = [&](int c) {
        if (c < nchunks && vrow) {
            const uint32_t dst = dst0 + (uint32_t)((c % NST) * STAGE_ELEMS) * 2u;
            const bf16 *s = xsrc_hi + c * KC;
            cp_async16(dst, s);
            cp_async16(dst + 64u, s + 32);
            cp_async16(dst + (uint32_t)(BCH * XLD) * 2u, s + lo_off);
            cp_async16(dst + (uint32_t)(BCH * XLD) * 2u + 64u, s + lo_off + 32);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) issue(c);

    // operands of the epilogue: in flight while the product streams
    constexpr int NIT = NB * 8 * BCH / (NWARP * 32);
    float pv[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        const int idx = threadIdx.x + i * NWARP * 32, r = idx / BCH, b2 = idx % BCH;
        pv[i] = (r < R && bc + b2 < Bpad) ? pre(r, bc + b2) : 0.f;
    }

    float acc[NB][4];
    const bf16 *wrow[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
        wrow[nb] = W + (size_t)min(nb * 8 + g, R - 1) * RS + kh * 32 + 2 * cq;   // rows >= R: clamped, discarded
    }
    // ldmatrix lane address inside a stage: A tile rows mb*16.., columns kh*32 + ks*16
    const uint32_t a_off = (uint32_t)((mb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * XLD + kh * 32 + ((lane >> 4) & 1) * 8) * 2u;
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<NST - 2>();
        __syncthreads();                       // chunk c landed for everyone; chunk c-1 fully consumed
        issue(c + NST - 1);
        const uint32_t st = xs_s + (uint32_t)((c % NST) * STAGE_ELEMS) * 2u + a_off;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t ah[4], al[4];
            ldsm_x4(ah, st + ks * 32u);
            ldsm_x4(al, st + (uint32_t)(BCH * XLD) * 2u + ks * 32u);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const bf16 *wp = wrow[nb] + c * KC + ks * 16;
                const uint32_t b0h = ldw32<WS>(wp), b1h = ldw32<WS>(wp + 8), b0l = ldw32<WS>(wp + KP), b1l = ldw32<WS>(wp + KP + 8);
                mma_bf16(acc[nb], ah, b0h, b1h);
                mma_bf16(acc[nb], ah, b0l, b1l);
                mma_bf16(acc[nb], al, b0h, b1h);
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        const int n = nb * 8 + 2 * cq, u = mb * 16 + g;
        red[(kh * RPAD + n) * RLD + u] = acc[nb][0];
        red[(kh * RPAD + n + 1) * RLD + u] = acc[nb][1];
        red[(kh * RPAD + n) * RLD + u + 8] = acc[nb][2];
        red[(kh * RPAD + n + 1) * RLD + u + 8] = acc[nb][3];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        const int idx = threadIdx.x + i * NWARP * 32, r = idx / BCH, b2 = idx % BCH;
        if (r < R && bc + b2 < Bpad) fin(r, bc + b2, red[r * RLD + b2] + red[(RPAD + r) * RLD + b2], pv[i]);
    }
    __syncthreads();
}

template <bool WS, typename XSrc, typename Pre, typename Fin>
__device__ __forceinline__ void rows_times_batch(const bf16 *W, int R, int K, int Bpad, int bc, XSrc xsrc, size_t lo_off,
                                                 bf16 *xs, float *red, Pre pre, Fin fin) {
    if (R <= 8) rows_times_batch_mma<1, WS>(W, R, K, Bpad, bc, xsrc, lo_off, xs, red, pre, fin);
    else if (R <= 16) rows_times_batch_mma<2, WS>(W, R, K, Bpad, bc, xsrc, lo_off, xs, red, pre, fin);
    else rows_times_batch_mma<3, WS>(W, R, K, Bpad, bc, xsrc, lo_off, xs, red, pre, fin);
}

// Monotonic-counter grid barrier (all CTAs are co-resident: cooperative launch), split into
// arrive / wait so that work which does not depend on the other CTAs can sit in between.
// Cheaper than cooperative_groups' grid.sync() and traps instead of hanging if a CTA never arrives.
__device__ __forceinline__ void grid_arrive(unsigned int *counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
    }
}
__device__ __forceinline__ void grid_wait(unsigned int *counter, unsigned int target) {
    if (threadIdx.x == 0) {
        unsigned int v, spin = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (++spin > (1u << 28)) __trap();
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// (value, index) packed so that atomicMax picks the larger value and, on ties, the SMALLER index
// (the reference's strict '>' scans keep the first maximum).
__device__ __forceinline__ unsigned long long pack_key(float v, int idx) {
    unsigned int u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)idx);
}
__device__ __forceinline__ void unpack_key(unsigned long long k, float &v, int &idx) {
    unsigned int u = (unsigned int)(k >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    v = __uint_as_float(u);
    idx = (int)(0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFu));
}

__global__ void __launch_bounds__(NWARP * 32, 1) tdt_decode_kernel(TdtParams p) {
    extern __shared__ __align__(16) float sm[];
    const int G = gridDim.x, g = blockIdx.x, tid = threadIdx.x;
    const int P = p.P, J = p.J, V = p.V, D = p.D, L = p.L, Bpad = p.Bpad;
    const int NO = V + D;
    const int UPC = (P + G - 1) / G;            // LSTM units per CTA
    const int u0 = min(g * UPC, P), u1 = min(u0 + UPC, P);
    const int JPC = (J + G - 1) / G;
    const int j0 = min(g * JPC, J), j1 = min(j0 + JPC, J);
    const int OPC = (NO + G - 1) / G;
    const int o0 = min(g * OPC, NO), o1 = min(o0 + OPC, NO);
    const int nU = u1 - u0;

    // ---- shared memory carve-up: [red][gates][ring][cell state][decode state][weights...]
    float *red = sm;                               // [2][RPAD][RLD] k-half partial sums
    float *gsm = red + 2 * RPAD * RLD;             // [RMAX][BCH] gate pre-activations / logits
    bf16 *xs = reinterpret_cast<bf16 *>(gsm + RMAX * BCH);                 // [NST][2][BCH][XLD] cp.async ring
    float *csm = reinterpret_cast<float *>(xs + (size_t)NST * STAGE_ELEMS); // [L][2][UPC][Bpad] LSTM cell state
    int *s_cur = reinterpret_cast<int *>(csm + (size_t)L * 2 * UPC * Bpad); // replicated decode state, [Bpad] each
    int *s_token = s_cur + Bpad, *s_tpos = s_token + Bpad, *s_active = s_tpos + Bpad, *s_ntok = s_active + Bpad;
    int *s_pend = s_ntok + Bpad;                   // slot of a token whose confidence is still pending (-1: none)
    bf16 *wbf = reinterpret_cast<bf16 *>(s_pend + Bpad);
    // LSTM weights arrive "unit-major" (row = unit*4 + gate, engine.cu), so this CTA's rows
    // [u0*4, u1*4) are one contiguous block: W_hh always lives in shared memory, W_ih of the
    // upper layers too when it fits (else its fragments are read from L2).
    const int RSP = 2 * (P + 4), RSJ = 2 * (J + 4);          // row strides (bf16) for K = P / K = J
    auto stage_rows = [&](bf16 *dst, const bf16 *src, int rows, int RS) {   // 16-byte copies (RS * 2 B is a multiple of 16)
        const int n16 = rows * RS / 8;
        for (int i = tid; i < n16; i += blockDim.x)
            reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
    };
    const bf16 *w_hh[PK_MAX_LSTM], *w_ih[PK_MAX_LSTM];
    {
        bf16 *cur = wbf;
        for (int l = 0; l < L; ++l) {
            stage_rows(cur, p.Whh[l] + (size_t)u0 * 4 * RSP, nU * 4, RSP);
            w_hh[l] = cur;
            cur += (size_t)UPC * 4 * RSP;
            w_ih[l] = nullptr;
            if (l > 0) {
                if (p.wih_in_smem) {
                    stage_rows(cur, p.Wih[l] + (size_t)u0 * 4 * RSP, nU * 4, RSP);
                    w_ih[l] = cur;
                    cur += (size_t)UPC * 4 * RSP;
                } else {
                    w_ih[l] = p.Wih[l] + (size_t)u0 * 4 * RSP;
                }
            }
        }
    }
    bf16 *w_p = wbf + 2 * (size_t)p.smem_lstm_floats;        // [JPC] rows, K = P
    stage_rows(w_p, p.Wp + (size_t)j0 * RSP, j1 - j0, RSP);
    const bf16 *w_o;                                         // [OPC] rows, K = J: shared if it fits
    if (p.out_in_smem) {
        bf16 *w_os = w_p + (size_t)JPC * RSP;
        stage_rows(w_os, p.Wout + (size_t)o0 * RSJ, o1 - o0, RSJ);
        w_o = w_os;
    } else {
        w_o = p.Wout + (size_t)o0 * RSJ;
    }
    for (int i = tid; i < L * 2 * UPC * Bpad; i += blockDim.x) csm[i] = 0.f;   // zero cell state (tdt.cpp:49-59)
    for (int b = tid; b < Bpad; b += blockDim.x) {                             // initial decode state
        s_cur[b] = 0;
        s_token[b] = V - 1;
        s_tpos[b] = 0;
        s_active[b] = b < p.n_utt ? 1 : 0;
        s_ntok[b] = 0;
        s_pend[b] = -1;
    }
    __syncthreads();

    // h: bf16 planes [hi|lo][L][2][Bpad][P] (two state planes per utterance); z: [hi|lo][Bpad][J]
    const size_t HS = (size_t)P * Bpad;
    bf16 *hb = reinterpret_cast<bf16 *>(p.hbuf), *zb = reinterpret_cast<bf16 *>(p.z);
    const size_t h_lo = (size_t)L * 2 * HS, z_lo = (size_t)J * Bpad;
    const size_t KB = (size_t)Bpad;      // one key buffer
    const size_t PB = (size_t)G * Bpad;  // one partial buffer
    unsigned int nbar = 0;
    // deferred confidence: 1 / sum_q lsum_q * exp(lmax_q - gmax) over the per-CTA partials of `buf`
    auto finalize_conf = [&](int buf) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int b = g + warp * G; b < p.n_utt; b += G * NWARP) {   // utterances owned by this CTA (b % G == g)
            const int slot = s_pend[b];
            if (slot < 0) continue;
            float gmax = -INFINITY;
            for (int q = lane; q < G; q += 32) gmax = fmaxf(gmax, p.pl_max[buf * PB + (size_t)q * Bpad + b]);
            gmax = warp_max(gmax);
            float s = 0.f;
            for (int q = lane; q < G; q += 32) {
                const float m = p.pl_max[buf * PB + (size_t)q * Bpad + b];
                if (m > -INFINITY) s += p.pl_sum[buf * PB + (size_t)q * Bpad + b] * expf(m - gmax);
            }
            s = warp_sum(s);
            if (lane == 0) p.t_conf[(size_t)b * p.cap + slot] = 1.0f / s;
        }
    };

    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
    auto tick = [&](int slot) {   // phase timing of CTA 0 (debug aid, p.dbg may be null)
        const long long now = clock64();
        tacc[slot] += now - tprev;
        tprev = now;
    };
    int step = 0;
    for (;; ++step) {
        const int kb = step % 3;
        if (g == 0)   // reset the key buffer of the NEXT step (last read two barriers ago)
            for (int b = tid; b < Bpad; b += blockDim.x) {
                p.key_lab[((step + 1) % 3) * KB + b] = 0ull;
                p.key_dur[((step + 1) % 3) * KB + b] = 0ull;
            }
        // ================= P1: LSTM layers =================
        for (int l = 0; l < L; ++l) {
            for (int bc = 0; bc < Bpad; bc += BCH) {
                const int R = nU * 4;
                if (R > 0) {
                    // gate pre-activations: W_hh . h_prev + (W_ih0 . E[token] + b0 | b_l)   [+ W_ih . h'_{l-1}]
                    auto pre_g = [&](int r, int b) {
                        const int u = u0 + (r >> 2), gt = r & 3;
                        return (l == 0) ? p.G0[(size_t)s_token[b] * 4 * P + gt * P + u] : p.bih[l][gt * P + u];
                    };
                    auto xprev = [&](int b) { return hb + ((size_t)(l * 2 + s_cur[b])) * HS + (size_t)b * P; };
                    rows_times_batch<true>(w_hh[l], R, P, Bpad, bc, xprev, h_lo, xs, red, pre_g,
                                           [&](int r, int b, float v, float e) { gsm[r * BCH + (b - bc)] = v + e; });
                    if (l > 0) {  // input part: W_ih . h'_{l-1}(new)
                        auto xh = [&](int b) { return hb + ((size_t)((l - 1) * 2 + (1 - s_cur[b]))) * HS + (size_t)b * P; };
                        auto nopre = [&](int, int) { return 0.f; };
                        auto fa = [&](int r, int b, float v, float) { gsm[r * BCH + (b - bc)] += v; };
                        if (p.wih_in_smem) rows_times_batch<true>(w_ih[l], R, P, Bpad, bc, xh, h_lo, xs, red, nopre, fa);
                        else rows_times_batch<false>(w_ih[l], R, P, Bpad, bc, xh, h_lo, xs, red, nopre, fa);
                    }
                    for (int idx = tid; idx < nU * BCH; idx += blockDim.x) {
                        const int ul = idx / BCH, bb = idx % BCH, b = bc + bb;
                        if (b >= Bpad) continue;
                        const float gi = gsm[(ul * 4 + 0) * BCH + bb], gf = gsm[(ul * 4 + 1) * BCH + bb];
                        const float gg = gsm[(ul * 4 + 2) * BCH + bb], go = gsm[(ul * 4 + 3) * BCH + bb];
                        const int cu = s_cur[b];
                        const float c_old = csm[((size_t)(l * 2 + cu) * UPC + ul) * Bpad + b];
                        const float c_new = sigmoidf_(gf) * c_old + sigmoidf_(gi) * tanhf(gg);
                        const float h_new = sigmoidf_(go) * tanhf(c_new);
                        csm[((size_t)(l * 2 + 1 - cu) * UPC + ul) * Bpad + b] = c_new;
                        store_split(hb, hb + h_lo, ((size_t)(l * 2 + 1 - cu)) * HS + (size_t)b * P + u0 + ul, h_new);
                    }
                }
            }
            tick(0);
            grid_arrive(p.bar);
            if (l == L - 1) {
                // confidences of the tokens emitted in the previous step: their partials were complete at
                // that step's last barrier, so this sits in the shadow of the barrier wait
                if (step > 0) finalize_conf((step - 1) % 3);
                __syncthreads();
                for (int b = tid; b < Bpad; b += blockDim.x) s_pend[b] = -1;
            }
            grid_wait(p.bar, G * (++nbar));
            tick(1);
        }
        // ================= P2: joint hidden z = relu(EP[t] + Wp . h') =================
        for (int bc = 0; bc < Bpad; bc += BCH)
            for (int rg = j0; rg < j1; rg += RMAX) {
                const int R = min(RMAX, j1 - rg);
                rows_times_batch<true>(
                    w_p + (size_t)(rg - j0) * RSP, R, P, Bpad, bc,
                    [&](int b) { return hb + ((size_t)((L - 1) * 2 + (1 - s_cur[b]))) * HS + (size_t)b * P; }, h_lo, xs, red,
                    [&](int r, int b) {
                        if (b >= p.n_utt) return 0.f;
                        const int T = p.row_off[b + 1] - p.row_off[b];
                        const int t = min(s_tpos[b], T - 1);
                        return p.EP[(size_t)(p.row_off[b] + t) * J + rg + r];
                    },
                    [&](int r, int b, float v, float e) { store_split(zb, zb + z_lo, (size_t)b * J + rg + r, fmaxf(v + e, 0.f)); });
            }
        tick(2);
        grid_arrive(p.bar);
        grid_wait(p.bar, G * (++nbar));
        tick(3);
        // ================= P3: logits -> per-CTA partials + global arg-max keys =================
        for (int bc = 0; bc < Bpad; bc += BCH) {
            float lmax = -INFINITY, lsum = 0.f, dmax = -INFINITY;
            int lidx = 0x7fffffff, didx = 0x7fffffff;
            for (int rg = o0; rg < o1; rg += RMAX) {
                const int R = min(RMAX, o1 - rg);
                auto xz = [&](int b) { return zb + (size_t)b * J; };
                auto pb = [&](int r, int) { return p.bout[rg + r]; };
                auto fl = [&](int r, int b, float v, float e) { gsm[r * BCH + (b - bc)] = v + e; };
                if (p.out_in_smem)
                    rows_times_batch<true>(w_o + (size_t)(rg - o0) * RSJ, R, J, Bpad, bc, xz, z_lo, xs, red, pb, fl);
                else
                    rows_times_batch<false>(w_o + (size_t)(rg - o0) * RSJ, R, J, Bpad, bc, xz, z_lo, xs, red, pb, fl);
                if (tid < BCH) {
                    for (int r = 0; r < R; ++r) {
                        const float v = gsm[r * BCH + tid];
                        const int n = rg + r;
                        if (n < V) {
                            if (v > lmax) {
                                lsum = lsum * expf(lmax - v) + 1.f;
                                lmax = v;
                                lidx = n;
                            } else {
                                lsum += expf(v - lmax);
                            }
                        } else if (v > dmax) {
                            dmax = v;
                            didx = n - V;
                        }
                    }
                }
                __syncthreads();
            }
            if (tid < BCH && bc + tid < Bpad) {
                const int b = bc + tid;
                p.pl_max[kb * PB + (size_t)g * Bpad + b] = lmax;
                p.pl_sum[kb * PB + (size_t)g * Bpad + b] = lsum;
                if (b < p.n_utt && s_active[b]) {
                    if (lmax > -INFINITY) atomicMax(&p.key_lab[kb * KB + b], pack_key(lmax, lidx));
                    if (dmax > -INFINITY) atomicMax(&p.key_dur[kb * KB + b], pack_key(dmax, didx));
                }
            }
        }
        tick(4);
        grid_arrive(p.bar);
        grid_wait(p.bar, G * (++nbar));
        tick(5);
        // ================= P4 (replicated in every CTA): state update =================
        int any = 0;
        for (int b = tid; b < p.n_utt; b += blockDim.x) {
            if (!s_active[b]) continue;
            float lmax, dmax;
            int lidx, didx;
            unpack_key(p.key_lab[kb * KB + b], lmax, lidx);
            unpack_key(p.key_dur[kb * KB + b], dmax, didx);
            const int T = p.row_off[b + 1] - p.row_off[b];
            const int skip = (didx < p.n_dur) ? p.durations[didx] : 1;
            int t = s_tpos[b];
            bool act = true;
            if (lidx == V - 1) {                 // blank: LSTM state reverts (cur unchanged)
                t += max(skip, 1);
            } else {
                const int n = s_ntok[b];
                if (n < p.cap && (b % G) == g) { // the owner CTA writes the token; confidence follows
                    int32_t *row = p.tok + (size_t)b * (1 + p.cap);
                    row[1 + n] = lidx;
                    p.t_start[(size_t)b * p.cap + n] = t;
                    p.t_end[(size_t)b * p.cap + n] = min(t + max(skip, 1) - 1, T - 1);
                    row[0] = n + 1;
                }
                if (n < p.cap) s_pend[b] = n;
                s_ntok[b] = n + 1;
                s_token[b] = lidx;
                s_cur[b] = 1 - s_cur[b];         // commit the new LSTM state
                t += skip;
                if (n + 1 >= p.cap) {
                    act = false;
                    if ((b % G) == g) p.overflow[b] = 1;
                }
            }
            s_tpos[b] = t;
            if (t >= T) act = false;
            s_active[b] = act ? 1 : 0;
            any |= act ? 1 : 0;
        }
        any = __syncthreads_or(any);
        tick(6);
        if (!any || step + 1 >= p.max_steps) break;
    }
    if (p.dbg && g == 0 && tid == 0) {
        for (int i = 0; i < 7; ++i) p.dbg[i] = tacc[i];
        p.dbg[7] = step + 1;
    }
    // confidences of the last step's tokens (all partials were visible at that step's last barrier)
    finalize_conf(step % 3);
}

__global__ void tdt_init_kernel(TdtParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *p.bar = 0u;
    if (b < 3 * p.Bpad) {
        p.key_lab[b] = 0ull;
        p.key_dur[b] = 0ull;
    }
    if (b < p.n_utt) {
        p.tok[(size_t)b * (1 + p.cap)] = 0;
        p.overflow[b] = 0;
    }
}

// fp32 rows [rows][K] -> pre-split rows [rows][2 * (K + 4)] = [hi: K+4][lo: K+4] bf16 (pads zero)
__global__ void tdt_split_rows_kernel(const float *__restrict__ src, int rows, int K, bf16 *__restrict__ dst) {
    const int KP = K + 4;
    const size_t n = (size_t)rows * KP;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / KP;
        const int k = (int)(i - r * KP);
        const float v = k < K ? src[r * K + k] : 0.f;
        const bf16 h = __float2bfloat16_rn(v);
        dst[r * 2 * KP + k] = h;
        dst[r * 2 * KP + KP + k] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

}  // namespace

void launch_tdt_split_rows(const float *src, int rows, int K, bf16 *dst, cudaStream_t st) {
    tdt_split_rows_kernel<<<256, 256, 0, st>>>(src, rows, K, dst);
}

size_t tdt_smem_bytes(const TdtParams &p, int grid, bool *out_in_smem, bool *wih_in_smem, int *lstm_floats) {
    const int UPC = (p.P + grid - 1) / grid, JPC = (p.J + grid - 1) / grid, OPC = (p.V + p.D + grid - 1) / grid;
    const size_t budget = 225 * 1024 / sizeof(float);
    size_t fixed = (size_t)2 * RPAD * RLD + RMAX * BCH + (size_t)NST * STAGE_ELEMS / 2 + (size_t)p.L * 2 * UPC * p.Bpad +
                   6 * (size_t)p.Bpad;
    // weight rows are bf16 hi/lo with a 4-element pad each: K + 4 floats per row
    const size_t hh = (size_t)p.L * UPC * 4 * (p.P + 4), ih = (size_t)(p.L - 1) * UPC * 4 * (p.P + 4);
    const size_t wp = (size_t)JPC * (p.P + 4), wo = (size_t)OPC * (p.J + 4);
    size_t total = fixed + hh + wp;                 // always resident
    *wih_in_smem = (ih == 0) || (total + ih <= budget);
    if (*wih_in_smem) total += ih;
    *lstm_floats = (int)(hh + (*wih_in_smem ? ih : 0));
    *out_in_smem = total + wo <= budget;
    if (*out_in_smem) total += wo;
    return total * sizeof(float);
}

cudaError_t launch_tdt_decode(TdtParams p, int num_sms, cudaStream_t st) {
    if (p.P % KC || p.J % KC) return cudaErrorInvalidValue;
    int grid = num_sms;
    // every CTA must own <= 5 LSTM units (RMAX = 20 gate rows)
    if ((p.P + grid - 1) / grid * 4 > RMAX) return cudaErrorInvalidConfiguration;
    bool out_in_smem, wih_in_smem;
    int lstm_floats;
    size_t smem = tdt_smem_bytes(p, grid, &out_in_smem, &wih_in_smem, &lstm_floats);
    p.out_in_smem = out_in_smem ? 1 : 0;
    p.wih_in_smem = wih_in_smem ? 1 : 0;
    p.smem_lstm_floats = lstm_floats;
    cudaError_t err = cudaFuncSetAttribute(tdt_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tdt_decode_kernel, NWARP * 32, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    tdt_init_kernel<<<(3 * p.Bpad + 127) / 128, 128, 0, st>>>(p);
    void *args[] = {&p};
    return cudaLaunchCooperativeKernel((void *)tdt_decode_kernel, dim3(grid), dim3(NWARP * 32), args, smem, st);
}

}  // namespace pk
