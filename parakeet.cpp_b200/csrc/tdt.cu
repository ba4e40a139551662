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
// only the tiny per-utterance vectors (h, z; [utterance][k]) travel through L2 (staged into
// shared memory by a cp.async ring; a warp lane is an utterance) between the phases of a step, separated by grid barriers:
//   P1 LSTM gates + cell (per layer) | P2 joint hidden | P3 logits -> per-CTA (max, sum-exp)
//   partials + atomicMax of packed (value, index) keys | P4 state update, replicated in every
//   CTA from the keys (no barrier before the next P1; confidences are finalised one phase later
//   from the partials, in a fixed order, by the CTA that owns the utterance).
// Three monotonic-counter grid barriers per step (cooperative launch guarantees co-residency).
// enc_proj(enc)+bias for all frames and the layer-0 input table W_ih.E[token]+b for all
// tokens are precomputed by GEMMs (engine.cu).
#include "kernels.h"

namespace pk {
namespace {

constexpr int RMAX = 20;   // rows accumulated per pass (5 LSTM units x 4 gates)
constexpr int NWARP = 8;
constexpr int BCH = 64;    // utterances per pass (2 per lane)

constexpr int KC = NWARP * 4;   // k-values staged per chunk (4 per warp)
constexpr int NST = 8;          // cp.async ring depth
constexpr int XLD = KC + 4;     // staged row stride (floats), CUDA-core path: 16 B aligned, conflict-free float4 reads
constexpr int XLDM = KC + 8;    // staged row stride, tensor-core path: conflict-free float2 A-fragment reads
constexpr int RPAD = 24;        // RMAX rounded up to whole 8-row MMA n-blocks

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool WS>
__device__ __forceinline__ float4 load_w4(const float *p) {
    if (WS) {
        float4 v;   // weights never change during the kernel: not volatile, so loads can be batched ahead of the FMAs
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"((uint32_t)__cvta_generic_to_shared(p)));
        return v;
    }
    return __ldg(reinterpret_cast<const float4 *>(p));
}

// out(r, b) = sum_k W[r][k] * x_b[k] for r < R (R <= RMAX), b in [bc, bc+64).
// W rows are contiguous [R][K] (shared memory when WS, else global).  xsrc(b) -> pointer to the
// K contiguous floats of utterance b (global; lives in L2).  x is streamed through a 4-deep
// cp.async ring in shared memory, 64 utterances x 32 k-values per chunk (16-byte copies that
// bypass L1; a warp fetches 4 full lines), so the L2 latency is paid once per phase; within a chunk warp w owns k = 4w..4w+3
// (K-split) and a lane owns utterances (lane, lane+32).  Partials are reduced across the 8
// warps through `red`.
// RB = compile-time bound on the rows of this call (the row loop is fully unrolled and the
// compiler if-converts `r < R`, so every unrolled row costs its FMAs whether it is live or not).
template <int RB, bool WS, typename XSrc, typename Fin>
__device__ __forceinline__ void rows_times_batch_rb(const float *W, int R, int K, int Bpad, int bc, XSrc xsrc,
                                                    float *xs, float *red, Fin fin) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = K / KC;
    // this thread copies 16-byte piece `pc` of staged rows b_a and b_a + 32 (a warp = 4 full 128 B lines)
    const int pc = threadIdx.x & 7, b_a = threadIdx.x >> 3;
    const bool va = (bc + b_a) < Bpad, vb = (bc + b_a + 32) < Bpad;
    const float *xa_src = xsrc(va ? bc + b_a : 0) + pc * 4;
    const float *xb_src = xsrc(vb ? bc + b_a + 32 : 0) + pc * 4;
    auto issue = [&](int c) {
        if (c < nchunks) {
            float *dst = xs + (size_t)(c % NST) * BCH * XLD + pc * 4;
            if (va) cp_async16(dst + b_a * XLD, xa_src + c * KC);
            if (vb) cp_async16(dst + (b_a + 32) * XLD, xb_src + c * KC);
        }
        cp_async_commit();
    };
    float acc[RB][2];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r][0] = acc[r][1] = 0.f;
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) issue(c);
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<NST - 2>();
        __syncthreads();                       // chunk c landed for everyone; chunk c-1 fully consumed
        issue(c + NST - 1);
        const float *xc = xs + (size_t)(c % NST) * BCH * XLD + warp * 4;
        const float4 xa = *reinterpret_cast<const float4 *>(xc + lane * XLD);
        const float4 xb = *reinterpret_cast<const float4 *>(xc + (lane + 32) * XLD);
        const int k = c * KC + warp * 4;
        // rows in groups of up to 5: issue the group's weight loads first, then its 40 FMAs, so the
        // shared-memory latency is paid once per group instead of once per row
#pragma unroll
        for (int r0 = 0; r0 < RB; r0 += 5) {
            float4 w[5];
#pragma unroll
            for (int j = 0; j < 5; ++j)
                if (r0 + j < RB) w[j] = load_w4<WS>(W + (size_t)min(r0 + j, R - 1) * K + k);
#pragma unroll
            for (int j = 0; j < 5; ++j)
                if (r0 + j < RB) {
                    const int r = r0 + j;
                    acc[r][0] = fmaf(w[j].x, xa.x, acc[r][0]);
                    acc[r][1] = fmaf(w[j].x, xb.x, acc[r][1]);
                    acc[r][0] = fmaf(w[j].y, xa.y, acc[r][0]);
                    acc[r][1] = fmaf(w[j].y, xb.y, acc[r][1]);
                    acc[r][0] = fmaf(w[j].z, xa.z, acc[r][0]);
                    acc[r][1] = fmaf(w[j].z, xb.z, acc[r][1]);
                    acc[r][0] = fmaf(w[j].w, xa.w, acc[r][0]);
                    acc[r][1] = fmaf(w[j].w, xb.w, acc[r][1]);
                }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int r = 0; r < RB; ++r)
        if (r < R) {
            red[(warp * RMAX + r) * BCH + lane] = acc[r][0];
            red[(warp * RMAX + r) * BCH + 32 + lane] = acc[r][1];
        }
    __syncthreads();
    for (int idx = threadIdx.x; idx < R * BCH; idx += blockDim.x) {
        const int r = idx / BCH, b2 = idx % BCH;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) s += red[(w * RMAX + r) * BCH + b2];
        if (bc + b2 < Bpad) fin(r, bc + b2, s);
    }
    __syncthreads();
}

template <bool WS, typename XSrc, typename Fin>
__device__ __forceinline__ void rows_times_batch(const float *W, int R, int K, int Bpad, int bc, XSrc xsrc,
                                                 float *xs, float *red, Fin fin) {
    if (R <= 5) rows_times_batch_rb<5, WS>(W, R, K, Bpad, bc, xsrc, xs, red, fin);
    else if (R <= 8) rows_times_batch_rb<8, WS>(W, R, K, Bpad, bc, xsrc, xs, red, fin);
    else if (R <= 12) rows_times_batch_rb<12, WS>(W, R, K, Bpad, bc, xsrc, xs, red, fin);
    else rows_times_batch_rb<RMAX, WS>(W, R, K, Bpad, bc, xsrc, xs, red, fin);
}

// ---- tensor-core variant for weights resident in shared memory --------------------------------
// Same contract as rows_times_batch_rb, on mma.sync.m16n8k16 with the bf16 hi/lo operand split of
// the encoder GEMMs (x_hi.W_hi + x_hi.W_lo + x_lo.W_hi, fp32 accumulate: ~16 mantissa bits).
//   M = utterances (64 = 4 blocks of 16), N = weight rows (NB blocks of 8), K in chunks of KC = 32.
//   W is pre-split at kernel start: row r = [hi: K+4 bf16][lo: K+4 bf16] (the +4 pad makes the
//   32-bit B-fragment reads of 8 rows conflict-free).  x arrives as fp32 [utterance][k] through the
//   same cp.async ring (row stride XLDM) and is split in registers while building A fragments.
//   Warp w owns utterance block (w & 3) and k-step (w >> 2) of every chunk; the two k-step halves
//   are added in `red` in a fixed order.
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2(float2 x, uint32_t &hi, uint32_t &lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x.x, x.y);
    float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(x.x - hf.x, x.y - hf.y);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = *reinterpret_cast<uint32_t *>(&l);
}
__device__ __forceinline__ uint32_t lds32(const bf16 *p) {
    uint32_t v;   // weights never change during the kernel: not volatile
    asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

template <int NB, typename XSrc, typename Fin>
__device__ __forceinline__ void rows_times_batch_mma(const bf16 *W, int R, int K, int Bpad, int bc, XSrc xsrc,
                                                     float *xs, float *red, Fin fin) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, cq = lane & 3;
    const int mb = warp & 3, ks = warp >> 2;
    const int KP = K + 4, RS = 2 * KP;
    const int nchunks = K / KC;
    const int pc = threadIdx.x & 7, b_a = threadIdx.x >> 3;
    const bool va = (bc + b_a) < Bpad, vb = (bc + b_a + 32) < Bpad;
    const float *xa_src = xsrc(va ? bc + b_a : 0) + pc * 4;
    const float *xb_src = xsrc(vb ? bc + b_a + 32 : 0) + pc * 4;
    auto issue = [&](int c) {
        if (c < nchunks) {
            float *dst = xs + (size_t)(c % NST) * BCH * XLDM + pc * 4;
            if (va) cp_async16(dst + b_a * XLDM, xa_src + c * KC);
            if (vb) cp_async16(dst + (b_a + 32) * XLDM, xb_src + c * KC);
        }
        cp_async_commit();
    };
    float acc[NB][4];
    const bf16 *wrow[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
        wrow[nb] = W + (size_t)min(nb * 8 + g, R - 1) * RS + ks * 16 + 2 * cq;   // rows >= R: clamped, discarded
    }
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) issue(c);
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<NST - 2>();
        __syncthreads();                       // chunk c landed for everyone; chunk c-1 fully consumed
        issue(c + NST - 1);
        const float *xc = xs + (size_t)(c % NST) * BCH * XLDM + (mb * 16 + g) * XLDM + ks * 16 + 2 * cq;
        uint32_t ah[4], al[4];
        split2(*reinterpret_cast<const float2 *>(xc), ah[0], al[0]);
        split2(*reinterpret_cast<const float2 *>(xc + 8 * XLDM), ah[1], al[1]);
        split2(*reinterpret_cast<const float2 *>(xc + 8), ah[2], al[2]);
        split2(*reinterpret_cast<const float2 *>(xc + 8 * XLDM + 8), ah[3], al[3]);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const bf16 *wp = wrow[nb] + c * KC;
            const uint32_t b0h = lds32(wp), b1h = lds32(wp + 8), b0l = lds32(wp + KP), b1l = lds32(wp + KP + 8);
            mma_bf16(acc[nb], ah, b0h, b1h);
            mma_bf16(acc[nb], ah, b0l, b1l);
            mma_bf16(acc[nb], al, b0h, b1h);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        const int n = nb * 8 + 2 * cq, u = mb * 16 + g;
        red[(ks * RPAD + n) * BCH + u] = acc[nb][0];
        red[(ks * RPAD + n + 1) * BCH + u] = acc[nb][1];
        red[(ks * RPAD + n) * BCH + u + 8] = acc[nb][2];
        red[(ks * RPAD + n + 1) * BCH + u + 8] = acc[nb][3];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < R * BCH; idx += blockDim.x) {
        const int r = idx / BCH, b2 = idx % BCH;
        const float s = red[r * BCH + b2] + red[(RPAD + r) * BCH + b2];
        if (bc + b2 < Bpad) fin(r, bc + b2, s);
    }
    __syncthreads();
}

template <typename XSrc, typename Fin>
__device__ __forceinline__ void rows_times_batch_s(const bf16 *W, int R, int K, int Bpad, int bc, XSrc xsrc, float *xs,
                                                   float *red, Fin fin) {
    if (R <= 8) rows_times_batch_mma<1>(W, R, K, Bpad, bc, xsrc, xs, red, fin);
    else if (R <= 16) rows_times_batch_mma<2>(W, R, K, Bpad, bc, xsrc, xs, red, fin);
    else rows_times_batch_mma<3>(W, R, K, Bpad, bc, xsrc, xs, red, fin);
}

// Monotonic-counter grid barrier (all CTAs are co-resident: cooperative launch).  Cheaper than
// cooperative_groups' grid.sync() and traps instead of hanging if a CTA never arrives.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
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

    // ---- shared memory carve-up: [red][gates][state][weights...]
    float *red = sm;                               // [NWARP][RMAX][BCH]
    float *gsm = red + NWARP * RMAX * BCH;         // [RMAX][BCH] gate pre-activations / logits
    float *xs = gsm + RMAX * BCH;                  // [NST][BCH][XLDM] cp.async ring for the x vectors
    int *s_cur = reinterpret_cast<int *>(xs + NST * BCH * XLDM);  // replicated decode state, [Bpad] each
    int *s_token = s_cur + Bpad, *s_tpos = s_token + Bpad, *s_active = s_tpos + Bpad, *s_ntok = s_active + Bpad;
    int *s_pend = s_ntok + Bpad;                   // slot of a token whose confidence is still pending (-1: none)
    float *wsm = reinterpret_cast<float *>(s_pend + Bpad);
    const int nU = u1 - u0;
    // LSTM weights arrive "unit-major" (row = unit*4 + gate, engine.cu), so this CTA's rows
    // [u0*4, u1*4) are one contiguous block: W_hh always lives in shared memory, W_ih of the
    // upper layers too when it fits (else it is streamed from L2).
    // Shared-memory weights are stored pre-split for the tensor-core path: row = [hi: K+4][lo: K+4] bf16
    // (= K+4 floats per row).
    auto stage_rows = [&](bf16 *dst, const float *src, int rows, int K) {
        const int KP = K + 4, RS = 2 * KP;
        for (int idx = tid; idx < rows * K; idx += blockDim.x) {
            const int r = idx / K, k = idx - r * K;
            const float v = src[idx];
            const bf16 h = __float2bfloat16_rn(v);
            dst[(size_t)r * RS + k] = h;
            dst[(size_t)r * RS + KP + k] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    };
    const int RSP = 2 * (P + 4), RSJ = 2 * (J + 4);          // smem row strides (bf16) for K = P / K = J
    const bf16 *w_hh[PK_MAX_LSTM], *w_ih_s[PK_MAX_LSTM];
    const float *w_ih_g[PK_MAX_LSTM];
    bf16 *wbf = reinterpret_cast<bf16 *>(wsm);
    {
        bf16 *cur = wbf;
        for (int l = 0; l < L; ++l) {
            stage_rows(cur, p.Whh[l] + (size_t)u0 * 4 * P, nU * 4, P);
            w_hh[l] = cur;
            cur += (size_t)UPC * 4 * RSP;
            w_ih_s[l] = nullptr;
            w_ih_g[l] = nullptr;
            if (l > 0) {
                if (p.wih_in_smem) {
                    stage_rows(cur, p.Wih[l] + (size_t)u0 * 4 * P, nU * 4, P);
                    w_ih_s[l] = cur;
                    cur += (size_t)UPC * 4 * RSP;
                } else {
                    w_ih_g[l] = p.Wih[l] + (size_t)u0 * 4 * P;
                }
            }
        }
    }
    bf16 *w_p = wbf + 2 * (size_t)p.smem_lstm_floats;        // [JPC] rows, K = P
    stage_rows(w_p, p.Wp + (size_t)j0 * P, j1 - j0, P);
    const bf16 *w_o_s = nullptr;                             // [OPC] rows, K = J: shared if it fits
    const float *w_o_g = nullptr;
    if (p.out_in_smem) {
        bf16 *w_os = w_p + (size_t)JPC * RSP;
        stage_rows(w_os, p.Wout + (size_t)o0 * J, o1 - o0, J);
        w_o_s = w_os;
    } else {
        w_o_g = p.Wout + (size_t)o0 * J;
    }
    for (int b = tid; b < Bpad; b += blockDim.x) {           // initial state (tdt.cpp:49-59)
        s_cur[b] = 0;
        s_token[b] = V - 1;
        s_tpos[b] = 0;
        s_active[b] = b < p.n_utt ? 1 : 0;
        s_ntok[b] = 0;
        s_pend[b] = -1;
    }
    __syncthreads();

    const size_t HS = (size_t)P * Bpad;  // one h/c plane
    // All CTAs read the same h / z lines in the same microsecond, and an L2 slice serialises
    // requests to one line: the vectors are therefore kept in NREP replicas (writers store every
    // replica, CTA g reads replica g % NREP), which spreads the readers of a line 8 ways.
    const int rep = g % TDT_NREP;
    const size_t HREP = (size_t)L * 2 * HS, ZREP = (size_t)J * Bpad;
    const float *h_rd = p.hbuf + (size_t)rep * HREP;
    const float *z_rd = p.z + (size_t)rep * ZREP;
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
                    rows_times_batch_s(
                        w_hh[l], R, P, Bpad, bc,
                        [&](int b) { return h_rd + ((size_t)(l * 2 + s_cur[b])) * HS + (size_t)b * P; }, xs, red,
                        [&](int r, int b, float v) { gsm[r * BCH + (b - bc)] = v; });
                    if (p.dbg_variant == 1) tick(0);
                    if (l > 0) {  // input part: W_ih . h'_{l-1}(new)
                        auto xh = [&](int b) { return h_rd + ((size_t)((l - 1) * 2 + (1 - s_cur[b]))) * HS + (size_t)b * P; };
                        auto fa = [&](int r, int b, float v) { gsm[r * BCH + (b - bc)] += v; };
                        if (p.wih_in_smem) rows_times_batch_s(w_ih_s[l], R, P, Bpad, bc, xh, xs, red, fa);
                        else rows_times_batch<false>(w_ih_g[l], R, P, Bpad, bc, xh, xs, red, fa);
                    }
                    __syncthreads();
                    for (int idx = tid; idx < nU * BCH; idx += blockDim.x) {
                        const int ul = idx / BCH, bb = idx % BCH, b = bc + bb;
                        if (b >= Bpad) continue;
                        const int u = u0 + ul;
                        float gi = gsm[(ul * 4 + 0) * BCH + bb], gf = gsm[(ul * 4 + 1) * BCH + bb];
                        float gg = gsm[(ul * 4 + 2) * BCH + bb], go = gsm[(ul * 4 + 3) * BCH + bb];
                        if (l == 0) {
                            const float *row = p.G0 + (size_t)s_token[b] * 4 * P;
                            gi += row[u]; gf += row[P + u]; gg += row[2 * P + u]; go += row[3 * P + u];
                        } else {
                            const float *bi = p.bih[l];
                            gi += bi[u]; gf += bi[P + u]; gg += bi[2 * P + u]; go += bi[3 * P + u];
                        }
                        const int cu = s_cur[b];
                        const float c_old = p.cbuf[((size_t)(l * 2 + cu)) * HS + (size_t)b * P + u];
                        const float c_new = sigmoidf_(gf) * c_old + sigmoidf_(gi) * tanhf(gg);
                        const float h_new = sigmoidf_(go) * tanhf(c_new);
                        p.cbuf[((size_t)(l * 2 + 1 - cu)) * HS + (size_t)b * P + u] = c_new;
                        for (int rr = 0; rr < TDT_NREP; ++rr)
                            p.hbuf[(size_t)rr * HREP + ((size_t)(l * 2 + 1 - cu)) * HS + (size_t)b * P + u] = h_new;
                    }
                    __syncthreads();
                }
            }
            tick(p.dbg_variant == 1 ? 1 : 0);
            grid_barrier(p.bar, G * (++nbar));
            tick(p.dbg_variant == 1 ? 4 : 1);
        }
        // confidences of the tokens emitted in the previous step (partials are complete now)
        if (step > 0) finalize_conf((step - 1) % 3);
        __syncthreads();
        for (int b = tid; b < Bpad; b += blockDim.x) s_pend[b] = -1;
        if (p.dbg_variant == 1) tick(2);
        // ================= P2: joint hidden z = relu(EP[t] + Wp . h') =================
        for (int bc = 0; bc < Bpad; bc += BCH)
            for (int rg = j0; rg < j1; rg += RMAX) {
                const int R = min(RMAX, j1 - rg);
                rows_times_batch_s(
                    w_p + (size_t)(rg - j0) * RSP, R, P, Bpad, bc,
                    [&](int b) { return h_rd + ((size_t)((L - 1) * 2 + (1 - s_cur[b]))) * HS + (size_t)b * P; }, xs, red,
                    [&](int r, int b, float v) {
                        float e = 0.f;
                        if (b < p.n_utt) {
                            const int T = p.row_off[b + 1] - p.row_off[b];
                            const int t = min(s_tpos[b], T - 1);
                            e = p.EP[(size_t)(p.row_off[b] + t) * J + rg + r];
                        }
                        const float zv = fmaxf(v + e, 0.f);
                        for (int rr = 0; rr < TDT_NREP; ++rr) p.z[(size_t)rr * ZREP + (size_t)b * J + rg + r] = zv;
                    });
            }
        tick(p.dbg_variant == 1 ? 3 : 2);
        grid_barrier(p.bar, G * (++nbar));
        tick(p.dbg_variant == 1 ? 4 : 3);
        // ================= P3: logits -> per-CTA partials + global arg-max keys =================
        for (int bc = 0; bc < Bpad; bc += BCH) {
            float lmax = -INFINITY, lsum = 0.f, dmax = -INFINITY;
            int lidx = 0x7fffffff, didx = 0x7fffffff;
            for (int rg = o0; rg < o1; rg += RMAX) {
                const int R = min(RMAX, o1 - rg);
                auto xz = [&](int b) { return z_rd + (size_t)b * J; };
                auto fl = [&](int r, int b, float v) { gsm[r * BCH + (b - bc)] = v + p.bout[rg + r]; };
                if (p.out_in_smem)
                    rows_times_batch_s(w_o_s + (size_t)(rg - o0) * RSJ, R, J, Bpad, bc, xz, xs, red, fl);
                else
                    rows_times_batch<false>(w_o_g + (size_t)(rg - o0) * J, R, J, Bpad, bc, xz, xs, red, fl);
                __syncthreads();
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
        grid_barrier(p.bar, G * (++nbar));
        tick(p.dbg_variant == 1 ? 4 : 5);
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
        tick(p.dbg_variant == 1 ? 4 : 6);
        if (!any || step + 1 >= p.max_steps) break;
    }
    if (p.dbg && g == 0 && tid == 0) {
        for (int i = 0; i < 7; ++i) p.dbg[i] = tacc[i];
        p.dbg[7] = step + 1;
    }
    // confidences of the last step's tokens: every CTA's partials must be visible first
    grid_barrier(p.bar, G * (++nbar));
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

}  // namespace

size_t tdt_smem_bytes(const TdtParams &p, int grid, bool *out_in_smem, bool *wih_in_smem, int *lstm_floats) {
    const int UPC = (p.P + grid - 1) / grid, JPC = (p.J + grid - 1) / grid, OPC = (p.V + p.D + grid - 1) / grid;
    const size_t budget = 225 * 1024 / sizeof(float);
    size_t fixed = (size_t)NWARP * RMAX * BCH + RMAX * BCH + (size_t)NST * BCH * XLDM + 6 * (size_t)p.Bpad;
    // shared-memory weight rows are bf16 hi/lo with a 4-element pad each: K + 4 floats per row
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
    if (p.P % 32 || p.J % 32) return cudaErrorInvalidValue;
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
