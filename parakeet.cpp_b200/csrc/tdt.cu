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
// only the tiny per-utterance vectors (h, z; stored [k][utterance] so a warp lane is an
// utterance) travel through L2 between the phases of a step, separated by grid barriers:
//   P1 LSTM gates + cell (per layer)   P2 joint hidden   P3 logits -> per-CTA partial
//   argmax / sum-exp                    P4 per-utterance reduction + state update.
// enc_proj(enc)+bias for all frames and the layer-0 input table W_ih.E[token]+b for all
// tokens are precomputed by GEMMs (engine.cu).
#include <cooperative_groups.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace pk {
namespace {

constexpr int RMAX = 20;   // rows accumulated per pass (5 LSTM units x 4 gates)
constexpr int NWARP = 8;
constexpr int BCH = 64;    // utterances per pass (2 per lane)

// out(r, b) = sum_k W[r][k] * x_b[k] for r < R (R <= RMAX), b in [bc, bc+64).
// W rows are contiguous [R][K] (shared or global); xsrc(b) -> pointer to x_b[0] with
// element stride Bpad.  K-split across the 8 warps, partials reduced through `red`.
template <typename XSrc, typename Fin>
__device__ __forceinline__ void rows_times_batch(const float *W, int R, int K, int Bpad, int bc, XSrc xsrc,
                                                 float *red, Fin fin) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kch = K / NWARP, k0 = warp * kch;
    const int b0 = bc + lane, b1 = bc + 32 + lane;
    const bool has0 = b0 < Bpad, has1 = b1 < Bpad;
    const float *x0 = xsrc(has0 ? b0 : 0), *x1 = xsrc(has1 ? b1 : 0);
    float acc[RMAX][2];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[r][0] = acc[r][1] = 0.f;
    for (int k = k0; k < k0 + kch; k += 4) {
        float xa[4], xb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xa[i] = x0[(size_t)(k + i) * Bpad];
            xb[i] = x1[(size_t)(k + i) * Bpad];
        }
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
            if (r < R) {
                const float4 w = *reinterpret_cast<const float4 *>(W + (size_t)r * K + k);
                acc[r][0] = fmaf(w.x, xa[0], acc[r][0]);
                acc[r][0] = fmaf(w.y, xa[1], acc[r][0]);
                acc[r][0] = fmaf(w.z, xa[2], acc[r][0]);
                acc[r][0] = fmaf(w.w, xa[3], acc[r][0]);
                acc[r][1] = fmaf(w.x, xb[0], acc[r][1]);
                acc[r][1] = fmaf(w.y, xb[1], acc[r][1]);
                acc[r][1] = fmaf(w.z, xb[2], acc[r][1]);
                acc[r][1] = fmaf(w.w, xb[3], acc[r][1]);
            }
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
        if (r < R) {
            red[(warp * RMAX + r) * BCH + lane] = acc[r][0];
            red[(warp * RMAX + r) * BCH + 32 + lane] = acc[r][1];
        }
    __syncthreads();
    for (int idx = threadIdx.x; idx < R * BCH; idx += blockDim.x) {
        const int r = idx / BCH, bb = idx % BCH;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) s += red[(w * RMAX + r) * BCH + bb];
        if (bc + bb < Bpad) fin(r, bc + bb, s);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NWARP * 32, 1) tdt_decode_kernel(TdtParams p) {
    cg::grid_group grid = cg::this_grid();
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

    // ---- shared memory carve-up: [red][gates][weights...]
    float *red = sm;                               // [NWARP][RMAX][BCH]
    float *gsm = red + NWARP * RMAX * BCH;         // [RMAX][BCH] gate pre-activations / logits
    float *wsm = gsm + RMAX * BCH;
    const int nU = u1 - u0;
    // per layer: Whh rows (nU*4, K=P); layers >= 1 also Wih rows
    float *w_hh[PK_MAX_LSTM], *w_ih[PK_MAX_LSTM];
    {
        float *cur = wsm;
        for (int l = 0; l < L; ++l) {
            w_hh[l] = cur;
            cur += (size_t)UPC * 4 * P;
            w_ih[l] = nullptr;
            if (l > 0) {
                w_ih[l] = cur;
                cur += (size_t)UPC * 4 * P;
            }
        }
        for (int l = 0; l < L; ++l)
            for (int idx = tid; idx < nU * 4 * P; idx += blockDim.x) {
                const int r = idx / P, k = idx % P;
                const int u = u0 + r / 4, gate = r % 4;
                w_hh[l][idx] = p.Whh[l][(size_t)(gate * P + u) * P + k];
                if (l > 0) w_ih[l][idx] = p.Wih[l][(size_t)(gate * P + u) * P + k];
            }
    }
    float *w_p = wsm + (size_t)p.smem_lstm_floats;          // [JPC][P]
    for (int idx = tid; idx < (j1 - j0) * P; idx += blockDim.x) w_p[idx] = p.Wp[(size_t)j0 * P + idx];
    const float *w_o;                                        // [OPC][J], shared if it fits
    if (p.out_in_smem) {
        float *w_os = w_p + (size_t)JPC * P;
        for (int idx = tid; idx < (o1 - o0) * J; idx += blockDim.x) w_os[idx] = p.Wout[(size_t)o0 * J + idx];
        w_o = w_os;
    } else {
        w_o = p.Wout + (size_t)o0 * J;
    }
    __syncthreads();

    const size_t HS = (size_t)P * Bpad;  // one h/c plane
    int step = 0;
    for (;; ++step) {
        if (g == 0 && tid == 0) p.n_active[(step + 1) % 3] = 0;
        // ================= P1: LSTM layers =================
        for (int l = 0; l < L; ++l) {
            for (int bc = 0; bc < Bpad; bc += BCH) {
                const int R = nU * 4;
                if (R > 0) {
                    // recurrent part: W_hh . h_l(current)
                    rows_times_batch(
                        w_hh[l], R, P, Bpad, bc,
                        [&](int b) { return p.hbuf + ((size_t)(l * 2 + p.cur[b])) * HS + b; }, red,
                        [&](int r, int b, float v) { gsm[r * BCH + (b - bc)] = v; });
                    if (l > 0) {  // input part: W_ih . h'_{l-1}(new)
                        rows_times_batch(
                            w_ih[l], R, P, Bpad, bc,
                            [&](int b) { return p.hbuf + ((size_t)((l - 1) * 2 + (1 - p.cur[b]))) * HS + b; },
                            red, [&](int r, int b, float v) { gsm[r * BCH + (b - bc)] += v; });
                    }
                    __syncthreads();
                    for (int idx = tid; idx < nU * BCH; idx += blockDim.x) {
                        const int ul = idx / BCH, bb = idx % BCH, b = bc + bb;
                        if (b >= Bpad) continue;
                        const int u = u0 + ul;
                        float gi = gsm[(ul * 4 + 0) * BCH + bb], gf = gsm[(ul * 4 + 1) * BCH + bb];
                        float gg = gsm[(ul * 4 + 2) * BCH + bb], go = gsm[(ul * 4 + 3) * BCH + bb];
                        if (l == 0) {
                            const float *row = p.G0 + (size_t)p.token[b] * 4 * P;
                            gi += row[u]; gf += row[P + u]; gg += row[2 * P + u]; go += row[3 * P + u];
                        } else {
                            const float *bi = p.bih[l];
                            gi += bi[u]; gf += bi[P + u]; gg += bi[2 * P + u]; go += bi[3 * P + u];
                        }
                        const int cu = p.cur[b];
                        const float c_old = p.cbuf[((size_t)(l * 2 + cu)) * HS + (size_t)u * Bpad + b];
                        const float c_new = sigmoidf_(gf) * c_old + sigmoidf_(gi) * tanhf(gg);
                        const float h_new = sigmoidf_(go) * tanhf(c_new);
                        p.cbuf[((size_t)(l * 2 + 1 - cu)) * HS + (size_t)u * Bpad + b] = c_new;
                        p.hbuf[((size_t)(l * 2 + 1 - cu)) * HS + (size_t)u * Bpad + b] = h_new;
                    }
                    __syncthreads();
                }
            }
            grid.sync();
        }
        // ================= P2: joint hidden z = relu(EP[t] + Wp . h') =================
        for (int bc = 0; bc < Bpad; bc += BCH)
            for (int rg = j0; rg < j1; rg += RMAX) {
                const int R = min(RMAX, j1 - rg);
                rows_times_batch(
                    w_p + (size_t)(rg - j0) * P, R, P, Bpad, bc,
                    [&](int b) { return p.hbuf + ((size_t)((L - 1) * 2 + (1 - p.cur[b]))) * HS + b; }, red,
                    [&](int r, int b, float v) {
                        float e = 0.f;
                        if (b < p.n_utt) {
                            const int T = p.row_off[b + 1] - p.row_off[b];
                            const int t = min(p.tpos[b], T - 1);
                            e = p.EP[(size_t)(p.row_off[b] + t) * J + rg + r];
                        }
                        p.z[(size_t)(rg + r) * Bpad + b] = fmaxf(v + e, 0.f);
                    });
            }
        grid.sync();
        // ================= P3: logits + per-CTA partial reductions =================
        for (int bc = 0; bc < Bpad; bc += BCH) {
            float lmax = -INFINITY, lsum = 0.f, dmax = -INFINITY;
            int lidx = 0x7fffffff, didx = 0x7fffffff;
            for (int rg = o0; rg < o1; rg += RMAX) {
                const int R = min(RMAX, o1 - rg);
                rows_times_batch(
                    w_o + (size_t)(rg - o0) * J, R, J, Bpad, bc, [&](int b) { return p.z + b; }, red,
                    [&](int r, int b, float v) { gsm[r * BCH + (b - bc)] = v + p.bout[rg + r]; });
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
                const size_t o = (size_t)g * Bpad + bc + tid;
                p.pl_max[o] = lmax; p.pl_idx[o] = lidx; p.pl_sum[o] = lsum;
                p.pd_max[o] = dmax; p.pd_idx[o] = didx;
            }
        }
        grid.sync();
        // ================= P4: per-utterance argmax + state update (one warp each) ============
        {
            const int warp = tid >> 5, lane = tid & 31;
            for (int b = g * NWARP + warp; b < p.n_utt; b += G * NWARP) {
                if (!p.active[b]) continue;
                float lmax = -INFINITY, dmax = -INFINITY;
                int lidx = 0x7fffffff, didx = 0x7fffffff;
                for (int q = lane; q < G; q += 32) {
                    const size_t o = (size_t)q * Bpad + b;
                    const float m = p.pl_max[o];
                    const int i = p.pl_idx[o];
                    if (m > lmax || (m == lmax && i < lidx)) { lmax = m; lidx = i; }
                    const float dm = p.pd_max[o];
                    const int di = p.pd_idx[o];
                    if (dm > dmax || (dm == dmax && di < didx)) { dmax = dm; didx = di; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float m = __shfl_xor_sync(0xffffffffu, lmax, o);
                    const int i = __shfl_xor_sync(0xffffffffu, lidx, o);
                    if (m > lmax || (m == lmax && i < lidx)) { lmax = m; lidx = i; }
                    const float dm = __shfl_xor_sync(0xffffffffu, dmax, o);
                    const int di = __shfl_xor_sync(0xffffffffu, didx, o);
                    if (dm > dmax || (dm == dmax && di < didx)) { dmax = dm; didx = di; }
                }
                float s = 0.f;
                for (int q = lane; q < G; q += 32) {
                    const size_t o = (size_t)q * Bpad + b;
                    const float m = p.pl_max[o];
                    if (m > -INFINITY) s += p.pl_sum[o] * expf(m - lmax);
                }
                s = warp_sum(s);
                if (lane == 0) {
                    const int T = p.row_off[b + 1] - p.row_off[b];
                    const int skip = (didx < p.n_dur) ? p.durations[didx] : 1;
                    int t = p.tpos[b];
                    bool act = true;
                    if (lidx == V - 1) {            // blank: state reverts (cur unchanged)
                        t += max(skip, 1);
                    } else {
                        const int n = p.ntok[b];
                        if (n < p.cap) {
                            int32_t *row = p.tok + (size_t)b * (1 + p.cap);
                            row[1 + n] = lidx;
                            p.t_start[(size_t)b * p.cap + n] = t;
                            p.t_end[(size_t)b * p.cap + n] = min(t + max(skip, 1) - 1, T - 1);
                            p.t_conf[(size_t)b * p.cap + n] = 1.0f / s;
                            row[0] = n + 1;
                        }
                        p.ntok[b] = n + 1;
                        p.token[b] = lidx;
                        p.cur[b] = 1 - p.cur[b];     // commit the new LSTM state
                        t += skip;
                        if (n + 1 >= p.cap) { act = false; p.overflow[b] = 1; }
                    }
                    p.tpos[b] = t;
                    if (t >= T) act = false;
                    p.active[b] = act ? 1 : 0;
                    if (act) atomicAdd(&p.n_active[step % 3], 1);
                }
            }
        }
        grid.sync();
        if (p.n_active[step % 3] == 0 || step + 1 >= p.max_steps) break;
    }
}

__global__ void tdt_init_kernel(TdtParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < 3) p.n_active[b] = 0;
    if (b >= p.Bpad) return;
    p.cur[b] = 0;
    p.token[b] = p.V - 1;   // SOS = blank (tdt.cpp:56-58)
    p.tpos[b] = 0;
    p.active[b] = b < p.n_utt ? 1 : 0;
    p.ntok[b] = 0;
    p.overflow[b] = 0;
    if (b < p.n_utt) p.tok[(size_t)b * (1 + p.cap)] = 0;
}

}  // namespace

size_t tdt_smem_bytes(const TdtParams &p, int grid, bool *out_in_smem, int *lstm_floats) {
    const int UPC = (p.P + grid - 1) / grid, JPC = (p.J + grid - 1) / grid, OPC = (p.V + p.D + grid - 1) / grid;
    size_t lstm = 0;
    for (int l = 0; l < p.L; ++l) lstm += (size_t)UPC * 4 * p.P * (l > 0 ? 2 : 1);
    size_t base = (size_t)NWARP * RMAX * BCH + RMAX * BCH + lstm + (size_t)JPC * p.P;
    size_t with_out = base + (size_t)OPC * p.J;
    *lstm_floats = (int)lstm;
    if (with_out * sizeof(float) <= 200 * 1024) {
        *out_in_smem = true;
        return with_out * sizeof(float);
    }
    *out_in_smem = false;
    return base * sizeof(float);
}

cudaError_t launch_tdt_decode(TdtParams p, int num_sms, cudaStream_t st) {
    if (p.P % 32 || p.J % 32) return cudaErrorInvalidValue;
    int grid = num_sms;
    // every CTA must own <= 5 LSTM units (RMAX = 20 gate rows)
    if ((p.P + grid - 1) / grid * 4 > RMAX) return cudaErrorInvalidConfiguration;
    bool out_in_smem;
    int lstm_floats;
    size_t smem = tdt_smem_bytes(p, grid, &out_in_smem, &lstm_floats);
    p.out_in_smem = out_in_smem ? 1 : 0;
    p.smem_lstm_floats = lstm_floats;
    cudaError_t err = cudaFuncSetAttribute(tdt_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tdt_decode_kernel, NWARP * 32, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    tdt_init_kernel<<<(p.Bpad + 127) / 128, 128, 0, st>>>(p);
    void *args[] = {&p};
    return cudaLaunchCooperativeKernel((void *)tdt_decode_kernel, dim3(grid), dim3(NWARP * 32), args, smem, st);
}

}  // namespace pk
