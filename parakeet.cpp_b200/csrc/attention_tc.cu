// attention_tc.cu -- K7 on tensor cores (head_dim 64): the same relative-position attention as
// attention.cu (reference src/encoder.cpp:111-178, rel_shift :85-109)
//     S[i,j] = ((q_i + u).k_j + (q_i + v).PP[i-j]) / sqrt(hd),  ctx_i = softmax_j(S[i,:]) V
// with every product on mma.sync.m16n8k16 (bf16 inputs, fp32 accumulate) using the same hi/lo
// operand split as the GEMMs (3 MMAs per product: hi.hi + hi.lo + lo.hi), i.e. ~16 mantissa bits,
// and an fp32 online softmax.  tcgen05 is not used here: per (utterance, head) the matrices are
// 126 x 126 x 64, far below a UMMA tile pipeline's break-even, and the rel_shift needs a per-row
// skew that is natural in registers/shared memory.
//
// One CTA = (64-query tile, head, utterance), 4 warps x 16 query rows.  Per 64-key tile a warp does
//   AC  = Qu . K^T                  16 x 64   (8 n-blocks x 4 k-steps x 3 MMAs)
//   G   = Qv . PPwin^T              16 x 80   window of relative positions i-j (10 x 4 x 3 MMAs)
//   S   = AC + skew(G)              G goes through a per-warp smem patch: S[r][jj] += G[r][r+63-jj]
//   online softmax, P -> bf16 hi/lo A-fragments (C-fragment layout == A-fragment layout)
//   O  += P . V                     16 x 64   (8 x 4 x 3 MMAs, V staged transposed)
#include "kernels.h"

namespace pk {
namespace {

constexpr int HD = 64, BQ = 64, BKV = 64, LDS_ = 72;   // LDS_: smem row stride in bf16 (conflict-free fragment loads)
constexpr int NPW = 128;                               // relative-position window rows per (q-tile, k-tile)
constexpr int LDG_ = 84;                               // G patch row stride (floats)
constexpr int TCA_THREADS = 128;

struct __align__(16) AttnSmem {
    bf16 qu_hi[BQ * LDS_], qu_lo[BQ * LDS_], qv_hi[BQ * LDS_], qv_lo[BQ * LDS_];
    bf16 k_hi[BKV * LDS_], k_lo[BKV * LDS_];
    bf16 vt_hi[HD * LDS_], vt_lo[HD * LDS_];          // [dim][key]
    bf16 pp_hi[NPW * LDS_], pp_lo[NPW * LDS_];
    float g[4][16 * LDG_];
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// A fragment (16 x 16, row-major source [row][LDS_]) of rows row0.. and columns k0..
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16 *base, int row0, int k0, int g, int c) {
    const bf16 *p = base + (row0 + g) * LDS_ + k0 + 2 * c;
    a[0] = *reinterpret_cast<const uint32_t *>(p);
    a[1] = *reinterpret_cast<const uint32_t *>(p + 8 * LDS_);
    a[2] = *reinterpret_cast<const uint32_t *>(p + 8);
    a[3] = *reinterpret_cast<const uint32_t *>(p + 8 * LDS_ + 8);
}
// B fragment (16 x 8, "col" layout) where B[k][n] = src[n0 + n][k0 + k] (src row-major [n][LDS_])
__device__ __forceinline__ void load_b(uint32_t (&b)[2], const bf16 *base, int n0, int k0, int g, int c) {
    const bf16 *p = base + (n0 + g) * LDS_ + k0 + 2 * c;
    b[0] = *reinterpret_cast<const uint32_t *>(p);
    b[1] = *reinterpret_cast<const uint32_t *>(p + 8);
}
__device__ __forceinline__ void split2(float x, float y, uint32_t &hi, uint32_t &lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = *reinterpret_cast<uint32_t *>(&l);
}
__device__ __forceinline__ void split_store(bf16 *hi, bf16 *lo, int idx, float x) {
    bf16 h = __float2bfloat16_rn(x);
    hi[idx] = h;
    lo[idx] = __float2bfloat16_rn(x - __bfloat162float(h));
}

__global__ void __launch_bounds__(TCA_THREADS)
relpos_attention_tc_kernel(const float *__restrict__ qkv, int ld_qkv, const int32_t *__restrict__ row_off,
                           const bf16 *__restrict__ pp_hi, const bf16 *__restrict__ pp_lo, int tmax,
                           const float *__restrict__ bias_u, const float *__restrict__ bias_v, int d_model, ActBuf out) {
    extern __shared__ __align__(16) uint8_t smraw[];
    AttnSmem &sm = *reinterpret_cast<AttnSmem *>(smraw);
    const int b = blockIdx.z, h = blockIdx.y;
    const int r0 = row_off[b], T = row_off[b + 1] - r0;
    const int i0 = blockIdx.x * BQ;
    if (i0 >= T) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;

    // ---- query tile: Qu = q + u_h, Qv = q + v_h, split to bf16 hi/lo
    for (int idx = tid; idx < BQ * HD; idx += TCA_THREADS) {
        const int i = idx / HD, k = idx % HD;
        float q = 0.f;
        if (i0 + i < T) q = qkv[(size_t)(r0 + i0 + i) * ld_qkv + h * HD + k];
        split_store(sm.qu_hi, sm.qu_lo, i * LDS_ + k, q + bias_u[h * HD + k]);
        split_store(sm.qv_hi, sm.qv_lo, i * LDS_ + k, q + bias_v[h * HD + k]);
    }

    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float oacc[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[nb][e] = 0.f;
    const int wrow = warp * 16;          // this warp's first query row inside the tile
    float *gs = sm.g[warp];

    for (int j0 = 0; j0 < T; j0 += BKV) {
        __syncthreads();                 // previous key tile fully consumed (and Q stores visible)
        for (int idx = tid; idx < BKV * HD; idx += TCA_THREADS) {
            const int j = idx / HD, k = idx % HD;
            float kv = 0.f, vv = 0.f;
            if (j0 + j < T) {
                const float *row = qkv + (size_t)(r0 + j0 + j) * ld_qkv + h * HD + k;
                kv = row[d_model];
                vv = row[2 * d_model];
            }
            split_store(sm.k_hi, sm.k_lo, j * LDS_ + k, kv);
            split_store(sm.vt_hi, sm.vt_lo, k * LDS_ + j, vv);
        }
        const int pmin = i0 - (j0 + BKV - 1);
        for (int idx = tid; idx < NPW * (HD / 2); idx += TCA_THREADS) {   // 2 bf16 per thread-iteration
            const int w = idx / (HD / 2), k2 = (idx % (HD / 2)) * 2;
            const int prow = pmin + w + tmax - 1;
            uint32_t vh = 0u, vl = 0u;
            if (prow >= 0 && prow < 2 * tmax - 1) {
                const size_t o = (size_t)prow * d_model + h * HD + k2;
                vh = *reinterpret_cast<const uint32_t *>(pp_hi + o);
                vl = *reinterpret_cast<const uint32_t *>(pp_lo + o);
            }
            *reinterpret_cast<uint32_t *>(sm.pp_hi + w * LDS_ + k2) = vh;
            *reinterpret_cast<uint32_t *>(sm.pp_lo + w * LDS_ + k2) = vl;
        }
        __syncthreads();

        // ---- AC = Qu K^T (16 x 64 per warp)
        float sacc[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
#pragma unroll
            for (int e = 0; e < 4; ++e) sacc[nb][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t ah[4], al[4];
            load_a(ah, sm.qu_hi, wrow, ks * 16, g, c);
            load_a(al, sm.qu_lo, wrow, ks * 16, g, c);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                uint32_t bh[2], bl[2];
                load_b(bh, sm.k_hi, nb * 8, ks * 16, g, c);
                load_b(bl, sm.k_lo, nb * 8, ks * 16, g, c);
                mma_bf16(sacc[nb], ah, bh);
                mma_bf16(sacc[nb], ah, bl);
                mma_bf16(sacc[nb], al, bh);
            }
        }
        // ---- G = Qv PPwin^T (16 x 80: window rows wrow .. wrow+79), through the smem patch, skewed into S
        {
            float gacc[10][4];
#pragma unroll
            for (int nb = 0; nb < 10; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) gacc[nb][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) {
                uint32_t ah[4], al[4];
                load_a(ah, sm.qv_hi, wrow, ks * 16, g, c);
                load_a(al, sm.qv_lo, wrow, ks * 16, g, c);
#pragma unroll
                for (int nb = 0; nb < 10; ++nb) {
                    uint32_t bh[2], bl[2];
                    load_b(bh, sm.pp_hi, wrow + nb * 8, ks * 16, g, c);
                    load_b(bl, sm.pp_lo, wrow + nb * 8, ks * 16, g, c);
                    mma_bf16(gacc[nb], ah, bh);
                    mma_bf16(gacc[nb], ah, bl);
                    mma_bf16(gacc[nb], al, bh);
                }
            }
#pragma unroll
            for (int nb = 0; nb < 10; ++nb) {
                *reinterpret_cast<float2 *>(gs + g * LDG_ + nb * 8 + 2 * c) = make_float2(gacc[nb][0], gacc[nb][1]);
                *reinterpret_cast<float2 *>(gs + (g + 8) * LDG_ + nb * 8 + 2 * c) = make_float2(gacc[nb][2], gacc[nb][3]);
            }
            __syncwarp();
            // S[r][jj] += G[r][r + 63 - jj]   (relative position i - j, window column n = w - wrow)
#pragma unroll
            for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int jj = nb * 8 + 2 * c + e;
                    sacc[nb][e] += gs[g * LDG_ + g + (BKV - 1) - jj];
                    sacc[nb][2 + e] += gs[(g + 8) * LDG_ + g + 8 + (BKV - 1) - jj];
                }
            __syncwarp();
        }
        // ---- scale, mask, online softmax (rows g and g+8 of this warp)
        float alpha[2];
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            float mx = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int jj = nb * 8 + 2 * c + e;
                    float s = sacc[nb][hrow * 2 + e] * 0.125f;
                    s = (j0 + jj < T) ? s : -INFINITY;
                    sacc[nb][hrow * 2 + e] = s;
                    mx = fmaxf(mx, s);
                }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run[hrow], mx);
            alpha[hrow] = (m_run[hrow] == -INFINITY) ? 0.f : expf(m_run[hrow] - m_new);
            float sum = 0.f;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float s = sacc[nb][hrow * 2 + e];
                    const float pexp = (s == -INFINITY) ? 0.f : expf(s - m_new);
                    sacc[nb][hrow * 2 + e] = pexp;
                    sum += pexp;
                }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            l_run[hrow] = l_run[hrow] * alpha[hrow] + sum;
            m_run[hrow] = m_new;
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            oacc[nb][0] *= alpha[0];
            oacc[nb][1] *= alpha[0];
            oacc[nb][2] *= alpha[1];
            oacc[nb][3] *= alpha[1];
        }
        // ---- O += P V : P's C-fragments of n-blocks (2kk, 2kk+1) are the A-fragment of k-step kk
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
            uint32_t ph[4], pl[4];
            split2(sacc[2 * kk][0], sacc[2 * kk][1], ph[0], pl[0]);
            split2(sacc[2 * kk][2], sacc[2 * kk][3], ph[1], pl[1]);
            split2(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1], ph[2], pl[2]);
            split2(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                uint32_t bh[2], bl[2];
                load_b(bh, sm.vt_hi, nb * 8, kk * 16, g, c);
                load_b(bl, sm.vt_lo, nb * 8, kk * 16, g, c);
                mma_bf16(oacc[nb], ph, bh);
                mma_bf16(oacc[nb], ph, bl);
                mma_bf16(oacc[nb], pl, bh);
            }
        }
    }
    // ---- normalise and store ctx[(row), h*HD + dim] (feeds the out_proj GEMM)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
        const int i = i0 + wrow + g + hrow * 8;
        if (i >= T) continue;
        const float inv = 1.0f / l_run[hrow];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const size_t idx = (size_t)(r0 + i) * d_model + h * HD + nb * 8 + 2 * c;
            const float x = oacc[nb][hrow * 2] * inv, y = oacc[nb][hrow * 2 + 1] * inv;
            if (out.f32) *reinterpret_cast<float2 *>(out.f32 + idx) = make_float2(x, y);
            if (out.hi) {
                uint32_t hi, lo;
                split2(x, y, hi, lo);
                *reinterpret_cast<uint32_t *>(out.hi + idx) = hi;
                if (out.lo) *reinterpret_cast<uint32_t *>(out.lo + idx) = lo;
            }
        }
    }
}

}  // namespace

bool launch_relpos_attention_tc(const float *qkv, int ld_qkv, const int32_t *row_off, int n_utt, int max_T, int n_heads,
                                int head_dim, const bf16 *pp_hi, const bf16 *pp_lo, int tmax, const float *bu,
                                const float *bv, int d_model, ActBuf out, cudaStream_t st) {
    if (head_dim != HD) return false;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(relpos_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(AttnSmem)) != cudaSuccess)
            return false;
        attr = true;
    }
    dim3 grid((max_T + BQ - 1) / BQ, n_heads, n_utt);
    relpos_attention_tc_kernel<<<grid, TCA_THREADS, sizeof(AttnSmem), st>>>(qkv, ld_qkv, row_off, pp_hi, pp_lo, tmax, bu,
                                                                          bv, d_model, out);
    return true;
}

}  // namespace pk
