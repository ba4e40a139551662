// attention_umma.cu -- K7 on the 5th-generation tensor cores: the relative-position attention of reference
// src/encoder.cpp:111-178 (rel_shift :85-109) for head_dim 64 and utterances of up to 128 encoder frames (a 10 s clip
// has 126), persistent CTAs that each serve one head and a share of the utterances:
//     S[i,j] = ((q_i + u) . k_j + (q_i + v) . PP[i - j]) / sqrt(hd),   ctx_i = softmax_j(S[i,:]) V
// Every product is a tcgen05.mma (UMMA M = 128, kind::f16, fp32 accumulators in TMEM) on the bf16 hi/lo operand split
// of the GEMMs (hi.hi + hi.lo + lo.hi):
//     AC = Qu . K^T        128 x 128 x 64      K tile by TMA (SWIZZLE_128B) from the k | v planes of the q/k/v GEMM
//     G  = Qv . PPwin^T    128 x 256 x 64      PPwin = the 255 relative positions i - j in [-127, 127] of this tile: ONE
//                                              TMA box of the projected position table, whatever the utterance
//     O  = P . V           128 x 64 x 128      V is [key][dim] in shared memory = an MN-major B operand (b_major bit of
//                                              the instruction descriptor); P is written by the softmax warps as a
//                                              K-major SWIZZLE_128B A operand (two 64-key atoms per plane)
// TMEM: S 128 columns | G 256 | O 64.  rel_shift is a per-row skew, S[i][j] += G[i][i - j + 127], which tcgen05.ld
// cannot express (every lane of a load reads the same columns); a warp therefore loads the 64 G columns that cover its 32
// rows for a 32-key chunk and each lane selects its own 32-wide window with a 5-stage barrel shifter in registers
// (186 selects per chunk) -- no shared-memory patch.  Softmax in base 2 on the whole row (one tile: no online rescaling),
// P unnormalised in [0, 1] split hi/lo, 1 / sum applied to O.
//
// Roles (320 threads): warp 0 TMA (position window once per CTA, K and V per utterance), warp 1 TMEM allocation + MMA issue,
// warps 2..9 the row work: two threads per query row (TMEM lane = row; each takes 64 of the 128 keys and 32 of the 64 output
// columns; row maximum and sum meet through 2 KB of shared memory and a 64-thread named barrier).  They also form
// Qu = q + pos_bias_u and Qv = q + pos_bias_v from the fp32 q of the projection GEMM (coalesced: 8 lanes per row) and write them
// as swizzled operand tiles, and they pass O through a 32 KB staging tile so that the ctx planes are stored in full 128-byte
// lines.  The CTA is persistent: it serves one head (the position window stays resident) and the utterances slot, slot +
// nslots, ...; K and V of the next utterance are fetched as soon as the products that read the current ones have retired, its
// q rows travel under P . V.  Shared memory: Qu/Qv 64 KB (reused for P) | K 32 | PP window 64 | V 32 | O staging 32.
// Longer utterances and other head sizes take the mma.sync kernel (attention_tc.cu).
#include <cuda.h>

#include <cstdio>

#include "kernels.h"
#include "tc_prims.cuh"

namespace pk {
namespace {

using namespace tc;

constexpr int AU_T = 128;                 // queries = keys = one UMMA tile
constexpr int AU_HD = 64;
constexpr int AU_THREADS = 320;
constexpr int AU_ROWTHR = 256;            // row threads (warps 2..9)
constexpr int AU_TILE = AU_T * AU_HD * 2; // one bf16 plane of a 128 x 64 tile: 16 KB
// Q (later P) 64 KB | K 32 KB | PP window 64 KB (resident: it depends on the head only) | V 32 KB | O staging 32 KB | row max / sum 2 KB
constexpr int AU_RED = 2 * 2 * AU_T * 4;
constexpr int AU_LAYOUT = 4 * AU_TILE + 2 * AU_TILE + 4 * AU_TILE + 2 * AU_TILE + 2 * AU_TILE + AU_RED + 128;
constexpr int AU_SMEM = 227 * 1024;       // (the slack after AU_LAYOUT absorbs the alignment of the base to 1024 B; checked in the kernel)
constexpr uint32_t AU_COL_S = 0, AU_COL_G = 128, AU_COL_O = 384, AU_TMEM_COLS = 512;
static_assert(AU_LAYOUT + 512 <= AU_SMEM, "shared memory budget");

__device__ __forceinline__ void tmem_ld64_issue(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// kind::f16 instruction descriptor with an MN-major B operand (bit 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int m, int n) { return umma_idesc_bf16(m, n) | (1u << 16); }

// measurement aid (dbg != 0): clock64 of CTA 0 per item -- row thread 64: [0] item start, [1] Q tiles written, [2] S ready, [3] P written,
// [4] O ready, [5] O stored; MMA thread: [6] S products issued, [7] P.V issued
__device__ long long g_au_tl[8][8];

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Persistent: CTA c serves head c % H and the utterances slot, slot + nslots, ... (slot = c / H, nslots = gridDim.x / H).
__global__ void __launch_bounds__(AU_THREADS, 1)
relpos_attention_umma_kernel(const __grid_constant__ CUtensorMap tmKV_hi, const __grid_constant__ CUtensorMap tmKV_lo,
                             const __grid_constant__ CUtensorMap tmPP_hi, const __grid_constant__ CUtensorMap tmPP_lo,
                             const float *__restrict__ q32, const float *__restrict__ pos_u, const float *__restrict__ pos_v,
                             const int32_t *__restrict__ row_off, int n_utt, int n_heads, int tmax, int d_model, ActBuf out, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    if ((base - smem_raw) + AU_LAYOUT > AU_SMEM) __trap();      // (the dynamic window starts 1024-aligned in practice)
    uint8_t *sQ = base;                          // Qu_hi | Qu_lo | Qv_hi | Qv_lo ; then P_hi (2 atoms) | P_lo (2 atoms)
    uint8_t *sK = sQ + 4 * AU_TILE;              // K_hi | K_lo
    uint8_t *sPP = sK + 2 * AU_TILE;             // PP_hi (256 rows) | PP_lo
    uint8_t *sV = sPP + 4 * AU_TILE;             // V_hi | V_lo   ([key][dim])
    uint8_t *sO = sV + 2 * AU_TILE;              // O_hi | O_lo staging (128-byte rows, swizzled)
    float *red = reinterpret_cast<float *>(sO + 2 * AU_TILE);   // [max | sum][half][row]
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(red) + AU_RED);
    uint64_t *bar_pp = bars, *k_full = bars + 1, *k_empty = bars + 2, *v_full = bars + 3, *v_empty = bars + 4, *q_full = bars + 5, *s_full = bars + 6,
             *p_full = bars + 7, *o_full = bars + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x % n_heads, slot = blockIdx.x / n_heads, nslots = gridDim.x / n_heads;
    if (threadIdx.x == 0) {
        mbar_init(bar_pp, 1);
        mbar_init(k_full, 1);
        mbar_init(k_empty, 1);
        mbar_init(v_full, 1);
        mbar_init(v_empty, 1);
        mbar_init(q_full, AU_ROWTHR);
        mbar_init(s_full, 1);
        mbar_init(p_full, AU_ROWTHR);
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(AU_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();
    // Every role walks the same item list: utterances slot, slot + nslots, ... that have frames.

    if (warp == 0) {
        // ===================== TMA: the position window once, then K and V of every item =====================
        if (elect_one()) {
            bool first = true;
            uint32_t n = 0;
            for (int b = slot; b < n_utt; b += nslots) {
                const int r0 = row_off[b];
                if (row_off[b + 1] - r0 <= 0) continue;
                if (first) {
                    // window row w = i - j + 127 is table row (i - j) + tmax - 1: rows tmax - 128 .. tmax + 127 (outside the table: zeros)
                    mbar_expect_tx(bar_pp, 4 * AU_TILE);
                    tma_load_2d(sPP, &tmPP_hi, bar_pp, h * AU_HD, tmax - AU_T);
                    tma_load_2d(sPP + 2 * AU_TILE, &tmPP_lo, bar_pp, h * AU_HD, tmax - AU_T);
                    first = false;
                }
                mbar_wait(k_empty, (n & 1u) ^ 1u);                      // the S products of item n - 1 have read K
                mbar_expect_tx(k_full, 2 * AU_TILE);
                tma_load_2d(sK, &tmKV_hi, k_full, h * AU_HD, r0);
                tma_load_2d(sK + AU_TILE, &tmKV_lo, k_full, h * AU_HD, r0);
                mbar_wait(v_empty, (n & 1u) ^ 1u);                      // P . V of item n - 1 has read V
                mbar_expect_tx(v_full, 2 * AU_TILE);
                tma_load_2d(sV, &tmKV_hi, v_full, d_model + h * AU_HD, r0);
                tma_load_2d(sV + AU_TILE, &tmKV_lo, v_full, d_model + h * AU_HD, r0);
                ++n;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issue =====================
        if (elect_one()) {
            const uint32_t q_s = smem_u32(sQ), k_s = smem_u32(sK), pp_s = smem_u32(sPP), v_s = smem_u32(sV);
            constexpr uint32_t id_s = umma_idesc_bf16(AU_T, 128), id_g = umma_idesc_bf16(AU_T, 256), id_o = umma_idesc_bf16_bmn(AU_T, AU_HD);
            uint32_t n = 0;
            for (int b = slot; b < n_utt; b += nslots) {
                if (row_off[b + 1] - row_off[b] <= 0) continue;
                const uint32_t par = n & 1u;
                if (n == 0) mbar_wait(bar_pp, 0);
                mbar_wait(q_full, par);
                mbar_wait(k_full, par);
                tcgen05_fence_after();
#pragma unroll
                for (int ks = 0; ks < AU_HD / UMMA_K; ++ks) {          // AC = Qu . K^T
                    const uint32_t ko = (uint32_t)ks * 32u;             // 16 bf16 = 32 B along K inside the 128-byte swizzle row
                    const uint64_t qh = umma_desc_sw128(q_s + ko), ql = umma_desc_sw128(q_s + AU_TILE + ko);
                    const uint64_t kh = umma_desc_sw128(k_s + ko), kl = umma_desc_sw128(k_s + AU_TILE + ko);
                    umma_bf16(tmem_base + AU_COL_S, qh, kh, id_s, ks != 0);
                    umma_bf16(tmem_base + AU_COL_S, qh, kl, id_s, 1);
                    umma_bf16(tmem_base + AU_COL_S, ql, kh, id_s, 1);
                }
#pragma unroll
                for (int ks = 0; ks < AU_HD / UMMA_K; ++ks) {          // G = Qv . PPwin^T
                    const uint32_t ko = (uint32_t)ks * 32u;
                    const uint64_t qh = umma_desc_sw128(q_s + 2 * AU_TILE + ko), ql = umma_desc_sw128(q_s + 3 * AU_TILE + ko);
                    const uint64_t ph = umma_desc_sw128(pp_s + ko), pl = umma_desc_sw128(pp_s + 2 * AU_TILE + ko);
                    umma_bf16(tmem_base + AU_COL_G, qh, ph, id_g, ks != 0);
                    umma_bf16(tmem_base + AU_COL_G, qh, pl, id_g, 1);
                    umma_bf16(tmem_base + AU_COL_G, ql, ph, id_g, 1);
                }
                umma_commit(s_full);                                    // S and G complete ...
                umma_commit(k_empty);                                   // ... and K (and the Q tiles) have been read
                if (dbg && blockIdx.x == 0 && n < 8) g_au_tl[n][6] = clock64();
                mbar_wait(p_full, par);
                mbar_wait(v_full, par);
                tcgen05_fence_after();
#pragma unroll
                for (int ks = 0; ks < AU_T / UMMA_K; ++ks) {           // O = P . V  (16 keys per instruction)
                    // A: P atom ks / 4 (64 keys), 32 B per k-step inside its 128-byte row.  B: V rows 16 ks .. 16 ks + 15, MN-major
                    const uint32_t pa = q_s + (uint32_t)(ks >> 2) * AU_TILE + (uint32_t)(ks & 3) * 32u;
                    const uint64_t ph = umma_desc_sw128(pa), pl = umma_desc_sw128(pa + 2 * AU_TILE);
                    const uint64_t vh = umma_desc_sw128(v_s + (uint32_t)ks * 2048u), vl = umma_desc_sw128(v_s + AU_TILE + (uint32_t)ks * 2048u);
                    umma_bf16(tmem_base + AU_COL_O, ph, vh, id_o, ks != 0);
                    umma_bf16(tmem_base + AU_COL_O, ph, vl, id_o, 1);
                    umma_bf16(tmem_base + AU_COL_O, pl, vh, id_o, 1);
                }
                umma_commit(o_full);
                umma_commit(v_empty);
                if (dbg && blockIdx.x == 0 && n < 8) g_au_tl[n][7] = clock64();
                ++n;
            }
        }
    } else {
        // ===================== the row work: thread = (query row, key / output-column half) =====================
        const int qd = warp & 3;                     // TMEM lane quarter of this warp
        const int hf = (warp - 2) >> 2;              // keys 64 hf .. 64 hf + 63, output columns 32 hf .. 32 hf + 31
        const int i = qd * 32 + lane;                // query row
        const int rt = threadIdx.x - 64;             // 0 .. 255: cooperative (coalesced) passes
        const int cch = rt & 7, crow = rt >> 3;      // those passes: 16-byte chunk of a 128-byte row, row crow + 32 k
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
        const uint32_t sw = (uint32_t)(i & 7);
        const uint32_t row_s = smem_u32(sQ) + (uint32_t)i * 128u;
        const uint32_t q_s = smem_u32(sQ), o_s = smem_u32(sO);
        float *red_max = red, *red_sum = red + 2 * AU_T;
        const bool tl_on = dbg && blockIdx.x == 0 && threadIdx.x == 64;
        // pos_bias_u / pos_bias_v of this thread's 8 dims (chunk cch of every row it prepares)
        const float4 u0 = __ldg(reinterpret_cast<const float4 *>(pos_u + h * AU_HD + 8 * cch)), u1 = __ldg(reinterpret_cast<const float4 *>(pos_u + h * AU_HD + 8 * cch + 4));
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(pos_v + h * AU_HD + 8 * cch)), v1 = __ldg(reinterpret_cast<const float4 *>(pos_v + h * AU_HD + 8 * cch + 4));
        float4 qa[4], qb[4];                         // q (fp32): dims 8 cch .. + 7 of rows crow + 32 k, fetched one item ahead (8 lanes = one 256-byte row)
        auto q_fetch = [&](int bb) {
            const int r0 = row_off[bb], T = row_off[bb + 1] - r0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int row = crow + 32 * k;
                qa[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                qb[k] = qa[k];
                if (row < T) {
                    const float *qp = q32 + (size_t)(r0 + row) * d_model + h * AU_HD + 8 * cch;
                    qa[k] = *reinterpret_cast<const float4 *>(qp);
                    qb[k] = *reinterpret_cast<const float4 *>(qp + 4);
                }
            }
        };
        // Qu = q + pos_bias_u, Qv = q + pos_bias_v as swizzled K-major tiles (rows past T: q = 0 was fetched -> finite scores, never stored)
        auto q_write = [&]() {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int row = crow + 32 * k;
                const float4 a = qa[k], bq = qb[k];
                uint32_t uh[4], ul[4], vh[4], vl[4];
                split_pair(a.x + u0.x, a.y + u0.y, uh[0], ul[0]);
                split_pair(a.z + u0.z, a.w + u0.w, uh[1], ul[1]);
                split_pair(bq.x + u1.x, bq.y + u1.y, uh[2], ul[2]);
                split_pair(bq.z + u1.z, bq.w + u1.w, uh[3], ul[3]);
                split_pair(a.x + v0.x, a.y + v0.y, vh[0], vl[0]);
                split_pair(a.z + v0.z, a.w + v0.w, vh[1], vl[1]);
                split_pair(bq.x + v1.x, bq.y + v1.y, vh[2], vl[2]);
                split_pair(bq.z + v1.z, bq.w + v1.w, vh[3], vl[3]);
                const uint32_t a0 = q_s + (uint32_t)row * 128u + (((uint32_t)cch ^ (uint32_t)(row & 7)) << 4);
                sts128(a0, uh[0], uh[1], uh[2], uh[3]);
                sts128(a0 + AU_TILE, ul[0], ul[1], ul[2], ul[3]);
                sts128(a0 + 2 * AU_TILE, vh[0], vh[1], vh[2], vh[3]);
                sts128(a0 + 3 * AU_TILE, vl[0], vl[1], vl[2], vl[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the UMMA reads
            mbar_arrive(q_full);
        };
        auto next_item = [&](int bb) {               // next utterance of this CTA with frames, or n_utt
            for (bb += nslots; bb < n_utt; bb += nslots)
                if (row_off[bb + 1] - row_off[bb] > 0) return bb;
            return n_utt;
        };
        int b = next_item(slot - nslots);
        if (b < n_utt) {
            q_fetch(b);
            q_write();
        }
        uint32_t n = 0;
        for (; b < n_utt; ++n) {
            const int r0 = row_off[b], T = row_off[b + 1] - r0;
            const uint32_t par = n & 1u;
            const int bn = next_item(b);
            long long *tl = (tl_on && n < 8) ? g_au_tl[n] : nullptr;
            if (tl) tl[0] = tl[1] = clock64();
            // ---- S = (AC + skew(G)) * scale for this thread's 64 keys
            mbar_wait(s_full, par);
            tcgen05_fence_after();
            if (tl) tl[2] = clock64();
            constexpr float kScale = 0.125f * 1.4426950408889634f;     // 1 / sqrt(64), folded with log2 e
            uint32_t su[2][32];                          // scores, then probabilities (fp32 bit patterns)
            float mx = -INFINITY;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int c = 2 * hf + cc;               // 32-key chunk
                uint32_t R[64];
                const int gbase = 32 * (qd - c) + 96;    // window columns gbase + lane + m, m = 31 - jj
                tmem_ld64_issue(lane_addr + AU_COL_G + (uint32_t)gbase, R);
                tmem_ld32_issue(lane_addr + AU_COL_S + (uint32_t)(32 * c), su[cc]);
                tmem_wait_ld();
                // barrel shifter: R[m] <- R[lane + m]
#pragma unroll
                for (int k = 0; k < 47; ++k) R[k] = (lane & 16) ? R[k + 16] : R[k];
#pragma unroll
                for (int k = 0; k < 39; ++k) R[k] = (lane & 8) ? R[k + 8] : R[k];
#pragma unroll
                for (int k = 0; k < 35; ++k) R[k] = (lane & 4) ? R[k + 4] : R[k];
#pragma unroll
                for (int k = 0; k < 33; ++k) R[k] = (lane & 2) ? R[k + 2] : R[k];
#pragma unroll
                for (int k = 0; k < 32; ++k) R[k] = (lane & 1) ? R[k + 1] : R[k];
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    float v = (__uint_as_float(su[cc][jj]) + __uint_as_float(R[31 - jj])) * kScale;
                    v = (32 * c + jj < T) ? v : -INFINITY;
                    su[cc][jj] = __float_as_uint(v);
                    mx = fmaxf(mx, v);
                }
            }
            // row maximum over both halves (the other half of this row lives in the warp with the same lane quarter)
            red_max[hf * AU_T + i] = mx;
            named_bar_sync(1 + qd, 64);
            mx = fmaxf(mx, red_max[(hf ^ 1) * AU_T + i]);   // key 0 is always valid: finite
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const float p = ex2_approx(__uint_as_float(su[cc][jj]) - mx);
                    su[cc][jj] = __float_as_uint(p);
                    sum += p;
                }
            red_sum[hf * AU_T + i] = sum;
            // ---- P (unnormalised) as the A operand of P . V: plane, 64-key atom hf, row i, 8 keys per 16-byte chunk
            // (the S products have retired: the Q tiles are dead)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8) {
                    uint32_t ph[4], pl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        split_pair(__uint_as_float(su[cc][8 * k8 + 2 * e]), __uint_as_float(su[cc][8 * k8 + 2 * e + 1]), ph[e], pl[e]);
                    const uint32_t off = (uint32_t)hf * AU_TILE + ((((uint32_t)(4 * cc + k8)) ^ sw) << 4);
                    sts128(row_s + off, ph[0], ph[1], ph[2], ph[3]);
                    sts128(row_s + 2 * AU_TILE + off, pl[0], pl[1], pl[2], pl[3]);
                }
            tcgen05_fence_before();                      // (this thread's TMEM reads are complete)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(p_full);
            if (tl) tl[3] = clock64();
            if (bn < n_utt) q_fetch(bn);                 // the next item's q rows travel under P . V
            // ---- O: out of TMEM, then the next item's Q tiles (P . V has read P), then the stores
            mbar_wait(o_full, par);
            tcgen05_fence_after();
            if (tl) tl[4] = clock64();
            uint32_t o[32];
            tmem_ld32_issue(lane_addr + AU_COL_O + (uint32_t)(32 * hf), o);
            tmem_wait_ld();
            tcgen05_fence_before();
            if (bn < n_utt) q_write();
            named_bar_sync(1 + qd, 64);                  // the other half's sum is in red_sum (and it has read this half's maximum)
            {
                const float inv = 1.0f / (sum + red_sum[(hf ^ 1) * AU_T + i]);
                // this thread's 32 columns -> bf16 hi / lo, 4 chunks of the row's 128-byte staging line per plane
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {
                    uint4 hh, ll;
                    split_pair(__uint_as_float(o[8 * cq]) * inv, __uint_as_float(o[8 * cq + 1]) * inv, hh.x, ll.x);
                    split_pair(__uint_as_float(o[8 * cq + 2]) * inv, __uint_as_float(o[8 * cq + 3]) * inv, hh.y, ll.y);
                    split_pair(__uint_as_float(o[8 * cq + 4]) * inv, __uint_as_float(o[8 * cq + 5]) * inv, hh.z, ll.z);
                    split_pair(__uint_as_float(o[8 * cq + 6]) * inv, __uint_as_float(o[8 * cq + 7]) * inv, hh.w, ll.w);
                    const uint32_t a0 = o_s + (uint32_t)i * 128u + ((((uint32_t)(4 * hf + cq)) ^ sw) << 4);
                    sts128(a0, hh.x, hh.y, hh.z, hh.w);
                    sts128(a0 + AU_TILE, ll.x, ll.y, ll.z, ll.w);
                }
            }
            named_bar_sync(5, AU_ROWTHR);                // the whole O tile is staged
#pragma unroll
            for (int k = 0; k < 4; ++k) {                // 8 lanes = one 128-byte line of a ctx plane
                const int row = crow + 32 * k;
                if (row < T) {
                    const uint32_t a0 = o_s + (uint32_t)row * 128u + (((uint32_t)cch ^ (uint32_t)(row & 7)) << 4);
                    const size_t idx = (size_t)(r0 + row) * d_model + h * AU_HD + 8 * cch;
                    *reinterpret_cast<uint4 *>(out.hi + idx) = lds128u(a0);
                    if (out.lo) *reinterpret_cast<uint4 *>(out.lo + idx) = lds128u(a0 + AU_TILE);
                }
            }
            if (tl) tl[5] = clock64();
            b = bn;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(AU_TMEM_COLS) : "memory");
    }
}

int g_au_dbg = 0;

}  // namespace

bool relpos_attention_umma_supported(int head_dim, int max_T) { return head_dim == AU_HD && max_T <= AU_T; }

void relpos_attention_umma_set_debug(int on) { g_au_dbg = on; }
void relpos_attention_umma_print_timeline(int n_items) {
    long long h[8][8];
    if (cudaMemcpyFromSymbol(h, g_au_tl, sizeof(h)) != cudaSuccess) return;
    for (int i = 0; i < n_items && i < 8; ++i)
        fprintf(stderr, "    item %d (CTA 0, cycles since the item's start): S ready %lld (S issued %lld)  P written %lld  O ready %lld (P.V issued %lld)  O stored %lld   next item +%lld\n", i,
                h[i][2] - h[i][0], h[i][6] - h[i][0], h[i][3] - h[i][0], h[i][4] - h[i][0], h[i][7] - h[i][0], h[i][5] - h[i][0],
                i + 1 < n_items ? h[i + 1][0] - h[i][5] : 0LL);
}

// kv: tensor maps of the [M][2 d] k | v planes (box 64 x 128); pp: of the [2 tmax - 1][d] projected position table (box 64 x 256)
bool launch_relpos_attention_umma(const float *q32, const float *pos_u, const float *pos_v, const TcOperand &kv, const TcOperand &pp,
                                  const int32_t *row_off, int n_utt, int max_T, int n_heads, int head_dim, int tmax, int d_model, int num_sms, ActBuf out,
                                  cudaStream_t st) {
    if (!relpos_attention_umma_supported(head_dim, max_T) || !q32 || !pos_u || !pos_v || !kv.has_lo || !pp.has_lo || kv.box_rows != 128 || pp.box_rows != 256 ||
        !out.hi || n_heads * head_dim != d_model || n_utt < 1 || n_heads > num_sms)
        return false;
    static PerDeviceFlag attr_flag;
    if (!attr_flag.cur()) {
        if (cudaFuncSetAttribute(relpos_attention_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AU_SMEM) != cudaSuccess) return false;
        attr_flag.cur() = true;
    }
    int nslots = num_sms / n_heads;              // CTAs per head: one CTA per SM, every CTA keeps its head's position window
    if (nslots > n_utt) nslots = n_utt;
    return launch_pdl(relpos_attention_umma_kernel, dim3((unsigned)(n_heads * nslots)), dim3(AU_THREADS), (size_t)AU_SMEM, st, kv.hi, kv.lo, pp.hi, pp.lo, q32,
                      pos_u, pos_v, row_off, n_utt, n_heads, tmax, d_model, out, g_au_dbg) == cudaSuccess;
}

}  // namespace pk
