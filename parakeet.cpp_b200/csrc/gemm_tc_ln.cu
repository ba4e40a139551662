// gemm_tc_ln.cu -- K5b: the residual GEMM of a Conformer sub-block with the LayerNorm that FOLLOWS it fused into
// the epilogue (and, at the end of a block, the chained pair final_norm_ -> next block's ffn1_.norm_):
//     v      = resid + alpha * (A[M,K] . W[N,K]^T + bias)            (FeedForward / attention out / conv pw2,
//                                                                      reference src/encoder.cpp:39-46, :182-185, :72-74)
//     y1     = LayerNorm_1(v)                                         (nn::LayerNorm, axiom operations.cpp:1796-1809:
//     y2     = LayerNorm_2(y1)            [optional]                   biased variance, eps inside the root)
//     out_f32 = v  or  y1 ;   planes = bf16 hi/lo split of the last LayerNorm's result (A operand of the next GEMM)
// It replaces the pair  gemm_tc_kernel<EPI_RESID_F32>  +  layernorm_kernel  (69 stand-alone LayerNorm launches per
// 110m step, each re-reading the fp32 residual stream from L2 and writing the operand planes).  MEASURED (DESIGN.md
// section 4): equal to the pair in isolation, not faster inside the step graph -> opt-in (PK_FUSE_LN=1).
//
// A LayerNorm row spans all N = d_model columns, i.e. CLN = N / 128 accumulator tiles that live in the TMEM of CLN
// different SMs.  The kernel therefore runs as thread-block CLUSTERS of CLN CTAs along N: cluster c walks the row
// blocks c, c + #clusters, ...; CTA rank r of the cluster owns the columns [128 r, 128 r + 128) of every block -- same
// TMA -> smem -> tcgen05.mma -> TMEM pipeline and warp roles as gemm_tc_kernel (warp 0 producer, warp 1 MMA issuer,
// warps 2..9 epilogue, double-buffered accumulator).  In the epilogue a warp holds a 32-row x 64-column slab of v in
// registers (one row per lane), reduces it to the pair (sum, centred sum of squares about its OWN mean), and writes that
// pair into the statistics table of EVERY CTA of the cluster through distributed shared memory with st.async: the write
// completes 8 transaction bytes on the DESTINATION's mbarrier (armed once per round by one thread of that CTA), so no
// fence and no arrive are needed.  After an acquire-wait on the local mbarrier a lane combines the 2 CLN partials of its
// row in a fixed order with Chan's parallel formula
//     mean = (sum_i s_i) / N,   M2 = sum_i M2_i + 64 sum_i (s_i / 64 - mean)^2,   var = M2 / N
// -- one exchange per LayerNorm, numerically equivalent to the reference's two passes (no E[x^2] - mean^2
// cancellation) and deterministic.  Two table slots / two mbarriers alternate, which is enough: a warp can only be one
// exchange ahead of the slowest warp of its cluster (it needs everybody's data to pass the exchange in between).
// Global traffic stays coalesced: the residual (fetched BEFORE the accumulator is waited for) and all stores go through a
// 2 KB per-warp staging tile (16 rows x 128 B, SWIZZLE_128B pattern) that transposes between "lane = row" and "8 lanes =
// one full 128-byte line"; bias / LayerNorm weights of the CTA's 128 columns sit in shared memory.
#include <cuda.h>

#include <cstdio>

#include "kernels.h"
#include "tc_prims.cuh"

namespace pk {
namespace {

using namespace tc;

constexpr int LN_BN = 128;
constexpr int LN_EPI_WARPS = 8;
constexpr int LN_THREADS = 64 + LN_EPI_WARPS * 32;
constexpr int LN_STGW = 2048;                     // per-warp staging tile: 32 rows x 64 B

template <int NPASS, int CLN>
struct LnCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int W_BYTES = LN_BN * BK * 2;
    static constexpr int PLANES = (NPASS == 3) ? 2 : 1;
    static constexpr int STAGE_BYTES = PLANES * (A_BYTES + W_BYTES);
    static constexpr int STG_BYTES = LN_EPI_WARPS * LN_STGW;
    static constexpr int NSRC = 2 * CLN;                               // partial statistics per row (two 64-column slabs per CTA)
    static constexpr int STATS_BYTES = 2 * BM * NSRC * 8;              // [slot][row][source] float2
    static constexpr int PAR_BYTES = 3 * LN_BN * 4;                    // bias | ln1 weight | ln1 bias of this CTA's 128 columns
    static constexpr int AVAIL = 227 * 1024 - STG_BYTES - STATS_BYTES - PAR_BYTES - 1024 - 256;
    static constexpr int STAGES = AVAIL / STAGE_BYTES > 8 ? 8 : AVAIL / STAGE_BYTES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STG_BYTES + STATS_BYTES + PAR_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static constexpr int TMEM_COLS = 2 * LN_BN;
    static_assert(STAGES >= 2, "pipeline depth");
};

struct LnEpiDev {                  // LnEpi as the kernel sees it
    const float *bias, *resid;
    float *out_f32;
    const float *w1, *b1, *w2, *b2;
    bf16 *hi, *lo;
    float alpha, eps;
    int out_ln1;
};

// 8 bytes into the shared memory of a CTA of the cluster; completion is signalled as 8 transaction bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_b64(uint32_t cluster_addr, uint64_t v, uint32_t cluster_mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(cluster_addr), "l"(v), "r"(cluster_mbar) : "memory");
}
// measurement aid (dbg != 0): clock64 of CTA 0's first epilogue warp per tile -- [0] tile start (residual loads issued), [1] accumulator
// ready, [2] TMEM drained + bias, [3] residual added, [4] statistics sent, [5] x stored, [6] statistics received, [7] normalised (+ second
// LayerNorm), [8] planes stored
__device__ long long g_ln_tl[8][16];
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {   // acquire at cluster scope: remote st.shared::cluster before the arrivals are visible
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}

// ---- the 2 KB staging tile of a warp: 16 rows x 128 B, 16-byte chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4) (the
// SWIZZLE_128B pattern: conflict-free for "lane = row" accesses and for the transposed ones, where 8 lanes cover one row).
// A warp's 32-row slab passes through it in two halves (h = 0, 1: rows 16 h .. 16 h + 15, held by lanes 16 h .. 16 h + 15), so
// that every global access of the epilogue is a FULL 128-byte line per 8 lanes (measured: 64-byte pieces made the x / plane
// stores of a tile cost 7.4k of the epilogue's 13k cycles).
__device__ __forceinline__ uint32_t stg_row(uint32_t stg, int lane, int c) { return stg + (uint32_t)((lane & 15) * 128 + ((c ^ (lane & 7)) << 4)); }
__device__ __forceinline__ uint32_t stg_tr(uint32_t stg, int lane, int it, int &row) {
    row = it * 4 + (lane >> 3);
    return stg + (uint32_t)(row * 128 + (((lane & 7) ^ (row & 7)) << 4));
}
// lanes of half h hold w[32] = 128 bytes of their row; the warp writes rows row0 + 16 h .. + 15 at base + row * ld_bytes
__device__ __forceinline__ void store_half128(uint32_t stg, int lane, int h, const uint32_t (&w)[32], uint8_t *base, size_t ld_bytes, int row0, int M) {
    if ((lane >> 4) == h) {
#pragma unroll
        for (int c = 0; c < 8; ++c) sts128(stg_row(stg, lane, c), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        int row;
        const uint4 v = lds128u(stg_tr(stg, lane, it, row));
        const int grow = row0 + 16 * h + row;
        if (grow < M) *reinterpret_cast<uint4 *>(base + (size_t)grow * ld_bytes + ((lane & 7) << 4)) = v;
    }
    __syncwarp();
}

template <int NPASS, int CLN, bool MC>
__global__ void __cluster_dims__(CLN, 1, 1) __launch_bounds__(LN_THREADS, 1)
gemm_tc_ln_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, int M, int K,
                  const __grid_constant__ LnEpiDev epi, int dbg) {
    using C = LnCfg<NPASS, CLN>;
    constexpr int N = CLN * LN_BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *staging = tiles + (size_t)C::STAGES * C::STAGE_BYTES;
    uint8_t *stats = staging + C::STG_BYTES;
    float *par = reinterpret_cast<float *>(stats + C::STATS_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(stats + C::STATS_BYTES + C::PAR_BYTES);
    uint64_t *full = bars, *empty = bars + C::STAGES, *acc_full = bars + 2 * C::STAGES, *acc_empty = acc_full + 2, *ln_bar = acc_empty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ln_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / CLN, ncl = gridDim.x / CLN;
    const int nkb = K / BK;
    const int tiles_m = (M + BM - 1) / BM;
    const int n0 = rank * LN_BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MC ? CLN : 1);             // MC: the stage is written by every CTA of the cluster, so all of them must have consumed it
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], LN_EPI_WARPS);
            mbar_init(&ln_bar[b], 1);                       // armed once per exchange round; the data arrives as transaction bytes
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Per-column epilogue operands of this CTA's 128 columns -> shared memory (weights: not written by the preceding grid).
    // Measured: as warp-uniform __ldg they queued behind the epilogue's own global stores (normalise: 2.0-2.6k cycles).
    if (threadIdx.x >= 64 && threadIdx.x < 64 + LN_BN) {
        const int i = threadIdx.x - 64;
        par[i] = epi.bias[n0 + i];
        par[LN_BN + i] = epi.w1[n0 + i];
        par[2 * LN_BN + i] = epi.b1[n0 + i];
    }
    tcgen05_fence_before();
    cluster_sync();                               // every CTA's barriers exist before any remote arrival (and `par` is visible)
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            uint32_t it = 0;
            for (int rb = cid; rb < tiles_m; rb += ncl) {
                const int m0 = rb * BM;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t *st = tiles + (size_t)s * C::STAGE_BYTES;
                    mbar_expect_tx(&full[s], C::STAGE_BYTES);
                    if (MC) {
                        // The A tile is the same for all CLN CTAs: each loads a quarter of its rows (tmA_* then have a 32-row
                        // box) and MULTICASTS it into the stage of every CTA -- 1/CLN of the L2 -> SM operand traffic for A.
                        constexpr int SL = BM / CLN, SLB = SL * BK * 2;
                        constexpr uint16_t mask = (uint16_t)((1u << CLN) - 1u);
                        tma_load_2d_mcast(st + rank * SLB, &tmA_hi, &full[s], kb * BK, m0 + rank * SL, mask);
                        if (NPASS == 3) tma_load_2d_mcast(st + C::A_BYTES + C::W_BYTES + rank * SLB, &tmA_lo, &full[s], kb * BK, m0 + rank * SL, mask);
                    } else {
                        tma_load_2d(st, &tmA_hi, &full[s], kb * BK, m0);
                        if (NPASS == 3) tma_load_2d(st + C::A_BYTES + C::W_BYTES, &tmA_lo, &full[s], kb * BK, m0);
                    }
                    tma_load_2d(st + C::A_BYTES, &tmW_hi, &full[s], kb * BK, n0);
                    if (NPASS == 3) tma_load_2d(st + 2 * C::A_BYTES + C::W_BYTES, &tmW_lo, &full[s], kb * BK, n0);
                }
            }
            pdl_trigger_late();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, LN_BN);
            uint32_t it = 0, tcount = 0;
            for (int rb = cid; rb < tiles_m; rb += ncl, ++tcount) {
                const uint32_t buf = tcount & 1, aph = (tcount >> 1) & 1;
                mbar_wait(&acc_empty[buf], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + buf * LN_BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tcgen05_fence_after();
                    const uint32_t st = smem_u32(tiles + (size_t)s * C::STAGE_BYTES);
                    const uint64_t a_hi = umma_desc_sw128(st), w_hi = umma_desc_sw128(st + C::A_BYTES);
                    const uint64_t a_lo = umma_desc_sw128(st + C::A_BYTES + C::W_BYTES);
                    const uint64_t w_lo = umma_desc_sw128(st + 2 * C::A_BYTES + C::W_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
                        umma_bf16(tmem_d, a_hi + koff, w_hi + koff, idesc, (kb | k) != 0);
                        if (NPASS == 3) {
                            umma_bf16(tmem_d, a_hi + koff, w_lo + koff, idesc, 1);
                            umma_bf16(tmem_d, a_lo + koff, w_hi + koff, idesc, 1);
                        }
                    }
                    if (MC) umma_commit_mcast(&empty[s], (uint16_t)((1u << CLN) - 1u));
                    else umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int ew = warp - 2;
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int half = ew >> 2;                    // which 64 of this CTA's 128 columns
        const uint32_t stg_s = smem_u32(staging) + (uint32_t)ew * LN_STGW;
        const uint32_t stats_s = smem_u32(stats);
        const int gcol0 = n0 + half * 64;            // first global column of this warp's slab
        const int trow = q * 32 + lane;              // this lane's row inside the 128-row block
        const int my_src = rank * 2 + half;
        uint32_t xr = 0;                             // statistics exchanges done so far (same count in every warp of the cluster)
        uint32_t peer_stats[CLN], peer_bar[CLN];     // this CTA's table / ln_bar[0] as seen in each CTA of the cluster (ln_bar[1] = + 8 bytes)
#pragma unroll
        for (int d = 0; d < CLN; ++d) {
            peer_stats[d] = mapa_rank(stats_s, (uint32_t)d);
            peer_bar[d] = mapa_rank(smem_u32(&ln_bar[0]), (uint32_t)d);
        }
        const bool tl_on = dbg && blockIdx.x == 0 && threadIdx.x == 64;
        // Statistics exchange, split so that independent stores can sit between the two halves.  send: the partial pair of this
        // lane's row goes to every CTA of the cluster by st.async -- the write completes 8 bytes of the transaction count of the
        // DESTINATION's mbarrier, so no fence / arrive is needed; one thread per CTA arms its barrier with the bytes of a round.
        auto stats_send = [&](const float (&v)[64]) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) s += v[j];
            const float ml = s * (1.0f / 64.0f);
            float m2 = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float dlt = v[j] - ml;
                m2 = fmaf(dlt, dlt, m2);
            }
            const uint32_t slot = xr & 1u;
            if (ew == 0 && lane == 0) mbar_expect_tx(&ln_bar[slot], (uint32_t)(LN_EPI_WARPS * CLN * 32 * 8));
            const uint32_t off = ((slot * BM + (uint32_t)trow) * C::NSRC + (uint32_t)my_src) * 8u;
            const uint64_t pair = ((uint64_t)__float_as_uint(m2) << 32) | (uint64_t)__float_as_uint(s);
#pragma unroll
            for (int d = 0; d < CLN; ++d) st_async_b64(peer_stats[d] + off, pair, peer_bar[d] + slot * 8u);
        };
        auto stats_recv = [&](float &mean, float &rstd) {
            const uint32_t slot = xr & 1u, par = (xr >> 1) & 1u;
            ++xr;
            mbar_wait_cluster(&ln_bar[slot], par);
            float ps[C::NSRC], pm[C::NSRC];
            const uint32_t rbase = stats_s + ((slot * BM + (uint32_t)trow) * C::NSRC) * 8u;
#pragma unroll
            for (int i = 0; i < C::NSRC; i += 2) {
                const float4 t = lds128(rbase + (uint32_t)i * 8u);
                ps[i] = t.x; pm[i] = t.y; ps[i + 1] = t.z; pm[i + 1] = t.w;
            }
            float tot = 0.f;
#pragma unroll
            for (int i = 0; i < C::NSRC; ++i) tot += ps[i];
            mean = tot * (1.0f / (float)N);
            float M2 = 0.f;
#pragma unroll
            for (int i = 0; i < C::NSRC; ++i) {
                const float dm = ps[i] * (1.0f / 64.0f) - mean;
                M2 += pm[i] + 64.0f * dm * dm;
            }
            rstd = rsqrtf(M2 * (1.0f / (float)N) + epi.eps);
        };
        const uint32_t par_s = smem_u32(par) + (uint32_t)(half * 64) * 4u;   // this warp's 64 columns of [bias | w1 | b1]
        auto normalise1 = [&](float (&v)[64], float mean, float rstd) {         // LayerNorm_1: operands from shared memory (broadcast reads)
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
                const float4 g = lds128(par_s + (uint32_t)(LN_BN + j) * 4u);
                const float4 bb = lds128(par_s + (uint32_t)(2 * LN_BN + j) * 4u);
                v[j] = (v[j] - mean) * rstd * g.x + bb.x;
                v[j + 1] = (v[j + 1] - mean) * rstd * g.y + bb.y;
                v[j + 2] = (v[j + 2] - mean) * rstd * g.z + bb.z;
                v[j + 3] = (v[j + 3] - mean) * rstd * g.w + bb.w;
            }
        };
        auto normalise2 = [&](float (&v)[64], float mean, float rstd, const float *__restrict__ w, const float *__restrict__ b) {
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
                const float4 g = __ldg(reinterpret_cast<const float4 *>(w + gcol0 + j));
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(b + gcol0 + j));
                v[j] = (v[j] - mean) * rstd * g.x + bb.x;
                v[j + 1] = (v[j + 1] - mean) * rstd * g.y + bb.y;
                v[j + 2] = (v[j + 2] - mean) * rstd * g.z + bb.z;
                v[j + 3] = (v[j + 3] - mean) * rstd * g.w + bb.w;
            }
        };
        const size_t ldf = (size_t)N * 4, ldp = (size_t)N * 2;
        uint32_t tcount = 0;
        for (int rb = cid; rb < tiles_m; rb += ncl, ++tcount) {
            const int row0 = rb * BM + q * 32;
            const uint32_t buf = tcount & 1, aph = (tcount >> 1) & 1;
            long long *tl = (tl_on && tcount < 8) ? g_ln_tl[tcount] : nullptr;
            if (tl) tl[0] = clock64();
            // residual slab: issued BEFORE the accumulator is waited for (it does not depend on it), in the coalesced pattern
            // (half h, 128-byte column group t2: lane -> row 16 h + it*4 + lane/8, 16-byte chunk lane%8)
            uint4 rr[2][2][4];
            if (epi.resid) {
                const uint8_t *rbp = reinterpret_cast<const uint8_t *>(epi.resid + gcol0) + ((lane & 7) << 4);
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int row = row0 + 16 * h + it * 4 + (lane >> 3);
#pragma unroll
                        for (int t2 = 0; t2 < 2; ++t2) {
                            rr[h][t2][it] = make_uint4(0u, 0u, 0u, 0u);
                            if (row < M) rr[h][t2][it] = *reinterpret_cast<const uint4 *>(rbp + (size_t)row * ldf + t2 * 128);
                        }
                    }
            }
            mbar_wait(&acc_full[buf], aph);
            tcgen05_fence_after();
            if (tl) tl[1] = clock64();
            const uint32_t taddr = tmem_base + buf * LN_BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
            float v[64];
            {
                uint32_t a0[32], a1[32];
                tmem_ld32_issue(taddr, a0);
                tmem_ld32_issue(taddr + 32u, a1);
                tmem_wait_ld();
                tcgen05_fence_before();              // both reads have landed: the MMA warp may refill this buffer
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b0 = lds128(par_s + (uint32_t)j * 4u);
                    const float4 b1 = lds128(par_s + (uint32_t)(32 + j) * 4u);
                    v[j] = (__uint_as_float(a0[j]) + b0.x) * epi.alpha;
                    v[j + 1] = (__uint_as_float(a0[j + 1]) + b0.y) * epi.alpha;
                    v[j + 2] = (__uint_as_float(a0[j + 2]) + b0.z) * epi.alpha;
                    v[j + 3] = (__uint_as_float(a0[j + 3]) + b0.w) * epi.alpha;
                    v[32 + j] = (__uint_as_float(a1[j]) + b1.x) * epi.alpha;
                    v[32 + j + 1] = (__uint_as_float(a1[j + 1]) + b1.y) * epi.alpha;
                    v[32 + j + 2] = (__uint_as_float(a1[j + 2]) + b1.z) * epi.alpha;
                    v[32 + j + 3] = (__uint_as_float(a1[j + 3]) + b1.w) * epi.alpha;
                }
            }
            if (tl) tl[2] = clock64();
            if (epi.resid) {                          // + residual, transposed to "lane = row" through the staging tile
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            int row;
                            const uint32_t a = stg_tr(stg_s, lane, it, row);
                            sts128(a, rr[h][t2][it].x, rr[h][t2][it].y, rr[h][t2][it].z, rr[h][t2][it].w);
                        }
                        __syncwarp();
                        if ((lane >> 4) == h) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float4 r4 = lds128(stg_row(stg_s, lane, c));
                                v[32 * t2 + 4 * c] += r4.x; v[32 * t2 + 4 * c + 1] += r4.y; v[32 * t2 + 4 * c + 2] += r4.z; v[32 * t2 + 4 * c + 3] += r4.w;
                            }
                        }
                        __syncwarp();
                    }
            }
            if (tl) tl[3] = clock64();
            auto store_f32 = [&]() {
                uint8_t *ob = reinterpret_cast<uint8_t *>(epi.out_f32 + gcol0);
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2) {
                    uint32_t w[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(v[32 * t2 + i]);
                    store_half128(stg_s, lane, 0, w, ob + t2 * 128, ldf, row0, M);
                    store_half128(stg_s, lane, 1, w, ob + t2 * 128, ldf, row0, M);
                }
            };
            float mean, rstd;
            stats_send(v);
            if (tl) tl[4] = clock64();
            if (!epi.out_ln1) store_f32();            // the residual stream keeps v (stored while the statistics travel)
            if (tl) tl[5] = clock64();
            stats_recv(mean, rstd);
            if (tl) tl[6] = clock64();
            normalise1(v, mean, rstd);
            if (epi.w2) {                             // block end: x = LayerNorm_1(v), planes = LayerNorm_2(x)
                stats_send(v);
                store_f32();
                stats_recv(mean, rstd);
                normalise2(v, mean, rstd, epi.w2, epi.b2);
            } else if (epi.out_ln1) {
                store_f32();
            }
            if (tl) tl[7] = clock64();
            if (epi.hi) {                             // operand planes of the next GEMM: 64 bf16 = 128 B per row and plane
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) split_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                store_half128(stg_s, lane, 0, hi, reinterpret_cast<uint8_t *>(epi.hi + gcol0), ldp, row0, M);
                store_half128(stg_s, lane, 1, hi, reinterpret_cast<uint8_t *>(epi.hi + gcol0), ldp, row0, M);
                if (epi.lo) {
                    store_half128(stg_s, lane, 0, lo, reinterpret_cast<uint8_t *>(epi.lo + gcol0), ldp, row0, M);
                    store_half128(stg_s, lane, 1, lo, reinterpret_cast<uint8_t *>(epi.lo + gcol0), ldp, row0, M);
                }
            }
            if (tl) tl[8] = clock64();
        }
    }
    tcgen05_fence_before();
    cluster_sync();                               // nobody leaves while a peer may still write its statistics table / barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    }
}

int g_ln_dbg = 0;
int g_ln_clusters = 0;     // clusters of the last launch (reported by the timeline dump)

template <int NPASS, int CLN, bool MC>
cudaError_t launch_ln_k(const TcOperand &A, const TcOperand &W, int M, int K, const LnEpiDev &ep, int num_sms, cudaStream_t st) {
    using C = LnCfg<NPASS, CLN>;
    static PerDeviceFlag attr_flag;
    static int max_clusters[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_flag.cur()) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_ln_kernel<NPASS, CLN, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        // how many clusters can be resident at once (GPC boundaries can leave a few SMs without a full cluster)
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3((unsigned)(num_sms / CLN * CLN));
        q.blockDim = dim3(LN_THREADS);
        q.dynamicSmemBytes = C::SMEM;
        int mc = 0;
        if (cudaOccupancyMaxActiveClusters(&mc, gemm_tc_ln_kernel<NPASS, CLN, MC>, &q) != cudaSuccess || mc < 1) {
            cudaGetLastError();
            mc = num_sms / CLN;
        }
        max_clusters[dev & 63] = mc;
        attr_flag.cur() = true;
    }
    const int tiles_m = (M + BM - 1) / BM;
    int ncl = max_clusters[dev & 63];
    if (ncl > num_sms / CLN) ncl = num_sms / CLN;
    if (ncl > tiles_m) ncl = tiles_m;
    g_ln_clusters = ncl;
    const CUtensorMap &alo = (NPASS == 3) ? A.lo : A.hi, &wlo = (NPASS == 3) ? W.lo : W.hi;
    return launch_pdl(gemm_tc_ln_kernel<NPASS, CLN, MC>, dim3((unsigned)(ncl * CLN)), dim3(LN_THREADS), C::SMEM, st, A.hi, alo, W.hi, wlo, M, K, ep, g_ln_dbg);
}

}  // namespace

bool gemm_tc_ln_supported(int N) { return N == 4 * LN_BN; }

void gemm_tc_ln_set_debug(int on) { g_ln_dbg = on; }
void gemm_tc_ln_print_timeline(int n_tiles) {
    long long h[8][16];
    if (cudaMemcpyFromSymbol(h, g_ln_tl, sizeof(h)) != cudaSuccess) return;
    fprintf(stderr, "    gemm_tc_ln: %d clusters; epilogue of CTA 0 warp 2, cycles since the tile's start\n", g_ln_clusters);
    for (int i = 0; i < n_tiles && i < 8; ++i) {
        fprintf(stderr, "    tile %d:", i);
        static const char *nm[9] = {"start", "acc ready", "tmem+bias", "resid", "stats sent", "x stored", "stats recv", "normalised", "planes"};
        for (int k = 1; k < 9; ++k) fprintf(stderr, "  %s %lld", nm[k], h[i][k] - h[i][0]);
        fprintf(stderr, "   (next tile starts %lld after)\n", i + 1 < n_tiles ? h[i + 1][0] - h[i][8] : 0LL);
    }
}

cudaError_t launch_gemm_tc_ln(const TcOperand &A, const TcOperand &W, int M, int N, int K, bool split3, const LnEpi &epi, int num_sms,
                              cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    const bool mc = A.box_rows == BM / 4;        // a 32-row box: the A tile is fetched in quarters and multicast across the cluster
    if (!gemm_tc_ln_supported(N) || K % BK != 0 || (A.box_rows != BM && !mc) || W.box_rows != LN_BN) return cudaErrorInvalidValue;
    if (split3 && !(A.has_lo && W.has_lo)) return cudaErrorInvalidValue;
    if (!epi.bias || !epi.out_f32 || !epi.ln1_w || !epi.ln1_b || (epi.ln2_w && !epi.ln2_b)) return cudaErrorInvalidValue;
    LnEpiDev d;
    d.bias = epi.bias; d.resid = epi.resid; d.out_f32 = epi.out_f32;
    d.w1 = epi.ln1_w; d.b1 = epi.ln1_b; d.w2 = epi.ln2_w; d.b2 = epi.ln2_b;
    d.hi = epi.planes.hi; d.lo = epi.planes.lo;
    d.alpha = epi.alpha; d.eps = epi.eps; d.out_ln1 = epi.out_ln1 ? 1 : 0;
    if (mc) return split3 ? launch_ln_k<3, 4, true>(A, W, M, K, d, num_sms, st) : launch_ln_k<1, 4, true>(A, W, M, K, d, num_sms, st);
    return split3 ? launch_ln_k<3, 4, false>(A, W, M, K, d, num_sms, st) : launch_ln_k<1, 4, false>(A, W, M, K, d, num_sms, st);
}

}  // namespace pk
