// gemm_tc_ln.cu -- K5b: the residual GEMM of a Conformer sub-block with the LayerNorm that FOLLOWS it fused into
// the epilogue (and, at the end of a block, the chained pair final_norm_ -> next block's ffn1_.norm_):
//     v      = resid + alpha * (A[M,K] . W[N,K]^T + bias)            (FeedForward / attention out / conv pw2,
//                                                                      reference src/encoder.cpp:39-46, :182-185, :72-74)
//     y1     = LayerNorm_1(v)                                         (nn::LayerNorm, axiom operations.cpp:1796-1809:
//     y2     = LayerNorm_2(y1)            [optional]                   biased variance, eps inside the root)
//     out_f32 = v  or  y1 ;   planes = bf16 hi/lo split of the last LayerNorm's result (A operand of the next GEMM)
// It replaces the pair  gemm_tc_kernel<EPI_RESID_F32>  +  layernorm_kernel  (86 stand-alone LayerNorm launches per
// 110m step, each re-reading the fp32 residual stream from L2 and writing the operand planes).
//
// A LayerNorm row spans all N = d_model columns, i.e. CLN = N / 128 accumulator tiles that live in the TMEM of CLN
// different SMs.  The kernel therefore runs as thread-block CLUSTERS of CLN CTAs along N: cluster c walks the row
// blocks c, c + #clusters, ...; CTA rank r of the cluster owns the columns [128 r, 128 r + 128) of every block -- same
// TMA -> smem -> tcgen05.mma -> TMEM pipeline and warp roles as gemm_tc_kernel (warp 0 producer, warp 1 MMA issuer,
// warps 2..9 epilogue, double-buffered accumulator).  In the epilogue a warp holds a 32-row x 64-column slab of v in
// registers (one row per lane), reduces it to the pair (sum, centred sum of squares about its OWN mean), and writes that
// pair into the statistics table of EVERY CTA of the cluster through distributed shared memory (st.shared::cluster),
// followed by one release-arrive per destination on that CTA's mbarrier.  After an acquire-wait on the local mbarrier a
// lane combines the 2 CLN partials of its row in a fixed order with Chan's parallel formula
//     mean = (sum_i s_i) / N,   M2 = sum_i M2_i + 64 sum_i (s_i / 64 - mean)^2,   var = M2 / N
// -- one exchange per LayerNorm, numerically equivalent to the reference's two passes (no E[x^2] - mean^2
// cancellation) and deterministic.  Two table slots / two mbarriers alternate, which is enough: a warp can only be one
// exchange ahead of the slowest warp of its cluster (it needs everybody's arrival to pass the exchange in between).
// Global traffic stays coalesced: residual loads and all stores go through a 2 KB per-warp staging tile (32 rows x 64
// B, 16-byte chunks XOR-swizzled) that transposes between "lane = row" and "4 lanes = 64 contiguous bytes of a row".
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
    static constexpr int AVAIL = 227 * 1024 - STG_BYTES - STATS_BYTES - 1024 - 256;
    static constexpr int STAGES = AVAIL / STAGE_BYTES > 8 ? 8 : AVAIL / STAGE_BYTES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STG_BYTES + STATS_BYTES + 1024 /*align*/ + 256 /*barriers*/;
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

__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
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

// ---- the 32-row x 64-byte staging tile of a warp.  Chunk c (16 B) of row r sits at r * 64 + ((c ^ ((r >> 1) & 3)) << 4):
// conflict-free for "lane = row" accesses (8 lanes of a phase: 2 row parities x 4 swizzle values) and for the transposed
// ones (8 lanes = 2 rows x 4 chunks).
__device__ __forceinline__ uint32_t stg_rowwise(uint32_t stg, int lane, int c) { return stg + (uint32_t)(lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)); }
__device__ __forceinline__ uint32_t stg_transposed(uint32_t stg, int lane, int it, int &row) {
    row = it * 8 + (lane >> 2);
    return stg + (uint32_t)(row * 64 + (((lane & 3) ^ ((row >> 1) & 3)) << 4));
}
// global tile (rows row0.., 64 B per row at base + row * ld_bytes) -> r[16] = the 64 bytes of THIS lane's row (zeros past M)
__device__ __forceinline__ void load_tile64(uint32_t stg, int lane, const uint8_t *base, size_t ld_bytes, int row0, int M, uint32_t (&r)[16]) {
    uint4 t[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int row = it * 8 + (lane >> 2);
        t[it] = make_uint4(0u, 0u, 0u, 0u);
        if (row0 + row < M) t[it] = *reinterpret_cast<const uint4 *>(base + (size_t)(row0 + row) * ld_bytes + ((lane & 3) << 4));
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        int row;
        const uint32_t a = stg_transposed(stg, lane, it, row);
        sts128(a, t[it].x, t[it].y, t[it].z, t[it].w);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 v = lds128u(stg_rowwise(stg, lane, c));
        r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
    __syncwarp();
}
__device__ __forceinline__ void store_tile64(uint32_t stg, int lane, const uint32_t (&r)[16], uint8_t *base, size_t ld_bytes, int row0, int M) {
#pragma unroll
    for (int c = 0; c < 4; ++c) sts128(stg_rowwise(stg, lane, c), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        int row;
        const uint4 v = lds128u(stg_transposed(stg, lane, it, row));
        if (row0 + row < M) *reinterpret_cast<uint4 *>(base + (size_t)(row0 + row) * ld_bytes + ((lane & 3) << 4)) = v;
    }
    __syncwarp();
}

template <int NPASS, int CLN>
__global__ void __cluster_dims__(CLN, 1, 1) __launch_bounds__(LN_THREADS, 1)
gemm_tc_ln_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, int M, int K,
                  const __grid_constant__ LnEpiDev epi) {
    using C = LnCfg<NPASS, CLN>;
    constexpr int N = CLN * LN_BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *staging = tiles + (size_t)C::STAGES * C::STAGE_BYTES;
    uint8_t *stats = staging + C::STG_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stats + C::STATS_BYTES);
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
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], LN_EPI_WARPS);
            mbar_init(&ln_bar[b], LN_EPI_WARPS * CLN);      // one arrival per epilogue warp of the cluster
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync();                               // every CTA's barriers exist before any remote arrival
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
                    tma_load_2d(st, &tmA_hi, &full[s], kb * BK, m0);
                    tma_load_2d(st + C::A_BYTES, &tmW_hi, &full[s], kb * BK, n0);
                    if (NPASS == 3) {
                        tma_load_2d(st + C::A_BYTES + C::W_BYTES, &tmA_lo, &full[s], kb * BK, m0);
                        tma_load_2d(st + 2 * C::A_BYTES + C::W_BYTES, &tmW_lo, &full[s], kb * BK, n0);
                    }
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
                    umma_commit(&empty[s]);
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
        // (mean, 1/std) of this lane's row of `v` over all N columns of the cluster
        auto row_stats = [&](const float (&v)[64], float &mean, float &rstd) {
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
            const uint32_t slot = xr & 1u, par = (xr >> 1) & 1u;
            ++xr;
            const uint32_t off = ((slot * BM + (uint32_t)trow) * C::NSRC + (uint32_t)my_src) * 8u;
#pragma unroll
            for (int d = 0; d < CLN; ++d) st_cluster_f32x2(peer_stats[d] + off, s, m2);
            fence_acq_rel_cluster();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int d = 0; d < CLN; ++d) mbar_arrive_cluster(peer_bar[d] + slot * 8u);
            }
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
        auto normalise = [&](float (&v)[64], float mean, float rstd, const float *__restrict__ w, const float *__restrict__ b) {
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
        uint32_t tcount = 0;
        for (int rb = cid; rb < tiles_m; rb += ncl, ++tcount) {
            const int row0 = rb * BM + q * 32;
            const uint32_t buf = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&acc_full[buf], aph);
            tcgen05_fence_after();
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
                    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(epi.bias + gcol0 + j));
                    const float4 b1 = __ldg(reinterpret_cast<const float4 *>(epi.bias + gcol0 + 32 + j));
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
            const size_t ldf = (size_t)N * 4, ldp = (size_t)N * 2;
            if (epi.resid) {                          // + residual (coalesced loads, transposed to "lane = row" through the staging tile)
                const uint8_t *rbp = reinterpret_cast<const uint8_t *>(epi.resid + gcol0);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    uint32_t r[16];
                    load_tile64(stg_s, lane, rbp + t * 64, ldf, row0, M, r);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[16 * t + i] += __uint_as_float(r[i]);
                }
            }
            auto store_f32 = [&]() {
                uint8_t *ob = reinterpret_cast<uint8_t *>(epi.out_f32 + gcol0);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    uint32_t r[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[16 * t + i]);
                    store_tile64(stg_s, lane, r, ob + t * 64, ldf, row0, M);
                }
            };
            float mean, rstd;
            if (!epi.out_ln1) store_f32();            // the residual stream keeps v
            row_stats(v, mean, rstd);
            normalise(v, mean, rstd, epi.w1, epi.b1);
            if (epi.out_ln1) store_f32();             // block end: the residual stream is LayerNorm_1(v)
            if (epi.w2) {
                row_stats(v, mean, rstd);
                normalise(v, mean, rstd, epi.w2, epi.b2);
            }
            if (epi.hi) {                             // operand planes of the next GEMM
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) split_pair(v[32 * t + 2 * i], v[32 * t + 2 * i + 1], hi[i], lo[i]);
                    store_tile64(stg_s, lane, hi, reinterpret_cast<uint8_t *>(epi.hi + gcol0) + t * 64, ldp, row0, M);
                    if (epi.lo) store_tile64(stg_s, lane, lo, reinterpret_cast<uint8_t *>(epi.lo + gcol0) + t * 64, ldp, row0, M);
                }
            }
        }
    }
    tcgen05_fence_before();
    cluster_sync();                               // nobody leaves while a peer may still write its statistics table / barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    }
}

template <int NPASS, int CLN>
cudaError_t launch_ln_k(const TcOperand &A, const TcOperand &W, int M, int K, const LnEpiDev &ep, int num_sms, cudaStream_t st) {
    using C = LnCfg<NPASS, CLN>;
    static PerDeviceFlag attr_flag;
    static int max_clusters[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_flag.cur()) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_ln_kernel<NPASS, CLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        // how many clusters can be resident at once (GPC boundaries can leave a few SMs without a full cluster)
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3((unsigned)(num_sms / CLN * CLN));
        q.blockDim = dim3(LN_THREADS);
        q.dynamicSmemBytes = C::SMEM;
        int mc = 0;
        if (cudaOccupancyMaxActiveClusters(&mc, gemm_tc_ln_kernel<NPASS, CLN>, &q) != cudaSuccess || mc < 1) {
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
    const CUtensorMap &alo = (NPASS == 3) ? A.lo : A.hi, &wlo = (NPASS == 3) ? W.lo : W.hi;
    return launch_pdl(gemm_tc_ln_kernel<NPASS, CLN>, dim3((unsigned)(ncl * CLN)), dim3(LN_THREADS), C::SMEM, st, A.hi, alo, W.hi, wlo, M, K, ep);
}

}  // namespace

bool gemm_tc_ln_supported(int N) { return N == 4 * LN_BN; }

cudaError_t launch_gemm_tc_ln(const TcOperand &A, const TcOperand &W, int M, int N, int K, bool split3, const LnEpi &epi, int num_sms,
                              cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    if (!gemm_tc_ln_supported(N) || K % BK != 0 || A.box_rows != BM || W.box_rows != LN_BN) return cudaErrorInvalidValue;
    if (split3 && !(A.has_lo && W.has_lo)) return cudaErrorInvalidValue;
    if (!epi.bias || !epi.out_f32 || !epi.ln1_w || !epi.ln1_b || (epi.ln2_w && !epi.ln2_b)) return cudaErrorInvalidValue;
    LnEpiDev d;
    d.bias = epi.bias; d.resid = epi.resid; d.out_f32 = epi.out_f32;
    d.w1 = epi.ln1_w; d.b1 = epi.ln1_b; d.w2 = epi.ln2_w; d.b2 = epi.ln2_b;
    d.hi = epi.planes.hi; d.lo = epi.planes.lo;
    d.alpha = epi.alpha; d.eps = epi.eps; d.out_ln1 = epi.out_ln1 ? 1 : 0;
    return split3 ? launch_ln_k<3, 4>(A, W, M, K, d, num_sms, st) : launch_ln_k<1, 4>(A, W, M, K, d, num_sms, st);
}

}  // namespace pk
