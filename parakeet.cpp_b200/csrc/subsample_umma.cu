// subsample_umma.cu -- K3 on the tensor cores: conv1_ (1 -> 256 channels, 3x3, stride 2, pad 1) + ReLU + dw1_ (depthwise 3x3,
// stride 2, pad 1) of ConvSubsampling::forward (reference src/encoder.cpp:219-241, conv2d semantics axiom
// operations.cpp:3133-3326), fused; replaces subsample_conv1_dw1_kernel (subsample.cu), which spends 0.45 ms of a 64 x 10 s
// step issuing 2.9 G fp32 FMAs for conv1_ alone.
//
// conv1_ has ONE input channel: per output position it is a 9-term dot product with each of the 256 filters, i.e. a GEMM
//     conv1[pos][c] = sum_k patch[pos][k] . w1[c][k],   k = 3 p + q < 9 (padded to K = 16)
// Work item = one dw1 output row (utterance b, t2): it needs the conv1 rows t1 = 2 t2 - 1 .. 2 t2 + 1, i.e. M = 3 x f1n = 120
// positions (mel 80 -> f1n 40) = one UMMA M = 128 tile; N = 256 channels, K = 16: three tcgen05.mma (hi.hi + hi.lo + lo.hi of the
// bf16 split, fp32 accumulate in TMEM) per item.  The middle conv1 row of an item is recomputed by its neighbour (1.5x the
// conv1 work) -- free on the tensor pipe.  The patches (A operand) are written by the CTA's own threads as a SWIZZLE_128B
// K-major tile of which only the first 32 bytes of a row (K = 16) are ever read; the filters (B operand, 256 rows) once per CTA.
// Roles (288 threads, persistent CTAs, items c, c + grid, ...): warps 0..7 = two threads per position (TMEM lane = position,
// each half takes 128 channels): bias + ReLU + validity (conv1 rows outside the utterance are dw1's zero padding) out of TMEM,
// through a [position][64 channels] fp32 staging tile per half (16-byte chunks XOR-swizzled by the position), then dw1 as 9
// float4 FMAs per (f2, 4 channels) and the bf16 hi/lo planes of the next GEMM's A operand; they also build the next item's
// patches while the tensor core works on the current one.  Warp 8 issues the MMAs (accumulator double buffered: 2 x 256 columns).
#include <cuda.h>

#include "kernels.h"
#include "tc_prims.cuh"

namespace pk {
namespace {

using namespace tc;

constexpr int SU_C = 256;                         // conv channels
constexpr int SU_THREADS = 288;
constexpr int SU_ROWTHR = 256;
constexpr int SU_A_PLANE = 128 * 128;             // a 128-row SWIZZLE_128B tile: 16 KB (only K = 16 of each row is used)
constexpr int SU_B_PLANE = 256 * 128;
constexpr int SU_STG = 128 * 256;                 // staging per half: 128 positions x 64 channels fp32
constexpr int SU_SMEM = 2 * 2 * SU_A_PLANE + 2 * SU_B_PLANE + 2 * SU_STG + 1024 + 128;
static_assert(SU_SMEM <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t *v) {
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

// 16 bf16 (hi or lo) of one operand row: chunks 0 / 1 of the row's 128-byte swizzle line
__device__ __forceinline__ void put_row16(uint32_t plane_s, int row, const uint32_t (&w)[8]) {
    const uint32_t r = plane_s + (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    sts128(r + ((0u ^ sw) << 4), w[0], w[1], w[2], w[3]);
    sts128(r + ((1u ^ sw) << 4), w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void split16(const float (&x)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) split_pair(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
}

__global__ void __launch_bounds__(SU_THREADS, 1)
subsample_conv1_dw1_umma_kernel(const float *__restrict__ feats, const int32_t *__restrict__ frame_off, const int32_t *__restrict__ s2_off,
                                int n_utt, int mel, const float *__restrict__ w1, const float *__restrict__ b1,
                                const float *__restrict__ wd, const float *__restrict__ bd, ActBuf out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;                               // 2 buffers x (hi | lo)
    uint8_t *sB = sA + 4 * SU_A_PLANE;                // filters: hi | lo, 256 rows
    uint8_t *sS = sB + 2 * SU_B_PLANE;                // staging: half 0 | half 1
    uint64_t *bars = reinterpret_cast<uint64_t *>(sS + 2 * SU_STG);
    uint64_t *a_full = bars, *acc_full = bars + 2, *acc_empty = bars + 4, *b_full = bars + 6;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f1n = (mel - 1) / 2 + 1, f2n = (f1n - 1) / 2 + 1;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 128);               // the 128 threads of half 0 build the patches
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], SU_ROWTHR);
        }
        mbar_init(b_full, SU_ROWTHR);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();
    // dw1 output rows of this launch: absolute row numbers it0 .. it0 + n_items (the offset arrays may be a view of a sub-range of the batch)
    const int it0 = s2_off[0], n_items = s2_off[n_utt] - it0;

    if (warp == 8) {
        // ===================== MMA issue =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, SU_C);
            const uint32_t a_s = smem_u32(sA), b_s = smem_u32(sB);
            const uint64_t bh = umma_desc_sw128(b_s), bl = umma_desc_sw128(b_s + SU_B_PLANE);
            mbar_wait(b_full, 0);
            uint32_t n = 0;
            for (int li = blockIdx.x; li < n_items; li += gridDim.x, ++n) {
                const uint32_t buf = n & 1u, ph = (n >> 1) & 1u;
                mbar_wait(&a_full[buf], ph);
                mbar_wait(&acc_empty[buf], ph ^ 1u);
                tcgen05_fence_after();
                const uint64_t ah = umma_desc_sw128(a_s + buf * 2 * SU_A_PLANE), al = umma_desc_sw128(a_s + buf * 2 * SU_A_PLANE + SU_A_PLANE);
                const uint32_t d = tmem_base + buf * SU_C;
                umma_bf16(d, ah, bh, idesc, 0);
                umma_bf16(d, ah, bl, idesc, 1);
                umma_bf16(d, al, bh, idesc, 1);
                umma_commit(&acc_full[buf]);          // accumulator ready; the patches of this buffer have been read
            }
        }
    } else {
        // ===================== positions: thread = (position, channel half) =====================
        const int qd = warp & 3, hf = warp >> 2;      // TMEM lane quarter; channels 128 hf .. 128 hf + 127
        const int pos = qd * 32 + lane;               // conv1 position r * f1n + f1 of the item (< 3 f1n; rows past that idle)
        const int pr = pos / f1n, pf = pos - pr * f1n;
        const int gt = threadIdx.x & 127;             // index inside the half (dw1 pass)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
        const uint32_t stg_s = smem_u32(sS) + (uint32_t)hf * SU_STG;
        // ---- filters -> B operand (row = channel), once
        {
            const int c = threadIdx.x;                // 0 .. 255
            float x[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = k < 9 ? w1[c * 9 + k] : 0.f;
            uint32_t hi[8], lo[8];
            split16(x, hi, lo);
            put_row16(smem_u32(sB), c, hi);
            put_row16(smem_u32(sB) + SU_B_PLANE, c, lo);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(b_full);
        }
        // item -> (utterance, t2): s2_off is the prefix sum of the t2 counts
        int ub = 0;
        auto locate = [&](int it, int &b, int &t2) {
            while (ub + 1 < n_utt && s2_off[ub + 1] <= it) ++ub;      // items of a CTA ascend
            b = ub;
            t2 = it - s2_off[ub];
        };
        // patches of item `it` -> A buffer (half 0 only: 128 threads, one position each)
        auto build = [&](int it, uint32_t buf) {
            int b, t2;
            locate(it, b, t2);
            const int F = frame_off[b + 1] - frame_off[b];
            const float *src = feats + (size_t)frame_off[b] * mel;
            float x[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = 0.f;
            if (pr < 3) {
                const int t1 = 2 * t2 - 1 + pr;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const int fr = 2 * t1 - 1 + p;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int col = 2 * pf - 1 + q;
                        if (fr >= 0 && fr < F && col >= 0 && col < mel) x[3 * p + q] = __ldg(src + (size_t)fr * mel + col);
                    }
                }
            }
            uint32_t hi[8], lo[8];
            split16(x, hi, lo);
            const uint32_t a_s = smem_u32(sA) + buf * 2 * SU_A_PLANE;
            put_row16(a_s, pos, hi);
            put_row16(a_s + SU_A_PLANE, pos, lo);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&a_full[buf]);
        };
        if (hf == 0 && (int)blockIdx.x < n_items) build(it0 + (int)blockIdx.x, 0);
        // dw1 pass: thread = (4-channel group cq of the 64-channel chunk, f2 = fg, fg + 8, fg + 16); its taps for both chunks stay in registers
        const int cq = gt & 15, fg = gt >> 4;
        float4 wt[2][9], bias2[2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int c = hf * 128 + ch * 64 + 4 * cq;
#pragma unroll
            for (int k = 0; k < 9; ++k)
                wt[ch][k] = make_float4(__ldg(wd + (size_t)c * 9 + k), __ldg(wd + (size_t)(c + 1) * 9 + k), __ldg(wd + (size_t)(c + 2) * 9 + k),
                                        __ldg(wd + (size_t)(c + 3) * 9 + k));
            bias2[ch] = __ldg(reinterpret_cast<const float4 *>(bd + c));
        }
        uint32_t n = 0;
        int ub_e = 0;                                  // (the epilogue walks the items with its own cursor)
        for (int li = blockIdx.x; li < n_items; li += gridDim.x, ++n) {
            const int it = it0 + li;
            const uint32_t buf = n & 1u, ph = (n >> 1) & 1u;
            if (hf == 0 && li + (int)gridDim.x < n_items) build(it + (int)gridDim.x, buf ^ 1u);   // (its previous use was consumed: acc_full of item n - 1 was waited for)
            int b = ub_e, t2;
            while (b + 1 < n_utt && s2_off[b + 1] <= it) ++b;
            ub_e = b;
            t2 = it - s2_off[b];
            const int F = frame_off[b + 1] - frame_off[b];
            const int t1n = (F - 1) / 2 + 1;
            const int t1 = 2 * t2 - 1 + pr;
            const bool valid = pr < 3 && t1 >= 0 && t1 < t1n;     // conv1 rows outside the utterance are dw1's zero padding
            mbar_wait(&acc_full[buf], ph);
            tcgen05_fence_after();
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {           // 64-channel chunks of this half
                const int c0 = hf * 128 + ch * 64;
                uint32_t v[64];
                tmem_ld64(lane_addr + buf * SU_C + (uint32_t)c0, v);
                tmem_wait_ld();
                if (ch == 1) {                         // both chunks are out of TMEM: the MMA warp may reuse this accumulator
                    tcgen05_fence_before();
                    mbar_arrive(&acc_empty[buf]);
                }
                // conv1 = relu(acc + b1) (0 outside the utterance) -> staging row `pos`
                const uint32_t srow = stg_s + (uint32_t)pos * 256u, sw = (uint32_t)(pos & 15);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 bb = __ldg(reinterpret_cast<const float4 *>(b1 + c0 + 4 * j));
                    float4 o;
                    o.x = valid ? fmaxf(__uint_as_float(v[4 * j]) + bb.x, 0.f) : 0.f;
                    o.y = valid ? fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f) : 0.f;
                    o.z = valid ? fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f) : 0.f;
                    o.w = valid ? fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f) : 0.f;
                    sts128(srow + (((uint32_t)j ^ sw) << 4), __float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w));
                }
                named_bar(1 + hf, 128);                // the chunk of all positions is staged
                // dw1: out[f2][c] = bd[c] + sum_{i,j} wd[c][3 i + j] . conv1[row i][2 f2 - 1 + j][c]
                {
                    const int c = c0 + 4 * cq;
                    for (int f2 = fg; f2 < f2n; f2 += 8) {
                        float4 acc = bias2[ch];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const int f1 = 2 * f2 - 1 + j;
                                if (f1 < 0 || f1 >= f1n) continue;
                                const int p = i * f1n + f1;
                                const float4 x = lds128(stg_s + (uint32_t)p * 256u + ((((uint32_t)cq) ^ (uint32_t)(p & 15)) << 4));
                                acc.x = fmaf(wt[ch][3 * i + j].x, x.x, acc.x);
                                acc.y = fmaf(wt[ch][3 * i + j].y, x.y, acc.y);
                                acc.z = fmaf(wt[ch][3 * i + j].z, x.z, acc.z);
                                acc.w = fmaf(wt[ch][3 * i + j].w, x.w, acc.w);
                            }
                        store_act4(out, ((size_t)it * f2n + f2) * SU_C + c, acc);
                    }
                }
                named_bar(1 + hf, 128);                // everybody has read the staging tile
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace

bool subsample_umma_supported(int mel, int C) { return C == SU_C && mel >= 3 && 3 * ((mel - 1) / 2 + 1) <= 128; }

bool launch_subsample_conv1_dw1_umma(const float *feats, const int32_t *frame_off, const int32_t *s2_off, int n_utt, int mel, int C, const float *w1,
                                     const float *b1, const float *wd, const float *bd, ActBuf out, int num_sms, cudaStream_t st) {
    if (!subsample_umma_supported(mel, C) || n_utt < 1) return false;
    static PerDeviceFlag attr_flag;
    if (!attr_flag.cur()) {
        if (cudaFuncSetAttribute(subsample_conv1_dw1_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SU_SMEM) != cudaSuccess) return false;
        attr_flag.cur() = true;
    }
    return launch_pdl(subsample_conv1_dw1_umma_kernel, dim3((unsigned)num_sms), dim3(SU_THREADS), (size_t)SU_SMEM, st, feats, frame_off, s2_off, n_utt, mel, w1, b1, wd,
                      bd, out) == cudaSuccess;
}

}  // namespace pk
