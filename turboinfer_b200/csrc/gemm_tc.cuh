// gemm_tc.cuh -- tcgen05 tensor-core GEMM for the batched side of the hot path (prefill / batched decode), where the
// work really is a dense contraction:  Y[M,N] = X[M,K] . dequant(Wq[K,N]).
//
// Replaces simd_gemm_float (src/core/tensor_engine.cpp:191-255, the M,N,K >= 32 branch of TensorEngine::matmul) on
// quantized weights.  Same arithmetic contract as the streaming GEMV (gemv.cuh), so a row of the GEMM is bit-identical
// to the GEMV of that row: each activation row is converted to 24-bit block fixed point and split into three signed
// 8-bit digit planes; the tensor cores multiply the INT8 planes with the INT8 / unpacked-INT4 weights (kind::i8, exact
// int32 accumulation in TMEM, one accumulator per digit); the epilogue recombines the digits in int64 and scales once.
//
//   CTA tile      128 tokens x 128 columns, K in steps of 128 bytes; 3 accumulators x 128 columns of TMEM
//   warp 0        producer: per stage 3 digit tiles of A (TMA boxes of 128 x 128 B, SWIZZLE_128B) + 1 weight tile of B
//                 (16 KiB, stored in HBM as its swizzled image: one 1-D bulk copy)
//   warp 1        TMEM allocation; one elected lane issues tcgen05.mma (M 128, N 128, K 32): 4 k-steps x 3 digits per stage,
//                 tcgen05.commit releases the stage / publishes the accumulators
//   warps 2..5    epilogue: tcgen05.ld 32x32b (lane = token row), digit recombination, scale, store
// Operands are K-major in shared memory; descriptors follow the sm_100 layout (start address, SBO = 1024 B between 8-row
// groups, version 1, layout type SWIZZLE_128B); the k-step inside the 128-byte swizzle atom advances the start address.
#pragma once
#include <cuda.h>

#include "gemv.cuh"

namespace tib {

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 128;   // BK in bytes = int8 elements
constexpr int kGemmStages = 3;
constexpr int kGemmThreads = 192;
constexpr int kGemmTileBytes = kGemmBM * kGemmBK;              // 16 KiB
constexpr int kGemmStageBytes = 4 * kGemmTileBytes;            // 3 A digit tiles + 1 B tile
constexpr size_t kGemmSmemBytes = (size_t)kGemmStages * kGemmStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct GemmArgs {
    int M, N, K;            // logical sizes
    int m_pad, k_pad;       // padded to 128
    int a_signed_b;         // 1: weights are signed bytes (INT8), 0: unsigned (INT4 nibbles stored as q + woff)
    int woff;
    const float* sx;        // [M] s_x of each row
    const long long* sxf;   // [M] sum of xf of each row
    const float* colscale;  // [N]
    const float* colzterm;  // [N] or nullptr
    float* y;               // [M][N]
    const float* resid;     // optional [M][N] added to the result (may alias y): the residual connections of the decoder
    const uint8_t* wt;      // weights, tile-major pre-swizzled (wtile_offset)
};

// ---- activations -> digit planes [3][m_pad][k_pad] (K-major rows), one block per row ------------------------
__global__ void gemm_digits_kernel(const float* x, int M, int K, int m_pad, int k_pad, int8_t* planes, float* sx_out, long long* sxf_out) {
    __shared__ float red[32];
    __shared__ long long redl[32];
    const int row = blockIdx.x;
    const float* xr = x + (size_t)row * K;
    float amax = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) amax = fmaxf(amax, fabsf(xr[k]));
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    amax = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) amax = fmaxf(amax, red[i]);
    const bool finite = amax > 0.f && amax < INFINITY;
    const float inv_s = finite ? kXQMax / amax : 0.f;   // the expressions of gemv_stage_x, so rows match the GEMV bit for bit
    const float s_x = finite ? amax / kXQMax : 0.f;
    long long sxf = 0;
    for (int k = threadIdx.x; k < k_pad; k += blockDim.x) {
        const int f = k < K ? __float2int_rn(xr[k] * inv_s) : 0;
        sxf += f;
        // signed digits, f = d2*65536 + d1*256 + d0 with every digit in [-128, 127]: the bytes of f + 0x808080, xor 0x80
        const int u = f + 0x808080;
        planes[((size_t)0 * m_pad + row) * k_pad + k] = (int8_t)((u & 0xFF) ^ 0x80);
        planes[((size_t)1 * m_pad + row) * k_pad + k] = (int8_t)(((u >> 8) & 0xFF) ^ 0x80);
        planes[((size_t)2 * m_pad + row) * k_pad + k] = (int8_t)(((u >> 16) & 0xFF) ^ 0x80);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sxf += __shfl_xor_sync(0xffffffffu, sxf, o);
    if ((threadIdx.x & 31) == 0) redl[threadIdx.x >> 5] = sxf;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += redl[i];
        sx_out[row] = s_x;
        sxf_out[row] = t;
    }
}

// ---- the B operand in HBM: tile-major, pre-swizzled ---------------------------------------------------------------
// A weight tile of a k-step (128 columns x 128 B of k) is stored as the exact 16 KiB SWIZZLE_128B image the MMA reads from
// shared memory, tiles ordered [column tile][k-step]: one 1-D bulk copy per tile and k-step.  (A TMA box over a row-major
// [N][K] copy fetches 128 separate 128-byte rows 4 KiB apart; streamed once from HBM that was bound by the TMA request
// rate at ~30 GB/s per SM -- measured, ~1 us per k-step whatever the pipeline depth.)
// SWIZZLE_128B: inside every 8-row x 128-byte atom the 16-byte chunk c of row r sits at chunk position c ^ (r & 7).
TIB_HD size_t sw128_offset(int row, int byte) {   // inside a tile whose rows are 128 B
    return (size_t)(row >> 3) * 1024 + (size_t)(row & 7) * 128 + (size_t)(((byte >> 4) ^ (row & 7)) << 4) + (byte & 15);
}
TIB_HD size_t wtile_offset(int n, int k, int k_pad) {
    return ((size_t)(n / kGemmBN) * (k_pad / kGemmBK) + k / kGemmBK) * kGemmTileBytes + sw128_offset(n % kGemmBN, k % kGemmBK);
}
// packed streaming layout -> tile-major bytes, same thread mapping as unpack_kernel
__global__ void unpack_kmajor_kernel(const uint8_t* packed, QLayout L, int k_pad, uint8_t* wk) {
    const Slab slab = make_slab(L, blockIdx.x);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kik = kitem_k(L.bits), nel = L.bits == 4 ? 8 : 4;
    for_each_quad_word(L, slab, warp, lane, [&](size_t off, int nl, int grp, int chunk, int ki, int widx) {
        const uint32_t word = *reinterpret_cast<const uint32_t*>(packed + off);
        for (int i = 0; i < nel; ++i) {
            int row, kk;
            kitem_word_elem(L.bits, nl, widx, i, row, kk);
            const int n = slab.col0 + 16 * grp + row, k = chunk * L.kc + ki * kik + kk;
            if (k >= k_pad) continue;
            // INT4: u = q + woff (0..15); INT8: two's complement q
            wk[wtile_offset(n, k, k_pad)] = (uint8_t)(L.bits == 4 ? (word >> (4 * i)) & 0xFu : (word >> (8 * i)) & 0xFFu);
        }
    });
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B, tile rows 128 B apart inside 1024-byte 8-row groups
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_i8(int a_signed, int b_signed, int bn = kGemmBN) {
    return (2u << 4)                        // D format S32
           | ((uint32_t)a_signed << 7)      // A: 0 unsigned / 1 signed 8-bit
           | ((uint32_t)b_signed << 10)     // B
           | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kGemmBM >> 4) << 24);   // K-major A and B (bits 15, 16 = 0)
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap map_a, const GemmArgs g) {
    extern __shared__ uint8_t gsm_raw[];
    const uint32_t base = (smem_u32(gsm_raw) + 1023u) & ~1023u;
    uint8_t* tiles = gsm_raw + (base - smem_u32(gsm_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)kGemmStages * kGemmStageBytes);
    uint64_t* empty = full + kGemmStages;
    uint64_t* tmem_full = empty + kGemmStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = blockIdx.x, mb = blockIdx.y;
    const int KB = g.k_pad / kGemmBK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kGemmStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int st = kb % kGemmStages, use = kb / kGemmStages;
                if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
                mbar_arrive_expect_tx(&full[st], kGemmStageBytes);
                uint8_t* sbase = tiles + (size_t)st * kGemmStageBytes;
                for (int d = 0; d < 3; ++d) tma_load_2d(sbase + d * kGemmTileBytes, &map_a, kb * kGemmBK, d * g.m_pad + mb * kGemmBM, &full[st]);
                bulk_g2s(sbase + 3 * kGemmTileBytes, g.wt + ((size_t)nb * KB + kb) * kGemmTileBytes, kGemmTileBytes, &full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id_s = umma_idesc_i8(1, g.a_signed_b);   // A = signed digit plane
            for (int kb = 0; kb < KB; ++kb) {
                const int st = kb % kGemmStages, use = kb / kGemmStages;
                mbar_wait(&full[st], use & 1);
                tc_fence_after();
                const uint32_t sbase = base + st * kGemmStageBytes;
#pragma unroll
                for (int k4 = 0; k4 < kGemmBK / 32; ++k4) {
                    const uint64_t bdesc = umma_desc_sw128(sbase + 3 * kGemmTileBytes + k4 * 32);
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const uint64_t adesc = umma_desc_sw128(sbase + d * kGemmTileBytes + k4 * 32);
                        tc_mma_i8(tmem + d * kGemmBN, adesc, bdesc, id_s, (kb | k4) != 0 ? 1u : 0u);
                    }
                }
                tc_commit(&empty[st]);   // the stage may be refilled once these MMAs have read it
            }
            tc_commit(tmem_full);
        }
    } else {
        // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; lane = token row of the tile
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = mb * kGemmBM + q * 32 + lane;
        const bool row_ok = row < g.M;
        const float sx = row_ok ? g.sx[row] : 0.f;
        const long long sxf = row_ok ? g.sxf[row] : 0;
        const long long offterm = (long long)g.woff * sxf;
        const float fsxf = (float)sxf;
        for (int c0 = 0; c0 < kGemmBN; c0 += 16) {
            uint32_t a[3][16];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * kGemmBN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(a[d][0]), "=r"(a[d][1]), "=r"(a[d][2]), "=r"(a[d][3]), "=r"(a[d][4]), "=r"(a[d][5]), "=r"(a[d][6]), "=r"(a[d][7]),
                      "=r"(a[d][8]), "=r"(a[d][9]), "=r"(a[d][10]), "=r"(a[d][11]), "=r"(a[d][12]), "=r"(a[d][13]), "=r"(a[d][14]), "=r"(a[d][15])
                    : "r"(taddr));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
                auto val = [&](int i, int n) -> float {
                    const long long t = ((long long)(int)a[2][i] << 16) + ((long long)(int)a[1][i] << 8) + (long long)(int)a[0][i] - offterm;
                    const int hi = (int)(t >> 23), lo = (int)(t & 0x7FFFFF);
                    const float tf = fmaf((float)hi, 8388608.0f, (float)lo);
                    const float zt = g.colzterm ? g.colzterm[n] : 0.f;
                    return (fmaf(zt, fsxf, tf) * sx) * g.colscale[n];   // the GEMV epilogue's expression
                };
                const int nbase = nb * kGemmBN + c0;
                float* yrow = g.y + (size_t)row * g.N;
                const float* rrow = g.resid ? g.resid + (size_t)row * g.N : nullptr;
                if ((g.N & 3) == 0 && nbase + 16 <= g.N) {   // 64 contiguous bytes per lane: four 16-byte stores
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float4 v = make_float4(val(i, nbase + i), val(i + 1, nbase + i + 1), val(i + 2, nbase + i + 2), val(i + 3, nbase + i + 3));
                        if (rrow) {
                            const float4 r = *reinterpret_cast<const float4*>(rrow + nbase + i);
                            v.x = r.x + v.x; v.y = r.y + v.y; v.z = r.z + v.z; v.w = r.w + v.w;
                        }
                        *reinterpret_cast<float4*>(yrow + nbase + i) = v;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int n = nbase + i;
                        if (n < g.N) {
                            float yv = val(i, n);
                            if (rrow) yv = rrow[n] + yv;
                            yrow[n] = yv;
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}


// ---- the same GEMM for M <= 32 rows (batched decode), operands swapped ---------------------------------------------
// With a handful of rows the kernel above would spend 128-row MMAs and 48 KiB of (mostly zero) activation tiles on 32 live
// rows.  Here the roles are swapped: the MMA's M side (128 rows) are 128 WEIGHT COLUMNS, its N side (32) are the batch
// rows, D[128 x 32] per digit plane (96 TMEM columns):
//   * per k-step a CTA loads 16 KiB of weights and the 32 live rows of each digit plane (12 KiB -- the same bytes for
//     every CTA, an L2 broadcast), both stored as ready-made swizzled tile images and fetched with 1-D bulk copies;
//   * the tensor work is 4x smaller than with padded rows, and the epilogue has all four warps busy with coalesced stores
//     (lane = weight column);
//   * 7 stages of 28 KiB: what bounds a CTA is the chain of TMA round trips of its k-loop (measured: with 3 stages a
//     32-step loop took 21 us whatever else the chip was doing), so the ring is as deep as shared memory allows;
//   * split-K: grid (N / 128, S); CTA (n, s) multiplies k-steps [s KB / S, (s + 1) KB / S).  The partial sums are exact
//     integers, so they are combined with 64-bit integer atomics in a workspace (order-independent: bit-identical to
//     S = 1); the CTA that arrives last on the tile's counter applies offset, scales and residual, stores the tile and
//     leaves workspace and counter zeroed for the next GEMM.  S is chosen on the host for GEMMs with few column tiles.
// warps 0..3: epilogue (TMEM lanes 32w.. = weight columns), warp 4: TMEM allocation + MMA issue, warp 5: TMA producer.
#ifndef TIB_SMALL_TILES
#define TIB_SMALL_TILES 1
#endif
// A pipeline stage holds kSmallTiles consecutive k-tiles (each: weights 16 KiB + digits 12 KiB, fetched as ready-made images).
// Stages of two or three k-tiles (fewer mbarrier polls per byte at the same ~170-196 KiB in flight) were measured and are no
// faster (7B, batch 32: 5.34 ms per step with 1 tile x 7 stages, 5.37 with 2 x 3, 5.48 with 3 x 2): the polls are not what paces
// the k-loop.  Kept as a compile-time knob.
constexpr int kSmallTiles = TIB_SMALL_TILES;
constexpr int kSmallRows = 32, kSmallThreads = 192, kSmallStages = kSmallTiles == 1 ? 7 : (kSmallTiles == 2 ? 3 : 2);
constexpr int kSmallPlaneBytes = kSmallRows * kGemmBK;                       // 4 KiB
constexpr int kSmallTileBytes = kGemmTileBytes + 3 * kSmallPlaneBytes;       // one k-tile: weights 16 KiB + digits 12 KiB
constexpr int kSmallStageBytes = kSmallTiles * kSmallTileBytes;
constexpr int kSmallTmemCols = 512;                                          // four accumulator sets of 3 x 32 columns
constexpr size_t kSmallSmemBytes = (size_t)kSmallStages * kSmallStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 2 * kSmallRows * 8;

// The digit planes of the 32 rows for this kernel: [k-step][96 rows = 32 d + m][128 B], each k-step's 12 KiB the exact
// SWIZZLE_128B image of the three plane tiles (one bulk copy per k-step)
TIB_HD size_t xtile_offset(int d, int m, int k) {
    return (size_t)(k / kGemmBK) * 3 * 4096 + sw128_offset(d * 32 + m, k % kGemmBK);
}
struct SplitKArgs {
    const uint8_t* xt;        // digit planes in xtile_offset order
    unsigned long long* ws;   // [32][n_pad] partial integer sums
    unsigned int* cnt;        // [N / 128] arrivals per tile
    int n_pad;
    long long* dbg;           // optional (microbenchmark): SM-clock stamps of CTA (0, 0)
};

__host__ __device__ constexpr uint32_t umma_idesc_i8_mn(int a_signed, int b_signed, int m, int n) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(kSmallThreads, 1)
gemm_i8_tc_small_kernel(const GemmArgs g, const SplitKArgs sk) {
    extern __shared__ uint8_t gsm_raw[];
    const uint32_t base = (smem_u32(gsm_raw) + 1023u) & ~1023u;
    uint8_t* tiles = gsm_raw + (base - smem_u32(gsm_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)kSmallStages * kSmallStageBytes);
    uint64_t* empty = full + kSmallStages;
    uint64_t* tmem_full = empty + kSmallStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
    int* flag = reinterpret_cast<int*>(tmem_ptr + 1);
    float* s_sx = reinterpret_cast<float*>(tiles + (size_t)kSmallStages * kSmallStageBytes + 256);
    long long* s_off = reinterpret_cast<long long*>(s_sx + kSmallRows);   // woff * sum(xf) per row; sum(xf) itself after it
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = blockIdx.x, S = gridDim.y;
    const int KB = g.k_pad / kGemmBK;
    const int kb0 = (int)((long long)blockIdx.y * KB / S), kb1 = (int)((long long)(blockIdx.y + 1) * KB / S);
    // Every CTA reads the SAME digit-plane tiles; marching through k in lockstep would make all of them hit the same L2
    // lines at the same moment (measured: ~1.2 us per k-step whatever the pipeline depth).  Each column tile therefore
    // starts its k-loop at a different offset and wraps around -- integer sums do not care about the order.
    const int nk = kb1 - kb0, rot = nk > 0 ? (int)((blockIdx.x * 7u) % (unsigned)nk) : 0;
    const int ns = (nk + kSmallTiles - 1) / kSmallTiles;   // pipeline stages of this CTA: stage i = k-tiles [i * kSmallTiles, ...) of its (rotated) range

    if (threadIdx.x == 0) {
        for (int i = 0; i < kSmallStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(kSmallTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Everything above touches shared memory and TMEM only: under programmatic dependent launch (the lockstep step's graph) it
    // overlaps the tail of the previous kernel; the activations, scales and split-K workspace are read after this point.
    pdl_wait_prior_grid();
    pdl_launch_dependents();
    if (threadIdx.x < kSmallRows) {
        const bool ok = (int)threadIdx.x < g.M;
        s_sx[threadIdx.x] = ok ? g.sx[threadIdx.x] : 0.f;
        s_off[threadIdx.x] = ok ? g.sxf[threadIdx.x] : 0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    // Both pipeline loops poll mbarriers, and a poll costs ~0.1-0.2 us even when the phase is long complete (measured,
    // scripts/microbench7.cu) -- as much as a whole 28 KiB k-step should take.  So four lanes probe the next four stages
    // at once (lane 0 blocking, the others non-blocking) and lane 0 then handles every consecutive stage that is ready.
    auto probe4 = [&](uint64_t* bars, int i, int phase_flip) -> int {
        const int j = i + lane;
        bool ok = false;
        if (lane < 4 && j < ns) {
            const int stj = j % kSmallStages;
            const uint32_t parj = (uint32_t)((j / kSmallStages) & 1) ^ (uint32_t)phase_flip;
            ok = lane == 0 ? mbar_try_wait(&bars[stj], parj) : mbar_test_wait(&bars[stj], parj);
        }
        const unsigned int mask = __ballot_sync(0xffffffffu, ok) & 0xFu;
        return __ffs(~mask) - 1;   // consecutive ready stages from i on (0..4)
    };
    if (warp == 5) {
        int i = 0;
        while (i < ns) {
            const int nready = probe4(empty, i, 1);   // a fresh mbarrier counts its "previous" phase as complete
            if (lane == 0) {
                for (int q = 0; q < nready; ++q) {
                    const int ii = i + q, st = ii % kSmallStages;
                    const int ntile = min(kSmallTiles, nk - ii * kSmallTiles);
                    if (sk.dbg && blockIdx.x == 0 && blockIdx.y == 0 && ii < 40) sk.dbg[64 + ii] = clock64();
                    mbar_arrive_expect_tx(&full[st], (uint32_t)(ntile * kSmallTileBytes));
                    for (int j = 0; j < ntile; ++j) {
                        const int tt = ii * kSmallTiles + j;
                        const int kb = kb0 + (tt + rot < nk ? tt + rot : tt + rot - nk);
                        uint8_t* sbase = tiles + (size_t)st * kSmallStageBytes + (size_t)j * kSmallTileBytes;
                        bulk_g2s(sbase, g.wt + ((size_t)nb * KB + kb) * kGemmTileBytes, kGemmTileBytes, &full[st]);
                        bulk_g2s(sbase + kGemmTileBytes, sk.xt + (size_t)kb * 3 * kSmallPlaneBytes, 3 * kSmallPlaneBytes, &full[st]);
                    }
                }
            }
            __syncwarp();
            i += nready;
        }
    } else if (warp == 4) {
        // A = weights (unsigned nibbles-in-bytes for INT4, signed for INT8); B = ALL THREE signed digit planes at once:
        // their 32-row tiles are contiguous in the stage, so one MMA with N = 96 fills the three accumulators
        // (columns 32 d + m) of the k4 step's accumulator set.
        const uint32_t id = umma_idesc_i8_mn(g.a_signed_b, 1, kGemmBN, 3 * kSmallRows);
        int i = 0;
        while (i < ns) {
            const int nready = probe4(full, i, 0);
            if (lane == 0 && nready > 0) {
                tc_fence_after();
                for (int q = 0; q < nready; ++q) {
                    const int ii = i + q, st = ii % kSmallStages;
                    const int ntile = min(kSmallTiles, nk - ii * kSmallTiles);
                    if (sk.dbg && blockIdx.x == 0 && blockIdx.y == 0 && ii < 40) sk.dbg[128 + ii] = clock64();
                    for (int j = 0; j < ntile; ++j) {
                        const uint32_t sbase = base + st * kSmallStageBytes + j * kSmallTileBytes;
#pragma unroll
                        for (int k4 = 0; k4 < kGemmBK / 32; ++k4) {
                            const uint64_t wdesc = umma_desc_sw128(sbase + k4 * 32);
                            const uint64_t xdesc = umma_desc_sw128(sbase + kGemmTileBytes + k4 * 32);
                            // one accumulator set per k4: four independent accumulation chains; the epilogue adds the sets
                            tc_mma_i8(tmem + k4 * 3 * kSmallRows, wdesc, xdesc, id, (ii | j) != 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(&empty[st]);   // the stage may be refilled once these MMAs have read it
                }
            }
            __syncwarp();
            i += nready;
        }
        if (lane == 0) tc_commit(tmem_full);
    } else {
        // epilogue, warps 0..3: TMEM lane = weight column of the tile, TMEM column = batch row (per digit plane)
        if (sk.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) sk.dbg[0] = clock64();
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (sk.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) sk.dbg[1] = clock64();
        const int n = nb * kGemmBN + warp * 32 + lane;
        const bool n_ok = n < g.N;
        const float cs = n_ok ? g.colscale[n] : 0.f;
        const float zt = (n_ok && g.colzterm) ? g.colzterm[n] : 0.f;
        auto finish = [&](int m, long long t) {   // t = sum over all k of u * xf for (row m, column n)
            const long long sxf = s_off[m];
            t -= (long long)g.woff * sxf;
            const int hi = (int)(t >> 23), lo = (int)(t & 0x7FFFFF);
            const float tf = fmaf((float)hi, 8388608.0f, (float)lo);
            float yv = (fmaf(zt, (float)sxf, tf) * s_sx[m]) * cs;   // the GEMV epilogue's expression
            if (g.resid) yv = g.resid[(size_t)m * g.N + n] + yv;
            g.y[(size_t)m * g.N + n] = yv;
        };
        for (int m0 = 0; m0 < kSmallRows; m0 += 16) {
            uint32_t a[3][16];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[d][i] = 0u;
#pragma unroll
                for (int set = 0; set < kGemmBK / 32; ++set) {
                    uint32_t b[16];
                    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(set * 3 * kSmallRows + d * kSmallRows + m0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
                          "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) a[d][i] += b[i];   // the four k4 chains of this digit plane
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int m = m0 + i;
                const long long t = ((long long)(int)a[2][i] << 16) + ((long long)(int)a[1][i] << 8) + (long long)(int)a[0][i];
                if (m < g.M && n_ok) {
                    if (S == 1) finish(m, t);
                    else atomicAdd(sk.ws + (size_t)m * sk.n_pad + n, (unsigned long long)t);
                }
            }
        }
        if (S > 1) {
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 0) {
                const unsigned int prev = atomicAdd(sk.cnt + nb, 1u);
                *flag = prev == (unsigned int)(S - 1) ? 1 : 0;
                if (*flag) sk.cnt[nb] = 0u;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (*flag) {   // every split of this tile has added its partial sums
                __threadfence();
                for (int m0 = 0; m0 < g.M; m0 += 16) {   // 16 loads in flight, then their 16 rows
                    unsigned long long v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = (m0 + i < g.M && n_ok) ? __ldcg(sk.ws + (size_t)(m0 + i) * sk.n_pad + n) : 0ull;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (m0 + i < g.M && n_ok) {
                            finish(m0 + i, (long long)v[i]);
                            sk.ws[(size_t)(m0 + i) * sk.n_pad + n] = 0ull;
                        }
                }
            }
        }
        if (sk.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) sk.dbg[2] = clock64();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kSmallTmemCols) : "memory");
    }
}

}  // namespace tib
