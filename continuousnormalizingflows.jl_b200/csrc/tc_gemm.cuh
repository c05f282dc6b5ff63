// tc_gemm.cuh -- tcgen05 / TMEM / TMA GEMM with fused epilogues for wide MLPs (sm_100a).
//
//   D[m][n] = sum_k A[m][k] * B[n][k]  (+ sum_k A2[m][k] * B2[n][k])      all operands bf16, K contiguous
//
// Forward / VJP chain / tangent / backprop GEMMs: A = activations (samples x K), B = weights (units x K).
// Weight-gradient GEMMs: A = cotangents transposed (n_out x samples), B = layer inputs transposed
// (n_in x samples), K = the sample index, two operand pairs (abar h' + g w') accumulated into one tile and
// the sample range cut into split-K slices whose tiles are added into per-slice fp32 buffers (no atomics:
// one CTA owns a (tile, slice) pair, so the sum order is fixed).
//
// One CTA computes a 128 x BN tile (BN = 128 or 256): warp 0 streams 64-wide K blocks of both operands into a
// shared-memory ring with TMA (cp.async.bulk.tensor, 128-byte swizzle), one elected thread of warp 1 issues
// tcgen05.mma (M = 128, N = BN, K = 16, bf16 x bf16 -> fp32) into a BN-column TMEM accumulator and releases
// ring slots with tcgen05.commit, and warps 2..9 run the epilogue straight out of TMEM (tcgen05.ld, one
// row and half of the columns per thread) -- bias + activation + sigma', the VJP's ".* d", the exact-trace
// contraction, the second-order tangent terms, the backprop ".* d + extra" -- so activations go to HBM once,
// in bf16, row-major for the next GEMM and (reverse sweep) transposed for the weight gradient.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tc.h"

namespace icnf {
namespace tc {

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.b32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30),
// SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version 1 in [46,48), SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// value -> (hi, lo) bf16 pair: hi = bf16(x), lo = bf16(x - hi); hi + lo carries 16 mantissa bits
__device__ __forceinline__ void split_bf16(float x, float& hi, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(x));
    lo = x - hi;
}

// store 32 consecutive columns of one row as bf16 (and their lo halves when SPLIT)
template <bool SPLIT>
__device__ __forceinline__ void store_row32(__nv_bfloat16* base, size_t row_off, int nb, int pitch, int lo_o,
                                            const float (&v)[32]) {
    uint32_t hp[16], lp[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        if constexpr (SPLIT) {
            float h0, l0, h1, l1;
            split_bf16(v[j], h0, l0);
            split_bf16(v[j + 1], h1, l1);
            hp[j / 2] = pack_bf16(h0, h1);
            lp[j / 2] = pack_bf16(l0, l1);
        } else {
            hp[j / 2] = pack_bf16(v[j], v[j + 1]);
        }
    }
    uint4* ho = reinterpret_cast<uint4*>(base + row_off + nb);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (nb + q * 8 < pitch) {
            ho[q] = make_uint4(hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]);
            if constexpr (SPLIT) {
                uint4* lo = reinterpret_cast<uint4*>(base + row_off + lo_o + nb);
                lo[q] = make_uint4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]);
            }
        }
    }
}
// load 32 consecutive columns of one row (hi + lo when SPLIT) as fp32
template <bool SPLIT>
__device__ __forceinline__ void load_row32(const __nv_bfloat16* base, size_t row_off, int nb, int pitch, int lo_o,
                                           float (&v)[32]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (nb + q * 8 < pitch) {
            const uint4 a4 = *reinterpret_cast<const uint4*>(base + row_off + nb + q * 8);
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
            uint32_t lw[4] = {0u, 0u, 0u, 0u};
            if constexpr (SPLIT) {
                const uint4 l4 = *reinterpret_cast<const uint4*>(base + row_off + lo_o + nb + q * 8);
                lw[0] = l4.x; lw[1] = l4.y; lw[2] = l4.z; lw[3] = l4.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&aw[e]);
                float x0 = __low2float(h2), x1 = __high2float(h2);
                if constexpr (SPLIT) {
                    const __nv_bfloat162 l2 = *reinterpret_cast<const __nv_bfloat162*>(&lw[e]);
                    x0 += __low2float(l2);
                    x1 += __high2float(l2);
                }
                v[q * 8 + e * 2] = x0;
                v[q * 8 + e * 2 + 1] = x1;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q * 8 + e] = 0.f;
        }
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// sigma'' / sigma' from (h, d): v .* sigma'' = g .* phi for the chain cotangent g = v .* sigma'
__device__ __forceinline__ float act_ratio_rt(int act, float h, float d) {
    switch (act) {
        case ICNF_ACT_SOFTPLUS: return 1.0f - d;
        case ICNF_ACT_TANH: return -2.0f * h;
        case ICNF_ACT_SIGMOID: return 1.0f - 2.0f * h;
        default: return 0.0f;
    }
}

// transposed copy of 32 consecutive columns of row m: outT[(nb + j) * ldT + m]; a warp writes 32 consecutive
// rows m, i.e. 64 contiguous bytes per column
template <bool SPLIT>
__device__ __forceinline__ void store_col32T(__nv_bfloat16* outT, long long ldT, int lo_T, int nb, int N, int m,
                                             const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (nb + j < N) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v[j]);
            __nv_bfloat16* p = outT + (long long)(nb + j) * ldT + m;
            p[0] = hi;
            if constexpr (SPLIT) p[lo_T] = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
        }
    }
}

// Persistent: the grid is one CTA per SM and every CTA walks work items blockIdx.x, + gridDim.x, ...
// A work item is a (tile, split-K slice) pair; tiles are ordered with the B-operand tiles fastest, so
// concurrently running CTAs share the same A rows in L2.  The accumulator is double-buffered in TMEM
// (2 x BN columns): while the epilogue warps drain item i the MMA warp already accumulates item i + 1, and
// the TMA warp runs ahead across item boundaries.
template <bool SPLIT, int BN>
__global__ void __launch_bounds__(TTHREADS, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2, TcArgs g) {
    if (g.done && *g.done) return;
    constexpr int NT_A = SPLIT ? 2 : 1;   // tiles per operand and stage
    constexpr int NSTAGE = stages(SPLIT, BN);
    constexpr int BTB = b_tile_bytes(BN);
    constexpr int STAGE_BYTES = stage_bytes(SPLIT, BN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // stage s: [A_hi | A_lo? | B_hi | B_lo?]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* tmem_full = empty + NSTAGE;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;      // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* sbias = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES + 256);   // [2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb0 = (g.K + TBK - 1) / TBK;
    const int nkb1 = g.K2 > 0 ? (g.K2 + TBK - 1) / TBK : 0;
    const int ntn = (g.N + BN - 1) / BN;
    const int ntiles = ((g.M + TBM - 1) / TBM) * ntn;
    const int nsl = g.nslices > 1 ? g.nslices : 1;
    const int nwork = ntiles * nsl;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        if (nkb1) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA2) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB2) : "memory");
        }
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
                const int tile = work / nsl, sl = work - tile * nsl;
                const int m0 = (tile / ntn) * TBM, n0 = (tile % ntn) * BN;
                const int nseg = (nkb1 && n0 < g.N2) ? 2 : 1;
                for (int seg = 0; seg < nseg; ++seg) {
                    const CUtensorMap* ma = seg ? &mapA2 : &mapA;
                    const CUtensorMap* mb = seg ? &mapB2 : &mapB;
                    const int nkb = seg ? nkb1 : nkb0;
                    const int lo_a = seg ? g.lo_a2 : g.lo_a, lo_b = seg ? g.lo_b2 : g.lo_b;
                    const int kbe = (int)((long long)nkb * (sl + 1) / nsl);
                    for (int kb = (int)((long long)nkb * sl / nsl); kb < kbe; ++kb, ++it) {
                        const int s = it % NSTAGE;
                        const uint32_t ph = (it / NSTAGE) & 1;
                        uint8_t* st = smem + s * STAGE_BYTES;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], STAGE_BYTES);
                        tma_load_2d(st, ma, &full[s], kb * TBK, m0);
                        if constexpr (SPLIT) tma_load_2d(st + A_TILE_BYTES, ma, &full[s], lo_a + kb * TBK, m0);
                        tma_load_2d(st + NT_A * A_TILE_BYTES, mb, &full[s], kb * TBK, n0);
                        if constexpr (SPLIT) tma_load_2d(st + NT_A * A_TILE_BYTES + BTB, mb, &full[s], lo_b + kb * TBK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TBM, BN);
            int it = 0, i = 0;
            for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++i) {
                const int tile = work / nsl, sl = work - tile * nsl;
                const int n0 = (tile % ntn) * BN;
                const int nseg = (nkb1 && n0 < g.N2) ? 2 : 1;
                const int as = i & 1;
                mbar_wait(&tmem_empty[as], ((i >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
                uint32_t started = 0;
                for (int seg = 0; seg < nseg; ++seg) {
                    const int nkb = seg ? nkb1 : nkb0;
                    const int kbe = (int)((long long)nkb * (sl + 1) / nsl);
                    for (int kb = (int)((long long)nkb * sl / nsl); kb < kbe; ++kb, ++it) {
                        const int s = it % NSTAGE;
                        const uint32_t ph = (it / NSTAGE) & 1;
                        uint8_t* st = smem + s * STAGE_BYTES;
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint64_t a_hi = make_desc(smem_u32(st));
                        const uint64_t b_hi = make_desc(smem_u32(st + NT_A * A_TILE_BYTES));
#pragma unroll
                        for (int k = 0; k < TBK / 16; ++k) {
                            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
                            tc_mma_bf16(tacc, a_hi + 2 * k, b_hi + 2 * k, idesc, started | (uint32_t)k);
                        }
                        started = 1;
                        if constexpr (SPLIT) {
                            // a b ~ a_hi b_hi + a_lo b_hi + a_hi b_lo  (lo x lo is below fp32 rounding)
                            const uint64_t a_lo = make_desc(smem_u32(st + A_TILE_BYTES));
                            const uint64_t b_lo = make_desc(smem_u32(st + NT_A * A_TILE_BYTES + BTB));
#pragma unroll
                            for (int k = 0; k < TBK / 16; ++k) tc_mma_bf16(tacc, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
                            for (int k = 0; k < TBK / 16; ++k) tc_mma_bf16(tacc, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                        }
                        tc_commit(&empty[s]);
                    }
                }
                tc_commit(&tmem_full[as]);
            }
        }
    } else {
        // ===== epilogue: warps 2..9, TMEM lane quarter = warp % 4, two warps per quarter (half the columns each) =====
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int pitch = SPLIT ? g.lo_o : g.ldo;   // columns of one half
        int i = 0;
        for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++i) {
            const int tile = work / nsl, sl = work - tile * nsl;
            const int as = i & 1;
            const int m0 = (tile / ntn) * TBM, n0 = (tile % ntn) * BN;
            float* sb = sbias + as * BN;
            {   // stage this tile's bias slice (double-buffered: item i - 1 may still be read by a slower warp)
                const int e = threadIdx.x - 64;   // 0..255
                if (e < BN) sb[e] = (g.bias && n0 + e < g.N) ? __ldg(g.bias + n0 + e) : 0.f;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(&tmem_full[as], (i >> 1) & 1);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            const bool row_ok = m < g.M;
            float rowsum = 0.f;
            for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
                if (n0 + c0 >= g.N) break;   // warp-uniform
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0), r);
                if (!row_ok) continue;
                const int nb = n0 + c0;
                const size_t row_off = (size_t)m * g.ldo;
                if (g.ep == TEP_ACT) {
                    float hv[32], dv[32];
                    if (g.act == ICNF_ACT_SOFTPLUS) {   // the default activation without a per-element switch
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            hv[j] = 0.f; dv[j] = 0.f;
                            if (nb + j < g.N) act_eval<ICNF_ACT_SOFTPLUS>(__uint_as_float(r[j]) + sb[c0 + j], hv[j], dv[j]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            hv[j] = 0.f; dv[j] = 0.f;
                            if (nb + j < g.N) act_eval_rt(g.act, __uint_as_float(r[j]) + sb[c0 + j], hv[j], dv[j]);
                        }
                    }
                    store_row32<SPLIT>(g.out0, row_off, nb, pitch, g.lo_o, hv);
                    store_row32<SPLIT>(g.out1, row_off, nb, pitch, g.lo_o, dv);
                    if (g.outT) store_col32T<SPLIT>(g.outT, g.ldT, g.lo_T, nb, g.N, m, hv);
                } else if (g.ep == TEP_MULD) {
                    float dv[32], gv[32];
                    load_row32<SPLIT>(g.aux, row_off, nb, pitch, g.lo_o, dv);
#pragma unroll
                    for (int j = 0; j < 32; ++j) gv[j] = (nb + j < g.N) ? __uint_as_float(r[j]) * dv[j] : 0.f;
                    store_row32<SPLIT>(g.out0, row_off, nb, pitch, g.lo_o, gv);
                    if (g.outT) store_col32T<SPLIT>(g.outT, g.ldT, g.lo_T, nb, g.N, m, gv);
                } else if (g.ep == TEP_TANGENT) {
                    float dv[32], gv[32], o0[32], o1[32];
                    load_row32<SPLIT>(g.aux, row_off, nb, pitch, g.lo_o, dv);
                    load_row32<SPLIT>(g.aux1, row_off, nb, pitch, g.lo_o, gv);
                    if (g.act == ICNF_ACT_SOFTPLUS) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float rr = (nb + j < g.N) ? __uint_as_float(r[j]) : 0.f;
                            o0[j] = rr * dv[j];
                            o1[j] = rr * gv[j] * (1.0f - dv[j]);
                        }
                    } else {
                        float hv[32];
                        load_row32<SPLIT>(g.aux2, row_off, nb, pitch, g.lo_o, hv);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float rr = (nb + j < g.N) ? __uint_as_float(r[j]) : 0.f;
                            o0[j] = rr * dv[j];
                            o1[j] = rr * gv[j] * act_ratio_rt(g.act, hv[j], dv[j]);
                        }
                    }
                    store_row32<SPLIT>(g.out0, row_off, nb, pitch, g.lo_o, o0);
                    store_row32<SPLIT>(g.out1, row_off, nb, pitch, g.lo_o, o1);
                    if (g.outT) store_col32T<SPLIT>(g.outT, g.ldT, g.lo_T, nb, g.N, m, o0);
                } else if (g.ep == TEP_MULADD) {
                    float dv[32], ax[32], o0[32];
                    load_row32<SPLIT>(g.aux, row_off, nb, pitch, g.lo_o, dv);
                    load_row32<SPLIT>(g.aux1, row_off, nb, pitch, g.lo_o, ax);
#pragma unroll
                    for (int j = 0; j < 32; ++j) o0[j] = (nb + j < g.N) ? fmaf(__uint_as_float(r[j]), dv[j], ax[j]) : 0.f;
                    store_row32<SPLIT>(g.out0, row_off, nb, pitch, g.lo_o, o0);
                    if (g.outT) store_col32T<SPLIT>(g.outT, g.ldT, g.lo_T, nb, g.N, m, o0);
                } else if (g.ep == TEP_WGRAD) {
                    float* base = g.out_f32 + (long long)sl * g.slice_stride + m;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (nb + j < g.N) {
                            float* p = base + (long long)(nb + j) * g.ldw;   // a warp covers 32 consecutive m: coalesced
                            *p += __uint_as_float(r[j]);
                        }
                    }
                } else if (g.ep == TEP_TRACE) {
                    float dv[32];
                    load_row32<SPLIT>(g.aux, row_off, nb, pitch, g.lo_o, dv);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < g.N) rowsum = fmaf(__uint_as_float(r[j]), dv[j], rowsum);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nb + j;
                        if (n < g.N && n < g.n_limit) {
                            float v = __uint_as_float(r[j]);
                            if (g.ep == TEP_LIN_SOA) v += sb[c0 + j];
                            g.out_f32[(size_t)n * g.M + m] = v;
                        }
                    }
                }
            }
            // this warp's TMEM reads of the tile are complete (tcgen05.wait::ld in tmem_ld32): hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            // exact trace: one part per (unit tile, column half), added in part order by the reader (no atomics)
            if (g.ep == TEP_TRACE && row_ok) g.out_f32[(size_t)((tile % ntn) * 2 + chalf) * g.M + m] = rowsum;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
    }
}

}  // namespace tc
}  // namespace icnf
