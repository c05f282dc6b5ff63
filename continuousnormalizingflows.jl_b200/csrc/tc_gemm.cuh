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
// 32 lanes x 16 consecutive fp32 columns (issue only: pair with tmem_wait_ld before the registers are read)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

constexpr int NC = 16;   // accumulator columns per epilogue chunk (one tcgen05.ld .x16 per thread)
// the WQ epilogue warps that share a TMEM lane quarter split the tile's bn / NC column chunks as evenly as possible:
// warp `cq` owns chunks [chunk_begin(bn, cq), chunk_begin(bn, cq + 1))
__host__ __device__ constexpr int chunk_begin(int bn, int cq) { return (bn / NC) * cq / WQ; }

// Row-major bf16 outputs leave the SM through TMA: the 32 lanes of an epilogue warp (= 32 consecutive rows) put their
// NC columns into the warp's staging box in shared memory (hi half, then lo half) and one lane issues the bulk tensor
// stores.  Storing from registers instead costs one 16-byte piece per lane and row, i.e. 32 half-filled sectors per
// instruction: measured, that made the epilogue of a 128 x 128 tile (15 000 cycles) longer than its main loop.
struct EpiStore {
    uint8_t* stage;          // this warp's staging: [32 rows][NC] bf16 hi, then the same for lo
    int lane;
    int m_warp;              // first row of this warp
    int lo_o;                // column offset of the lo half in the output rows
    // direct mode (GEMM chains): every lane stores its row's NC columns straight from registers -- one full 32-byte
    // sector per half.  A bulk tensor store queues behind the producer's operand loads in the SM's TMA unit, so a warp
    // that recycles one staging box per chunk waits out that queue for every chunk: measured, the epilogue of a tile
    // then takes twice as long as its main loop and sets the pace of the whole chain.
    bool direct = false;
    bool row_ok = false;
    long long row_off = 0;   // m * ldo
};
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
template <bool SPLIT>
__device__ __forceinline__ void store_rows(const EpiStore& es, const CUtensorMap* map, __nv_bfloat16* base, int nb, const float (&v)[NC]) {
    uint32_t hp[NC / 2], lp[NC / 2];
#pragma unroll
    for (int j = 0; j < NC; j += 2) {
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
    if (es.direct) {
        if (es.row_ok) {
            uint4* hrow = reinterpret_cast<uint4*>(base + es.row_off + nb);
#pragma unroll
            for (int q = 0; q < NC / 8; ++q) __stcg(hrow + q, make_uint4(hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]));
            if constexpr (SPLIT) {
                uint4* lrow = reinterpret_cast<uint4*>(base + es.row_off + es.lo_o + nb);
#pragma unroll
                for (int q = 0; q < NC / 8; ++q) __stcg(lrow + q, make_uint4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]));
            }
        }
        return;
    }
    // the previous boxes of this warp have been read out of the staging buffer
    if (es.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    uint4* hrow = reinterpret_cast<uint4*>(es.stage + es.lane * (NC * 2));
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) hrow[q] = make_uint4(hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]);
    if constexpr (SPLIT) {
        uint4* lrow = reinterpret_cast<uint4*>(es.stage + 32 * NC * 2 + es.lane * (NC * 2));
#pragma unroll
        for (int q = 0; q < NC / 8; ++q) lrow[q] = make_uint4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (es.lane == 0) {
        tma_store_2d(map, es.stage, nb, es.m_warp);                                   // rows >= M are clipped by the tensor map
        if constexpr (SPLIT) tma_store_2d(map, es.stage + 32 * NC * 2, es.lo_o + nb, es.m_warp);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
// load NC consecutive columns of one row (hi + lo when SPLIT) as fp32
template <bool SPLIT>
__device__ __forceinline__ void load_row(const __nv_bfloat16* base, size_t row_off, int nb, int pitch, int lo_o, float (&v)[NC]) {
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
        if (nb + q * 8 < pitch) {
            const uint4 a4 = __ldcg(reinterpret_cast<const uint4*>(base + row_off + nb + q * 8));   // L2: another SM may have written it in this launch (chains)
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
            uint32_t lw[4] = {0u, 0u, 0u, 0u};
            if constexpr (SPLIT) {
                const uint4 l4 = __ldcg(reinterpret_cast<const uint4*>(base + row_off + lo_o + nb + q * 8));
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

// the same NC columns as raw bf16 pairs (x[0..1] = hi half, x[2..3] = lo half), for a load issued one chunk ahead of its use
struct RawRow { uint4 x[4]; };
template <bool SPLIT>
__device__ __forceinline__ void load_row_raw(const __nv_bfloat16* base, size_t row_off, int nb, int pitch, int lo_o, RawRow& o) {
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
        o.x[q] = make_uint4(0u, 0u, 0u, 0u);
        o.x[2 + q] = make_uint4(0u, 0u, 0u, 0u);
        if (nb + q * 8 < pitch) {
            o.x[q] = __ldcg(reinterpret_cast<const uint4*>(base + row_off + nb + q * 8));
            if constexpr (SPLIT) o.x[2 + q] = __ldcg(reinterpret_cast<const uint4*>(base + row_off + lo_o + nb + q * 8));
        }
    }
}
template <bool SPLIT>
__device__ __forceinline__ void raw_to_f32(const RawRow& o, float (&v)[NC]) {
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
        const uint32_t aw[4] = {o.x[q].x, o.x[q].y, o.x[q].z, o.x[q].w};
        const uint32_t lw[4] = {o.x[2 + q].x, o.x[2 + q].y, o.x[2 + q].z, o.x[2 + q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x0 = __uint_as_float(aw[e] << 16), x1 = __uint_as_float(aw[e] & 0xffff0000u);
            if constexpr (SPLIT) { x0 += __uint_as_float(lw[e] << 16); x1 += __uint_as_float(lw[e] & 0xffff0000u); }
            v[q * 8 + e * 2] = x0;
            v[q * 8 + e * 2 + 1] = x1;
        }
    }
}

__device__ __forceinline__ void trace_event(long long* base, int off, int tag) {
    if (base && blockIdx.x == 0) {
        long long* trace = base + off;
        const long long n = trace[0];
        if (n < 4000) { trace[1 + 2 * n] = tag; trace[2 + 2 * n] = clock64(); trace[0] = n + 1; }
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

// sigma'(a) from h = sigma(a)
__device__ __forceinline__ float act_deriv_from_h(int act, float h) {
    switch (act) {
        case ICNF_ACT_SOFTPLUS: return 1.0f - __expf(-h);     // h = log(1 + e^a)  =>  e^-h = 1 - sigmoid(a)
        case ICNF_ACT_TANH: return 1.0f - h * h;
        case ICNF_ACT_SIGMOID: return h * (1.0f - h);
        default: return 1.0f;
    }
}
// the sigma' values of a chunk: loaded, or derived from the loaded h
template <bool SPLIT>
__device__ __forceinline__ void load_deriv(const TcArgs& g, size_t row_off, int nb, int pitch, float (&dv)[NC], const RawRow* pre = nullptr) {
    if (pre) raw_to_f32<SPLIT>(*pre, dv);
    else load_row<SPLIT>(g.aux, row_off, nb, pitch, g.lo_o, dv);
    if (g.aux_is_h) {
        if (g.act == ICNF_ACT_SOFTPLUS) {
#pragma unroll
            for (int j = 0; j < NC; ++j) dv[j] = 1.0f - __expf(-dv[j]);
        } else {
#pragma unroll
            for (int j = 0; j < NC; ++j) dv[j] = act_deriv_from_h(g.act, dv[j]);
        }
    }
}

// transposed copy of NC consecutive columns of row m: outT[(nb + j) * ldT + m]; a warp writes 32 consecutive
// rows m, i.e. 64 contiguous bytes per column
template <bool SPLIT>
__device__ __forceinline__ void store_colT(__nv_bfloat16* outT, long long ldT, int lo_T, int nb, int N, int m, const float (&v)[NC]) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        if (nb + j < N) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v[j]);
            __nv_bfloat16* p = outT + (long long)(nb + j) * ldT + m;
            p[0] = hi;
            if constexpr (SPLIT) p[lo_T] = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
        }
    }
}

// epilogue of NC consecutive accumulator columns (units nb .. nb + NC - 1) of row m; `sbc` = the bias of those columns.
// Every lane of the warp calls it (the staged stores are warp-wide); rows beyond M compute on zeros and store nothing.
template <bool SPLIT>
__device__ __forceinline__ void tc_epilogue(const TcArgs& g, const CUtensorMap* mapO0, const CUtensorMap* mapO1, const EpiStore& es,
                                            const uint32_t (&r)[NC], bool row_ok, int m, int nb, const float* sbc, int sl, int pitch,
                                            float& rowsum, const RawRow* pre = nullptr) {   // pre: g.aux's columns of this chunk, already loaded
    const size_t row_off = (size_t)m * g.ldo;
    auto zero = [](float (&v)[NC]) {
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = 0.f;
    };
    if (g.ep == TEP_ACT) {
        float hv[NC], dv[NC];
        if (g.act == ICNF_ACT_SOFTPLUS) {   // the default activation without a per-element switch
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                hv[j] = 0.f; dv[j] = 0.f;
                if (nb + j < g.N) act_eval<ICNF_ACT_SOFTPLUS>(__uint_as_float(r[j]) + sbc[j], hv[j], dv[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                hv[j] = 0.f; dv[j] = 0.f;
                if (nb + j < g.N) act_eval_rt(g.act, __uint_as_float(r[j]) + sbc[j], hv[j], dv[j]);
            }
        }
        store_rows<SPLIT>(es, mapO0, g.out0, nb, hv);
        if (g.out1) store_rows<SPLIT>(es, mapO1, g.out1, nb, dv);
        if (g.outT && row_ok) store_colT<SPLIT>(g.outT, g.ldT, g.lo_T, nb, g.N, m, hv);
    } else if (g.ep == TEP_MULD) {
        float dv[NC], gv[NC];
        if (row_ok) load_deriv<SPLIT>(g, row_off, nb, pitch, dv, pre); else zero(dv);
#pragma unroll
        for (int j = 0; j < NC; ++j) gv[j] = (nb + j < g.N) ? __uint_as_float(r[j]) * dv[j] : 0.f;
        store_rows<SPLIT>(es, mapO0, g.out0, nb, gv);
        if (g.outT && row_ok) store_colT<SPLIT>(g.outT, g.ldT, g.lo_T, nb, g.N, m, gv);
    } else if (g.ep == TEP_TANGENT) {
        float dv[NC], gv[NC], o0[NC], o1[NC];
        float hv[NC];
        if (row_ok) {
            load_row<SPLIT>(g.aux1, row_off, nb, pitch, g.lo_o, gv);
            if (g.act != ICNF_ACT_SOFTPLUS) load_row<SPLIT>(g.aux2, row_off, nb, pitch, g.lo_o, hv);   // phi needs h itself
            load_deriv<SPLIT>(g, row_off, nb, pitch, dv, pre);
        } else { zero(dv); zero(gv); zero(hv); }
        if (g.act == ICNF_ACT_SOFTPLUS) {
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float rr = (nb + j < g.N) ? __uint_as_float(r[j]) : 0.f;
                o0[j] = rr * dv[j];
                o1[j] = rr * gv[j] * (1.0f - dv[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float rr = (nb + j < g.N) ? __uint_as_float(r[j]) : 0.f;
                o0[j] = rr * dv[j];
                o1[j] = rr * gv[j] * act_ratio_rt(g.act, hv[j], dv[j]);
            }
        }
        store_rows<SPLIT>(es, mapO0, g.out0, nb, o0);
        store_rows<SPLIT>(es, mapO1, g.out1, nb, o1);
        if (g.outT && row_ok) store_colT<SPLIT>(g.outT, g.ldT, g.lo_T, nb, g.N, m, o0);
    } else if (g.ep == TEP_MULADD) {
        float dv[NC], ax[NC], o0[NC];
        if (row_ok) { load_row<SPLIT>(g.aux1, row_off, nb, pitch, g.lo_o, ax); load_deriv<SPLIT>(g, row_off, nb, pitch, dv, pre); }
        else { zero(dv); zero(ax); }
#pragma unroll
        for (int j = 0; j < NC; ++j) o0[j] = (nb + j < g.N) ? fmaf(__uint_as_float(r[j]), dv[j], ax[j]) : 0.f;
        store_rows<SPLIT>(es, mapO0, g.out0, nb, o0);
        if (g.outT && row_ok) store_colT<SPLIT>(g.outT, g.ldT, g.lo_T, nb, g.N, m, o0);
    } else if (!row_ok) {
        return;
    } else if (g.ep == TEP_WGRAD) {
        // read-modify-write of this (tile, slice)'s region: all loads of the chunk in flight before the first store (a naive
        // "*p += acc" loop serialises dependent global round trips per chunk: measured 2x on the whole GEMM)
        float* base = g.out_f32 + (long long)sl * g.slice_stride + m;
        float old[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) old[j] = (nb + j < g.N) ? __ldcg(base + (long long)(nb + j) * g.ldw) : 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (nb + j < g.N) __stcg(base + (long long)(nb + j) * g.ldw, old[j] + __uint_as_float(r[j]));   // a warp covers 32 consecutive m: coalesced
    } else if (g.ep == TEP_TRACE) {
        float dv[NC];
        if (pre) raw_to_f32<SPLIT>(*pre, dv);
        else load_row<SPLIT>(g.aux, row_off, nb, pitch, g.lo_o, dv);
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (nb + j < g.N) rowsum = fmaf(__uint_as_float(r[j]), dv[j], rowsum);
    } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int n = nb + j;
            if (n < g.N && n < g.n_limit) {
                float v = __uint_as_float(r[j]);
                if (g.ep == TEP_LIN_SOA) v += sbc[j];
                g.out_f32[(size_t)n * g.M + m] = v;
            }
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
                   const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2,
                   const __grid_constant__ CUtensorMap mapO0, const __grid_constant__ CUtensorMap mapO1, TcArgs g) {
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
    uint8_t* sstage = smem + NSTAGE * STAGE_BYTES + 256 + 2 * 256 * 4;            // [epilogue warp][EPI_STAGE_BYTES]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb0 = (g.K + TBK - 1) / TBK;
    const int nkb1 = g.K2 > 0 ? (g.K2 + TBK - 1) / TBK : 0;
    const int ntn = (g.N + BN - 1) / BN;
    const int ntiles = ((g.M + TBM - 1) / TBM) * ntn;
    const int nsl = g.nslices > 1 ? g.nslices : 1;
    const int nwork = ntiles * nsl;

    if (warp == TMA_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        if (g.out0) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapO0) : "memory");
        if (g.out1) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapO1) : "memory");
        if (nkb1) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA2) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB2) : "memory");
        }
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4 * WQ); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == TMA_WARP) {
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
                        trace_event(g.trace, 0, 1000 + it);
                        mbar_expect_tx(&full[s], STAGE_BYTES);
                        tma_load_2d(st, ma, &full[s], kb * TBK, m0);
                        if constexpr (SPLIT) tma_load_2d(st + A_TILE_BYTES, ma, &full[s], lo_a + kb * TBK, m0);
                        tma_load_2d(st + NT_A * A_TILE_BYTES, mb, &full[s], kb * TBK, n0);
                        if constexpr (SPLIT) tma_load_2d(st + NT_A * A_TILE_BYTES + BTB, mb, &full[s], lo_b + kb * TBK, n0);
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
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
                        trace_event(g.trace, 8192, 2000 + it);
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
                        trace_event(g.trace, 8192, 3000 + it);
                    }
                }
                tc_commit(&tmem_full[as]);
            }
        }
    } else {
        // ===== epilogue: warps 0 .. 4 WQ - 1, TMEM lane quarter = warp % 4, WQ warps per quarter share the tile's columns =====
        const int q = warp & 3;
        const int cq = warp >> 2;              // which of the WQ warps of this TMEM lane quarter
        EpiStore es{sstage + warp * EPI_STAGE_BYTES, lane, 0, g.lo_o};
        const int pitch = SPLIT ? g.lo_o : g.ldo;   // columns of one half
        int i = 0;
        for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++i) {
            const int tile = work / nsl, sl = work - tile * nsl;
            const int as = i & 1;
            const int m0 = (tile / ntn) * TBM, n0 = (tile % ntn) * BN;
            float* sb = sbias + as * BN;
            {   // stage this tile's bias slice (double-buffered: item i - 1 may still be read by a slower warp)
                const int e = threadIdx.x;         // 0 .. EPI_THREADS - 1
                for (int c = e; c < BN; c += EPI_THREADS) sb[c] = (g.bias && n0 + c < g.N) ? __ldg(g.bias + n0 + c) : 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }
            mbar_wait(&tmem_full[as], (i >> 1) & 1);
            tc_fence_after();
            if (threadIdx.x == 0) trace_event(g.trace, 16384, 4000 + i);
            const int m = m0 + q * 32 + lane;
            const bool row_ok = m < g.M;
            es.m_warp = m0 + q * 32;
            float rowsum = 0.f;
            // chunks of NC columns, double-buffered in registers: the TMEM load of chunk c + 1 is in flight while chunk c
            // is processed (tcgen05.wait::ld waits for every outstanding load, so it sits after the processing)
            const int cbeg = chunk_begin(BN, cq) * NC, nch = chunk_begin(BN, cq + 1) - chunk_begin(BN, cq);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            uint32_t ra[NC], rb[NC];
            if (n0 + cbeg < g.N) {
                tmem_ld16(trow + (uint32_t)cbeg, ra);
                tmem_wait_ld();
            }
            for (int ci = 0; ci < nch; ci += 2) {   // two chunks per trip: the register buffers alternate without dynamic indexing
                const int c0 = cbeg + ci * NC;
                if (n0 + c0 >= g.N) break;   // warp-uniform
                const bool more1 = (ci + 1 < nch) && (n0 + c0 + NC < g.N);
                if (more1) tmem_ld16(trow + (uint32_t)(c0 + NC), rb);
                if (threadIdx.x == 0) trace_event(g.trace, 16384, 6000 + ci);
                tc_epilogue<SPLIT>(g, &mapO0, &mapO1, es, ra, row_ok, m, n0 + c0, sb + c0, sl, pitch, rowsum);
                if (threadIdx.x == 0) trace_event(g.trace, 16384, 7000 + ci);
                if (!more1) break;
                tmem_wait_ld();
                const bool more2 = (ci + 2 < nch) && (n0 + c0 + 2 * NC < g.N);
                if (more2) tmem_ld16(trow + (uint32_t)(c0 + 2 * NC), ra);
                if (threadIdx.x == 0) trace_event(g.trace, 16384, 6000 + ci + 1);
                tc_epilogue<SPLIT>(g, &mapO0, &mapO1, es, rb, row_ok, m, n0 + c0 + NC, sb + c0 + NC, sl, pitch, rowsum);
                if (threadIdx.x == 0) trace_event(g.trace, 16384, 7000 + ci + 1);
                if (more2) tmem_wait_ld();
            }
            // this warp's TMEM reads of the tile are complete (tcgen05.wait::ld above): hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            if (threadIdx.x == 0) trace_event(g.trace, 16384, 5000 + i);
            // exact trace: one part per (unit tile, column quarter), added in part order by the reader (no atomics)
            if (g.ep == TEP_TRACE && row_ok) g.out_f32[(size_t)((tile % ntn) * WQ + cq) * g.M + m] = rowsum;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's bulk stores have completed
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster on neighbouring SMs compute one 256 x 256 tile.  Each CTA
// loads ITS 128 rows of A and ITS 128 rows of B per K block (the same 64 KB per stage as the 128 x 128 kernel), the
// leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256), which reads both CTAs' shared memory and writes
// 128 accumulator rows x 256 columns into EACH CTA's TMEM; both CTAs run the epilogue on their own rows.  Operand
// bytes pulled through L2 per MAC are half those of the 128 x 128 kernel -- these GEMMs are bound by the SMs' share
// of L2 bandwidth (213 MB of tile loads per 8192 x 512 x 785 layer), not by the tensor pipe.
// Barriers: full[s] lives in the leader (both CTAs' TMA loads complete_tx on it); empty[s] and tmem_full[a] are
// per CTA and are signalled by one multicast tcgen05.commit; tmem_empty[a] lives in the leader and collects the
// 32 epilogue warps of the pair.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same location in the pair's even (leader) CTA
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(leader_bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all prior MMAs have completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

constexpr int PBN = 256;   // unit columns of a pair tile (= TMEM columns per accumulator in each CTA)

template <bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TTHREADS, 1)
    tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                        const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2,
                        const __grid_constant__ CUtensorMap mapO0, const __grid_constant__ CUtensorMap mapO1, TcArgs g) {
    if (g.done && *g.done) return;   // uniform over the grid
    constexpr int NT_A = SPLIT ? 2 : 1;
    constexpr int NSTAGE = stages(SPLIT, 128);
    constexpr int BTB = b_tile_bytes(128);
    constexpr int STAGE_BYTES = stage_bytes(SPLIT, 128);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* tmem_full = empty + NSTAGE;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;      // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* sbias = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES + 256);   // [2][PBN]
    uint8_t* sstage = smem + NSTAGE * STAGE_BYTES + 256 + 2 * 256 * 4;            // [epilogue warp][EPI_STAGE_BYTES]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int nkb0 = (g.K + TBK - 1) / TBK;
    const int nkb1 = g.K2 > 0 ? (g.K2 + TBK - 1) / TBK : 0;
    const int ntn = (g.N + PBN - 1) / PBN;
    const int ntiles = ((g.M + 2 * TBM - 1) / (2 * TBM)) * ntn;
    const int nsl = g.nslices > 1 ? g.nslices : 1;
    const int nwork = ntiles * nsl;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == TMA_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        if (g.out0) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapO0) : "memory");
        if (g.out1) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapO1) : "memory");
        if (nkb1) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA2) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB2) : "memory");
        }
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8 * WQ); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(2 * PBN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();     // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == TMA_WARP) {
        // ===== TMA producer (both CTAs: each loads its own halves) =====
        if (lane == 0) {
            int it = 0;
            for (int work = pair; work < nwork; work += npairs) {
                const int tile = work / nsl, sl = work - tile * nsl;
                const int m0 = (tile / ntn) * (2 * TBM) + (int)rank * TBM;
                const int n0 = (tile % ntn) * PBN;
                const int nl = n0 + (int)rank * 128;          // this CTA's half of the B rows
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
                        trace_event(g.trace, 0, 1000 + it);
                        if (leader) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
                        tma_load_2d_pair(st, ma, &full[s], kb * TBK, m0);
                        if constexpr (SPLIT) tma_load_2d_pair(st + A_TILE_BYTES, ma, &full[s], lo_a + kb * TBK, m0);
                        tma_load_2d_pair(st + NT_A * A_TILE_BYTES, mb, &full[s], kb * TBK, nl);
                        if constexpr (SPLIT) tma_load_2d_pair(st + NT_A * A_TILE_BYTES + BTB, mb, &full[s], lo_b + kb * TBK, nl);
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(2 * TBM, PBN);
            int it = 0, i = 0;
            for (int work = pair; work < nwork; work += npairs, ++i) {
                const int tile = work / nsl, sl = work - tile * nsl;
                const int n0 = (tile % ntn) * PBN;
                const int nseg = (nkb1 && n0 < g.N2) ? 2 : 1;
                const int as = i & 1;
                mbar_wait(&tmem_empty[as], ((i >> 1) & 1) ^ 1);   // all 32 epilogue warps of the pair have drained it
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * PBN);
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
                        trace_event(g.trace, 8192, 2000 + it);
                        const uint64_t a_hi = make_desc(smem_u32(st));
                        const uint64_t b_hi = make_desc(smem_u32(st + NT_A * A_TILE_BYTES));
#pragma unroll
                        for (int k = 0; k < TBK / 16; ++k) tc_mma_bf16_pair(tacc, a_hi + 2 * k, b_hi + 2 * k, idesc, started | (uint32_t)k);
                        started = 1;
                        if constexpr (SPLIT) {
                            const uint64_t a_lo = make_desc(smem_u32(st + A_TILE_BYTES));
                            const uint64_t b_lo = make_desc(smem_u32(st + NT_A * A_TILE_BYTES + BTB));
#pragma unroll
                            for (int k = 0; k < TBK / 16; ++k) tc_mma_bf16_pair(tacc, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
                            for (int k = 0; k < TBK / 16; ++k) tc_mma_bf16_pair(tacc, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                        }
                        tc_commit_pair(&empty[s]);     // frees slot s in both CTAs
                        trace_event(g.trace, 8192, 3000 + it);
                    }
                }
                tc_commit_pair(&tmem_full[as]);        // wakes the epilogue warps of both CTAs
            }
        }
    } else {
        // ===== epilogue: warps 0 .. 4 WQ - 1 of both CTAs, each CTA on its own 128 accumulator rows =====
        const int q = warp & 3;
        const int cq = warp >> 2;              // which of the WQ warps of this TMEM lane quarter
        EpiStore es{sstage + warp * EPI_STAGE_BYTES, lane, 0, g.lo_o};
        const int pitch = SPLIT ? g.lo_o : g.ldo;
        int i = 0;
        for (int work = pair; work < nwork; work += npairs, ++i) {
            const int tile = work / nsl, sl = work - tile * nsl;
            const int as = i & 1;
            const int m0 = (tile / ntn) * (2 * TBM) + (int)rank * TBM;
            const int n0 = (tile % ntn) * PBN;
            float* sb = sbias + as * PBN;
            {
                const int e = threadIdx.x;         // 0 .. EPI_THREADS - 1
                for (int c = e; c < PBN; c += EPI_THREADS) sb[c] = (g.bias && n0 + c < g.N) ? __ldg(g.bias + n0 + c) : 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }
            mbar_wait(&tmem_full[as], (i >> 1) & 1);
            tc_fence_after();
            if (threadIdx.x == 0) trace_event(g.trace, 16384, 4000 + i);
            const int m = m0 + q * 32 + lane;
            const bool row_ok = m < g.M;
            es.m_warp = m0 + q * 32;
            float rowsum = 0.f;
            // chunks of NC columns, double-buffered in registers: the TMEM load of chunk c + 1 is in flight while chunk c
            // is processed (tcgen05.wait::ld waits for every outstanding load, so it sits after the processing)
            const int cbeg = chunk_begin(PBN, cq) * NC, nch = chunk_begin(PBN, cq + 1) - chunk_begin(PBN, cq);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PBN);
            uint32_t ra[NC], rb[NC];
            if (n0 + cbeg < g.N) {
                tmem_ld16(trow + (uint32_t)cbeg, ra);
                tmem_wait_ld();
            }
            for (int ci = 0; ci < nch; ci += 2) {   // two chunks per trip: the register buffers alternate without dynamic indexing
                const int c0 = cbeg + ci * NC;
                if (n0 + c0 >= g.N) break;   // warp-uniform
                const bool more1 = (ci + 1 < nch) && (n0 + c0 + NC < g.N);
                if (more1) tmem_ld16(trow + (uint32_t)(c0 + NC), rb);
                tc_epilogue<SPLIT>(g, &mapO0, &mapO1, es, ra, row_ok, m, n0 + c0, sb + c0, sl, pitch, rowsum);
                if (!more1) break;
                tmem_wait_ld();
                const bool more2 = (ci + 2 < nch) && (n0 + c0 + 2 * NC < g.N);
                if (more2) tmem_ld16(trow + (uint32_t)(c0 + 2 * NC), ra);
                tc_epilogue<SPLIT>(g, &mapO0, &mapO1, es, rb, row_ok, m, n0 + c0 + NC, sb + c0 + NC, sl, pitch, rowsum);
                if (more2) tmem_wait_ld();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
            if (threadIdx.x == 0) trace_event(g.trace, 16384, 5000 + i);
            if (g.ep == TEP_TRACE && row_ok) g.out_f32[(size_t)((tile % ntn) * WQ + cq) * g.M + m] = rowsum;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's bulk stores have completed
    }
    tc_fence_before();
    cluster_sync_all();     // neither CTA may free its TMEM (or exit) while the pair's MMAs / remote arrivals are in flight
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * PBN));
    }
}

}  // namespace tc
}  // namespace icnf
