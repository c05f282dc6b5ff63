// tc_chain.cuh -- a chain of dependent tcgen05 GEMMs in ONE persistent kernel (see tc.h, gemm_chain).
//
// Same tile pipeline as tc_gemm_kernel<SPLIT, 128> (TMA ring -> tcgen05.mma into a double-buffered TMEM accumulator
// -> 12 epilogue warps), but the work list is the concatenation of the tiles of up to CHAIN_MAXG GEMMs whose
// parameters (tensor maps included) travel in the kernel parameter block.  Dependencies are device-side counters:
//   row counter (g, r)  += 1 per epilogue warp and item of GEMM g that has published its part of row tile r
//   all counter (g)     += 1 per epilogue warp and item of GEMM g
// A consumer's TMA producer spins on the counters of the GEMMs it reads (acquire), then issues its loads; the MMA and
// epilogue roles are ordered behind it through the CTA's mbarriers.  CTAs take items blockIdx.x, + gridDim.x, ... in
// list order and an item only waits on EARLIER items, so the chain cannot deadlock once the grid is resident (one CTA
// per SM).  Two counter sets alternate between launches: a launch zeroes the set of the next one.
#pragma once
#include "tc_gemm.cuh"

namespace icnf {
namespace tc {

struct alignas(128) ChainGemm {
    CUtensorMap mapA, mapB, mapA2, mapB2, mapO0, mapO1;
    CUtensorMap mapAs, mapA2s;   // clusters: the A operand in slices of 128 / cl rows (each CTA loads one slice and multicasts it)
    TcArgs g;
    int work_begin, nwork;   // this GEMM's slice of the work list
    int ntn, nsl;            // unit tiles per row tile, split-K slices
    int ntg;                 // groups of `cl` unit tiles per row tile = work items per (row tile, slice)
    int dep_row[2], dep_all[2];
    unsigned row_target[2], all_target[2];
};
struct ChainParams {
    int ngemm, nwork;
    unsigned* flags;        // [CHAIN_MAXG][row_stride] row counters, then [CHAIN_MAXG] whole-GEMM counters
    unsigned* flags_next;   // the other set: zeroed by this launch
    int row_stride, nflags;
    // row-tile super-groups (chains whose GEMMs all run over the same sample rows and have no whole-GEMM dependency): the
    // list is ordered [super-group][GEMM][tile], so that what a GEMM writes is read by the next one while it is still in
    // L2, with ~3 rounds of work between an item and the items it waits for.  rg = row tiles per super-group (0: off),
    // ntm = row tiles in total, row_items = sum over the GEMMs of (unit tiles x slices) per row tile
    int rg, ntm, row_items;
    // clusters of cl CTAs (1, 2 or 4) take the cl unit tiles of one group: they share the group's A tile, which every CTA loads
    // a 1 / cl slice of and multicasts to the others -- fewer operand bytes through L2 per MAC
    int cl;
    const int* done;
    long long* trace;
    int direct_stores;      // row-major outputs: 1 = 16-byte stores from registers (short chains: latency), 0 = bulk tensor stores (long chains: throughput)
    int nacc, nacc_log2;    // TMEM accumulators in rotation: 2 (default) or 4 (knob 4, ICNF_CHAIN_NACC)
    int dbg;                // development knob (ICNF_CHAIN_DBG): 1 = skip row stores, 2 = identity activation, 4 = skip publish fence, 8 = skip TMEM loads
    ChainGemm gm[CHAIN_MAXG];
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// (bounded: ~2^26 polls of >= 64 ns is several seconds, far beyond any chain; a launch that could not make progress then
// finishes with wrong numbers instead of hanging the device)
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
    long long spins = 0;
    while (ld_acquire_u32(p) < target) {
        if (++spins > (1LL << 26)) break;
        __nanosleep(64);
    }
}


// ---- lean epilogue of the chain kernel ---------------------------------------------------------------------------
// One instantiation per epilogue kind (switch per work item): the kinds no longer share one register allocation, and
// the row-major outputs are built as packed bf16 pairs.  The general tc_epilogue (tc_gemm.cuh) spilled ~1.5 KB per
// thread at 128 registers; with ~220 KB of the SM's L1/shared array carved out as shared memory those spills went to
// L2, and the 9 local loads of a 16-column chunk (~8000 cycles) made the epilogue of a tile twice as long as its main
// loop -- which, through the row-tile dependencies, set the pace of the whole chain.
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16(a, b);
    lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ float raw_lo16(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float raw_hi16(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// word jp (columns 2 jp, 2 jp + 1) of a RawRow half
__device__ __forceinline__ uint32_t raw_word(const uint4 (&x)[4], int idx) {
    const uint4& v = x[idx >> 2];
    return (idx & 3) == 0 ? v.x : (idx & 3) == 1 ? v.y : (idx & 3) == 2 ? v.z : v.w;
}
template <bool SPLIT>
__device__ __forceinline__ void raw_pair(const RawRow& r, int jp, float& a, float& b) {
    const uint32_t hw = raw_word(r.x, jp);
    a = raw_lo16(hw); b = raw_hi16(hw);
    if constexpr (SPLIT) { const uint32_t lw = raw_word(r.x, 8 + jp); a += raw_lo16(lw); b += raw_hi16(lw); }
}
// One row's NC columns as packed pairs, through the warp's staging box and a bulk tensor store (warp-collective: every
// lane calls it; rows beyond M are clipped by the tensor map).  Measured against 16-byte stores straight from the
// registers (one half-filled sector per lane and instruction): the bulk stores are 20 % faster on the large-batch shapes.
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
struct RowDst {            // where a row-major output row of this lane goes
    const CUtensorMap* map; uint32_t stage; int lo_o, m_warp;        // bulk tensor store
    __nv_bfloat16* base; long long row_off; bool direct, row_ok;     // direct stores
};
template <bool SPLIT>
__device__ __forceinline__ void store_row_packed(const RowDst& d, int nb, const uint32_t (&hi)[NC / 2], const uint32_t (&lo)[NC / 2]) {
    if (d.direct) {   // warp-uniform
        if (d.row_ok) {
            uint4* hrow = reinterpret_cast<uint4*>(d.base + d.row_off + nb);
            __stcg(hrow, make_uint4(hi[0], hi[1], hi[2], hi[3]));
            __stcg(hrow + 1, make_uint4(hi[4], hi[5], hi[6], hi[7]));
            if constexpr (SPLIT) {
                uint4* lrow = reinterpret_cast<uint4*>(d.base + d.row_off + d.lo_o + nb);
                __stcg(lrow, make_uint4(lo[0], lo[1], lo[2], lo[3]));
                __stcg(lrow + 1, make_uint4(lo[4], lo[5], lo[6], lo[7]));
            }
        }
        return;
    }
    const CUtensorMap* map = d.map;
    const uint32_t stage = d.stage;
    const int lo_o = d.lo_o, m_warp = d.m_warp;
    const int lane = threadIdx.x & 31;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous boxes have left the staging buffer
    __syncwarp();
    const uint32_t row = stage + (uint32_t)lane * (NC * 2);
    sts_v4(row, hi[0], hi[1], hi[2], hi[3]);
    sts_v4(row + 16, hi[4], hi[5], hi[6], hi[7]);
    if constexpr (SPLIT) {
        sts_v4(row + 32 * NC * 2, lo[0], lo[1], lo[2], lo[3]);
        sts_v4(row + 32 * NC * 2 + 16, lo[4], lo[5], lo[6], lo[7]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(stage), "r"(nb), "r"(m_warp) : "memory");
        if constexpr (SPLIT)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(stage + 32 * NC * 2), "r"(lo_o + nb), "r"(m_warp) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
// transposed copy outT[(nb + j) * ldT + m] (a warp = 32 consecutive m: 64 contiguous bytes per column)
template <bool SPLIT>
__device__ __forceinline__ void store_colT_packed(__nv_bfloat16* outT, long long ldT, int lo_T, int nb, int nvalid, int m,
                                                  const uint32_t (&hi)[NC / 2], const uint32_t (&lo)[NC / 2]) {
    unsigned short* base = reinterpret_cast<unsigned short*>(outT) + (long long)nb * ldT + m;
#pragma unroll
    for (int jp = 0; jp < NC / 2; ++jp) {
        if (2 * jp < nvalid) {
            base[(long long)(2 * jp) * ldT] = (unsigned short)(hi[jp] & 0xffffu);
            if constexpr (SPLIT) base[(long long)(2 * jp) * ldT + lo_T] = (unsigned short)(lo[jp] & 0xffffu);
        }
        if (2 * jp + 1 < nvalid) {
            base[(long long)(2 * jp + 1) * ldT] = (unsigned short)(hi[jp] >> 16);
            if constexpr (SPLIT) base[(long long)(2 * jp + 1) * ldT + lo_T] = (unsigned short)(lo[jp] >> 16);
        }
    }
}
__device__ __forceinline__ float deriv_of(int act, bool is_h, float x) {   // x = sigma' itself, or h
    if (!is_h) return x;
    if (act == ICNF_ACT_SOFTPLUS) return 1.0f - __expf(-x);
    return act_deriv_from_h(act, x);
}

// What the epilogue of one work item needs, per warp, in SHARED memory: the epilogue kinds are real functions (own
// register allocation each), and neither their arguments nor the GEMM's parameter block may travel through local memory
// or generic loads -- with ~220 KB of the SM's L1/shared array carved out as shared memory, local memory lives in L2.
struct EpiShared {
    __nv_bfloat16 *out0, *out1, *outT;
    const __nv_bfloat16 *aux, *aux1, *aux2;
    float* out_f32;
    long long ldT, slice_stride;
    int N, M, act, aux_is_h, lo_o, ldo, lo_T, n_limit, ldw, pitch;
    int m_base, nb0, nch, sl;
    uint32_t trow, sb_addr;
    int dbg;
    const CUtensorMap *map0, *map1;   // row-major outputs (parameter space)
    int direct;
    uint32_t stage_addr;              // this warp's staging boxes: [32 rows][NC] bf16 hi, then lo
};
__shared__ EpiShared g_epi[4 * WQ];
// bias of each epilogue warp's own columns (a private slice per warp, written and read by that warp only: no cross-warp
// traffic, no double buffering, nothing for racecheck to flag)
__shared__ float g_bias[4 * WQ][((128 / NC + WQ - 1) / WQ) * NC];

struct EpiItem {          // registers (rebuilt from g_epi inside each epilogue function)
    const EpiShared* g;
    int m, nb0, nch, sl, pitch;
    bool row_ok;
    long long row_off;
};
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// chunk kinds: EP = TcEpilogue; F = ACT: also write sigma'
template <bool SPLIT, int EP, bool F>
__device__ __forceinline__ void epi_chunk(const EpiItem& it, const uint32_t (&r)[NC], int nb, int nvalid, uint32_t sbc,
                                          const RawRow* xr, float& rowsum) {
    const EpiShared& g = *it.g;
    uint32_t o_hi[NC / 2], o_lo[NC / 2];
    if constexpr (EP == TEP_ACT) {
        uint32_t d_hi[NC / 2], d_lo[NC / 2];
        const bool sp = g.act == ICNF_ACT_SOFTPLUS;
#pragma unroll
        for (int jp = 0; jp < NC / 2; ++jp) {
            float h0, h1, d0, d1;
            const float a0 = __uint_as_float(r[2 * jp]) + lds_f32(sbc + 8 * jp), a1 = __uint_as_float(r[2 * jp + 1]) + lds_f32(sbc + 8 * jp + 4);
            if (g.dbg & 2) { h0 = a0; h1 = a1; d0 = 1.f; d1 = 1.f; }
            else if (sp) { act_eval<ICNF_ACT_SOFTPLUS>(a0, h0, d0); act_eval<ICNF_ACT_SOFTPLUS>(a1, h1, d1); }
            else { act_eval_rt(g.act, a0, h0, d0); act_eval_rt(g.act, a1, h1, d1); }
            if (2 * jp >= nvalid) { h0 = 0.f; d0 = 0.f; }
            if (2 * jp + 1 >= nvalid) { h1 = 0.f; d1 = 0.f; }
            split_pack2(h0, h1, o_hi[jp], o_lo[jp]);
            if constexpr (F) split_pack2(d0, d1, d_hi[jp], d_lo[jp]);
        }
        if (!(g.dbg & 1)) {
            store_row_packed<SPLIT>(RowDst{g.map0, g.stage_addr, g.lo_o, g.m_base, g.out0, it.row_off, g.direct != 0, it.row_ok}, nb, o_hi, o_lo);
            if constexpr (F) store_row_packed<SPLIT>(RowDst{g.map1, g.stage_addr, g.lo_o, g.m_base, g.out1, it.row_off, g.direct != 0, it.row_ok}, nb, d_hi, d_lo);
            if (g.outT && it.row_ok) store_colT_packed<SPLIT>(g.outT, g.ldT, g.lo_T, nb, nvalid, it.m, o_hi, o_lo);
        }
    } else if constexpr (EP == TEP_MULD) {
#pragma unroll
        for (int jp = 0; jp < NC / 2; ++jp) {
            float x0, x1;
            raw_pair<SPLIT>(*xr, jp, x0, x1);
            float v0 = __uint_as_float(r[2 * jp]) * deriv_of(g.act, g.aux_is_h != 0, x0);
            float v1 = __uint_as_float(r[2 * jp + 1]) * deriv_of(g.act, g.aux_is_h != 0, x1);
            if (2 * jp >= nvalid) v0 = 0.f;
            if (2 * jp + 1 >= nvalid) v1 = 0.f;
            split_pack2(v0, v1, o_hi[jp], o_lo[jp]);
        }
        store_row_packed<SPLIT>(RowDst{g.map0, g.stage_addr, g.lo_o, g.m_base, g.out0, it.row_off, g.direct != 0, it.row_ok}, nb, o_hi, o_lo);
        if (g.outT && it.row_ok) store_colT_packed<SPLIT>(g.outT, g.ldT, g.lo_T, nb, nvalid, it.m, o_hi, o_lo);
    } else if constexpr (EP == TEP_TANGENT) {
        RawRow gr, hr;
        const bool need_h = g.act != ICNF_ACT_SOFTPLUS && !g.aux_is_h;
        if (it.row_ok) {
            load_row_raw<SPLIT>(g.aux1, (size_t)it.row_off, nb, it.pitch, g.lo_o, gr);
            if (need_h) load_row_raw<SPLIT>(g.aux2, (size_t)it.row_off, nb, it.pitch, g.lo_o, hr);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) { gr.x[q] = make_uint4(0u, 0u, 0u, 0u); hr.x[q] = make_uint4(0u, 0u, 0u, 0u); }
        }
        uint32_t p_hi[NC / 2], p_lo[NC / 2];
#pragma unroll
        for (int jp = 0; jp < NC / 2; ++jp) {
            float x0, x1, g0, g1, h0 = 0.f, h1 = 0.f;
            raw_pair<SPLIT>(*xr, jp, x0, x1);
            raw_pair<SPLIT>(gr, jp, g0, g1);
            if (g.aux_is_h) { h0 = x0; h1 = x1; }
            else if (need_h) raw_pair<SPLIT>(hr, jp, h0, h1);
            const float d0 = deriv_of(g.act, g.aux_is_h != 0, x0), d1 = deriv_of(g.act, g.aux_is_h != 0, x1);
            const float r0 = (2 * jp < nvalid) ? __uint_as_float(r[2 * jp]) : 0.f;
            const float r1 = (2 * jp + 1 < nvalid) ? __uint_as_float(r[2 * jp + 1]) : 0.f;
            float f0, f1;
            if (g.act == ICNF_ACT_SOFTPLUS) { f0 = 1.0f - d0; f1 = 1.0f - d1; }
            else { f0 = act_ratio_rt(g.act, h0, d0); f1 = act_ratio_rt(g.act, h1, d1); }
            split_pack2(r0 * d0, r1 * d1, o_hi[jp], o_lo[jp]);
            split_pack2(r0 * g0 * f0, r1 * g1 * f1, p_hi[jp], p_lo[jp]);
        }
        store_row_packed<SPLIT>(RowDst{g.map0, g.stage_addr, g.lo_o, g.m_base, g.out0, it.row_off, g.direct != 0, it.row_ok}, nb, o_hi, o_lo);
        store_row_packed<SPLIT>(RowDst{g.map1, g.stage_addr, g.lo_o, g.m_base, g.out1, it.row_off, g.direct != 0, it.row_ok}, nb, p_hi, p_lo);
        if (g.outT && it.row_ok) store_colT_packed<SPLIT>(g.outT, g.ldT, g.lo_T, nb, nvalid, it.m, o_hi, o_lo);
    } else if constexpr (EP == TEP_MULADD) {
        RawRow ar;
        if (it.row_ok) load_row_raw<SPLIT>(g.aux1, (size_t)it.row_off, nb, it.pitch, g.lo_o, ar);
        else {
#pragma unroll
            for (int q = 0; q < 4; ++q) ar.x[q] = make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int jp = 0; jp < NC / 2; ++jp) {
            float x0, x1, a0, a1;
            raw_pair<SPLIT>(*xr, jp, x0, x1);
            raw_pair<SPLIT>(ar, jp, a0, a1);
            float v0 = fmaf(__uint_as_float(r[2 * jp]), deriv_of(g.act, g.aux_is_h != 0, x0), a0);
            float v1 = fmaf(__uint_as_float(r[2 * jp + 1]), deriv_of(g.act, g.aux_is_h != 0, x1), a1);
            if (2 * jp >= nvalid) v0 = 0.f;
            if (2 * jp + 1 >= nvalid) v1 = 0.f;
            split_pack2(v0, v1, o_hi[jp], o_lo[jp]);
        }
        store_row_packed<SPLIT>(RowDst{g.map0, g.stage_addr, g.lo_o, g.m_base, g.out0, it.row_off, g.direct != 0, it.row_ok}, nb, o_hi, o_lo);
        if (g.outT && it.row_ok) store_colT_packed<SPLIT>(g.outT, g.ldT, g.lo_T, nb, nvalid, it.m, o_hi, o_lo);
    } else if constexpr (EP == TEP_TRACE) {
        if (!it.row_ok) return;
#pragma unroll
        for (int jp = 0; jp < NC / 2; ++jp) {
            float x0, x1;
            raw_pair<SPLIT>(*xr, jp, x0, x1);
            if (2 * jp < nvalid) rowsum = fmaf(__uint_as_float(r[2 * jp]), x0, rowsum);
            if (2 * jp + 1 < nvalid) rowsum = fmaf(__uint_as_float(r[2 * jp + 1]), x1, rowsum);
        }
    } else if constexpr (EP == TEP_WGRAD) {
        if (!it.row_ok) return;
        float* base = g.out_f32 + (long long)it.sl * g.slice_stride + it.m;
        float old[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) old[j] = (j < nvalid) ? __ldcg(base + (long long)(nb + j) * g.ldw) : 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (j < nvalid) __stcg(base + (long long)(nb + j) * g.ldw, old[j] + __uint_as_float(r[j]));
    } else {   // TEP_LIN_SOA / TEP_PLAIN_SOA
        if (!it.row_ok) return;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int n = nb + j;
            if (j < nvalid && n < g.n_limit) {
                float v = __uint_as_float(r[j]);
                if constexpr (EP == TEP_LIN_SOA) v += lds_f32(sbc + 4 * j);
                g.out_f32[(size_t)n * g.M + it.m] = v;
            }
        }
    }
}

// all chunks of one work item for this warp.  A real function per epilogue kind (not inlined): each kind gets its own
// register allocation instead of one allocation for the union of all kinds inside the persistent kernel body.
template <bool SPLIT, int EP, bool F>
__device__ __noinline__ float epi_item(int warp) {
    const EpiShared& g = g_epi[warp];
    constexpr bool AUX = EP == TEP_MULD || EP == TEP_TANGENT || EP == TEP_MULADD || EP == TEP_TRACE;
    EpiItem it;
    it.g = &g;
    it.m = g.m_base + (int)(threadIdx.x & 31);
    it.nb0 = g.nb0; it.nch = g.nch; it.sl = g.sl; it.pitch = g.pitch;
    it.row_ok = it.m < g.M;
    it.row_off = (long long)it.m * g.ldo;
    const int N = g.N;
    const uint32_t trow = g.trow, sb_addr = g.sb_addr;
    float rowsum = 0.f;
#pragma unroll 1
    for (int ci = 0; ci < it.nch; ++ci) {
        const int nb = it.nb0 + ci * NC;
        if (nb >= N) break;   // warp-uniform
        uint32_t r[NC];
        RawRow xr;
        if (!(g.dbg & 8)) tmem_ld16(trow + (uint32_t)(ci * NC), r);
        else {
#pragma unroll
            for (int j = 0; j < NC; ++j) r[j] = 0x3f800000u;
        }
        if constexpr (AUX) {   // in flight together with the accumulator load
            if (it.row_ok) load_row_raw<SPLIT>(g.aux, (size_t)it.row_off, nb, it.pitch, g.lo_o, xr);
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) xr.x[q] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        tmem_wait_ld();
        epi_chunk<SPLIT, EP, F>(it, r, nb, min(NC, N - nb), sb_addr + (uint32_t)(ci * NC * 4), &xr, rowsum);
    }
    return rowsum;
}

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// arrive on the barrier at this offset in every CTA of the cluster once all prior MMAs of THIS CTA have completed
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
struct WorkItem { int gi, mt, nt, sl; };
__device__ __forceinline__ WorkItem decode_work(const ChainParams& P, int work, int& gi_hint) {
    WorkItem w;
    if (P.rg == 0) {
        while (work >= P.gm[gi_hint].work_begin + P.gm[gi_hint].nwork) ++gi_hint;
        const ChainGemm& cg = P.gm[gi_hint];
        const int local = work - cg.work_begin;
        const int tile = local / cg.nsl;
        w.gi = gi_hint; w.sl = local - tile * cg.nsl; w.mt = tile / cg.ntg; w.nt = tile - w.mt * cg.ntg;
        return w;
    }
    const int per_sg = P.rg * P.row_items;
    const int sg = work / per_sg;
    int rem = work - sg * per_sg;
    const int rows = min(P.rg, P.ntm - sg * P.rg);
    int g = 0;
    for (; g < P.ngemm - 1; ++g) {
        const int cnt = rows * P.gm[g].ntg * P.gm[g].nsl;
        if (rem < cnt) break;
        rem -= cnt;
    }
    const ChainGemm& cg = P.gm[g];
    const int tile = rem / cg.nsl;
    w.gi = g; w.sl = rem - tile * cg.nsl;
    const int r = tile / cg.ntg;
    w.mt = sg * P.rg + r; w.nt = tile - r * cg.ntg;
    return w;
}

template <bool SPLIT>
__global__ void __launch_bounds__(TTHREADS, 1) tc_chain_kernel(const __grid_constant__ ChainParams P) {
    constexpr int BN = 128;
    constexpr int NT_A = SPLIT ? 2 : 1;
    constexpr int NSTAGE = stages(SPLIT, BN);
    constexpr int BTB = b_tile_bytes(BN);
    constexpr int STAGE_BYTES = stage_bytes(SPLIT, BN);
    // the next launch's counters (the previous launch, which used them, has completed: stream order)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.nflags; i += gridDim.x * blockDim.x) P.flags_next[i] = 0u;
    if (P.done && *P.done) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty = full + NSTAGE;
    // up to four accumulators in rotation (all 512 TMEM columns are allocated; one CTA per SM).  Two is the default: with
    // four the MMA thread runs further ahead, which measured +0 % on the main-loop-bound config 4 and -1 % on the
    // epilogue-bound config 5 (scripts/ab_chain.py, knob 4 / ICNF_CHAIN_NACC)
    constexpr int NACC = 4;
    uint64_t* tmem_full = empty + NSTAGE;         // [NACC]
    uint64_t* tmem_empty = tmem_full + NACC;      // [NACC]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + NACC);
    uint8_t* sstage = smem + NSTAGE * STAGE_BYTES + 256 + 2 * 256 * 4;            // [epilogue warp][EPI_STAGE_BYTES]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* all_flags = P.flags + CHAIN_MAXG * P.row_stride;

    if (warp == TMA_WARP && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)P.cl); }   // a slot is free when EVERY CTA of the cluster has consumed it
        for (int a = 0; a < NACC; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4 * WQ); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(NACC * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    if (P.cl > 1) cluster_sync_all();      // the peers' barriers are initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int cl = P.cl;
    const int crank = cl > 1 ? (int)cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << cl) - 1u);
    const int work0 = blockIdx.x / cl, wstride = gridDim.x / cl;     // a cluster walks the item list together
    const int arows = TBM / cl;                                      // rows of the A tile this CTA loads (and multicasts)

    if (warp == TMA_WARP) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0, gi = 0;
            for (int work = work0; work < P.nwork; work += wstride) {
                const WorkItem wi = decode_work(P, work, gi);
                const ChainGemm& cg = P.gm[wi.gi];
                const TcArgs& g = cg.g;
                const int nsl = cg.nsl;
                const int sl = wi.sl, mt = wi.mt;
                const int m0 = mt * TBM, n0 = (wi.nt * cl + crank) * BN;      // (a unit tile beyond N: zero-filled loads, no epilogue)
                const int nkb0 = (g.K + TBK - 1) / TBK;
                const int nkb1 = g.K2 > 0 ? (g.K2 + TBK - 1) / TBK : 0;
                const int nseg = (nkb1 && wi.nt * cl * BN < g.N2) ? 2 : 1;   // cluster-uniform: the CTAs of a cluster step through the same K blocks
                // ---- dependencies: everything this item reads has been published
                bool waited = false;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (cg.dep_all[k] >= 0) { wait_counter(all_flags + cg.dep_all[k], cg.all_target[k]); waited = true; }
                    if (cg.dep_row[k] >= 0) { wait_counter(P.flags + cg.dep_row[k] * P.row_stride + mt, cg.row_target[k]); waited = true; }
                }
                if (waited) asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> async-proxy (TMA) reads
                for (int seg = 0; seg < nseg; ++seg) {
                    const CUtensorMap* ma = cl > 1 ? (seg ? &cg.mapA2s : &cg.mapAs) : (seg ? &cg.mapA2 : &cg.mapA);
                    const CUtensorMap* mb = seg ? &cg.mapB2 : &cg.mapB;
                    const int nkb = seg ? nkb1 : nkb0;
                    const int lo_a = seg ? g.lo_a2 : g.lo_a, lo_b = seg ? g.lo_b2 : g.lo_b;
                    const int kbe = (int)((long long)nkb * (sl + 1) / nsl);
                    for (int kb = (int)((long long)nkb * sl / nsl); kb < kbe; ++kb, ++it) {
                        const int s = it % NSTAGE;
                        const uint32_t ph = (it / NSTAGE) & 1;
                        uint8_t* st = smem + s * STAGE_BYTES;
                        mbar_wait(&empty[s], ph ^ 1);
                        trace_event(P.trace, 0, 1000 + (it % 1000));
                        mbar_expect_tx(&full[s], STAGE_BYTES);
                        if (cl > 1) {   // this CTA's slice of the A tile, to every CTA of the cluster
                            const int ao = crank * arows * TBK * 2;
                            tma_load_2d_mc(st + ao, ma, &full[s], kb * TBK, m0 + crank * arows, cmask);
                            if constexpr (SPLIT) tma_load_2d_mc(st + A_TILE_BYTES + ao, ma, &full[s], lo_a + kb * TBK, m0 + crank * arows, cmask);
                        } else {
                            tma_load_2d(st, ma, &full[s], kb * TBK, m0);
                            if constexpr (SPLIT) tma_load_2d(st + A_TILE_BYTES, ma, &full[s], lo_a + kb * TBK, m0);
                        }
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
            int it = 0, i = 0, gi = 0;
            for (int work = work0; work < P.nwork; work += wstride, ++i) {
                const WorkItem wi = decode_work(P, work, gi);
                const ChainGemm& cg = P.gm[wi.gi];
                const TcArgs& g = cg.g;
                const int nsl = cg.nsl;
                const int sl = wi.sl;
                const int nkb0 = (g.K + TBK - 1) / TBK;
                const int nkb1 = g.K2 > 0 ? (g.K2 + TBK - 1) / TBK : 0;
                const int nseg = (nkb1 && wi.nt * cl * BN < g.N2) ? 2 : 1;
                const int as = i & (P.nacc - 1);
                mbar_wait(&tmem_empty[as], ((i >> P.nacc_log2) & 1) ^ 1);   // the epilogue has drained this accumulator
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
                        for (int k = 0; k < TBK / 16; ++k) tc_mma_bf16(tacc, a_hi + 2 * k, b_hi + 2 * k, idesc, started | (uint32_t)k);
                        started = 1;
                        if constexpr (SPLIT) {
                            const uint64_t a_lo = make_desc(smem_u32(st + A_TILE_BYTES));
                            const uint64_t b_lo = make_desc(smem_u32(st + NT_A * A_TILE_BYTES + BTB));
#pragma unroll
                            for (int k = 0; k < TBK / 16; ++k) tc_mma_bf16(tacc, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
                            for (int k = 0; k < TBK / 16; ++k) tc_mma_bf16(tacc, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                        }
                        if (cl > 1) tc_commit_mc(&empty[s], cmask);
                        else tc_commit(&empty[s]);
                        trace_event(P.trace, 8192, 3000 + (it % 1000));
                    }
                }
                tc_commit(&tmem_full[as]);
            }
        }
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3;
        const int cq = warp >> 2;
        int i = 0, gi = 0;
        for (int work = work0; work < P.nwork; work += wstride, ++i) {
            const WorkItem wi = decode_work(P, work, gi);
            const ChainGemm& cg = P.gm[wi.gi];
            const TcArgs& g = cg.g;
            const int sl = wi.sl, mt = wi.mt;
            const int as = i & (P.nacc - 1);
            const int m0 = mt * TBM, nt = wi.nt * cl + crank, n0 = nt * BN;
            const int pitch = SPLIT ? g.lo_o : g.ldo;
            float* sb = g_bias[warp];
            const int cbeg = chunk_begin(BN, cq) * NC, nch = chunk_begin(BN, cq + 1) - chunk_begin(BN, cq);
            // every warp stages the bias of ITS columns in its own slice (no CTA-wide barrier: the 12 epilogue warps run
            // out of step); the loads are issued before the wait on the accumulator
            float bv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = cbeg + lane + 32 * u;
                bv[u] = (g.bias && lane + 32 * u < nch * NC && n0 + c < g.N) ? __ldg(g.bias + n0 + c) : 0.f;
            }
            mbar_wait(&tmem_full[as], (i >> P.nacc_log2) & 1);
            tc_fence_after();
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (lane + 32 * u < nch * NC) sb[lane + 32 * u] = bv[u];
            __syncwarp();
            if (threadIdx.x == 0) trace_event(P.trace, 16384, 4000 + (i % 1000));
            const int m = m0 + q * 32 + lane;
            const bool row_ok = m < g.M;
            if (lane == 0) {   // this item's parameters: parameter bank -> this warp's shared-memory block
                EpiShared& e = g_epi[warp];
                e.out0 = g.out0; e.out1 = g.out1; e.outT = g.outT; e.aux = g.aux; e.aux1 = g.aux1; e.aux2 = g.aux2; e.out_f32 = g.out_f32;
                e.ldT = g.ldT; e.slice_stride = g.slice_stride;
                e.N = g.N; e.M = g.M; e.act = g.act; e.aux_is_h = g.aux_is_h; e.lo_o = g.lo_o; e.ldo = g.ldo; e.lo_T = g.lo_T;
                e.n_limit = g.n_limit; e.ldw = g.ldw; e.pitch = pitch;
                e.m_base = m0 + q * 32; e.nb0 = n0 + cbeg; e.nch = nch; e.sl = sl;
                e.trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + cbeg);
                e.sb_addr = smem_u32(sb);
                e.dbg = P.dbg; e.direct = P.direct_stores;
                e.map0 = &cg.mapO0; e.map1 = &cg.mapO1; e.stage_addr = smem_u32(sstage + warp * EPI_STAGE_BYTES);
            }
            __syncwarp();
            float rowsum = 0.f;
            const int ep = g.ep;
            const bool with_d = g.out1 != nullptr;
            switch (ep) {   // warp-uniform
                case TEP_ACT: rowsum = with_d ? epi_item<SPLIT, TEP_ACT, true>(warp) : epi_item<SPLIT, TEP_ACT, false>(warp); break;
                case TEP_MULD: rowsum = epi_item<SPLIT, TEP_MULD, false>(warp); break;
                case TEP_TANGENT: rowsum = epi_item<SPLIT, TEP_TANGENT, false>(warp); break;
                case TEP_MULADD: rowsum = epi_item<SPLIT, TEP_MULADD, false>(warp); break;
                case TEP_TRACE: rowsum = epi_item<SPLIT, TEP_TRACE, false>(warp); break;
                case TEP_WGRAD: rowsum = epi_item<SPLIT, TEP_WGRAD, false>(warp); break;
                case TEP_LIN_SOA: rowsum = epi_item<SPLIT, TEP_LIN_SOA, false>(warp); break;
                default: rowsum = epi_item<SPLIT, TEP_PLAIN_SOA, false>(warp); break;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            const bool tile_valid = nt < cg.ntn;      // clusters: the last group of a row tile may hold fewer than cl unit tiles
            if (g.ep == TEP_TRACE && row_ok && tile_valid) g.out_f32[(size_t)(nt * WQ + cq) * g.M + m] = rowsum;
            // ---- publish: this warp's stores of the item are complete and visible device-wide, then count it
            // (one fence by one lane: the other lanes' stores are ordered before it by the warp barrier, and the fence is
            // cumulative; 32 lanes each running __threadfence + two releasing reductions cost a fifth of the kernel)
            __syncwarp();
            if (lane == 0 && tile_valid) {
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's bulk stores have completed
                asm volatile("fence.proxy.async;" ::: "memory");
                if (!(P.dbg & 4)) asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(P.flags + wi.gi * P.row_stride + mt), "r"(1u) : "memory");
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(all_flags + wi.gi), "r"(1u) : "memory");
            }
            if (threadIdx.x == 0) trace_event(P.trace, 16384, 5000 + (i % 1000));
        }
    }
    tc_fence_before();
    if (cl > 1) cluster_sync_all();     // no CTA leaves while a peer may still multicast into it or arrive on its barriers
    else __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NACC * BN));
    }
}

}  // namespace tc
}  // namespace icnf
