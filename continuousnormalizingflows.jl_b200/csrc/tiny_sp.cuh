// tiny_sp.cuh -- sample-parallel backward kernel of the tiny family.
//
// Two phases per Tsit5 stage, separated by CTA barriers:
//
//   main pass   one thread = one sample, exactly like the forward kernels: weights arrive as a
//               __grid_constant__ kernel parameter (uniform registers feed FFMA2), no shuffles,
//               no shared-memory weight reads.  The thread runs forward, the VJP chain, the
//               tangent pass and the backprop of the second-order reverse sweep and leaves the
//               vectors the weight gradient needs in shared memory, as rows of float2 pairs
//               indexed [pair-row][sample].
//   dW phase    the weight gradient is a sum of rank-1 updates over samples,
//                   dW_l = sum_b  abar_l[b] fin_l[b]' + g_l[b] w_{l-1}[b]',
//               i.e. a small GEMM whose K dimension is the sample index.  The CTA's threads
//               re-partition into (block, sample-group) pairs; a block is a 12 x 4 register tile
//               of one layer's dW (plus its bias column), accumulated in registers over all
//               stages, steps and samples with FFMA2 fed by 128-bit shared-memory loads, and
//               reduced across groups / CTAs only once, at the end of the kernel.
//
// The weights are passed TWICE (sw, sw2: same contents).  Forward / VJP chain read sw, tangent pass /
// backprop read sw2.  With a single copy the compiler merges the constant-bank loads of the two
// passes that touch the same weight, keeps ~200 weights live in vector registers between them and
// spills (1.4 KB of stack at 128 registers); with two copies every weight is a uniform-register
// operand with a short live range (226 -> 128 registers, no spills).
//
// Math: identical to tiny::rhs_reverse (derivation there and in DESIGN.md); stage inputs come
// from the forward solve's checkpoints (ckpt_index).  Reference: gradient of loss,
// src/core/icnf.jl:628-649 through the solve of src/core/base_icnf.jl:134-140 (Zygote +
// SciMLSensitivity in the reference, icnf.jl:90-99).
#pragma once
#include "tiny.cuh"

namespace icnf {
namespace tiny {

#ifndef ICNF_SP_MAXT
#define ICNF_SP_MAXT 224   // threads (= samples) per CTA, chosen per launch as a multiple of 32 up to this
#endif
#ifndef ICNF_SP_MINB
#define ICNF_SP_MINB 2
#endif
#ifndef ICNF_SP_MAXREG
#define ICNF_SP_MAXREG 128   // 136 and 144 registers measured slower (225 vs 193 us: the second CTA no longer fits the SM)
#endif

template <class N, bool EXACT>
struct SPCfg {
    static constexpr int NL = N::NL, NH = N::NL - 1, D = N::D;
    // pair-row counts: inputs of layer l plus the constant 1 that multiplies the bias; outputs of layer l;
    // tangent inputs (only inputs that carry a z-derivative)
    __host__ __device__ static constexpr int fin_p(int l) { return (N::n(l) + 2) / 2; }
    __host__ __device__ static constexpr int out_p(int l) { return (N::n(l + 1) + 1) / 2; }
    __host__ __device__ static constexpr int wt_p(int l) { return (N::kz(l) + 1) / 2; }
    __host__ __device__ static constexpr int sum_fin(int l) { int o = 0; for (int i = 0; i < l; ++i) o += fin_p(i); return o; }
    __host__ __device__ static constexpr int sum_out(int l) { int o = 0; for (int i = 0; i < l; ++i) o += out_p(i); return o; }
    __host__ __device__ static constexpr int sum_wt(int l) { int o = 0; for (int i = 0; i < l; ++i) o += wt_p(i); return o; }
    // record layout in pair-rows
    __host__ __device__ static constexpr int FIN(int l) { return sum_fin(l); }
    static constexpr int GG0 = sum_fin(NL);
    __host__ __device__ static constexpr int GG(int l) { return GG0 + sum_out(l); }
    static constexpr int AB0 = GG0 + sum_out(NL);
    __host__ __device__ static constexpr int AB(int l) { return AB0 + sum_out(l); }
    static constexpr int WT0 = AB0 + sum_out(NL);
    __host__ __device__ static constexpr int WT(int l) { return WT0 + sum_wt(l); }
    static constexpr int CC0 = WT0 + sum_wt(NL);
    __host__ __device__ static constexpr int CC(int l) { return EXACT ? CC0 + sum_out(l) : AB(l); }   // one probe: c lives where abar will
    static constexpr int KB0 = CC0 + (EXACT ? sum_out(NH) : 0);
    static constexpr int XN0 = KB0 + 3 * D;                  // stage cotangents: 6 D' float rows = 3 D' pair-rows
    static constexpr int NPR = XN0 + (D + 1) / 2;            // prefetched next stage input: D' floats per sample
    // ---- dW blocks: LCH L pair-rows x 2 R pair-rows.  "normal" orientation: L = layer outputs (abar / g),
    //      R = layer inputs (fin / w); "transposed": L = inputs, R = outputs (cheaper when n_out is tiny)
    static constexpr int LCH = 6;
    __host__ __device__ static constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
    __host__ __device__ static constexpr int nblk_n(int l) { return cdiv(out_p(l), LCH) * cdiv(fin_p(l), 2); }
    __host__ __device__ static constexpr int nblk_t(int l) { return cdiv(fin_p(l), LCH) * cdiv(out_p(l), 2); }
    __host__ __device__ static constexpr bool transposed(int l) { return nblk_t(l) < nblk_n(l); }
    __host__ __device__ static constexpr int nblk(int l) { return transposed(l) ? nblk_t(l) : nblk_n(l); }
    __host__ __device__ static constexpr int blkoff(int l) { int o = 0; for (int i = 0; i < l; ++i) o += nblk(i); return o; }
    static constexpr int NBLK = blkoff(NL);
    struct Blk { int Lb[2], Rb[2], nL[2], nR[2]; };
    __host__ __device__ static constexpr Blk desc(int blk) {
        Blk b{{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int l = 0; l < NL; ++l) {
            const int off = blkoff(l), nb = nblk(l);
            if (blk < off || blk >= off + nb) continue;
            const int i = blk - off;
            const int lo = out_p(l), fi = fin_p(l), wt = wt_p(l);
            if (!transposed(l)) {
                const int nrc = cdiv(fi, 2), lc = i / nrc, rc = i - lc * nrc;
                b.Lb[0] = AB(l) + lc * LCH; b.Lb[1] = GG(l) + lc * LCH;
                b.nL[0] = b.nL[1] = (lo - lc * LCH < LCH) ? lo - lc * LCH : LCH;
                b.Rb[0] = FIN(l) + 2 * rc; b.nR[0] = (fi - 2 * rc < 2) ? fi - 2 * rc : 2;
                b.Rb[1] = WT(l) + 2 * rc;
                const int w = (wt - 2 * rc < 2) ? wt - 2 * rc : 2;
                b.nR[1] = w > 0 ? w : 0;
            } else {
                const int nrc = cdiv(lo, 2), lc = i / nrc, rc = i - lc * nrc;
                b.Lb[0] = FIN(l) + lc * LCH; b.nL[0] = (fi - lc * LCH < LCH) ? fi - lc * LCH : LCH;
                b.Lb[1] = WT(l) + lc * LCH;
                const int w = (wt - lc * LCH < LCH) ? wt - lc * LCH : LCH;
                b.nL[1] = w > 0 ? w : 0;
                b.Rb[0] = AB(l) + 2 * rc; b.Rb[1] = GG(l) + 2 * rc;
                b.nR[0] = b.nR[1] = (lo - 2 * rc < 2) ? lo - 2 * rc : 2;
            }
        }
        if (b.nL[1] == 0) b.nR[1] = 0;
        if (b.nR[1] == 0) b.nL[1] = 0;
        return b;
    }
    // issue slots one sample pair costs a thread of this block (128-bit loads + FFMA2): the launcher hands
    // out the CTA's threads to the blocks in proportion, so that every warp leaves the dW phase together
    __host__ __device__ static constexpr int cost(int blk) {
        const Blk b = desc(blk);
        int c = 4;
        for (int t = 0; t < 2; ++t)
            if (b.nR[t] > 0) c += b.nL[t] + b.nR[t] + 4 * b.nL[t] * b.nR[t];
        return c;
    }
    // samples per pair-row = largest CTA size: compile time, so every record access has an immediate offset;
    // the largest multiple of 32 (<= ICNF_SP_MAXT) that leaves room for ICNF_SP_MINB CTAs per SM
    __host__ __device__ static constexpr int pitch() {
        int p = ICNF_SP_MAXT;
        while (p > 32 && (size_t)NPR * p * 8 > (size_t)(227 * 1024) / ICNF_SP_MINB - 1024) p -= 32;
        return p;
    }
    static constexpr int PITCH = pitch();
    __host__ __device__ static constexpr size_t smem_bytes(int max_groups) {
        size_t rec = (size_t)NPR * PITCH * 8;
        size_t red = (size_t)max_groups * N::NP * 4;
        return rec > red ? rec : red;
    }
};

// store a vector held as float2 pairs (n valid entries, optional constant 1 at index n) as pair-rows
template <int NS, int n, bool ONE, int NM>
__device__ __forceinline__ void sp_store(float2* rec, int row0, const float2 (&v)[NM]) {
    constexpr int np = ONE ? (n + 2) / 2 : (n + 1) / 2;
#pragma unroll
    for (int p = 0; p < np; ++p) {
        float2 o;
        o.x = (2 * p < n) ? v[(2 * p < n) ? p : 0].x : ((ONE && 2 * p == n) ? 1.f : 0.f);
        o.y = (2 * p + 1 < n) ? v[(2 * p + 1 < n) ? p : 0].y : ((ONE && 2 * p + 1 == n) ? 1.f : 0.f);
        rec[(row0 + p) * NS] = o;
    }
}

// thread ranges of the dW blocks inside a CTA: block b owns threads [first[b], first[b + 1])
struct SPPlan {
    unsigned short first[34];
    unsigned short max_groups;
};

template <class N, bool EXACT>
__global__ void __maxnreg__(ICNF_SP_MAXREG)
    backward_sp_kernel(const __grid_constant__ WBlock<N> sw, const __grid_constant__ WBlock<N> sw2,
                       const __grid_constant__ SPPlan plan, BackwardArgs a) {
    using C = SPCfg<N, EXACT>;
    constexpr int NL = N::NL, NH = NL - 1, D = N::D;
    extern __shared__ __align__(16) float smem[];
    constexpr int NS = C::PITCH;               // row pitch; the CTA has blockDim.x <= NS threads (= samples per tile)
    const int NT_ = blockDim.x, tid = threadIdx.x;
    float2* rec = reinterpret_cast<float2*>(smem) + tid;      // this thread's column; row r at rec[r * NS]
    float* kbm = smem + (size_t)C::KB0 * NS * 2 + tid;        // stage cotangents, float rows [6 D'][NS]
    float* xnm = smem + (size_t)C::XN0 * NS * 2 + tid * D;    // next stage input of this thread (cp.async target)
    const unsigned xn_sa = (unsigned)__cvta_generic_to_shared(xnm);
    // asynchronous global -> shared copy of one stage input: no register holds the value while it is in
    // flight (a register prefetch was spilled at once and stalled on its own load)
    auto prefetch_x = [&](const float* src) __attribute__((always_inline)) {
#pragma unroll
        for (int j = 0; j < D; ++j)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(xn_sa + 4u * j), "l"(src + j) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const float4* rec4 = reinterpret_cast<const float4*>(smem);
    // a failed forward solve has no gradient: skip the sweep, return NaN (see tiny::backward_kernel)
    const int nsteps = (a.stats->status == ICNF_OK) ? a.stats->naccept : 0;

    // ---- dW-phase role of this thread: block `blk`, sample group `grp` of NG
    static_assert(C::NBLK <= 33, "dW blocks per CTA");
    int blk = 0;
    while (blk < C::NBLK && tid >= (int)plan.first[blk + 1]) ++blk;
    const bool dw_active = blk < C::NBLK;
    const int NG = dw_active ? (int)plan.first[blk + 1] - (int)plan.first[blk] : 1;
    const int grp = dw_active ? tid - (int)plan.first[blk] : 0;
    // the block descriptor stays packed in two registers between stages (pair-row bases; valid counts, group)
    static_assert(C::NPR < 256 && ICNF_SP_MAXT <= 256, "descriptor packing");
    unsigned desc0, desc1;
    {
        const typename C::Blk bd = C::desc(dw_active ? blk : 0);
        desc0 = (unsigned)bd.Lb[0] | ((unsigned)bd.Lb[1] << 8) | ((unsigned)bd.Rb[0] << 16) | ((unsigned)bd.Rb[1] << 24);
        desc1 = (unsigned)bd.nL[0] | ((unsigned)bd.nL[1] << 4) | ((unsigned)bd.nR[0] << 8) | ((unsigned)bd.nR[1] << 12) |
                ((unsigned)grp << 16) | ((unsigned)(NG - 1) << 24);
        if (!dw_active) desc1 = 0u;   // no terms: nR = 0
    }
    constexpr int half = NS >> 1;   // float4 units per pair-row
    float2 acc[2 * C::LCH][2];
#pragma unroll
    for (int i = 0; i < 2 * C::LCH; ++i) { acc[i][0] = make_float2(0.f, 0.f); acc[i][1] = make_float2(0.f, 0.f); }

    // one L pair-row against the block's R pair-rows, two samples per 128-bit load
    auto dw_row = [&](const float4 l4, const float4 r0, const float4 r1, bool two, float2 (&a0)[2], float2 (&a1)[2])
                      __attribute__((always_inline)) {
        a0[0] = __ffma2_rn(bc2(l4.x), make_float2(r0.x, r0.y), a0[0]);
        a1[0] = __ffma2_rn(bc2(l4.y), make_float2(r0.x, r0.y), a1[0]);
        a0[0] = __ffma2_rn(bc2(l4.z), make_float2(r0.z, r0.w), a0[0]);
        a1[0] = __ffma2_rn(bc2(l4.w), make_float2(r0.z, r0.w), a1[0]);
        if (two) {
            a0[1] = __ffma2_rn(bc2(l4.x), make_float2(r1.x, r1.y), a0[1]);
            a1[1] = __ffma2_rn(bc2(l4.y), make_float2(r1.x, r1.y), a1[1]);
            a0[1] = __ffma2_rn(bc2(l4.z), make_float2(r1.z, r1.w), a0[1]);
            a1[1] = __ffma2_rn(bc2(l4.w), make_float2(r1.z, r1.w), a1[1]);
        }
    };
    auto dw_phase = [&](bool termA, int nduo) __attribute__((always_inline)) {
        const int Lb[2] = {(int)(desc0 & 255u), (int)((desc0 >> 8) & 255u)};
        const int Rb[2] = {(int)((desc0 >> 16) & 255u), (int)(desc0 >> 24)};
        const int nL[2] = {(int)(desc1 & 15u), (int)((desc1 >> 4) & 15u)};
        const int nR[2] = {(int)((desc1 >> 8) & 15u), (int)((desc1 >> 12) & 15u)};
        const int grp_ = (int)((desc1 >> 16) & 255u), NG_ = (int)(desc1 >> 24) + 1;
        if ((nR[0] | nR[1]) == 0) return;
        for (int d = grp_; d < nduo; d += NG_) {
#pragma unroll
            for (int term = 0; term < 2; ++term) {
                if (term == 0 && !termA) continue;
                if (nR[term] == 0) continue;
                const float4* lp = rec4 + Lb[term] * half + d;
                const float4* rp = rec4 + Rb[term] * half + d;
                const float4 r0 = rp[0];
                const bool two = nR[term] > 1;
                float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (two) r1 = rp[half];
                if (nL[term] == C::LCH) {
                    if (two) {
#pragma unroll
                        for (int i = 0; i < C::LCH; ++i) dw_row(lp[i * half], r0, r1, true, acc[2 * i], acc[2 * i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < C::LCH; ++i) dw_row(lp[i * half], r0, r1, false, acc[2 * i], acc[2 * i + 1]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < C::LCH; ++i)
                        if (i < nL[term]) dw_row(lp[i * half], r0, r1, two, acc[2 * i], acc[2 * i + 1]);
                }
            }
        }
    };

    const int64_t ntiles = (a.B + NT_ - 1) / NT_;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t braw = tile * NT_ + tid;
        const bool valid = braw < a.B;
        const int64_t b = valid ? braw : a.B - 1;
        const float wgt = valid ? a.inv_denominator : 0.f;
        const int nvalid = (int)min((int64_t)NT_, a.B - tile * NT_);
        const int nduo = (nvalid + 1) >> 1;

        float eps[D], x[N::n(0)];
        if (a.mode != ICNF_TEST) {
            if (a.eps_kind == ICNF_EPS_SUPPLIED) {
#pragma unroll
                for (int j = 0; j < D; ++j) eps[j] = __ldg(a.eps + b * D + j);
            } else {
#pragma unroll
                for (int q = 0; q < (D + 3) / 4; ++q) {
                    float o[4];
                    philox_draw4(a.eps_kind, a.seed, PHILOX_STREAM_EPS, a.sample_offset + b, q, o);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (q * 4 + r < D) eps[q * 4 + r] = o[r];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) eps[j] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < N::C; ++c) x[D + N::TIN + c] = __ldg(a.ys + b * N::C + c);

        float zbar[D];
        {
            const float* zf = a.ckpt + ckpt_index<N>(nsteps, a.B, b, 0);
            float za = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                zbar[j] = zf[j];
                if (j >= a.nvars) za = fmaf(zf[j], zf[j], za);
            }
            if (a.reg_a) {
                const float s = a.squared ? 2.0f * a.lam3 : (za > 0.f ? a.lam3 * rsqrtf(za) : 0.f);
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j >= a.nvars) zbar[j] = fmaf(s, zf[j], zbar[j]);
            }
#pragma unroll
            for (int j = 0; j < D; ++j) zbar[j] *= wgt;
        }
        if (nsteps > 0) prefetch_x(a.ckpt + ckpt_index<N>(nsteps - 1, a.B, b, 5));
        for (int step = nsteps - 1; step >= 0; --step) {
            const float t = a.steps[step].t, h = a.steps[step].dt;
            for (int i = 0; i < 6; ++i) {
                const float c = h * c_a[6][i];
#pragma unroll
                for (int j = 0; j < D; ++j) kbm[(i * D + j) * NS] = c * zbar[j];
            }
            for (int i = 5; i >= 0; --i) {
                float zb[D];
                asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    x[j] = xnm[j];
                    zb[j] = kbm[(i * D + j) * NS];
                }
                {   // prefetch the next stage input (previous stage, or stage 5 of the previous step): its
                    // latency hides behind this stage's work instead of stalling the whole CTA after the barrier
                    const int ni = i > 0 ? i - 1 : 5, nstep = i > 0 ? step : step - 1;
                    if (nstep >= 0) prefetch_x(a.ckpt + ckpt_index<N>(nstep, a.B, b, ni));
                }
                if constexpr (N::TIN) x[D] = fmaf(c_c[i], h, t);
                const float hb = h * c_a[6][i];
                const float cl = hb * wgt, cE = a.reg_e ? cl * a.lam1 : 0.f, cn = a.reg_n ? cl * a.lam2 : 0.f;

                // ---------------- main pass: forward
                Acts<N> A;
                forward<N>(sw, x, A);
                {
                    float2 xin[(N::n(0) + 1) / 2];
#pragma unroll
                    for (int p = 0; p < (N::n(0) + 1) / 2; ++p)
                        xin[p] = make_float2(x[2 * p], (2 * p + 1 < N::n(0)) ? x[(2 * p + 1 < N::n(0)) ? 2 * p + 1 : 0] : 0.f);
                    sp_store<NS, N::n(0), true>(rec, C::FIN(0), xin);
                }
                static_for<0, NH>([&](auto lc) __attribute__((always_inline)) {
                    constexpr int l = decltype(lc)::value;
                    sp_store<NS, N::n(l + 1), true>(rec, C::FIN(l + 1), A.h[l]);
                });
                if (!EXACT && cE != 0.f) {
                    float zz = 0.f;
#pragma unroll
                    for (int j = 0; j < D; ++j) zz = fmaf(EL(A.h[NL - 1], j), EL(A.h[NL - 1], j), zz);
                    const float s = a.squared ? 2.0f * cE : (zz > 0.f ? cE * rsqrtf(zz) : 0.f);
#pragma unroll
                    for (int j = 0; j < D; ++j) zb[j] = fmaf(s, EL(A.h[NL - 1], j), zb[j]);
                }
                float2 ab[N::NP2];
#pragma unroll
                for (int c = 0; c < (D + 1) / 2; ++c) ab[c] = make_float2(zb[2 * c], (2 * c + 1 < D) ? zb[(2 * c + 1 < D) ? 2 * c + 1 : 0] : 0.f);
                sp_store<NS, D, false>(rec, C::AB(NL - 1), ab);

                // ---------------- probes: VJP chain, cotangent on q, tangent pass
                constexpr int NPROBE = EXACT ? D : 1;
#pragma unroll 1
                for (int p = 0; p < NPROBE; ++p) {
                    float probe[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) probe[j] = EXACT ? ((j == p) ? 1.f : 0.f) : eps[j];
                    float2 g[N::NP2];
#pragma unroll
                    for (int c = 0; c < (D + 1) / 2; ++c) g[c] = make_float2(probe[2 * c], (2 * c + 1 < D) ? probe[(2 * c + 1 < D) ? 2 * c + 1 : 0] : 0.f);
                    sp_store<NS, D, false>(rec, C::GG(NL - 1), g);
                    float q[D];
                    static_rfor<NL>([&](auto lc) __attribute__((always_inline)) {
                        constexpr int l = decltype(lc)::value;
                        float2 s[N::NP2];
                        wt_matvec<N, l>(sw, g, s);
                        if constexpr (l > 0) {
#pragma unroll
                            for (int c = 0; c < C::out_p(l - 1); ++c) {
                                const float2 d2 = A.d[l - 1][c];
                                g[c] = __fmul2_rn(s[c], d2);
                                rec[(C::GG(l - 1) + c) * NS] = g[c];
                                rec[(C::CC(l - 1) + c) * NS] = __fmul2_rn(s[c], act_dd2<N::ACT>(A.h[l - 1][c], d2));
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < D; ++k) q[k] = EL(s, k);
                        }
                    });
                    float qb[D];
                    if (EXACT) {
#pragma unroll
                        for (int j = 0; j < D; ++j) qb[j] = (j == p) ? -cl : 0.f;
                    } else {
                        float qq = 0.f;
#pragma unroll
                        for (int j = 0; j < D; ++j) qq = fmaf(q[j], q[j], qq);
                        const float s = (cn != 0.f) ? (a.squared ? 2.0f * cn : (qq > 0.f ? cn * rsqrtf(qq) : 0.f)) : 0.f;
#pragma unroll
                        for (int j = 0; j < D; ++j) qb[j] = fmaf(s, q[j], -cl * eps[j]);
                    }
                    float2 wv[N::NP2];
#pragma unroll
                    for (int c = 0; c < (D + 1) / 2; ++c) wv[c] = make_float2(qb[2 * c], (2 * c + 1 < D) ? qb[(2 * c + 1 < D) ? 2 * c + 1 : 0] : 0.f);
                    sp_store<NS, D, false>(rec, C::WT(0), wv);
                    static_for<0, NH>([&](auto lc) __attribute__((always_inline)) {
                        constexpr int l = decltype(lc)::value;
                        constexpr int kk = N::kz(l), np = C::out_p(l);
                        float2 r[np];
#pragma unroll
                        for (int k = 0; k < kk; ++k) {
                            const float wk = EL(wv, k);
#pragma unroll
                            for (int jp = 0; jp < np; ++jp) {
                                if (k == 0) r[jp] = __fmul2_rn(WPAIR(N, sw2, l, jp, k), bc2(wk));
                                else r[jp] = __ffma2_rn(WPAIR(N, sw2, l, jp, k), bc2(wk), r[jp]);
                            }
                        }
#pragma unroll
                        for (int jp = 0; jp < np; ++jp) {
                            const float2 c2 = rec[(C::CC(l) + jp) * NS];
                            wv[jp] = __fmul2_rn(r[jp], A.d[l][jp]);
                            rec[(C::WT(l + 1) + jp) * NS] = wv[jp];
                            float2 ax = __fmul2_rn(r[jp], c2);
                            if (EXACT && p > 0) ax = __fadd2_rn(ax, rec[(C::AB(l) + jp) * NS]);
                            rec[(C::AB(l) + jp) * NS] = ax;
                        }
                    });
                    if (EXACT && p < NPROBE - 1) {
                        __syncthreads();
                        dw_phase(false, nduo);
                        __syncthreads();
                    }
                }

                // ---------------- backprop with output cotangent zb and the second-order extras
                float sbar[D];
                static_rfor<NL>([&](auto lc) __attribute__((always_inline)) {
                    constexpr int l = decltype(lc)::value;
                    float2 hbv[N::NP2];
                    wt_matvec<N, l>(sw2, ab, hbv);
                    if constexpr (l > 0) {
#pragma unroll
                        for (int c = 0; c < C::out_p(l - 1); ++c) {
                            ab[c] = __ffma2_rn(hbv[c], A.d[l - 1][c], rec[(C::AB(l - 1) + c) * NS]);
                            rec[(C::AB(l - 1) + c) * NS] = ab[c];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < D; ++k) sbar[k] = EL(hbv, k);
                    }
                });
#pragma unroll
                for (int j = 0; j < D; ++j) zbar[j] += sbar[j];
                for (int jj = 0; jj < i; ++jj) {
                    const float c = h * c_a[i][jj];
#pragma unroll
                    for (int j = 0; j < D; ++j) kbm[(jj * D + j) * NS] = fmaf(c, sbar[j], kbm[(jj * D + j) * NS]);
                }
                // ---------------- weight gradient of this stage
                __syncthreads();
                dw_phase(true, nduo);
                __syncthreads();
            }
        }
        if (a.dxs && valid) {
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (j < a.nvars) a.dxs[b * a.nvars + j] = (a.stats->status == ICNF_OK) ? zbar[j] : __int_as_float(0x7fc00000);
        }
    }

    // ---- reduce the register tiles over the sample groups, write this CTA's partial gradient
    __syncthreads();
    float* red = smem;   // [max_groups][NP]
    const int NGM = plan.max_groups;
    for (int i = tid; i < NGM * N::NP; i += NT_) red[i] = 0.f;
    __syncthreads();
    int blk2 = 0;   // recomputed: the role variables are not kept live across the main loop
    while (blk2 < C::NBLK && tid >= (int)plan.first[blk2 + 1]) ++blk2;
    const int grp2 = (int)((desc1 >> 16) & 255u);
    if (blk2 < C::NBLK) {
        static_for<0, NL>([&](auto lc) __attribute__((always_inline)) {
            constexpr int l = decltype(lc)::value;
            constexpr int nb = C::nblk(l), off = C::blkoff(l), nin = N::n(l), nout = N::n(l + 1);
            if (blk2 >= off && blk2 < off + nb) {
                const int i = blk2 - off;
                constexpr int nrc = C::transposed(l) ? C::cdiv(C::out_p(l), 2) : C::cdiv(C::fin_p(l), 2);
                const int lc_ = i / nrc, rc = i - lc_ * nrc;
#pragma unroll
                for (int li = 0; li < 2 * C::LCH; ++li) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int lidx = lc_ * 2 * C::LCH + li;          // index along L
                            const int ridx = rc * 4 + 2 * c + e;             // index along R
                            const int j = C::transposed(l) ? ridx : lidx;    // output unit
                            const int k = C::transposed(l) ? lidx : ridx;    // input (k == nin: bias)
                            const float v = e ? acc[li][c].y : acc[li][c].x;
                            if (j < nout && k <= nin)
                                red[grp2 * N::NP + N::toff(l) + (k < nin ? k * nout + j : nin * nout + j)] = v;
                        }
                    }
                }
            }
        });
    }
    __syncthreads();
    float* gp = a.gpartial + (int64_t)blockIdx.x * N::NP;
    for (int p = tid; p < N::NP; p += NT_) {
        float s = 0.f;
        for (int w = 0; w < NGM; ++w) s += red[w * N::NP + p];
        gp[p] = (a.stats->status == ICNF_OK) ? s : __int_as_float(0x7fc00000);
    }
}

}  // namespace tiny
}  // namespace icnf
