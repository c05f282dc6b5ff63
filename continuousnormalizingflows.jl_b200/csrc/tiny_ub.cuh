// tiny_ub.cuh -- unit-parallel backward kernel of the tiny family.
//
// The thread-per-sample backward (tiny.cuh) keeps the whole second-order state of a sample
// in one thread (255 registers + spills) and must reduce every sample's rank-1 weight
// gradient across the warp at every RK stage.  Here G lanes (4, 8, 16 or 32) cooperate on
// one sample instead: lane g owns hidden units j = g, g + G, ... of every hidden layer and
// the matching ROWS of the weight gradient.  A layer's input vector is all-gathered inside
// the lane group with shuffles, the owned rows of dW accumulate in registers over all
// stages, steps and samples, and lanes are only summed once, at the end of the kernel.
// Weights are staged in shared memory once per CTA, as rows (for W h) and as columns (for
// W' g), so a lane reads its slice with 128-bit loads.
//
// Math: identical to tiny::rhs_reverse / backward_kernel (derivation there and in DESIGN.md);
// reference: gradient of loss, src/core/icnf.jl:628-649 through the solve of
// src/core/base_icnf.jl:134-140 (Zygote + SciMLSensitivity in the reference, icnf.jl:90-99).
#pragma once
#include "tiny.cuh"

namespace icnf {
namespace tiny {

template <class N>
struct UBCfg {
    static constexpr int hmax() {
        int m = 1;
        for (int l = 1; l < N::NL; ++l) m = N::n(l) > m ? N::n(l) : m;
        return m;
    }
    // lanes per sample: at most 3 owned units per lane and hidden layer
    static constexpr int G = hmax() <= 12 ? 4 : hmax() <= 24 ? 8 : hmax() <= 48 ? 16 : 32;
    static constexpr int NH = N::NL - 1;                                  // hidden layers
    __host__ __device__ static constexpr int U(int l) { return (N::n(l + 1) + G - 1) / G; }   // owned units of hidden layer l
    static constexpr int umax() {
        int m = 1;
        for (int l = 0; l < NH; ++l) m = U(l) > m ? U(l) : m;
        return m;
    }
    static constexpr int UMAX = umax();
    __host__ __device__ static constexpr int pad4(int x) { return (x + 3) & ~3; }
    // shared-memory weight layout
    //   rows   Wr_l[j][k]  (hidden layers l = 0..NH-1), pitch pad4(n_l), then bias b_l[j]
    //   cols   Wc_l[k][j]  (l = 1..NH-1: W_l seen from its inputs), pitch pad4(n_{l+1})
    //   last   WL[k][i]    (columns of the output layer, k = hidden unit, i < D'), pitch pad4(D'), then b_L
    __host__ __device__ static constexpr int rpitch(int l) { return pad4(N::n(l)); }
    __host__ __device__ static constexpr int roff(int l) {
        int o = 0;
        for (int i = 0; i < l; ++i) o += N::n(i + 1) * rpitch(i) + pad4(N::n(i + 1));
        return o;
    }
    __host__ __device__ static constexpr int boff(int l) { return roff(l) + N::n(l + 1) * rpitch(l); }
    __host__ __device__ static constexpr int cpitch(int l) { return pad4(N::n(l + 1)); }
    __host__ __device__ static constexpr int coff(int l) {   // l >= 1
        int o = roff(NH);
        for (int i = 1; i < l; ++i) o += N::n(i) * cpitch(i);
        return o;
    }
    static constexpr int lpitch = pad4(N::D);
    static constexpr int loff = coff(NH);
    static constexpr int lboff = loff + N::n(N::NL - 1) * lpitch;
    static constexpr int WSM = lboff + pad4(N::D);
    // gradient accumulators per lane.  Hidden layer l, owned unit u: KP(l) float2 pairs over the inputs
    // (accW) and one bias scalar (accB); output layer: owned columns (accL) and its bias.
    __host__ __device__ static constexpr int KP(int l) { return (N::n(l) + 1) / 2; }
    __host__ __device__ static constexpr int woff2(int l) {   // offset into accW (float2 units)
        int o = 0;
        for (int i = 0; i < l; ++i) o += U(i) * KP(i);
        return o;
    }
    static constexpr int NACCW = woff2(NH);
    __host__ __device__ static constexpr int boffa(int l) {
        int o = 0;
        for (int i = 0; i < l; ++i) o += U(i);
        return o;
    }
    static constexpr int NACCB = boffa(NH);
    static constexpr int NACCL = U(NH - 1) * N::D + N::D;
    static constexpr int NP2 = (N::NMAX + 1) / 2;
    static constexpr int SPB = NT / G;   // samples per CTA
};

// all-gather of a vector distributed over the lane group: full[k] = own[k / G] of lane (k % G)
template <int G, int n, int UM, int NM>
__device__ __forceinline__ void group_gather(const float (&own)[UM], float2 (&full)[NM], int gbase) {
#pragma unroll
    for (int k = 0; k < n; ++k) {
        const float v = __shfl_sync(0xffffffffu, own[k / G], gbase + (k % G));
        if (k & 1) full[k >> 1].y = v;
        else full[k >> 1].x = v;
    }
    if (n & 1) full[n >> 1].y = 0.f;
}
// dot product of a shared-memory weight slice with a vector held as float2 pairs (FFMA2)
template <int n, int NM>
__device__ __forceinline__ float dot_pairs(const float* wrow, const float2 (&vec)[NM]) {
    // two independent accumulator chains: a lane owns only 2-3 units per layer, so the
    // dot products themselves must supply the instruction-level parallelism
    const float2* w2 = reinterpret_cast<const float2*>(wrow);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int kp = 0; kp < (n + 1) / 2; ++kp) {
        if (kp & 1) acc1 = __ffma2_rn(w2[kp], vec[kp], acc1);
        else acc0 = __ffma2_rn(w2[kp], vec[kp], acc0);
    }
    return (acc0.x + acc1.x) + (acc0.y + acc1.y);
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <class N>
struct UState {
    using C = UBCfg<N>;
    float h[C::NH][C::UMAX], d[C::NH][C::UMAX];        // owned activations
    float2 fin[C::NH][C::NP2];                          // gathered input vector of hidden layer l as pairs (l = 0: [z; t; ys])
};

// forward on owned units; zdot replicated in every lane of the group
template <class N>
__device__ __forceinline__ void ub_forward(const float* sw, const float (&x)[N::n(0)], int g, int gbase, UState<N>& S,
                                           float (&zdot)[N::D]) {
    using C = UBCfg<N>;
    static_for<0, C::NH>([&](auto lc) __attribute__((always_inline)) {
        constexpr int l = decltype(lc)::value;
        constexpr int nin = N::n(l), nout = N::n(l + 1);
        if constexpr (l > 0) {
            group_gather<C::G, nin>(S.h[l - 1], S.fin[l], gbase);
        } else {
#pragma unroll
            for (int kp = 0; kp < (nin + 1) / 2; ++kp)
                S.fin[0][kp] = make_float2(x[2 * kp], (2 * kp + 1 < nin) ? x[2 * kp + 1] : 0.f);
        }
#pragma unroll
        for (int u = 0; u < C::U(l); ++u) {
            const int j = g + C::G * u;
            float hval = 0.f, dval = 0.f;
            if (j < nout) {
                const float a = sw[C::boff(l) + j] + dot_pairs<nin>(sw + C::roff(l) + j * C::rpitch(l), S.fin[l]);
                act_eval<N::ACT>(a, hval, dval);
            }
            S.h[l][u] = hval;
            S.d[l][u] = dval;
        }
    });
    // output layer: partial sums over owned hidden units, then all-reduce
    constexpr int lh = C::NH - 1;
#pragma unroll
    for (int i = 0; i < N::D; ++i) zdot[i] = 0.f;
#pragma unroll
    for (int u = 0; u < C::U(lh); ++u) {
        const int k = g + C::G * u;
        if (k < N::n(N::NL - 1)) {
            const float* col = sw + C::loff + k * C::lpitch;
#pragma unroll
            for (int i = 0; i < N::D; ++i) zdot[i] = fmaf(col[i], S.h[lh][u], zdot[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < N::D; ++i) zdot[i] = group_sum<C::G>(zdot[i]) + sw[C::lboff + i];
}

// out[u'] (owned units k of hidden layer l-1) = sum_j W_l[j, k] * vec[j]   with vec gathered over the group (l >= 1)
template <class N, int l>
__device__ __forceinline__ void ub_wt_hidden(const float* sw, const float (&vec_own)[UBCfg<N>::UMAX], int g, int gbase,
                                             float (&out)[UBCfg<N>::UMAX]) {
    using C = UBCfg<N>;
    constexpr int nout = N::n(l + 1), nin = N::n(l);
    float2 full[C::NP2];
    group_gather<C::G, nout>(vec_own, full, gbase);
#pragma unroll
    for (int u = 0; u < C::U(l - 1); ++u) {
        const int k = g + C::G * u;
        out[u] = (k < nin) ? dot_pairs<nout>(sw + C::coff(l) + k * C::cpitch(l), full) : 0.f;
    }
}

#ifndef ICNF_UB_MINB
#define ICNF_UB_MINB 2   // 2 CTAs per SM: no spills at ~235 registers beats 3 CTAs with spills (sweep in profiles/README.md)
#endif
template <class N, bool EXACT>
__global__ void __launch_bounds__(NT, ICNF_UB_MINB) backward_ub_kernel(BackwardArgs a) {
    using C = UBCfg<N>;
    constexpr int G = C::G, NH = C::NH, lh = NH - 1, D = N::D;
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                                   // weights, C::WSM floats
    float* sKB = smem + C::WSM;                         // stage cotangents [6 * D][SPB]
    // ---- stage the weights (rows, columns, output layer) from the native ComponentArray layout
    for (int idx = threadIdx.x; idx < C::WSM; idx += NT) sw[idx] = 0.f;
    __syncthreads();
    static_for<0, N::NL>([&](auto lc) __attribute__((always_inline)) {
        constexpr int l = decltype(lc)::value;
        constexpr int nin = N::n(l), nout = N::n(l + 1);
        for (int e = threadIdx.x; e < nin * nout; e += NT) {
            const int k = e / nout, j = e - k * nout;
            const float wv = __ldg(a.theta + N::toff(l) + e);
            if constexpr (l < NH) sw[C::roff(l) + j * C::rpitch(l) + k] = wv;
            if constexpr (l >= 1 && l < NH) sw[C::coff(l) + k * C::cpitch(l) + j] = wv;
            if constexpr (l == N::NL - 1) sw[C::loff + k * C::lpitch + j] = wv;
        }
        for (int j = threadIdx.x; j < nout; j += NT) {
            const float bv = __ldg(a.theta + N::toff(l) + nin * nout + j);
            if constexpr (l < NH) sw[C::boff(l) + j] = bv;
            else sw[C::lboff + j] = bv;
        }
    });
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int g = lane % G, gbase = lane - g;
    const int slot = threadIdx.x / G;                   // sample slot inside the CTA
    const int nsteps = a.stats->naccept;
    float2 accW[C::NACCW];
    float accB[C::NACCB], accL[C::NACCL];
#pragma unroll
    for (int i = 0; i < C::NACCW; ++i) accW[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < C::NACCB; ++i) accB[i] = 0.f;
#pragma unroll
    for (int i = 0; i < C::NACCL; ++i) accL[i] = 0.f;

    const int64_t stride = (int64_t)gridDim.x * C::SPB;
    const int64_t nloop = (a.B + stride - 1) / stride;
    for (int64_t it = 0; it < nloop; ++it) {
        const int64_t braw = it * stride + (int64_t)blockIdx.x * C::SPB + slot;
        const bool valid = braw < a.B;
        const int64_t b = valid ? braw : a.B - 1;
        const float wgt = valid ? a.inv_denominator : 0.f;

        float eps[D], x[N::n(0)];
        if (a.mode != ICNF_TEST) {
            if (a.eps_kind == ICNF_EPS_SUPPLIED) {
#pragma unroll
                for (int j = 0; j < D; ++j) eps[j] = __ldg(a.eps + b * D + j);
            } else {
#pragma unroll
                for (int blk = 0; blk < (D + 3) / 4; ++blk) {
                    float o[4];
                    philox_draw4(a.eps_kind, a.seed, PHILOX_STREAM_EPS, a.sample_offset + b, blk, o);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (blk * 4 + r < D) eps[blk * 4 + r] = o[r];
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
        const float lbar = wgt;
        const float Ebar = a.reg_e ? a.lam1 * wgt : 0.f;
        const float nbar = a.reg_n ? a.lam2 * wgt : 0.f;

        UState<N> S;
        for (int step = nsteps - 1; step >= 0; --step) {
            const float t = a.steps[step].t, h = a.steps[step].dt;
            __syncwarp();
            if (g == 0) {
                for (int i = 0; i < 6; ++i) {
                    const float c = h * c_a[6][i];
#pragma unroll
                    for (int j = 0; j < D; ++j) sKB[(i * D + j) * C::SPB + slot] = c * zbar[j];
                }
            }
            __syncwarp();

            for (int i = 5; i >= 0; --i) {
                float kb[D], zdot[D];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    x[j] = a.ckpt[ckpt_index<N>(step, a.B, b, i) + j];
                    kb[j] = sKB[(i * D + j) * C::SPB + slot];
                }
                if constexpr (N::TIN) x[D] = fmaf(c_c[i], h, t);
                const float hb = h * c_a[6][i];
                const float cl = hb * lbar, cE = hb * Ebar, cn = hb * nbar;

                ub_forward<N>(sw, x, g, gbase, S, zdot);
                // cotangent on zdot
                float zb[D];
#pragma unroll
                for (int j = 0; j < D; ++j) zb[j] = kb[j];
                if (!EXACT && cE != 0.f) {
                    float zz = 0.f;
#pragma unroll
                    for (int j = 0; j < D; ++j) zz = fmaf(zdot[j], zdot[j], zz);
                    const float s = a.squared ? 2.0f * cE : (zz > 0.f ? cE * rsqrtf(zz) : 0.f);
#pragma unroll
                    for (int j = 0; j < D; ++j) zb[j] = fmaf(s, zdot[j], zb[j]);
                }
                float aex[NH][C::UMAX];
#pragma unroll
                for (int l = 0; l < NH; ++l)
#pragma unroll
                    for (int u = 0; u < C::UMAX; ++u) aex[l][u] = 0.f;

                const int nprobe = EXACT ? D : 1;
                for (int p = 0; p < nprobe; ++p) {
                    float probe[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) probe[j] = EXACT ? ((j == p) ? 1.f : 0.f) : eps[j];
                    // ---- VJP chain: v, gch on owned units
                    float v[NH][C::UMAX], gch[NH][C::UMAX];
#pragma unroll
                    for (int u = 0; u < C::U(lh); ++u) {
                        const int k = g + G * u;
                        float s = 0.f;
                        if (k < N::n(N::NL - 1)) {
                            const float* col = sw + C::loff + k * C::lpitch;
#pragma unroll
                            for (int j = 0; j < D; ++j) s = fmaf(col[j], probe[j], s);
                        }
                        v[lh][u] = s;
                        gch[lh][u] = s * S.d[lh][u];
                    }
                    static_rfor<NH>([&](auto lc) __attribute__((always_inline)) {
                        constexpr int l = decltype(lc)::value;   // hidden layer l, l >= 1 handled here
                        if constexpr (l >= 1) {
                            ub_wt_hidden<N, l>(sw, gch[l], g, gbase, v[l - 1]);
#pragma unroll
                            for (int u = 0; u < C::U(l - 1); ++u) gch[l - 1][u] = v[l - 1][u] * S.d[l - 1][u];
                        }
                    });
                    float q[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) q[j] = 0.f;
#pragma unroll
                    for (int u = 0; u < C::U(0); ++u) {
                        const int k = g + G * u;
                        if (k < N::n(1)) {
                            const float* row = sw + C::roff(0) + k * C::rpitch(0);
#pragma unroll
                            for (int j = 0; j < D; ++j) q[j] = fmaf(row[j], gch[0][u], q[j]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < D; ++j) q[j] = group_sum<G>(q[j]);
                    // ---- cotangent on q
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
                    // ---- tangent pass on owned units; the chain's weight gradient g_l w_l' goes straight
                    //      into the owned rows
                    float wv[NH][C::UMAX];
                    static_for<0, NH>([&](auto lc) __attribute__((always_inline)) {
                        constexpr int l = decltype(lc)::value;
                        constexpr int nin = N::n(l), nout = N::n(l + 1);
                        // tangent input of this layer as pairs: qb (first D' inputs, zero elsewhere) or gathered w_{l-1}
                        float2 wfull[C::NP2];
                        if constexpr (l > 0) {
                            group_gather<G, nin>(wv[l - 1], wfull, gbase);
                        } else {
#pragma unroll
                            for (int kp = 0; kp < (nin + 1) / 2; ++kp)
                                wfull[kp] = make_float2(2 * kp < D ? qb[2 * kp < D ? 2 * kp : 0] : 0.f,
                                                        2 * kp + 1 < D ? qb[2 * kp + 1 < D ? 2 * kp + 1 : 0] : 0.f);
                        }
                        constexpr int kpn = (l == 0) ? (D + 1) / 2 : (nin + 1) / 2;   // pairs that can be non-zero
#pragma unroll
                        for (int u = 0; u < C::U(l); ++u) {
                            const int j = g + G * u;
                            float r = 0.f;
                            if (j < nout) {
                                const float2* w2 = reinterpret_cast<const float2*>(sw + C::roff(l) + j * C::rpitch(l));
                                float2 racc = make_float2(0.f, 0.f), racc1 = make_float2(0.f, 0.f);
#pragma unroll
                                for (int kp = 0; kp < kpn; ++kp) {
                                    if (kp & 1) racc1 = __ffma2_rn(w2[kp], wfull[kp], racc1);
                                    else racc = __ffma2_rn(w2[kp], wfull[kp], racc);
                                    accW[C::woff2(l) + u * C::KP(l) + kp] =
                                        __ffma2_rn(bc2(gch[l][u]), wfull[kp], accW[C::woff2(l) + u * C::KP(l) + kp]);
                                }
                                r = (racc.x + racc1.x) + (racc.y + racc1.y);
                            }
                            wv[l][u] = r * S.d[l][u];
                            aex[l][u] = fmaf(r * v[l][u], act_dd<N::ACT>(S.h[l][u], S.d[l][u]), aex[l][u]);
                        }
                    });
                    // output layer's chain gradient: dW_L[i, k] += probe_i * wv[lh][k] on owned columns k
#pragma unroll
                    for (int u = 0; u < C::U(lh); ++u)
#pragma unroll
                        for (int j = 0; j < D; ++j) accL[u * D + j] = fmaf(probe[j], wv[lh][u], accL[u * D + j]);
                }

                // ---- backprop with output cotangent zb
                float ab[NH][C::UMAX];
#pragma unroll
                for (int u = 0; u < C::U(lh); ++u) {
                    const int k = g + G * u;
                    float s = 0.f;
                    if (k < N::n(N::NL - 1)) {
                        const float* col = sw + C::loff + k * C::lpitch;
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            s = fmaf(col[j], zb[j], s);
                            accL[u * D + j] = fmaf(zb[j], S.h[lh][u], accL[u * D + j]);
                        }
                    }
                    ab[lh][u] = fmaf(s, S.d[lh][u], aex[lh][u]);
                }
                if (g == 0) {
#pragma unroll
                    for (int j = 0; j < D; ++j) accL[C::U(lh) * D + j] += zb[j];
                }
                static_rfor<NH>([&](auto lc) __attribute__((always_inline)) {
                    constexpr int l = decltype(lc)::value;
                    // owned rows of dW_l and db_l
#pragma unroll
                    for (int u = 0; u < C::U(l); ++u) {
                        if (g + G * u < N::n(l + 1)) {
#pragma unroll
                            for (int kp = 0; kp < C::KP(l); ++kp)
                                accW[C::woff2(l) + u * C::KP(l) + kp] =
                                    __ffma2_rn(bc2(ab[l][u]), S.fin[l][kp], accW[C::woff2(l) + u * C::KP(l) + kp]);
                            accB[C::boffa(l) + u] += ab[l][u];
                        }
                    }
                    if constexpr (l >= 1) {
                        float hb2[C::UMAX];
                        ub_wt_hidden<N, l>(sw, ab[l], g, gbase, hb2);
#pragma unroll
                        for (int u = 0; u < C::U(l - 1); ++u) ab[l - 1][u] = fmaf(hb2[u], S.d[l - 1][u], aex[l - 1][u]);
                    }
                });
                float sbar[D];
#pragma unroll
                for (int j = 0; j < D; ++j) sbar[j] = 0.f;
#pragma unroll
                for (int u = 0; u < C::U(0); ++u) {
                    const int k = g + G * u;
                    if (k < N::n(1)) {
                        const float* row = sw + C::roff(0) + k * C::rpitch(0);
#pragma unroll
                        for (int j = 0; j < D; ++j) sbar[j] = fmaf(row[j], ab[0][u], sbar[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    sbar[j] = group_sum<G>(sbar[j]);
                    zbar[j] += sbar[j];
                }
                __syncwarp();
                if (g == 0) {
                    for (int jj = 0; jj < i; ++jj) {
                        const float c = h * c_a[i][jj];
#pragma unroll
                        for (int j = 0; j < D; ++j) sKB[(jj * D + j) * C::SPB + slot] = fmaf(c, sbar[j], sKB[(jj * D + j) * C::SPB + slot]);
                    }
                }
                __syncwarp();
            }
        }
        if (a.dxs && valid && g == 0) {
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (j < a.nvars) a.dxs[b * a.nvars + j] = zbar[j];
        }
    }

    // ---- sum the accumulators over the samples of the warp (lanes with the same g), then over warps
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < C::NACCW; ++i) {
            accW[i].x += __shfl_xor_sync(0xffffffffu, accW[i].x, o);
            accW[i].y += __shfl_xor_sync(0xffffffffu, accW[i].y, o);
        }
#pragma unroll
        for (int i = 0; i < C::NACCB; ++i) accB[i] += __shfl_xor_sync(0xffffffffu, accB[i], o);
#pragma unroll
        for (int i = 0; i < C::NACCL; ++i) accL[i] += __shfl_xor_sync(0xffffffffu, accL[i], o);
    }
    __syncthreads();
    float* sg = smem;   // [NT / 32][NP], reuses the weight / stage area
    const int wid = threadIdx.x >> 5;
    if (lane < G) {
        static_for<0, NH>([&](auto lc) __attribute__((always_inline)) {
            constexpr int l = decltype(lc)::value;
            constexpr int nin = N::n(l), nout = N::n(l + 1);
#pragma unroll
            for (int u = 0; u < C::U(l); ++u) {
                const int j = g + G * u;
                if (j < nout) {
#pragma unroll
                    for (int k = 0; k < nin; ++k) {
                        const float2 pr = accW[C::woff2(l) + u * C::KP(l) + (k >> 1)];
                        sg[wid * N::NP + N::toff(l) + k * nout + j] = (k & 1) ? pr.y : pr.x;
                    }
                    sg[wid * N::NP + N::toff(l) + nin * nout + j] = accB[C::boffa(l) + u];
                }
            }
        });
        constexpr int nl = N::NL - 1, nk = N::n(nl);
#pragma unroll
        for (int u = 0; u < C::U(lh); ++u) {
            const int k = g + G * u;
            if (k < nk) {
#pragma unroll
                for (int j = 0; j < D; ++j) sg[wid * N::NP + N::toff(nl) + k * D + j] = accL[u * D + j];
            }
        }
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < D; ++j) sg[wid * N::NP + N::toff(nl) + nk * D + j] = accL[C::U(lh) * D + j];
        }
    }
    __syncthreads();
    float* gp = a.gpartial + (int64_t)blockIdx.x * N::NP;
    for (int p = threadIdx.x; p < N::NP; p += NT) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += sg[w * N::NP + p];
        gp[p] = s;
    }
}

}  // namespace tiny
}  // namespace icnf
