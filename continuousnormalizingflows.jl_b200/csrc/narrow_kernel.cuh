// narrow_kernel.cuh -- "narrow" fast path of the generic family: the whole exact-trace (TestMode) Tsit5 solve of a narrow MLP
// with two hidden layers (every width <= 128, D' <= 32: BASELINE config 3, ICNF(nvariables = 3..7), 3-64-64-2) in ONE
// persistent cooperative kernel, any widths at run time.
//
// What it computes (reference, paths relative to the reference root): augmented_f in TestMode, src/core/icnf.jl:297-316,
// with the exact trace that utils.jl:35-54 obtains from D' one-hot pullbacks, here in closed form
//     tr J = d2' (W2 .* (W1z W3)') d1                      (the matrix comes from generic.cu's on_params)
// integrated by Tsit5 (base_sol, src/core/base_icnf.jl:134-140) with the whole-batch RMS error norm, fused with
// inference_prob (:247-296), inference_sol (:158-172), reg_z_aug (:106-132) and generate_sol (:185-194).
//
// Design.  A CTA of 8 warps owns a tile of 128 samples for a whole step attempt: all six Tsit5 stages of the tile run
// out of shared memory and registers, and only the state and the FSAL derivative cross HBM (2 x 2 x S floats per
// sample and attempt).  The weights (transposed, padded) stay in shared memory for the life of the kernel.  A layer is a
// small SGEMM: warp w owns output units [w jt, (w + 1) jt), lane l owns samples 4 l .. 4 l + 3; per input k the warp
// reads its units' weights with broadcast 128-bit loads and its samples' activations with one conflict-free 128-bit load,
// and issues packed FP32 FMAs (fma.rn.f32x2) on (unit pair) x sample accumulators.  The second hidden layer accumulates
// the trace contraction (A d1) next to (W2 h1) in the same k loop.  The thread that computes a z-row's derivative also
// owns that row's Tsit5 state (k1..k5, the solution and error sums) in registers, so stage combination, error norm and
// accept / reject never leave the kernel; the controller is the tiny family's (one grid-wide reduction per attempt).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>

#pragma once
#include "tiny.cuh"
#include "tiny_vcabm.cuh"
#include "narrow.h"

namespace icnf {
namespace narrow {

namespace cg = cooperative_groups;
using tiny::c_a;
using tiny::c_bt;
using tiny::c_c;

constexpr int NW = 8, NTHR = NW * 32, SPT = 4, NS = 32 * SPT;

struct Layout {            // shared-memory offsets in floats
    int n0, n1, n2, D, C, tin, act;
    int jt1, jt2, jt3;     // output units per warp of layers 1..3
    int ld12, ld3;         // padded row length of the transposed weights: 8 * JP / 8 * JP3
    int w1, b1, w2, at, b2, w3, b3, x, h1, d1, h2, red, total;
    // Hutchinson modes (exact = 0): transposed-role copies of the weights for the VJP chain (W3b[i][unit of layer 2],
    // W2b[unit of layer 2][unit of layer 1], W1b[unit of layer 1][z-row]), the probe tile, the zdot tile; no sigma' arrays
    // (derived from h) and no trace matrix
    int exact, w3b, w2b, w1b, eps, zdt;
};

struct Params {
    SolveArgs a;
    const float* amat;     // exact-trace matrix, (j, k) at k * n2 + j  (generic.cu g_trace_matrix_kernel)
    Layout L;
    int nvars, adaptive;
    long long woff[3], boff[3];
};

__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }

// acc[jp][s] (+)= sum_k Wt[k][warp slot 2 jp, 2 jp + 1] * act[k][4 lane + s];  DUAL: the same for (Wt2, act2) -> acc2
template <int JP, bool DUAL>
__device__ __forceinline__ void layer_mm(const float* wt, const float* wt2, int ldw,
                                         const float* act, const float* act2, int nin, int lane,
                                         float2 (&acc)[JP / 2][SPT], float2 (&acc2)[JP / 2][SPT]) {
#pragma unroll
    for (int p = 0; p < JP / 2; ++p)
#pragma unroll
        for (int s = 0; s < SPT; ++s) { acc[p][s] = f2(0.f, 0.f); if (DUAL) acc2[p][s] = f2(0.f, 0.f); }
    const float* ap = act + 4 * lane;
    const float* ap2 = DUAL ? act2 + 4 * lane : nullptr;
    // pointers advance by one input per trip, so that every load of the body has a constant offset
#pragma unroll 4
    for (int k = 0; k < nin; ++k, wt += ldw, ap += NS) {
        const float4 a = *reinterpret_cast<const float4*>(ap);
        const float2 as[SPT] = {f2(a.x, a.x), f2(a.y, a.y), f2(a.z, a.z), f2(a.w, a.w)};
#pragma unroll
        for (int p = 0; p < JP / 2; ++p) {
            const float2 wv = *reinterpret_cast<const float2*>(wt + 2 * p);
#pragma unroll
            for (int s = 0; s < SPT; ++s) acc[p][s] = __ffma2_rn(wv, as[s], acc[p][s]);
        }
        if constexpr (DUAL) {
            const float4 b = *reinterpret_cast<const float4*>(ap2);
            const float2 bs[SPT] = {f2(b.x, b.x), f2(b.y, b.y), f2(b.z, b.z), f2(b.w, b.w)};
#pragma unroll
            for (int p = 0; p < JP / 2; ++p) {
                const float2 wv = *reinterpret_cast<const float2*>(wt2 + 2 * p);
#pragma unroll
                for (int s = 0; s < SPT; ++s) acc2[p][s] = __ffma2_rn(wv, bs[s], acc2[p][s]);
            }
            wt2 += ldw; ap2 += NS;
        }
    }
}

__device__ __forceinline__ float comp(const float2& v, int odd) { return odd ? v.y : v.x; }
// ACT >= 0: compile-time activation; ACT < 0: the configured one, dispatched per element (rare activations)
template <int ACT>
__device__ __forceinline__ void act_any(int act, float a, float& h, float& d) {
    if constexpr (ACT >= 0) act_eval<ACT>(a, h, d);
    else act_eval_rt(act, a, h, d);
}
// sigma'(a) from h = sigma(a) (the Hutchinson path stores no sigma' arrays)
template <int ACT>
__device__ __forceinline__ float deriv_from_h(int act, float h) {
    const int a = ACT >= 0 ? ACT : act;
    if (a == ICNF_ACT_SOFTPLUS) return 1.0f - __expf(-h);
    if (a == ICNF_ACT_TANH) return 1.0f - h * h;
    if (a == ICNF_ACT_SIGMOID) return h * (1.0f - h);
    return 1.0f;
}
// the probe of one tile (constant over the solve): supplied matrix or the in-kernel Philox draw of the other families
static __device__ __noinline__ void fill_eps_tile(const SolveArgs& a, float* tile, long long tile0, int D) {
    const int nblk = (D + 3) / 4;
    for (int idx = threadIdx.x; idx < NS * nblk; idx += NTHR) {
        const int s = idx % NS, blk = idx / NS;
        const long long b = tile0 + s;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (b < a.B) {
            if (a.eps_kind == ICNF_EPS_SUPPLIED) {
                for (int r = 0; r < 4; ++r) if (4 * blk + r < D) o[r] = __ldg(a.eps + b * D + 4 * blk + r);
            } else {
                philox_draw4(a.eps_kind, a.seed, PHILOX_STREAM_EPS, a.sample_offset + b, blk, o);
            }
        }
        for (int r = 0; r < 4; ++r) if (4 * blk + r < D) tile[(4 * blk + r) * NS + s] = o[r];
    }
}
// first touch of a state row (inference_prob / generate_prob): out of line, it is cold and its Gaussian draw is large
static __device__ __noinline__ float input_value(const SolveArgs& a, long long b, int j, int D, int nvars) {
    const int S = D + 3;
    if (a.in_kind == IN_U0) return __ldg(a.in + b * S + j);
    if (a.in_kind == IN_XS) return (j < nvars) ? __ldg(a.in + b * nvars + j) : 0.f;
    if (a.in_kind == IN_Z0) return __ldg(a.in + b * D + j);
    float o[4];
    philox_draw4(ICNF_EPS_GAUSSIAN, a.seed, PHILOX_STREAM_BASE, a.sample_offset + b, j >> 2, o);
    return o[j & 3];
}

// Hutchinson RHS (TrainMode): forward pass, then the VJP chain of the probe through transposed-role weight copies,
//   v2 = W3' eps, g2 = v2 .* s'(a2);  v1 = W2' g2, g1 = v1 .* s'(a1);  q = W1z' g1 = eps' J,
// and per sample l' = -q . eps, E' = |zdot|, n' = |q| (icnf.jl:517-536, :184-251).  zd = zdot rows in owner layout; the three
// per-sample values are left in sm[L.red + {0, 1, 2} * NS + sample].  Eight CTA barriers.
template <int JP, int JP3, int ACT>
__device__ __forceinline__ void rhs_tile_hutch(const Layout& L, float* sm, int warp, int lane, int reg_e, int reg_n, int squared,
                                               float (&zd)[JP3][SPT]) {
    float2 acc[JP / 2][SPT], dummy[JP / 2][SPT];
    // ---- forward
    layer_mm<JP, false>(sm + L.w1 + warp * JP, nullptr, L.ld12, sm + L.x, nullptr, L.n0, lane, acc, dummy);
#pragma unroll
    for (int r = 0; r < JP; ++r) {
        const int j = warp * L.jt1 + r;
        if (r < L.jt1 && j < L.n1) {
            const float b = sm[L.b1 + warp * JP + r];
            float h[SPT], d[SPT];
#pragma unroll
            for (int s = 0; s < SPT; ++s) act_any<ACT>(L.act, comp(acc[r >> 1][s], r & 1) + b, h[s], d[s]);
            *reinterpret_cast<float4*>(sm + L.h1 + j * NS + 4 * lane) = make_float4(h[0], h[1], h[2], h[3]);
        }
    }
    __syncthreads();
    layer_mm<JP, false>(sm + L.w2 + warp * JP, nullptr, L.ld12, sm + L.h1, nullptr, L.n1, lane, acc, dummy);
#pragma unroll
    for (int r = 0; r < JP; ++r) {
        const int j = warp * L.jt2 + r;
        if (r < L.jt2 && j < L.n2) {
            const float b = sm[L.b2 + warp * JP + r];
            float h[SPT], d[SPT];
#pragma unroll
            for (int s = 0; s < SPT; ++s) act_any<ACT>(L.act, comp(acc[r >> 1][s], r & 1) + b, h[s], d[s]);
            *reinterpret_cast<float4*>(sm + L.h2 + j * NS + 4 * lane) = make_float4(h[0], h[1], h[2], h[3]);
        }
    }
    __syncthreads();
    float2 acc3[JP3 / 2][SPT], dummy3[JP3 / 2][SPT];
    layer_mm<JP3, false>(sm + L.w3 + warp * JP3, nullptr, L.ld3, sm + L.h2, nullptr, L.n2, lane, acc3, dummy3);
#pragma unroll
    for (int r = 0; r < JP3; ++r) {
        const float b = sm[L.b3 + warp * JP3 + r];
        const int j = warp * L.jt3 + r;
#pragma unroll
        for (int s = 0; s < SPT; ++s) zd[r][s] = comp(acc3[r >> 1][s], r & 1) + b;
        if (r < L.jt3 && j < L.D) *reinterpret_cast<float4*>(sm + L.zdt + j * NS + 4 * lane) = make_float4(zd[r][0], zd[r][1], zd[r][2], zd[r][3]);
    }
    __syncthreads();            // every warp has read h2 (layer 3) before it is overwritten with g2
    // ---- VJP chain
    layer_mm<JP, false>(sm + L.w3b + warp * JP, nullptr, L.ld12, sm + L.eps, nullptr, L.D, lane, acc, dummy);
#pragma unroll
    for (int r = 0; r < JP; ++r) {
        const int j = warp * L.jt2 + r;
        if (r < L.jt2 && j < L.n2) {
            float* hp = sm + L.h2 + j * NS + 4 * lane;
            const float4 h = *reinterpret_cast<const float4*>(hp);
            *reinterpret_cast<float4*>(hp) = make_float4(comp(acc[r >> 1][0], r & 1) * deriv_from_h<ACT>(L.act, h.x),
                                                         comp(acc[r >> 1][1], r & 1) * deriv_from_h<ACT>(L.act, h.y),
                                                         comp(acc[r >> 1][2], r & 1) * deriv_from_h<ACT>(L.act, h.z),
                                                         comp(acc[r >> 1][3], r & 1) * deriv_from_h<ACT>(L.act, h.w));
        }
    }
    __syncthreads();
    layer_mm<JP, false>(sm + L.w2b + warp * JP, nullptr, L.ld12, sm + L.h2, nullptr, L.n2, lane, acc, dummy);
#pragma unroll
    for (int r = 0; r < JP; ++r) {
        const int j = warp * L.jt1 + r;
        if (r < L.jt1 && j < L.n1) {
            float* hp = sm + L.h1 + j * NS + 4 * lane;
            const float4 h = *reinterpret_cast<const float4*>(hp);
            *reinterpret_cast<float4*>(hp) = make_float4(comp(acc[r >> 1][0], r & 1) * deriv_from_h<ACT>(L.act, h.x),
                                                         comp(acc[r >> 1][1], r & 1) * deriv_from_h<ACT>(L.act, h.y),
                                                         comp(acc[r >> 1][2], r & 1) * deriv_from_h<ACT>(L.act, h.z),
                                                         comp(acc[r >> 1][3], r & 1) * deriv_from_h<ACT>(L.act, h.w));
        }
    }
    __syncthreads();
    layer_mm<JP3, false>(sm + L.w1b + warp * JP3, nullptr, L.ld3, sm + L.h1, nullptr, L.n1, lane, acc3, dummy3);
#pragma unroll
    for (int r = 0; r < JP3; ++r) {
        const int j = warp * L.jt3 + r;     // q row j: into the (dead) stage-input rows of the x tile
        if (r < L.jt3 && j < L.D)
            *reinterpret_cast<float4*>(sm + L.x + j * NS + 4 * lane) =
                make_float4(comp(acc3[r >> 1][0], r & 1), comp(acc3[r >> 1][1], r & 1), comp(acc3[r >> 1][2], r & 1), comp(acc3[r >> 1][3], r & 1));
    }
    __syncthreads();
    if (threadIdx.x < NS) {      // one thread per sample: the three reductions over the D' rows, in row order
        const int s = threadIdx.x;
        float s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int i = 0; i < L.D; ++i) {
            const float q = sm[L.x + i * NS + s], e = sm[L.eps + i * NS + s], z = sm[L.zdt + i * NS + s];
            s1 = fmaf(q, e, s1); s2 = fmaf(q, q, s2); s3 = fmaf(z, z, s3);
        }
        sm[L.red + s] = -s1;
        sm[L.red + NS + s] = reg_e ? tiny::vec_norm(s3, squared) : 0.f;
        sm[L.red + 2 * NS + s] = reg_n ? tiny::vec_norm(s2, squared) : 0.f;
    }
    __syncthreads();
}

// One RHS evaluation of the tile whose network input sits in sm[L.x].  zd[r][s] = derivative of z-row (warp jt3 + r) for
// sample 4 lane + s (owner layout); the per-sample trace is left in sm[L.red + sample] by warp 0 (read it after the call's
// final barrier).  Three CTA barriers.
template <int JP, int JP3, int ACT>
__device__ __forceinline__ void rhs_tile(const Layout& L, float* sm, int warp, int lane, float (&zd)[JP3][SPT]) {
    float2 acc[JP / 2][SPT], acc2[JP / 2][SPT];
    // ---- layer 1: h1, d1
    layer_mm<JP, false>(sm + L.w1 + warp * JP, nullptr, L.ld12, sm + L.x, nullptr, L.n0, lane, acc, acc2);
#pragma unroll
    for (int r = 0; r < JP; ++r) {
        const int j = warp * L.jt1 + r;
        if (r < L.jt1 && j < L.n1) {
            const float b = sm[L.b1 + warp * JP + r];
            float h[SPT], d[SPT];
#pragma unroll
            for (int s = 0; s < SPT; ++s) act_any<ACT>(L.act, comp(acc[r >> 1][s], r & 1) + b, h[s], d[s]);
            *reinterpret_cast<float4*>(sm + L.h1 + j * NS + 4 * lane) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(sm + L.d1 + j * NS + 4 * lane) = make_float4(d[0], d[1], d[2], d[3]);
        }
    }
    __syncthreads();
    // ---- layer 2 with the trace contraction: a2 = W2 h1 + b2, t = A d1;  h2 = act(a2), trace partial += act'(a2) .* t
    layer_mm<JP, true>(sm + L.w2 + warp * JP, sm + L.at + warp * JP, L.ld12, sm + L.h1, sm + L.d1, L.n1, lane, acc, acc2);
    float trp[SPT] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < JP; ++r) {
        const int j = warp * L.jt2 + r;
        if (r < L.jt2 && j < L.n2) {
            const float b = sm[L.b2 + warp * JP + r];
            float h[SPT], d[SPT];
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                act_any<ACT>(L.act, comp(acc[r >> 1][s], r & 1) + b, h[s], d[s]);
                trp[s] = fmaf(d[s], comp(acc2[r >> 1][s], r & 1), trp[s]);
            }
            *reinterpret_cast<float4*>(sm + L.h2 + j * NS + 4 * lane) = make_float4(h[0], h[1], h[2], h[3]);
        }
    }
    *reinterpret_cast<float4*>(sm + L.red + (1 + warp) * NS + 4 * lane) = make_float4(trp[0], trp[1], trp[2], trp[3]);
    __syncthreads();
    // ---- layer 3 (linear): zdot rows of this warp; warp 0 also adds up the trace partials (fixed order)
    float2 acc3[JP3 / 2][SPT], dummy[JP3 / 2][SPT];
    layer_mm<JP3, false>(sm + L.w3 + warp * JP3, nullptr, L.ld3, sm + L.h2, nullptr, L.n2, lane, acc3, dummy);
#pragma unroll
    for (int r = 0; r < JP3; ++r) {
        const float b = sm[L.b3 + warp * JP3 + r];
#pragma unroll
        for (int s = 0; s < SPT; ++s) zd[r][s] = comp(acc3[r >> 1][s], r & 1) + b;
    }
    if (warp == 0) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float4 v = *reinterpret_cast<const float4*>(sm + L.red + (1 + w) * NS + 4 * lane);
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        *reinterpret_cast<float4*>(sm + L.red + 4 * lane) = t;
    }
    __syncthreads();
}

// weights -> shared memory (transposed, padded per warp; the padding reads as zero)
template <int JP, int JP3, bool EXACT>
__device__ __forceinline__ void stage_weights(const Params& P, float* sm) {
    const SolveArgs& a = P.a;
    const Layout& L = P.L;
    const int D = L.D;
    for (int i = threadIdx.x; i < L.x; i += NTHR) sm[i] = 0.f;
    __syncthreads();
    {
        const float* th = a.theta;
        for (int i = threadIdx.x; i < L.n0 * L.n1; i += NTHR) {       // W1 (j, k) at k * n1 + j
            const int k = i / L.n1, j = i - k * L.n1;
            sm[L.w1 + k * L.ld12 + (j / L.jt1) * JP + (j % L.jt1)] = th[P.woff[0] + i];
        }
        for (int i = threadIdx.x; i < L.n1 * L.n2; i += NTHR) {
            const int k = i / L.n2, j = i - k * L.n2;
            const int col = (j / L.jt2) * JP + (j % L.jt2);
            sm[L.w2 + k * L.ld12 + col] = th[P.woff[1] + i];
            if constexpr (EXACT) sm[L.at + k * L.ld12 + col] = P.amat[i];
            else sm[L.w2b + j * L.ld12 + (k / L.jt1) * JP + (k % L.jt1)] = th[P.woff[1] + i];     // W2b[unit of layer 2][unit of layer 1]
        }
        if constexpr (!EXACT) {
            for (int i = threadIdx.x; i < L.n2 * D; i += NTHR) {       // W3 (r, j) at j * D + r  ->  W3b[r][unit j of layer 2]
                const int j = i / D, r = i - j * D;
                sm[L.w3b + r * L.ld12 + (j / L.jt2) * JP + (j % L.jt2)] = th[P.woff[2] + i];
            }
            for (int i = threadIdx.x; i < D * L.n1; i += NTHR) {       // W1 (j, r) at r * n1 + j, r < D'  ->  W1b[unit j of layer 1][z-row r]
                const int r = i / L.n1, j = i - r * L.n1;
                sm[L.w1b + j * L.ld3 + (r / L.jt3) * JP3 + (r % L.jt3)] = th[P.woff[0] + i];
            }
        }
        for (int i = threadIdx.x; i < L.n2 * D; i += NTHR) {
            const int k = i / D, j = i - k * D;
            sm[L.w3 + k * L.ld3 + (j / L.jt3) * JP3 + (j % L.jt3)] = th[P.woff[2] + i];
        }
        for (int j = threadIdx.x; j < L.n1; j += NTHR) sm[L.b1 + (j / L.jt1) * JP + (j % L.jt1)] = th[P.boff[0] + j];
        for (int j = threadIdx.x; j < L.n2; j += NTHR) sm[L.b2 + (j / L.jt2) * JP + (j % L.jt2)] = th[P.boff[1] + j];
        for (int j = threadIdx.x; j < D; j += NTHR) sm[L.b3 + (j / L.jt3) * JP3 + (j % L.jt3)] = th[P.boff[2] + j];
    }
    __syncthreads();

}

// readout: one thread per sample (inference_sol, reg_z_aug, generate_sol); E = n = 0 in TestMode.  The caller has made the
// final state visible grid-wide.
template <bool EXACT>
__device__ __forceinline__ void readout(const Params& P, int cur, float span, int nacc) {
    const SolveArgs& a = P.a;
    const Layout& L = P.L;
    const int D = L.D, S = D + 3;
    const long long B = a.B;
    for (long long b = (long long)blockIdx.x * NTHR + threadIdx.x; b < B; b += (long long)gridDim.x * NTHR) {
        float zz = 0.f, za = 0.f, l = 0.f;
        const bool moved = span > 0.0f;
        for (int j = 0; j < D; ++j) {
            const float v = moved ? a.wu[cur][(long long)j * B + b] : input_value(a, b, j, D, P.nvars);
            zz = fmaf(v, v, zz);
            if (j >= P.nvars) za = fmaf(v, v, za);
            if (a.out_u) a.out_u[b * S + j] = v;
            if (a.out_x && j < P.nvars) a.out_x[b * P.nvars + j] = v;
            if (a.ckpt) a.ckpt[(((long long)nacc * 6) * D + j) * B + b] = v;     // slot `nacc`, stage 0 = the final state
        }
        float E = 0.f, n = 0.f;
        if (moved) {
            l = a.wu[cur][(long long)D * B + b];
            if constexpr (!EXACT) { E = a.wu[cur][(long long)(D + 1) * B + b]; n = a.wu[cur][(long long)(D + 2) * B + b]; }
        } else if (a.in_kind == IN_U0) {
            l = a.in[b * S + D]; E = a.in[b * S + D + 1]; n = a.in[b * S + D + 2];
        }
        if (a.out_u) { a.out_u[b * S + D] = l; a.out_u[b * S + D + 1] = E; a.out_u[b * S + D + 2] = n; }
        const float logp = -0.91893853320467274178f * (float)D - 0.5f * zz - l;
        const float Aa = a.reg_a ? tiny::vec_norm(za, a.squared) : 0.0f;
        if (a.out_logp) a.out_logp[b] = logp;
        if (a.out_regs) { a.out_regs[b * 3] = E; a.out_regs[b * 3 + 1] = n; a.out_regs[b * 3 + 2] = Aa; }
        if (a.out_lossterm) a.out_lossterm[b] = -logp + a.lam1 * E + a.lam2 * n + a.lam3 * Aa;
    }
}

template <int JP, int JP3, int ACT, bool EXACT>
__global__ void __launch_bounds__(NTHR, 1) solve_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) float sm[];
    __shared__ double sred[2 * NW + 2];
    const SolveArgs& a = P.a;
    const Layout& L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = L.D, S = D + 3;
    const long long B = a.B;
    tiny::GridReducer red{cg::this_grid(), a.partials, sred, 0, &a.xg, 0u, false};
    if (P.adaptive && a.xg.nranks > 1) red.seq = *reinterpret_cast<const volatile unsigned*>(a.xg.peers.p[a.xg.rank]);

    stage_weights<JP, JP3, EXACT>(P, sm);

    const float tdir = (a.t1 >= a.t0) ? 1.0f : -1.0f;
    const float span = fabsf(a.t1 - a.t0);
    const Controller ctl = a.ctl;
    const double inv_count = 1.0 / ((double)(a.norm_B > 0 ? a.norm_B : B) * (double)S);
    const long long ntiles = (B + NS - 1) / NS;
    enum { P_INIT = 0, P_PROBE = 1, P_STEP = 2 };
    int phase = P_INIT, cur = 0;
    int nacc = 0, nrej = 0, nf = 0, status = ICNF_OK, attempts = 0, fixed_step = 0;
    float t = a.t0, dt = (a.dt > 0.0f) ? fminf(a.dt, span) : 0.0f, dt0 = 0.0f, d1n = 0.0f;
    float qold = ctl.qoldinit, dt_last = 0.0f, hmag = 0.0f;
    bool last = false;
    // rows of the state this thread owns: j = warp * jt3 + r, r < jt3 (and j < D); samples 4 lane .. 4 lane + 3 of the tile
    bool own[JP3];
#pragma unroll
    for (int r = 0; r < JP3; ++r) own[r] = r < L.jt3 && warp * L.jt3 + r < D;

    while (span > 0.0f) {
        float h = 0.0f;
        if (phase == P_STEP) {
            if (P.adaptive) {
                const float remaining = fabsf(a.t1 - t);
                if (remaining <= 1e-7f * fmaxf(1.0f, fabsf(a.t1))) break;
                last = dt >= remaining * (1.0f - 1e-6f);
                hmag = last ? remaining : dt;
                h = tdir * hmag;
                if (!(hmag > 0.0f) || t + h == t) { status = ICNF_ERR_DT_UNDERFLOW; break; }
                if (++attempts > ctl.max_steps || (a.ckpt && nacc >= a.max_ckpt_steps)) { status = ICNF_ERR_MAX_STEPS; break; }
            } else {
                if (fixed_step >= a.nsteps) break;
                const float tb = fminf(span, fixed_step * a.dt);
                hmag = fminf(a.dt, span - tb);
                h = tdir * hmag;
                t = a.t0 + tdir * tb;
                last = false;
            }
        } else if (phase == P_PROBE) {
            h = tdir * dt0;
        }
        const int nstage = (phase == P_STEP) ? 6 : 1;
        double acc_a = 0.0, acc_b = 0.0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long b0 = tile * NS + 4 * lane;          // first of this lane's four samples
            const bool in4 = b0 + 3 < B;                        // all four exist (B is not required to be a multiple of 4)
            auto ld4 = [&](const float* base, int row) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* p = base + (long long)row * B + b0;
                if (in4 && ((B & 3) == 0)) v = *reinterpret_cast<const float4*>(p);
                else { if (b0 < B) v.x = p[0]; if (b0 + 1 < B) v.y = p[1]; if (b0 + 2 < B) v.z = p[2]; if (b0 + 3 < B) v.w = p[3]; }
                return v;
            };
            auto st4 = [&](float* base, int row, float4 v) {
                float* p = base + (long long)row * B + b0;
                if (in4 && ((B & 3) == 0)) *reinterpret_cast<float4*>(p) = v;
                else { if (b0 < B) p[0] = v.x; if (b0 + 1 < B) p[1] = v.y; if (b0 + 2 < B) p[2] = v.z; if (b0 + 3 < B) p[3] = v.w; }
            };
            // ---- state of the tile: z rows (owners), l (warp 0)
            float z[JP3][SPT], k1[JP3][SPT], K[4][JP3][SPT], zs[JP3][SPT], ze[JP3][SPT];
            constexpr int NX = EXACT ? 1 : 3;     // rows l (, E, n) of the state: warp 0, lane = four samples
            float lv[NX][SPT], kl1[NX][SPT], sl[NX][SPT], el[NX][SPT];
#pragma unroll
            for (int x = 0; x < NX; ++x)
#pragma unroll
                for (int s = 0; s < SPT; ++s) { lv[x][s] = 0.f; kl1[x][s] = 0.f; sl[x][s] = 0.f; el[x][s] = 0.f; }
            if (phase == P_INIT) {
#pragma unroll
                for (int r = 0; r < JP3; ++r) {
                    const int j = warp * L.jt3 + r;
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        const long long b = b0 + s;
                        float v = 0.f;
                        if (own[r] && b < B) v = input_value(a, b, j, D, P.nvars);
                        z[r][s] = v;
                    }
                }
                if (warp == 0 && a.in_kind == IN_U0) {
#pragma unroll
                    for (int x = 0; x < NX; ++x)
#pragma unroll
                        for (int s = 0; s < SPT; ++s) if (b0 + s < B) lv[x][s] = __ldg(a.in + (b0 + s) * S + D + x);
                }
            } else {
#pragma unroll
                for (int r = 0; r < JP3; ++r) {
                    const int j = warp * L.jt3 + r;
                    float4 u = make_float4(0.f, 0.f, 0.f, 0.f), k = u;
                    if (own[r]) { u = ld4(a.wu[cur], j); k = ld4(a.wk[cur], j); }
                    z[r][0] = u.x; z[r][1] = u.y; z[r][2] = u.z; z[r][3] = u.w;
                    k1[r][0] = k.x; k1[r][1] = k.y; k1[r][2] = k.z; k1[r][3] = k.w;
                }
                if (warp == 0) {
#pragma unroll
                    for (int x = 0; x < NX; ++x) {
                        const float4 u = ld4(a.wu[cur], D + x), k = ld4(a.wk[cur], D + x);
                        lv[x][0] = u.x; lv[x][1] = u.y; lv[x][2] = u.z; lv[x][3] = u.w;
                        kl1[x][0] = k.x; kl1[x][1] = k.y; kl1[x][2] = k.z; kl1[x][3] = k.w;
                    }
                }
            }
            // conditioning rows of the network input are constant over the stages
            for (int i = threadIdx.x; i < L.C * NS; i += NTHR) {
                const int c = i / NS, s = i - c * NS;
                const long long b = tile * NS + s;
                sm[L.x + (D + L.tin + c) * NS + s] = b < B ? __ldg(a.ys + b * L.C + c) : 0.f;
            }
            if constexpr (!EXACT) fill_eps_tile(a, sm + L.eps, tile * NS, D);
            // training checkpoints: the inputs of the six stages of the step, ckpt[slot][stage][D'][B] (a rejected attempt's
            // slot is simply rewritten)
            auto ckpt_stage = [&](int stage) { return a.ckpt + ((long long)nacc * 6 + stage) * D * B; };
            if (phase == P_STEP) {
                if (a.ckpt) {
#pragma unroll
                    for (int r = 0; r < JP3; ++r)
                        if (own[r]) st4(ckpt_stage(0), warp * L.jt3 + r, make_float4(z[r][0], z[r][1], z[r][2], z[r][3]));
                }
#pragma unroll
                for (int r = 0; r < JP3; ++r)
#pragma unroll
                    for (int s = 0; s < SPT; ++s) { zs[r][s] = c_a[6][0] * k1[r][s]; ze[r][s] = c_bt[0] * k1[r][s]; }
#pragma unroll
                for (int x = 0; x < NX; ++x)
#pragma unroll
                    for (int s = 0; s < SPT; ++s) { sl[x][s] = c_a[6][0] * kl1[x][s]; el[x][s] = c_bt[0] * kl1[x][s]; }
            }
            float zd[JP3][SPT], kl[NX][SPT];
#pragma unroll
            for (int x = 0; x < NX; ++x)
#pragma unroll
                for (int s = 0; s < SPT; ++s) kl[x][s] = 0.f;
            for (int sidx = 0; sidx < nstage; ++sidx) {
                const int i = sidx + 1;   // Tsit5 stage (1..6) when stepping
                // ---- stage input -> shared memory
                float tt = t;
                if (phase == P_PROBE) tt = t + h;
                else if (phase == P_STEP) tt = (i == 6 && last) ? a.t1 : fmaf(c_c[i], h, t);
#pragma unroll
                for (int r = 0; r < JP3; ++r) {
                    if (!own[r]) continue;
                    float xi[SPT];
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        float v = z[r][s];
                        if (phase == P_PROBE) v = fmaf(h, k1[r][s], v);
                        else if (phase == P_STEP) {
                            if (i == 6) v = fmaf(h, zs[r][s], v);      // u_new (b = a7: the FSAL stage)
                            else {
                                v = fmaf(h * c_a[i][0], k1[r][s], v);
#pragma unroll
                                for (int jj = 1; jj < 5; ++jj)
                                    if (jj < i) v = fmaf(h * c_a[i][jj], K[jj - 1][r][s], v);
                            }
                        }
                        xi[s] = v;
                    }
                    *reinterpret_cast<float4*>(sm + L.x + (warp * L.jt3 + r) * NS + 4 * lane) = make_float4(xi[0], xi[1], xi[2], xi[3]);
                    if (a.ckpt && phase == P_STEP && i < 6) st4(ckpt_stage(i), warp * L.jt3 + r, make_float4(xi[0], xi[1], xi[2], xi[3]));
                }
                if (L.tin && warp == 1) *reinterpret_cast<float4*>(sm + L.x + D * NS + 4 * lane) = make_float4(tt, tt, tt, tt);
                __syncthreads();
                if constexpr (EXACT) {
                    rhs_tile<JP, JP3, ACT>(L, sm, warp, lane, zd);
                    if (warp == 0) {
                        const float4 tr = *reinterpret_cast<const float4*>(sm + L.red + 4 * lane);
                        kl[0][0] = -tr.x; kl[0][1] = -tr.y; kl[0][2] = -tr.z; kl[0][3] = -tr.w;
                    }
                } else {
                    rhs_tile_hutch<JP, JP3, ACT>(L, sm, warp, lane, a.reg_e, a.reg_n, a.squared, zd);
                    if (warp == 0) {
#pragma unroll
                        for (int x = 0; x < NX; ++x) {
                            const float4 v = *reinterpret_cast<const float4*>(sm + L.red + x * NS + 4 * lane);
                            kl[x][0] = v.x; kl[x][1] = v.y; kl[x][2] = v.z; kl[x][3] = v.w;
                        }
                    }
                }
                // ---- stage bookkeeping (stages 1..5 of a step: k2..k6)
                if (phase == P_STEP && i < 6) {
                    const float bi = c_a[6][i], bti = c_bt[i];
#pragma unroll
                    for (int r = 0; r < JP3; ++r)
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            if (i < 5) K[i - 1][r][s] = zd[r][s];
                            zs[r][s] = fmaf(bi, zd[r][s], zs[r][s]);
                            ze[r][s] = fmaf(bti, zd[r][s], ze[r][s]);
                        }
#pragma unroll
                    for (int x = 0; x < NX; ++x)
#pragma unroll
                        for (int s = 0; s < SPT; ++s) { sl[x][s] = fmaf(bi, kl[x][s], sl[x][s]); el[x][s] = fmaf(bti, kl[x][s], el[x][s]); }
                }
            }
            // ---- epilogue of the phase for this tile
            if (phase == P_INIT) {
#pragma unroll
                for (int r = 0; r < JP3; ++r) {
                    if (!own[r]) continue;
                    const int j = warp * L.jt3 + r;
                    st4(a.wu[0], j, make_float4(z[r][0], z[r][1], z[r][2], z[r][3]));
                    st4(a.wk[0], j, make_float4(zd[r][0], zd[r][1], zd[r][2], zd[r][3]));
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        if (b0 + s >= B) continue;
                        const float sk = ctl.abstol + fabsf(z[r][s]) * ctl.reltol;
                        acc_a += (double)((z[r][s] / sk) * (z[r][s] / sk));
                        acc_b += (double)((zd[r][s] / sk) * (zd[r][s] / sk));
                    }
                }
                if (warp == 0) {
#pragma unroll
                    for (int x = 0; x < NX; ++x) {
                        st4(a.wu[0], D + x, make_float4(lv[x][0], lv[x][1], lv[x][2], lv[x][3]));
                        st4(a.wk[0], D + x, make_float4(kl[x][0], kl[x][1], kl[x][2], kl[x][3]));
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            if (b0 + s >= B) continue;
                            const float sk = ctl.abstol + fabsf(lv[x][s]) * ctl.reltol;
                            acc_a += (double)((lv[x][s] / sk) * (lv[x][s] / sk));
                            acc_b += (double)((kl[x][s] / sk) * (kl[x][s] / sk));
                        }
                    }
                }
            } else if (phase == P_PROBE) {
#pragma unroll
                for (int r = 0; r < JP3; ++r) {
                    if (!own[r]) continue;
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        if (b0 + s >= B) continue;
                        const float sk = ctl.abstol + fabsf(z[r][s]) * ctl.reltol;
                        const float df = (zd[r][s] - k1[r][s]) / sk;
                        acc_a += (double)(df * df);
                    }
                }
                if (warp == 0) {
#pragma unroll
                    for (int x = 0; x < NX; ++x)
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            if (b0 + s >= B) continue;
                            const float sk = ctl.abstol + fabsf(lv[x][s]) * ctl.reltol;
                            const float df = (kl[x][s] - kl1[x][s]) / sk;
                            acc_a += (double)(df * df);
                        }
                }
            } else {
                // zd / kl hold the FSAL stage k7 = f(u_new)
#pragma unroll
                for (int r = 0; r < JP3; ++r) {
                    if (!own[r]) continue;
                    const int j = warp * L.jt3 + r;
                    float zn[SPT];
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        zn[s] = fmaf(h, zs[r][s], z[r][s]);
                        if (b0 + s < B) {
                            const float e = h * fmaf(c_bt[6], zd[r][s], ze[r][s]);
                            const float sk = ctl.abstol + fmaxf(fabsf(z[r][s]), fabsf(zn[s])) * ctl.reltol;
                            const float q = e / sk;
                            acc_a += (double)(q * q);
                        }
                    }
                    st4(a.wu[cur ^ 1], j, make_float4(zn[0], zn[1], zn[2], zn[3]));
                    st4(a.wk[cur ^ 1], j, make_float4(zd[r][0], zd[r][1], zd[r][2], zd[r][3]));
                }
                if (warp == 0) {
#pragma unroll
                    for (int x = 0; x < NX; ++x) {
                        float ln[SPT];
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            ln[s] = fmaf(h, sl[x][s], lv[x][s]);
                            if (b0 + s < B) {
                                const float e = h * fmaf(c_bt[6], kl[x][s], el[x][s]);
                                const float sk = ctl.abstol + fmaxf(fabsf(lv[x][s]), fabsf(ln[s])) * ctl.reltol;
                                const float q = e / sk;
                                acc_a += (double)(q * q);
                            }
                        }
                        st4(a.wu[cur ^ 1], D + x, make_float4(ln[0], ln[1], ln[2], ln[3]));
                        st4(a.wk[cur ^ 1], D + x, make_float4(kl[x][0], kl[x][1], kl[x][2], kl[x][3]));
                    }
                }
            }
        }
        // ---- control (identical arithmetic in every thread; the two extra rows E, n of the state are identically zero
        // in TestMode: they contribute nothing to the sums but count in the mean, as in the other families)
        if (!P.adaptive) {
            // fixed steps: a tile's state is written and read by the same threads of the same CTA (tile -> CTA is static),
            // so the steps need no grid-wide synchronisation at all
            if (phase == P_INIT) { nf = 1; phase = P_STEP; continue; }
            if (a.steps && blockIdx.x == 0 && threadIdx.x == 0) { a.steps[nacc].t = t; a.steps[nacc].dt = h; }
            nf += 6; nacc++; dt_last = h; fixed_step++; cur ^= 1;
            t = (fixed_step >= a.nsteps) ? a.t1 : t + h;
            continue;
        }
        double ta, tb;
        red.sum2(acc_a, acc_b, ta, tb, true);
        if (phase == P_INIT) {
            nf = 1;
            if (a.dt > 0.0f) phase = P_STEP;
            else {
                const float d0 = (float)sqrt(ta * inv_count);
                d1n = (float)sqrt(tb * inv_count);
                dt0 = (d0 < 1e-5f || d1n < 1e-5f) ? 1e-6f : 0.01f * d0 / d1n;
                dt0 = fminf(dt0, span);
                phase = P_PROBE;
            }
        } else if (phase == P_PROBE) {
            nf += 1;
            const float d2 = (float)sqrt(ta * inv_count) / dt0;
            const float dm = fmaxf(d1n, d2);
            const float dt1 = (dm <= 1e-15f) ? fmaxf(1e-6f, dt0 * 1e-3f) : exp10f(-(2.0f + log10f(dm)) / 6.0f);
            dt = fminf(fminf(100.0f * dt0, dt1), span);
            phase = P_STEP;
        } else {
            nf += 6;
            const float eest = (float)sqrt(ta * inv_count);
            if (!isfinite(eest)) { status = ICNF_ERR_NONFINITE; break; }
            const float q11 = eest > 0.0f ? powf(eest, ctl.beta1) : 0.0f;
            float q = q11 / powf(qold, ctl.beta2);
            q = fmaxf(1.0f / ctl.qmax, fminf(1.0f / ctl.qmin, q / ctl.gamma));
            if (eest <= 1.0f) {
                if (a.steps && blockIdx.x == 0 && threadIdx.x == 0) { a.steps[nacc].t = t; a.steps[nacc].dt = h; }
                nacc++;
                dt_last = h;
                t = last ? a.t1 : t + h;
                cur ^= 1;
                if (q >= ctl.qsteady_min && q <= ctl.qsteady_max) q = 1.0f;
                qold = fmaxf(eest, ctl.qoldinit);
                dt = hmag / q;
            } else {
                nrej++;
                dt = hmag / fminf(1.0f / ctl.qmin, q11 / ctl.gamma);
            }
        }
    }

    __threadfence();
    red.grid.sync();
    readout<EXACT>(P, cur, span, nacc);
    if (P.adaptive && a.xg.nranks > 1 && blockIdx.x == 0 && threadIdx.x == 0)
        *reinterpret_cast<volatile unsigned*>(a.xg.peers.p[a.xg.rank]) = red.seq;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats) {
        a.stats->naccept = nacc;
        a.stats->nreject = nrej;
        a.stats->nf = nf;
        a.stats->status = status;
        a.stats->t_final = span > 0.0f ? t : a.t1;
        a.stats->dt_last = dt_last;
    }
}

// ------------------------------------------------------------------ VCABM (the reference's default alg) on the tile layout
// Variable-order Adams PECE as in tiny_vcabm.cuh (same coefficients, controller and order selection: oracle.vcabm_solve), with
// the RHS evaluated per 128-sample tile by the routines above.  Per attempt and tile: predictor from the stored Phi*(n-1)
// (HBM, [2][13][S][B], double-buffered against rejections), RHS, corrector, the four error estimates, and -- speculatively --
// the final evaluation at the corrected state.  Rows are owned as in solve_kernel (z rows by the thread that computes their
// derivative, the l / E / n rows by warp 0).
template <int JP, int JP3, int ACT, bool EXACT>
__global__ void __launch_bounds__(NTHR, 1) solve_vcabm_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) float sm[];
    __shared__ double sred[2 * NW + 2];
    __shared__ float s_beta[tiny::VC_MAXK + 2], s_g[tiny::VC_MAXK + 2];
    constexpr int NX = EXACT ? 1 : 3;
    constexpr int VK = tiny::VC_MAXK;
    const SolveArgs& a = P.a;
    const Layout& L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = L.D, S = D + 3;
    const long long B = a.B;
    tiny::GridReducer red{cg::this_grid(), a.partials, sred, 0, &a.xg, 0u, false};
    if (a.xg.nranks > 1) red.seq = *reinterpret_cast<const volatile unsigned*>(a.xg.peers.p[a.xg.rank]);
    stage_weights<JP, JP3, EXACT>(P, sm);

    const float tdir = (a.t1 >= a.t0) ? 1.0f : -1.0f;
    const float span = fabsf(a.t1 - a.t0);
    const Controller ctl = a.ctl;
    const double inv_count = 1.0 / ((double)(a.norm_B > 0 ? a.norm_B : B) * (double)S);
    const long long ntiles = (B + NS - 1) / NS;
    enum { P_INIT = 0, P_PROBE = 1, P_STEP = 2 };
    int phase = P_INIT, cur = 0;
    int nacc = 0, nrej = 0, nf = 0, status = ICNF_OK, attempts = 0, k = 1, nstep = 0;
    float hist[VK + 2];
#pragma unroll
    for (int i = 0; i < VK + 2; ++i) hist[i] = 0.0f;
    float t = a.t0, dt = (a.dt > 0.0f) ? fminf(a.dt, span) : 0.0f, dt0 = 0.0f, d1n = 0.0f, dt_last = 0.0f;
    bool last = false;
    bool own[JP3];
#pragma unroll
    for (int r = 0; r < JP3; ++r) own[r] = r < L.jt3 && warp * L.jt3 + r < D;

    while (span > 0.0f) {
        float h = 0.0f;
        int kk = 1;
        if (phase == P_STEP) {
            const float remaining = fabsf(a.t1 - t);
            if (remaining <= 1e-7f * fmaxf(1.0f, fabsf(a.t1))) break;
            last = dt >= remaining * (1.0f - 1e-6f);
            h = tdir * (last ? remaining : dt);
            if (!(fabsf(h) > 0.0f) || t + h == t) { status = ICNF_ERR_DT_UNDERFLOW; break; }
            if (++attempts > ctl.max_steps) { status = ICNF_ERR_MAX_STEPS; break; }
            kk = min(k + 1, nstep + 1);
            __syncthreads();
            if (threadIdx.x == 0) {
                float dts[VK + 3];
                dts[0] = h;
                for (int i = 0; i < VK + 2; ++i) dts[i + 1] = hist[i];
                tiny::vcabm_coefficients(dts, k, kk, s_beta, s_g);
            }
            __syncthreads();
        } else if (phase == P_PROBE) {
            h = tdir * dt0;
        }
        const int neval = (phase == P_STEP) ? 2 : 1;
        double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long b0 = tile * NS + 4 * lane;
            const bool in4 = b0 + 3 < B;
            auto ld4 = [&](const float* base, long long row) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* p = base + row * B + b0;
                if (in4 && ((B & 3) == 0)) v = *reinterpret_cast<const float4*>(p);
                else { if (b0 < B) v.x = p[0]; if (b0 + 1 < B) v.y = p[1]; if (b0 + 2 < B) v.z = p[2]; if (b0 + 3 < B) v.w = p[3]; }
                return v;
            };
            auto st4 = [&](float* base, long long row, const float (&v)[SPT]) {
                float* p = base + row * B + b0;
                if (in4 && ((B & 3) == 0)) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
                else { if (b0 < B) p[0] = v[0]; if (b0 + 1 < B) p[1] = v[1]; if (b0 + 2 < B) p[2] = v[2]; if (b0 + 3 < B) p[3] = v[3]; }
            };
            auto hrow = [&](int par, int j, int row) { return ((long long)par * (VK + 1) + j) * S + row; };   // row index into vc_hist
            auto get4 = [&](const float4& v, float (&o)[SPT]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; };
            // rows of this thread: slots 0 .. JP3-1 = z rows (owners), slots JP3 .. JP3+NX-1 = l (, E, n) rows (warp 0)
            constexpr int NR = JP3 + NX;
            bool mine[NR];
            int grow[NR];
#pragma unroll
            for (int r = 0; r < JP3; ++r) { mine[r] = own[r]; grow[r] = warp * L.jt3 + r; }
#pragma unroll
            for (int x = 0; x < NX; ++x) { mine[JP3 + x] = warp == 0; grow[JP3 + x] = D + x; }
            float u[NR][SPT], fn[NR][SPT], pv[NR][SPT], fo[NR][SPT];
#pragma unroll
            for (int r = 0; r < NR; ++r)
#pragma unroll
                for (int s = 0; s < SPT; ++s) { u[r][s] = 0.f; fn[r][s] = 0.f; pv[r][s] = 0.f; fo[r][s] = 0.f; }
            if (phase == P_INIT) {
#pragma unroll
                for (int r = 0; r < JP3; ++r)
#pragma unroll
                    for (int s = 0; s < SPT; ++s)
                        if (mine[r] && b0 + s < B) u[r][s] = input_value(a, b0 + s, grow[r], D, P.nvars);
                if (warp == 0 && a.in_kind == IN_U0) {
#pragma unroll
                    for (int x = 0; x < NX; ++x)
#pragma unroll
                        for (int s = 0; s < SPT; ++s) if (b0 + s < B) u[JP3 + x][s] = __ldg(a.in + (b0 + s) * S + D + x);
                }
            } else {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if (mine[r]) { get4(ld4(a.wu[cur], grow[r]), u[r]); get4(ld4(a.wk[cur], grow[r]), fn[r]); }
            }
            for (int i = threadIdx.x; i < L.C * NS; i += NTHR) {
                const int c = i / NS, s = i - c * NS;
                const long long b = tile * NS + s;
                sm[L.x + (D + L.tin + c) * NS + s] = b < B ? __ldg(a.ys + b * L.C + c) : 0.f;
            }
            if constexpr (!EXACT) fill_eps_tile(a, sm + L.eps, tile * NS, D);
            // ---- predictor (P_STEP); the evaluation point of the other phases
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                if (!mine[r]) continue;
                if (phase == P_INIT) {
#pragma unroll
                    for (int s = 0; s < SPT; ++s) pv[r][s] = u[r][s];
                } else if (phase == P_PROBE) {
#pragma unroll
                    for (int s = 0; s < SPT; ++s) pv[r][s] = fmaf(h, fn[r][s], u[r][s]);
                } else {
                    float phi[SPT];
                    const float hg0 = h * s_g[0];
#pragma unroll
                    for (int s = 0; s < SPT; ++s) { phi[s] = fn[r][s]; pv[r][s] = fmaf(hg0, fn[r][s], u[r][s]); }
                    st4(a.vc_hist, hrow(cur ^ 1, 0, grow[r]), fn[r]);
                    for (int j = 1; j < kk; ++j) {
                        const float bj = s_beta[j], hg = (j < k) ? h * s_g[j] : 0.0f;
                        float hp[SPT], ps[SPT];
                        get4(ld4(a.vc_hist, hrow(cur, j - 1, grow[r])), hp);
#pragma unroll
                        for (int s = 0; s < SPT; ++s) { phi[s] -= hp[s]; ps[s] = bj * phi[s]; pv[r][s] = fmaf(hg, ps[s], pv[r][s]); }
                        st4(a.vc_hist, hrow(cur ^ 1, j, grow[r]), ps);
                    }
                }
            }
            float un[NR][SPT];
            for (int ev = 0; ev < neval; ++ev) {
                // ---- evaluation point -> shared memory, RHS
                const float tt = (phase == P_INIT) ? t : (ev == 1 && last) ? a.t1 : t + h;
#pragma unroll
                for (int r = 0; r < JP3; ++r)
                    if (mine[r]) {
                        const float (&y)[SPT] = (ev == 0) ? pv[r] : un[r];
                        *reinterpret_cast<float4*>(sm + L.x + grow[r] * NS + 4 * lane) = make_float4(y[0], y[1], y[2], y[3]);
                    }
                if (L.tin && warp == 1) *reinterpret_cast<float4*>(sm + L.x + D * NS + 4 * lane) = make_float4(tt, tt, tt, tt);
                __syncthreads();
                float zd[JP3][SPT];
                if constexpr (EXACT) rhs_tile<JP, JP3, ACT>(L, sm, warp, lane, zd);
                else rhs_tile_hutch<JP, JP3, ACT>(L, sm, warp, lane, a.reg_e, a.reg_n, a.squared, zd);
#pragma unroll
                for (int r = 0; r < JP3; ++r)
#pragma unroll
                    for (int s = 0; s < SPT; ++s) fo[r][s] = zd[r][s];
                if (warp == 0) {
#pragma unroll
                    for (int x = 0; x < NX; ++x) {
                        const float4 v = *reinterpret_cast<const float4*>(sm + L.red + x * NS + 4 * lane);
                        const float sgn = EXACT ? -1.0f : 1.0f;      // rhs_tile leaves the trace, rhs_tile_hutch the derivative of l
                        fo[JP3 + x][0] = sgn * v.x; fo[JP3 + x][1] = sgn * v.y; fo[JP3 + x][2] = sgn * v.z; fo[JP3 + x][3] = sgn * v.w;
                    }
                }
                if (phase != P_STEP || ev == 1) break;
                // ---- corrector and error estimates (fo = f at the predictor)
                const float hgk = h * s_g[k], hdg = h * (s_g[k] - s_g[k - 1]);
                const float hm1 = h * tiny::c_gstar[k - 1], hm2 = (k >= 2) ? h * tiny::c_gstar[k - 2] : 0.0f;
                const bool has_p1 = (k < VK) && (kk >= k + 1);
                const float hp1 = has_p1 ? h * tiny::c_gstar[k + 1] : 0.0f;
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    if (!mine[r]) continue;
                    float php[SPT], pm1[SPT], pm2[SPT];
#pragma unroll
                    for (int s = 0; s < SPT; ++s) { php[s] = fo[r][s]; pm1[s] = 0.f; pm2[s] = 0.f; }
                    for (int j = 1; j <= k; ++j) {
                        float hp[SPT];
                        get4(ld4(a.vc_hist, hrow(cur ^ 1, j - 1, grow[r])), hp);
#pragma unroll
                        for (int s = 0; s < SPT; ++s) { pm2[s] = pm1[s]; pm1[s] = php[s]; php[s] -= hp[s]; }
                    }
                    float hk[SPT] = {0.f, 0.f, 0.f, 0.f};
                    if (has_p1) get4(ld4(a.vc_hist, hrow(cur ^ 1, k, grow[r])), hk);
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        un[r][s] = fmaf(hgk, php[s], pv[r][s]);
                        if (b0 + s < B) {
                            const float sk = ctl.abstol + fmaxf(fabsf(u[r][s]), fabsf(un[r][s])) * ctl.reltol;
                            const float q0 = hdg * php[s] / sk, q1 = hm1 * pm1[s] / sk, q2 = hm2 * pm2[s] / sk;
                            e0 += (double)(q0 * q0); e1 += (double)(q1 * q1); e2 += (double)(q2 * q2);
                            if (has_p1) { const float q3 = hp1 * (php[s] - hk[s]) / sk; e3 += (double)(q3 * q3); }
                        }
                    }
                }
            }
            // ---- per-phase epilogue of the tile
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                if (!mine[r]) continue;
                if (phase == P_INIT) {
                    st4(a.wu[0], grow[r], u[r]);
                    st4(a.wk[0], grow[r], fo[r]);
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        if (b0 + s >= B) continue;
                        const float sk = ctl.abstol + fabsf(u[r][s]) * ctl.reltol;
                        e0 += (double)((u[r][s] / sk) * (u[r][s] / sk));
                        e1 += (double)((fo[r][s] / sk) * (fo[r][s] / sk));
                    }
                } else if (phase == P_PROBE) {
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        if (b0 + s >= B) continue;
                        const float sk = ctl.abstol + fabsf(u[r][s]) * ctl.reltol;
                        const float df = (fo[r][s] - fn[r][s]) / sk;
                        e0 += (double)(df * df);
                    }
                } else {
                    st4(a.wu[cur ^ 1], grow[r], un[r]);
                    st4(a.wk[cur ^ 1], grow[r], fo[r]);
                }
            }
        }
        double t0s, t1s, t2s = 0.0, t3s = 0.0;
        red.sum2(e0, e1, t0s, t1s, true);
        if (phase == P_STEP) red.sum2(e2, e3, t2s, t3s, true);
        // ---- control: tiny_vcabm.cuh, identical arithmetic in every thread
        if (phase == P_INIT) {
            nf = 1;
            if (a.dt > 0.0f) phase = P_STEP;
            else {
                const float d0 = (float)sqrt(t0s * inv_count);
                d1n = (float)sqrt(t1s * inv_count);
                dt0 = (d0 < 1e-5f || d1n < 1e-5f) ? 1e-6f : 0.01f * d0 / d1n;
                dt0 = fminf(dt0, span);
                phase = P_PROBE;
            }
        } else if (phase == P_PROBE) {
            nf += 1;
            const float d2 = (float)sqrt(t0s * inv_count) / dt0;
            const float dm = fmaxf(d1n, d2);
            const float dt1 = (dm <= 1e-15f) ? fmaxf(1e-6f, dt0 * 1e-3f) : exp10f(-(2.0f + log10f(dm)) / 2.0f);
            dt = fminf(fminf(100.0f * dt0, dt1), span);
            phase = P_STEP;
        } else {
            float eest = (float)sqrt(t0s * inv_count);
            if (!isfinite(eest)) { status = ICNF_ERR_NONFINITE; break; }
            const float hmag = fabsf(h);
            nf += 2;
            if (eest > 1.0f) {
                nrej++;
                const float q = powf(eest, 1.0f / (float)(k + 1)) / ctl.gamma;
                dt = hmag / fminf(1.0f / ctl.qmin, fmaxf(1.0f / ctl.qmax, q));
            } else {
                int knew = k;
                if (nstep + 1 <= 4 || k < 3) knew = min(min(k + 1, 3), VK);
                else {
                    const float errm1 = (float)sqrt(t1s * inv_count), errm2 = (float)sqrt(t2s * inv_count);
                    if (fmaxf(errm1, errm2) <= eest) knew = k - 1;
                    else if (k < VK && kk >= k + 1) {
                        const float errp1 = (float)sqrt(t3s * inv_count);
                        if (errp1 < eest) { knew = k + 1; eest = errp1; }
                    }
                }
                nacc++;
                dt_last = h;
                t = last ? a.t1 : t + h;
                cur ^= 1;
#pragma unroll
                for (int i = VK + 1; i > 0; --i) hist[i] = hist[i - 1];
                hist[0] = h;
                nstep++;
                k = knew;
                float q = eest > 0.0f ? powf(eest, 1.0f / (float)(k + 1)) / ctl.gamma : 0.0f;
                q = fmaxf(1.0f / ctl.qmax, fminf(1.0f / ctl.qmin, q));
                if (q >= ctl.qsteady_min && q <= ctl.qsteady_max) q = 1.0f;
                dt = hmag / q;
            }
        }
    }
    __threadfence();
    red.grid.sync();
    readout<EXACT>(P, cur, span, 0);
    if (a.xg.nranks > 1 && blockIdx.x == 0 && threadIdx.x == 0)
        *reinterpret_cast<volatile unsigned*>(a.xg.peers.p[a.xg.rank]) = red.seq;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats) {
        a.stats->naccept = nacc;
        a.stats->nreject = nrej;
        a.stats->nf = nf;
        a.stats->status = status;
        a.stats->t_final = span > 0.0f ? t : a.t1;
        a.stats->dt_last = dt_last;
    }
}


template <int JP, int JP3, int ACT, bool EXACT>
static cudaError_t launch(const Params& P, int grid, size_t smem, cudaStream_t st) {
    static size_t attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr[dev & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(solve_kernel<JP, JP3, ACT, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr[dev & 63] = smem;
    }
    void* args[] = {(void*)&P};
    return cudaLaunchCooperativeKernel((const void*)solve_kernel<JP, JP3, ACT, EXACT>, dim3(grid), dim3(NTHR), args, smem, st);
}


template <int JP, int JP3, int ACT, bool EXACT>
static cudaError_t launch_vcabm(const Params& P, int grid, size_t smem, cudaStream_t st) {
    static size_t attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr[dev & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(solve_vcabm_kernel<JP, JP3, ACT, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr[dev & 63] = smem;
    }
    void* args[] = {(void*)&P};
    return cudaLaunchCooperativeKernel((const void*)solve_vcabm_kernel<JP, JP3, ACT, EXACT>, dim3(grid), dim3(NTHR), args, smem, st);
}

// one translation unit per (JP3, ACT, EXACT): the kernel is large and ptxas takes ~15 s per instantiation
#define ICNF_NARROW_INSTANCE(NAME, JP3V, ACTV, EXACTV)                                                            \
    cudaError_t NAME(const Params& P, int JP, int grid, size_t smem, cudaStream_t st) {                           \
        switch (JP) {                                                                                             \
            case 4: return launch<4, JP3V, ACTV, EXACTV>(P, grid, smem, st);                                      \
            case 8: return launch<8, JP3V, ACTV, EXACTV>(P, grid, smem, st);                                      \
            case 10: return launch<10, JP3V, ACTV, EXACTV>(P, grid, smem, st);                                    \
            case 16: return launch<16, JP3V, ACTV, EXACTV>(P, grid, smem, st);                                    \
            default: return cudaErrorInvalidConfiguration;                                                        \
        }                                                                                                         \
    }

// VCABM: run-time activation only (it is the compatibility path for default `sol_kwargs`, not a throughput path)
#define ICNF_NARROW_INSTANCE_VCABM(NAME, JP3V, EXACTV)                                                            \
    cudaError_t NAME(const Params& P, int JP, int grid, size_t smem, cudaStream_t st) {                           \
        switch (JP) {                                                                                             \
            case 4: return launch_vcabm<4, JP3V, -1, EXACTV>(P, grid, smem, st);                                  \
            case 8: return launch_vcabm<8, JP3V, -1, EXACTV>(P, grid, smem, st);                                  \
            case 10: return launch_vcabm<10, JP3V, -1, EXACTV>(P, grid, smem, st);                                \
            case 16: return launch_vcabm<16, JP3V, -1, EXACTV>(P, grid, smem, st);                                \
            default: return cudaErrorInvalidConfiguration;                                                        \
        }                                                                                                         \
    }

}  // namespace narrow
}  // namespace icnf
