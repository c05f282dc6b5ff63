// common.cuh -- shared device helpers: activations, Philox4x32-10, Tsit5 tableau,
// argument blocks passed from the C ABI (api.cu) to the kernel families.
//
// Reference behaviour implemented here (paths relative to the reference root):
//   NNlib.softplus / tanh / sigmoid as used by the Dense layers (src/core/icnf.jl:67-71)
//   Tsit5 coefficients and controller: SURVEY.md Appendix A (third-party OrdinaryDiffEq)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/icnf_b200.h"

namespace icnf {

// ---------------------------------------------------------------- static_for
template <int I, int N, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
// descending: N-1 ... 0
template <int I, class F>
__host__ __device__ __forceinline__ void static_rfor(F&& f) {
    if constexpr (I > 0) {
        f(std::integral_constant<int, I - 1>{});
        static_rfor<I - 1>(f);
    }
}

// ---------------------------------------------------------------- activations
// act_eval returns h = sigma(a) and d = sigma'(a); act_dd returns sigma''(a) from (h, d).
template <int ACT>
__device__ __forceinline__ void act_eval(float a, float& h, float& d) {
    if constexpr (ACT == ICNF_ACT_SOFTPLUS) {
        // softplus(a) = max(a,0) + log1p(exp(-|a|)); sigmoid(a) = [a>=0 ? 1 : e] / (1+e)
        float e = __expf(-fabsf(a));
        float t = 1.0f + e;
        float r = __fdividef(1.0f, t);
        h = fmaxf(a, 0.0f) + __logf(t);
        d = (a >= 0.0f) ? r : e * r;
    } else if constexpr (ACT == ICNF_ACT_TANH) {
        // tanh(a) = sign(a) (1 - e) / (1 + e), e = exp(-2|a|)
        float e = __expf(-2.0f * fabsf(a));
        float th = __fdividef(1.0f - e, 1.0f + e);
        h = copysignf(th, a);
        d = 1.0f - h * h;
    } else if constexpr (ACT == ICNF_ACT_SIGMOID) {
        float e = __expf(-fabsf(a));
        float r = __fdividef(1.0f, 1.0f + e);
        h = (a >= 0.0f) ? r : e * r;
        d = h * (1.0f - h);
    } else {
        h = a;
        d = 1.0f;
    }
}
template <int ACT>
__device__ __forceinline__ float act_dd(float h, float d) {
    if constexpr (ACT == ICNF_ACT_SOFTPLUS) return d * (1.0f - d);
    else if constexpr (ACT == ICNF_ACT_TANH) return -2.0f * h * d;
    else if constexpr (ACT == ICNF_ACT_SIGMOID) return d * (1.0f - 2.0f * h);
    else return 0.0f;
}
// runtime-dispatched versions for the generic kernels
__device__ __forceinline__ void act_eval_rt(int act, float a, float& h, float& d) {
    switch (act) {
        case ICNF_ACT_SOFTPLUS: act_eval<ICNF_ACT_SOFTPLUS>(a, h, d); break;
        case ICNF_ACT_TANH: act_eval<ICNF_ACT_TANH>(a, h, d); break;
        case ICNF_ACT_SIGMOID: act_eval<ICNF_ACT_SIGMOID>(a, h, d); break;
        default: h = a; d = 1.0f; break;
    }
}
__device__ __forceinline__ float act_dd_rt(int act, float h, float d) {
    switch (act) {
        case ICNF_ACT_SOFTPLUS: return d * (1.0f - d);
        case ICNF_ACT_TANH: return -2.0f * h * d;
        case ICNF_ACT_SIGMOID: return d * (1.0f - 2.0f * h);
        default: return 0.0f;
    }
}

// ---------------------------------------------------------------- Philox4x32-10
// Same draw spec as oracle/philox.py: word (row r, sample b) =
// philox(ctr = (b_lo, b_hi, r / 4, stream), key = (seed_lo, seed_hi))[r % 4].
constexpr uint32_t PHILOX_STREAM_EPS = 0x45505331u;
constexpr uint32_t PHILOX_STREAM_BASE = 0x42415345u;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        if (r != 9) { k.x += 0x9E3779B9u; k.y += 0xBB67AE85u; }
    }
    return c;
}
__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint32_t stream, int64_t sample, int blk) {
    uint64_t b = (uint64_t)sample;
    return philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(b >> 32), (uint32_t)blk, stream),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// four values (rows 4*blk .. 4*blk+3) of the draw for one sample
__device__ __forceinline__ void philox_draw4(int kind, uint64_t seed, uint32_t stream, int64_t sample,
                                             int blk, float out[4]) {
    uint4 w = philox_block(seed, stream, sample, blk);
    if (kind == ICNF_EPS_RADEMACHER) {
        out[0] = (w.x >> 31) ? 1.0f : -1.0f;
        out[1] = (w.y >> 31) ? 1.0f : -1.0f;
        out[2] = (w.z >> 31) ? 1.0f : -1.0f;
        out[3] = (w.w >> 31) ? 1.0f : -1.0f;
    } else {
        const float s = 5.9604644775390625e-08f;  // 2^-24
        float u1 = ((float)(w.x >> 8) + 1.0f) * s, u2 = (float)(w.y >> 8) * s;
        float r = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincosf(6.2831853071795864769f * u2, &sn, &cs);
        out[0] = r * cs; out[1] = r * sn;
        u1 = ((float)(w.z >> 8) + 1.0f) * s; u2 = (float)(w.w >> 8) * s;
        r = sqrtf(-2.0f * logf(u1));
        sincosf(6.2831853071795864769f * u2, &sn, &cs);
        out[2] = r * cs; out[3] = r * sn;
    }
}

// ---------------------------------------------------------------- Tsit5 tableau
struct Tsit5 {
    static constexpr int NS = 6;  // stages that enter the solution (k7 is FSAL / error only)
    __host__ __device__ static constexpr float c(int i) {
        constexpr float v[7] = {0.0f, 0.161f, 0.327f, 0.9f, 0.9800255409045097f, 1.0f, 1.0f};
        return v[i];
    }
    // a(i, j): stage i (0-based, 1..6) uses k_j, j < i.  Row 6 is the solution weights b.
    __host__ __device__ static constexpr float a(int i, int j) {
        constexpr float v[7][6] = {
            {0, 0, 0, 0, 0, 0},
            {0.161f, 0, 0, 0, 0, 0},
            {-0.008480655492356989f, 0.335480655492357f, 0, 0, 0, 0},
            {2.8971530571054935f, -6.359448489975075f, 4.3622954328695815f, 0, 0, 0},
            {5.325864828439257f, -11.748883564062828f, 7.4955393428898365f, -0.09249506636175525f, 0, 0},
            {5.86145544294642f, -12.92096931784711f, 8.159367898576159f, -0.071584973281401f,
             -0.028269050394068383f, 0},
            {0.09646076681806523f, 0.01f, 0.4798896504144996f, 1.379008574103742f, -3.290069515436081f,
             2.324710524099774f}};
        return v[i][j];
    }
    __host__ __device__ static constexpr float b(int j) { return a(6, j); }
    __host__ __device__ static constexpr float bt(int j) {
        constexpr float v[7] = {-0.00178001105222577714f, -0.0008164344596567469f, 0.007880878010261995f,
                                -0.1447110071732629f, 0.5823571654525552f, -0.45808210592918697f,
                                0.015151515151515152f};
        return v[j];
    }
};

// ---------------------------------------------------------------- argument blocks
constexpr int IN_U0 = 0;      // input has S rows (full state)
constexpr int IN_XS = 1;      // input has nvars rows; augmented rows and the 3 extra rows start at 0
constexpr int IN_Z0 = 2;      // input has D' rows (generate); the 3 extra rows start at 0
constexpr int IN_Z0_DRAW = 3; // z0 ~ N(0, I) drawn in-kernel (generate without a supplied base sample)

// ---- data-parallel exchange buffers (one per rank, mapped into every peer: CUDA IPC / peer access over NVLink) ----
// Layout (bytes) of a rank's buffer, R = XG_MAXR rank slots:
//   [0, 64)                                               header: word 0 = sequence number of the error-norm exchange
//   XG_OFF + ((par * R + r) * XG_FLOATS + p) * 4          gradient slice of rank r, parity par
//   XF_OFF + ((par * R + r) * XG_CTAS + c) * 4            flag of CTA c of rank r (= epoch when its slice is complete)
//   XN_OFF + (par * R + r) * 32                           error-norm record of rank r: {double a, double b, unsigned seq}
constexpr int XG_FLOATS = 4096, XG_CTAS = XG_FLOATS / 32, XG_MAXR = 16;
constexpr size_t XG_OFF = 64;
constexpr size_t XF_OFF = XG_OFF + (size_t)2 * XG_MAXR * XG_FLOATS * 4;
constexpr size_t XN_OFF = XF_OFF + (size_t)2 * XG_MAXR * XG_CTAS * 4;
constexpr size_t XBUF_BYTES = XN_OFF + (size_t)2 * XG_MAXR * 32 + 1024;

struct PeerTable { unsigned char* p[XG_MAXR]; };

// in-solve exchange of the adaptive controller's error sums: nranks <= 1 = off
struct NormXchg {
    PeerTable peers;
    int nranks, rank;
};

struct Controller {
    float reltol, abstol, beta1, beta2, gamma, qmin, qmax, qsteady_min, qsteady_max, qoldinit;
    int max_steps;
};

// written by the device loop, read by the host (or left on the device for _dev calls)
struct DevStats {
    int naccept, nreject, nf, status;
    float t_final, dt_last;
};

struct StepRec { float t, dt; };

struct SolveArgs {
    const float* theta;      // device, native ComponentArray order
    const float* in;         // see in_kind
    const float* eps;        // D' x B or null
    const float* ys;         // C x B or null
    float* out_u;            // S x B or null
    float* out_logp;         // B or null
    float* out_regs;         // 3 x B or null
    float* out_x;            // nvars x B or null (generate)
    float* out_lossterm;     // per-sample loss terms, or null
    float* out_loss;         // scalar loss = loss_scale * sum of the per-sample terms (families that fuse the sum), or null
    float loss_scale;
    float* ckpt;             // [step][B][D'] z at the start of every accepted step (training), or null
    StepRec* steps;          // accepted (t, dt), or null
    DevStats* stats;         // device
    // adaptive work space
    float* wu[2];            // S x B state double buffer
    float* wk[2];            // S x B k1 / k7 double buffer
    double* partials;        // [2][gridDim.x] squared-error partial sums
    int64_t B;
    int64_t sample_offset;
    uint64_t seed;
    int in_kind, eps_kind;
    int mode;                // icnf_mode
    int reg_e, reg_n, reg_a, squared;
    float lam1, lam2, lam3;
    float t0, t1;
    int nsteps;              // fixed-step count
    float dt;                // |dt| of a fixed step, or initial |dt| when adaptive (0 = automatic)
    int max_ckpt_steps;      // capacity of ckpt/steps
    Controller ctl;
    // data-parallel exact mode: the error norm is taken over the GLOBAL batch (one scalar pair per step attempt
    // exchanged through NVLink peer memory inside the device loop), so N shards take the steps of the unsharded solve
    NormXchg xg;
    int64_t norm_B;          // batch size in the error norm's mean (0 = B)
    // VCABM (alg = ICNF_ALG_VCABM): history of the modified divided differences, [2][13][S][B] floats
    float* vc_hist;
    int alg;
};

struct RhsArgs {
    const float* theta;
    const float* u;
    const float* eps;
    const float* ys;
    float* du;
    int64_t B;
    int mode, reg_e, reg_n, squared;
    float t;
};

struct BackwardArgs {
    const float* theta;
    const float* ckpt;       // [nsteps+1][B][D']: z at step starts, last = z(t_end)
    const StepRec* steps;    // [nsteps]
    const DevStats* stats;   // naccept = number of steps (device-resident; read by the kernel)
    const float* eps;
    const float* ys;
    float* gpartial;         // [n_partials][NP] per-warp/per-block partial gradients
    float* dxs;              // nvars x B or null
    int64_t B;
    int64_t sample_offset;
    uint64_t seed;
    int eps_kind;
    int mode, reg_e, reg_n, reg_a, squared;
    float lam1, lam2, lam3;
    float inv_denominator;   // 1 / global_batch
    int nvars;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace icnf
