// tiny_vcabm.cuh -- VCABM for the tiny family: variable-step, variable-order Adams-Bashforth-Moulton PECE, the
// reference's default `alg` (src/core/icnf.jl:89; third-party OrdinaryDiffEqAdamsBashforthMoulton, not vendored).
//
// The algorithm is the published one that package implements -- Hairer, Norsett, Wanner, "Solving ODEs I", III.5:
// modified divided differences Phi_j / Phi*_j, coefficients beta_j / g_j for arbitrary step-size histories, predictor of
// order k, one evaluation, corrector of order k + 1, error = difference of the two correctors; Shampine-Gordon order
// selection with the constant-step coefficients gamma*_j -- restated in oracle/icnf_oracle.py (vcabm_solve), which this
// kernel follows operation for operation.  [3P, from memory]: the rule for the first steps and the I-controller constants.
//
// One thread integrates one sample (as in tiny.cuh); one cooperative persistent kernel per solve.  Per step attempt:
// predictor from the stored Phi*(n-1), RHS, corrector, the four error estimates (orders k-2 .. k+1), and -- speculatively,
// rejections are rare -- the final PECE evaluation at the corrected state; then ONE pair of grid-wide reductions and the
// same controller arithmetic in every thread.  The history Phi*_j (j < 13, S rows per sample) lives in HBM, double
// buffered so that a rejected attempt leaves it untouched; coefficients are computed once per attempt and CTA.
#pragma once
#include "tiny.cuh"

namespace icnf {
namespace tiny {

constexpr int VC_MAXK = 12;
__constant__ float c_gstar[13] = {1.0f, -0.5f, -0.083333333333333333f, -0.041666666666666667f, -0.026388888888888889f,
                                  -0.01875f, -0.014269179894179894f, -0.011367394179894180f, -0.0093565365961199295f,
                                  -0.0078925540123456790f, -0.0067858499846542430f, -0.0059240564123376623f,
                                  -0.0052366692823494197f};

// beta_j (j < kk) and g_j (j <= k) for the step dts[0] after the earlier steps dts[1..]  (oracle: vcabm_coefficients)
__device__ inline void vcabm_coefficients(const float* dts, int k, int kk, float* beta, float* g) {
    const double h = dts[0];
    double b = 1.0, xi = h, xi0 = 0.0;
    beta[0] = 1.0f;
    for (int j = 1; j < kk; ++j) {
        xi0 += (double)dts[j];
        b = b * xi / xi0;
        beta[j] = (float)b;
        xi += (double)dts[j];
    }
    double c[VC_MAXK + 3];
    g[0] = 1.0f;
    for (int q = 0; q < k + 1; ++q) c[q] = 1.0 / ((double)(q + 1) * (double)(q + 2));
    if (k >= 1) g[1] = (float)c[0];
    double span = h;
    for (int j = 2; j <= k; ++j) {
        span += (double)dts[j - 1];
        for (int q = 0; q < k + 1 - (j - 1); ++q) c[q] = c[q] - c[q + 1] * h / span;
        g[j] = (float)c[0];
    }
}

template <class N, bool EXACT>
__global__ void __launch_bounds__(NTA, NTA_MINB) solve_vcabm_kernel(const __grid_constant__ WBlock<N> sw, const __grid_constant__ SolveArgs a,
                                                                     int nvars) {
    constexpr int S = N::S, D = N::D;
    __shared__ double sred[2 * (NTA / 32) + 2];
    __shared__ float s_beta[VC_MAXK + 2], s_g[VC_MAXK + 2];
    GridReducer red{cg::this_grid(), a.partials, sred, 0, &a.xg, 0u, false};
    if (a.xg.nranks > 1) red.seq = *reinterpret_cast<const volatile unsigned*>(a.xg.peers.p[a.xg.rank]);

    const float tdir = (a.t1 >= a.t0) ? 1.0f : -1.0f;
    const float span = fabsf(a.t1 - a.t0);
    const int64_t B = a.B;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t b0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const Controller ctl = a.ctl;
    const double inv_count = 1.0 / ((double)(a.norm_B > 0 ? a.norm_B : B) * (double)S);
    auto hidx = [&](int par, int j, int row, int64_t b) { return (((int64_t)par * (VC_MAXK + 1) + j) * S + row) * B + b; };
    enum { P_INIT = 0, P_PROBE = 1, P_STEP = 2 };
    int phase = P_INIT, cur = 0;
    int nacc = 0, nrej = 0, nf = 0, status = ICNF_OK, attempts = 0;
    int k = 1, nstep = 0;
    float hist[VC_MAXK + 2];
#pragma unroll
    for (int i = 0; i < VC_MAXK + 2; ++i) hist[i] = 0.0f;
    float t = a.t0, dt = (a.dt > 0.0f) ? fminf(a.dt, span) : 0.0f, dt0 = 0.0f, d1 = 0.0f, dt_last = 0.0f;
    bool last = false;

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
                float dts[VC_MAXK + 3];
                dts[0] = h;
                for (int i = 0; i < VC_MAXK + 2; ++i) dts[i + 1] = hist[i];
                vcabm_coefficients(dts, k, kk, s_beta, s_g);
            }
            __syncthreads();
        } else if (phase == P_PROBE) {
            h = tdir * dt0;
        }
        double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;   // P_STEP: squared scaled errors of orders k, k-1, k-2, k+1
        for (int64_t b = b0; b < B; b += stride) {
            Sample<N> sm;
            load_sample_consts<N>(a, b, sm);
            float u[S], fn[S], fo[S];
            if (phase == P_INIT) {
                float z0[D];
                load_state<N>(a, b, nvars, z0, u[D], u[D + 1], u[D + 2]);
#pragma unroll
                for (int j = 0; j < D; ++j) u[j] = z0[j];
            } else {
#pragma unroll
                for (int r = 0; r < S; ++r) { u[r] = a.wu[cur][(int64_t)r * B + b]; fn[r] = a.wk[cur][(int64_t)r * B + b]; }
            }
            auto eval = [&](const float (&y)[S], float tt, float (&out)[S]) {
#pragma unroll
                for (int j = 0; j < D; ++j) sm.x[j] = y[j];
                if constexpr (N::TIN) sm.x[D] = tt;
                float kz[D];
                rhs_eval<N, EXACT>(sw, sm.x, sm.eps, a.reg_e, a.reg_n, a.squared, kz, out[D], out[D + 1], out[D + 2]);
#pragma unroll
                for (int j = 0; j < D; ++j) out[j] = kz[j];
            };
            if (phase == P_INIT) {
                eval(u, t, fo);
#pragma unroll
                for (int r = 0; r < S; ++r) {
                    a.wu[0][(int64_t)r * B + b] = u[r];
                    a.wk[0][(int64_t)r * B + b] = fo[r];
                    const float sk = ctl.abstol + fabsf(u[r]) * ctl.reltol;
                    e0 += (double)((u[r] / sk) * (u[r] / sk));
                    e1 += (double)((fo[r] / sk) * (fo[r] / sk));
                }
            } else if (phase == P_PROBE) {
                float y[S];
#pragma unroll
                for (int r = 0; r < S; ++r) y[r] = fmaf(h, fn[r], u[r]);
                eval(y, t + h, fo);
#pragma unroll
                for (int r = 0; r < S; ++r) {
                    const float sk = ctl.abstol + fabsf(u[r]) * ctl.reltol;
                    const float df = (fo[r] - fn[r]) / sk;
                    e0 += (double)(df * df);
                }
            } else {
                // ---- predictor of order k; Phi*_j(n) for j < kk go to the other history buffer
                float phi[S], p[S];
                const float hg0 = h * s_g[0];
#pragma unroll
                for (int r = 0; r < S; ++r) {
                    phi[r] = fn[r];
                    p[r] = fmaf(hg0, fn[r], u[r]);
                    a.vc_hist[hidx(cur ^ 1, 0, r, b)] = fn[r];
                }
                for (int j = 1; j < kk; ++j) {
                    const float bj = s_beta[j], hg = (j < k) ? h * s_g[j] : 0.0f;
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        phi[r] -= a.vc_hist[hidx(cur, j - 1, r, b)];
                        const float ps = bj * phi[r];
                        a.vc_hist[hidx(cur ^ 1, j, r, b)] = ps;
                        p[r] = fmaf(hg, ps, p[r]);
                    }
                }
                float fp[S];
                eval(p, t + h, fp);
                // ---- Phi^p_j(n+1) = Phi^p_{j-1}(n+1) - Phi*_{j-1}(n); corrector with j = k
                float php[S], pm1[S], pm2[S];
#pragma unroll
                for (int r = 0; r < S; ++r) { php[r] = fp[r]; pm1[r] = 0.0f; pm2[r] = 0.0f; }
                for (int j = 1; j <= k; ++j) {
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        pm2[r] = pm1[r];
                        pm1[r] = php[r];
                        php[r] -= a.vc_hist[hidx(cur ^ 1, j - 1, r, b)];
                    }
                }
                // php = Phi^p_k, pm1 = Phi^p_{k-1}, pm2 = Phi^p_{k-2} (k >= 2)
                const float hgk = h * s_g[k], hdg = h * (s_g[k] - s_g[k - 1]);
                const float hm1 = h * c_gstar[k - 1], hm2 = (k >= 2) ? h * c_gstar[k - 2] : 0.0f;
                const bool has_p1 = (k < VC_MAXK) && (kk >= k + 1);
                const float hp1 = has_p1 ? h * c_gstar[k + 1] : 0.0f;
                float un[S];
#pragma unroll
                for (int r = 0; r < S; ++r) {
                    un[r] = fmaf(hgk, php[r], p[r]);
                    const float sk = ctl.abstol + fmaxf(fabsf(u[r]), fabsf(un[r])) * ctl.reltol;
                    const float q0 = hdg * php[r] / sk, q1 = hm1 * pm1[r] / sk, q2 = hm2 * pm2[r] / sk;
                    e0 += (double)(q0 * q0); e1 += (double)(q1 * q1); e2 += (double)(q2 * q2);
                    if (has_p1) {
                        const float q3 = hp1 * (php[r] - a.vc_hist[hidx(cur ^ 1, k, r, b)]) / sk;
                        e3 += (double)(q3 * q3);
                    }
                }
                // ---- final evaluation (the E of PECE), speculative: kept only if the step is accepted
                eval(un, last ? a.t1 : t + h, fo);
#pragma unroll
                for (int r = 0; r < S; ++r) { a.wu[cur ^ 1][(int64_t)r * B + b] = un[r]; a.wk[cur ^ 1][(int64_t)r * B + b] = fo[r]; }
            }
        }
        double t0s, t1s, t2s = 0.0, t3s = 0.0;
        red.sum2(e0, e1, t0s, t1s, true);
        if (phase == P_STEP) red.sum2(e2, e3, t2s, t3s, true);
        // ---- control (identical arithmetic in every thread)
        if (phase == P_INIT) {
            nf = 1;
            if (a.dt > 0.0f) phase = P_STEP;
            else {
                const float d0 = (float)sqrt(t0s * inv_count);
                d1 = (float)sqrt(t1s * inv_count);
                dt0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
                dt0 = fminf(dt0, span);
                phase = P_PROBE;
            }
        } else if (phase == P_PROBE) {
            nf += 1;
            const float d2 = (float)sqrt(t0s * inv_count) / dt0;
            const float dm = fmaxf(d1, d2);
            // starting step of an order-1 method: exponent 1 / (order + 1) = 1 / 2
            const float dt1 = (dm <= 1e-15f) ? fmaxf(1e-6f, dt0 * 1e-3f) : exp10f(-(2.0f + log10f(dm)) / 2.0f);
            dt = fminf(fminf(100.0f * dt0, dt1), span);
            phase = P_STEP;
        } else {
            float eest = (float)sqrt(t0s * inv_count);
            if (!isfinite(eest)) { status = ICNF_ERR_NONFINITE; break; }
            const float hmag = fabsf(h);
            if (eest > 1.0f) {
                nf += 2;      // the speculative final evaluation is counted where it is spent
                nrej++;
                const float q = powf(eest, 1.0f / (float)(k + 1)) / ctl.gamma;
                dt = hmag / fminf(1.0f / ctl.qmin, fmaxf(1.0f / ctl.qmax, q));
            } else {
                nf += 2;
                int knew = k;
                if (nstep + 1 <= 4 || k < 3) knew = min(min(k + 1, 3), VC_MAXK);
                else {
                    const float errm1 = (float)sqrt(t1s * inv_count), errm2 = (float)sqrt(t2s * inv_count);
                    if (fmaxf(errm1, errm2) <= eest) knew = k - 1;
                    else if (k < VC_MAXK && kk >= k + 1) {
                        const float errp1 = (float)sqrt(t3s * inv_count);
                        if (errp1 < eest) { knew = k + 1; eest = errp1; }
                    }
                }
                nacc++;
                dt_last = h;
                t = last ? a.t1 : t + h;
                cur ^= 1;
#pragma unroll
                for (int i = VC_MAXK + 1; i > 0; --i) hist[i] = hist[i - 1];
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

    // ---- readout
    double loss_local = 0.0;
    for (int64_t b = b0; b < B; b += stride) {
        float z[D], l, E, n;
        if (span > 0.0f) {
#pragma unroll
            for (int j = 0; j < D; ++j) z[j] = a.wu[cur][(int64_t)j * B + b];
            l = a.wu[cur][(int64_t)D * B + b]; E = a.wu[cur][(int64_t)(D + 1) * B + b]; n = a.wu[cur][(int64_t)(D + 2) * B + b];
        } else {
            load_state<N>(a, b, nvars, z, l, E, n);
        }
        loss_local += (double)write_outputs<N>(a, b, nvars, z, l, E, n);
    }
    if (a.out_loss) {
        double tot, unused;
        red.sum2(loss_local, 0.0, tot, unused);
        if (blockIdx.x == 0 && threadIdx.x == 0) a.out_loss[0] = (float)(tot * (double)a.loss_scale);
    }
    if (a.xg.nranks > 1 && blockIdx.x == 0 && threadIdx.x == 0)
        *reinterpret_cast<volatile unsigned*>(a.xg.peers.p[a.xg.rank]) = red.seq;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats) {
        a.stats->naccept = nacc;
        a.stats->nreject = nrej;
        a.stats->nf = nf;
        a.stats->status = status;
        a.stats->t_final = t;
        a.stats->dt_last = dt_last;
    }
}

}  // namespace tiny
}  // namespace icnf
