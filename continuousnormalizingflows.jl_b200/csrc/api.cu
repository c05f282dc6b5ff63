// api.cu -- the C ABI of libicnf_b200.so (include/icnf_b200.h): handle, buffers,
// host<->device staging, kernel-family dispatch, and the two small reduction
// kernels (loss mean, partial-gradient sum).  No CPU fallback lives here: every
// entry point either launches CUDA kernels or returns an error code.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "family.h"

namespace icnf {

std::vector<const Family*>& tiny_registry() {
    static std::vector<const Family*> reg;
    return reg;
}
const Family* generic_family();  // generic.cu

// ---------------------------------------------------------------- small kernels
// dtheta[p] = sum over partial rows, fixed order (bit-reproducible).  One CTA per 32
// parameters: 8 warps each sum a contiguous slice of the rows (coalesced across p), then
// the 8 partial sums are added in order.
__global__ void __launch_bounds__(256) reduce_grad_kernel(const float* __restrict__ partial, int nrows, int np,
                                                        float* __restrict__ out) {
    __shared__ float sh[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const int per = (nrows + 7) / 8;
    const int r0 = w * per, r1 = min(nrows, r0 + per);
    float s0 = 0.f, s1 = 0.f;
    if (p < np) {
        int r = r0;
        for (; r + 1 < r1; r += 2) {
            s0 += partial[(size_t)r * np + p];
            s1 += partial[(size_t)(r + 1) * np + p];
        }
        if (r < r1) s0 += partial[(size_t)r * np + p];
    }
    sh[w][lane] = s0 + s1;
    __syncthreads();
    if (w == 0 && p < np) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sh[i][lane];
        out[p] = s;
    }
}

// ---- data-parallel exchange through NVLink peer memory (small gradients) -------------------------------------------
// Every rank owns one exchange buffer (layout: common.cuh); `peers[r]` is rank r's buffer mapped into this process
// (CUDA IPC between processes, direct peer access inside one process).
// One kernel = the local gradient reduction AND the all-reduce: CTA c sums parameters [32 c, 32 c + 32) over the
// partial rows (fixed order), stores the slice into EVERY rank's exchange buffer (remote stores over NVLink), raises
// its flag there, waits for the same slice of every other rank to land in its own buffer and adds the slices in rank
// order -- every rank ends with bit-identical sums.  Element np of the vector is the scalar loss.
__global__ void __launch_bounds__(256) reduce_grad_xchg_kernel(const float* __restrict__ partial, int nrows, int np,
                                                             const float* __restrict__ loss_in, float* __restrict__ dtheta,
                                                             float* __restrict__ loss_out, PeerTable peers, int nranks, int rank,
                                                             unsigned epoch, int* __restrict__ timeout_flag) {
    __shared__ float sh[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x, p = c * 32 + lane;
    const int par = (int)(epoch & 1u);
    const int per = (nrows + 7) / 8;
    const int r0 = w * per, r1 = min(nrows, r0 + per);
    float s0 = 0.f, s1 = 0.f;
    if (p < np) {
        int r = r0;
        for (; r + 1 < r1; r += 2) {
            s0 += partial[(size_t)r * np + p];
            s1 += partial[(size_t)(r + 1) * np + p];
        }
        if (r < r1) s0 += partial[(size_t)r * np + p];
    }
    sh[w][lane] = s0 + s1;
    __syncthreads();
    if (w != 0) return;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][lane];
    if (p == np) s = loss_in ? loss_in[0] : 0.f;
    // publish this slice to every rank (own buffer included), then the flag
    const size_t goff = XG_OFF + ((size_t)(par * nranks + rank) * XG_FLOATS + p) * 4;
    const size_t foff = XF_OFF + ((size_t)(par * nranks + rank) * XG_CTAS + c) * 4;
    if (p <= np)
        for (int r = 0; r < nranks; ++r) *reinterpret_cast<volatile float*>(peers.p[r] + goff) = s;
    __threadfence_system();
    __syncwarp();
    if (lane < nranks) *reinterpret_cast<volatile unsigned*>(peers.p[lane] + foff) = epoch;
    // wait for every rank's slice in MY buffer
    unsigned char* mine = peers.p[rank];
    bool ok = true;
    if (lane < nranks) {
        const volatile unsigned* f = reinterpret_cast<const volatile unsigned*>(mine + XF_OFF + ((size_t)(par * nranks + lane) * XG_CTAS + c) * 4);
        long long spins = 0;
        while (*f != epoch) {
            if (++spins > (1LL << 26)) { ok = false; break; }   // a peer never arrived (~ seconds): report, do not hang the GPU
            __nanosleep(64);
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    __threadfence_system();
    float tot = 0.f;
    if (p <= np)
        for (int r = 0; r < nranks; ++r)
            tot += *reinterpret_cast<const volatile float*>(mine + XG_OFF + ((size_t)(par * nranks + r) * XG_FLOATS + p) * 4);
    if (!ok) { tot = __int_as_float(0x7fc00000); if (lane == 0 && timeout_flag) *timeout_flag = 1; }
    if (p < np) dtheta[p] = tot;
    else if (p == np && loss_out) loss_out[0] = tot;
}

// out[0] = scale * sum(x[0..n)) with a fixed reduction tree; single CTA, double accumulation
__global__ void sum_kernel(const float* __restrict__ x, long long n, float scale, float* __restrict__ out) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += (double)x[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        out[0] = (float)(t * (double)scale);
    }
}

// OptimiserChain(WeightDecay(lambda), Adam(eta, (beta1, beta2), epsilon)) of the MLJ adapter
// (src/exts/mlj_ext/core_icnf.jl:17-24), Optimisers.jl semantics: g += lambda * theta, then Adam.
__global__ void adam_step_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float eta, float beta1, float beta2, float epsilon,
                                 float lambda, float bp1, float bp2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = fmaf(lambda, theta[i], grad[i]);
    const float mi = beta1 * m[i] + (1.0f - beta1) * g;
    const float vi = beta2 * v[i] + (1.0f - beta2) * g * g;
    m[i] = mi;
    v[i] = vi;
    theta[i] -= eta * (mi / (1.0f - bp1)) / (sqrtf(vi / (1.0f - bp2)) + epsilon);
}

// FP32 FMA-pipe microbenchmark (the roofline denominator of the narrow-MLP kernels, which
// MEASURED_PEAKS.json does not carry): 16 independent FFMA chains per thread.
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;   // never true; keeps the chain alive
}

}  // namespace icnf

using namespace icnf;

// ---------------------------------------------------------------- handle
namespace {

// NVTX range over one entry point (tracing): shows up in Nsight timelines, costs nothing without a tool attached
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return (T*)p; }
};

}  // namespace

struct icnf_handle {
    icnf_config cfg;
    const Family* fam = nullptr;
    void* ws = nullptr;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::vector<float> theta_host;
    bool have_params = false;
    DevBuf vc_hist;   // VCABM: modified divided differences of the last 13 steps
    DevBuf theta_dev, in, eps, ys, out0, out1, out2, wu0, wu1, wk0, wk1, partials, ckpt, steps, stats, gpartial,
        lossterm, scalar, dtheta, dxs;
    DevStats* stats_host = nullptr;  // pinned
    float* scalar_host = nullptr;    // pinned
    float* grad_host = nullptr;      // pinned staging for the gradient: a copy into the caller's pageable buffer would block per copy
    size_t grad_host_cap = 0;
    struct Group* grp = nullptr;     // data-parallel communicator (icnf_group_join / icnf_create_group), or null
    bool dp_active = false;          // inside icnf_loss_grad_dp*
    bool dp_defer_reduce = false;    // the data-parallel step fuses the partial-gradient reduction with the exchange
    int dp_nrows = 0;                // rows of gpartial left by the last backward launch
    DevBuf dpbuf;                    // [dtheta; loss] of the data-parallel step
    long long launches = 0;
    bool profiling = false;
    // ordering between the two kinds of entry points: `_dev` calls run on the caller's stream, host-pointer calls on
    // the handle's private stream, and both use the handle's work buffers.  Every `_dev` call records `dev_done`;
    // every host-pointer call makes the private stream wait for it first.
    cudaEvent_t dev_done = nullptr;
    bool dev_pending = false;
    cudaEvent_t prof_ev[4][2] = {};
    bool prof_used[4] = {false, false, false, false};
    std::string err;
    // slot: 0 forward solve (or rhs), 1 loss sum, 2 backward, 3 gradient reduce
    void prof_begin(int slot, cudaStream_t st) {
        if (!profiling) return;
        if (!prof_ev[slot][0]) { cudaEventCreate(&prof_ev[slot][0]); cudaEventCreate(&prof_ev[slot][1]); }
        cudaEventRecord(prof_ev[slot][0], st);
    }
    void prof_end(int slot, cudaStream_t st) {
        if (!profiling) return;
        cudaEventRecord(prof_ev[slot][1], st);
        prof_used[slot] = true;
    }

    int D() const { return cfg.nvars + cfg.naug; }
    int S() const { return D() + 3; }
    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what) {
        if (e == cudaErrorNotSupported)
            return fail(ICNF_ERR_UNSUPPORTED, "%s: this kernel family does not implement the request", what);
        return fail(ICNF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
};

#define CK(h, call)                                                  \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return (h)->cuda_fail(e__, #call);   \
    } while (0)

namespace {

Controller make_controller(const icnf_solver* s) {
    Controller c;
    c.reltol = (s && s->reltol > 0) ? s->reltol : 1e-4f;
    c.abstol = (s && s->abstol > 0) ? s->abstol : 1e-4f;
    c.beta1 = (s && s->beta1 > 0) ? s->beta1 : 7.0f / 50.0f;
    c.beta2 = (s && s->beta2 > 0) ? s->beta2 : 2.0f / 25.0f;
    c.gamma = (s && s->gamma > 0) ? s->gamma : 0.9f;
    c.qmin = (s && s->qmin > 0) ? s->qmin : 0.2f;
    c.qmax = (s && s->qmax > 0) ? s->qmax : 10.0f;
    c.qsteady_min = (s && s->qsteady_min > 0) ? s->qsteady_min : 1.0f;
    c.qsteady_max = (s && s->qsteady_max > 0) ? s->qsteady_max : 1.2f;
    c.qoldinit = (s && s->qoldinit > 0) ? s->qoldinit : 1e-4f;
    c.max_steps = (s && s->max_steps > 0) ? s->max_steps : 100000;
    return c;
}

struct ModeFlags {
    bool exact;
    int reg_e, reg_n, reg_a;
};
ModeFlags mode_flags(const icnf_config& c, int mode) {
    ModeFlags f;
    f.exact = (mode == ICNF_TEST);
    const bool reg = (mode == ICNF_TRAIN_REG);
    f.reg_e = reg && c.lambda1 != 0.0f;
    f.reg_n = reg && c.lambda2 != 0.0f;
    f.reg_a = reg && c.lambda3 != 0.0f && c.naug != 0;
    return f;
}

// end of a `_dev` entry point: later host-pointer calls must not touch the work buffers before this point
int mark_dev(icnf_handle* h, cudaStream_t st) {
    if (st == h->stream) return ICNF_OK;   // host entry points reuse the `_dev` bodies on the private stream
    if (!h->dev_done) CK(h, cudaEventCreateWithFlags(&h->dev_done, cudaEventDisableTiming));
    CK(h, cudaEventRecord(h->dev_done, st));
    h->dev_pending = true;
    return ICNF_OK;
}
// start of a host-pointer entry point
int join_dev(icnf_handle* h) {
    if (!h->dev_pending) return ICNF_OK;
    CK(h, cudaStreamWaitEvent(h->stream, h->dev_done, 0));
    h->dev_pending = false;
    return ICNF_OK;
}

int validate_common(icnf_handle* h, int mode, int64_t B) {
    if (!h) return ICNF_ERR_INVALID;
    if (!h->have_params) return h->fail(ICNF_ERR_INVALID, "icnf_set_params has not been called");
    if (mode < ICNF_TEST || mode > ICNF_TRAIN_NOREG) return h->fail(ICNF_ERR_INVALID, "bad mode %d", mode);
    if (B < 0) return h->fail(ICNF_ERR_INVALID, "negative batch");
    CK(h, cudaSetDevice(h->device));
    return ICNF_OK;
}

// Number of fixed steps for |t1 - t0| at step dt (last step clipped to land on t1).
int fixed_steps(float t0, float t1, float dt) {
    double span = std::fabs((double)t1 - (double)t0);
    if (span == 0.0) return 0;
    return (int)std::ceil(span / (double)dt - 1e-6);
}

bool group_wants_global_norm(const icnf_handle* h);
void group_fill_xchg(const icnf_handle* h, NormXchg& x);

struct SolveRequest {
    int mode;
    const icnf_solver* sol;
    float t0, t1;
    const float* in;
    int in_kind;
    const icnf_noise* noise;
    const float* eps;
    const float* ys;
    float *out_u, *out_logp, *out_regs, *out_x, *out_lossterm;
    bool want_ckpt;
    int64_t B;
    int64_t global_B = 0;        // data-parallel step: size of the unsharded batch (error norm of the exact mode)
    float* out_loss = nullptr;   // scalar loss (fused by the family when it can)
    float loss_scale = 0.f;
    bool loss_fused = false;     // set by enqueue_solve
};

// Enqueue one solve on `st`.  All pointers are device pointers.  Statistics land in
// h->stats (device).
int enqueue_solve(icnf_handle* h, SolveRequest& r, cudaStream_t st) {
    const icnf_config& c = h->cfg;
    const ModeFlags mf = mode_flags(c, r.mode);
    const int D = h->D(), S = h->S();
    const bool adaptive = r.sol ? (r.sol->adaptive != 0) : true;
    int eps_kind = r.noise ? r.noise->kind : ICNF_EPS_SUPPLIED;
    if (r.mode != ICNF_TEST && eps_kind == ICNF_EPS_SUPPLIED && !r.eps)
        return h->fail(ICNF_ERR_INVALID, "mode needs a Hutchinson probe: pass eps or a noise kind");
    if (eps_kind < ICNF_EPS_SUPPLIED || eps_kind > ICNF_EPS_RADEMACHER) return h->fail(ICNF_ERR_INVALID, "bad noise kind");
    if (c.ncond && !r.ys) return h->fail(ICNF_ERR_INVALID, "conditioned flow needs ys");
    if (!adaptive && !(r.sol && r.sol->dt > 0)) return h->fail(ICNF_ERR_INVALID, "fixed-step solve needs dt > 0");
    if (r.sol && r.sol->alg != ICNF_ALG_TSIT5 && r.sol->alg != ICNF_ALG_VCABM) return h->fail(ICNF_ERR_INVALID, "unknown alg %d", r.sol->alg);
    if (r.sol && r.sol->alg == ICNF_ALG_VCABM && !adaptive) return h->fail(ICNF_ERR_INVALID, "VCABM is a variable-step method: adaptive must be 1");

    SolveArgs a;
    memset(&a, 0, sizeof a);
    a.theta = h->theta_dev.as<float>();
    a.in = r.in; a.eps = r.eps; a.ys = r.ys;
    a.out_u = r.out_u; a.out_logp = r.out_logp; a.out_regs = r.out_regs; a.out_x = r.out_x;
    a.out_lossterm = r.out_lossterm;
    a.B = r.B;
    a.sample_offset = r.noise ? r.noise->sample_offset : 0;
    a.seed = r.noise ? r.noise->seed : 0;
    a.in_kind = r.in_kind; a.eps_kind = eps_kind; a.mode = r.mode;
    a.reg_e = mf.reg_e; a.reg_n = mf.reg_n; a.reg_a = mf.reg_a; a.squared = c.reg_squared;
    a.lam1 = c.lambda1; a.lam2 = c.lambda2; a.lam3 = c.lambda3;
    a.t0 = r.t0; a.t1 = r.t1;
    a.ctl = make_controller(r.sol);
    CK(h, h->stats.reserve(sizeof(DevStats)));
    a.stats = h->stats.as<DevStats>();

    if (r.B == 0) {
        CK(h, cudaMemsetAsync(a.stats, 0, sizeof(DevStats), st));
        return ICNF_OK;
    }
    if (!adaptive) {
        a.nsteps = fixed_steps(r.t0, r.t1, r.sol->dt);
        a.dt = r.sol->dt;
        if (r.want_ckpt) {
            a.max_ckpt_steps = a.nsteps;
            CK(h, h->ckpt.reserve(sizeof(float) * (size_t)(a.nsteps + 1) * r.B * D * std::max(1, h->fam->ckpt_stages)));
            CK(h, h->steps.reserve(sizeof(StepRec) * (size_t)(a.nsteps + 1)));
            a.ckpt = h->ckpt.as<float>();
            a.steps = h->steps.as<StepRec>();
        }
        h->prof_begin(0, st);
        cudaError_t e = h->fam->solve_fixed(h->ws, h->theta_host.data(), a, c.nvars, mf.exact, h->sm_count, st);
        h->prof_end(0, st);
        if (e != cudaSuccess) return h->cuda_fail(e, "solve_fixed launch");
        h->launches++;
        return ICNF_OK;
    }
    // VCABM (the reference's default alg): adaptive, inference-side only -- a solve that records checkpoints for the
    // reverse sweep integrates with Tsit5 (the sweep differentiates discrete Runge-Kutta steps)
    const bool vcabm = r.sol && r.sol->alg == ICNF_ALG_VCABM && !r.want_ckpt;
    if (vcabm && !h->fam->solve_vcabm)
        return h->fail(ICNF_ERR_UNSUPPORTED, "alg = VCABM is served by the single-launch solves (tiny family, narrow path); this network runs on the %s family: pass alg = Tsit5",
                       h->fam->name);
    // adaptive: cooperative persistent kernel
    const int grid_cap = h->fam->adaptive_max_grid(mf.exact, h->sm_count);
    if (grid_cap <= 0) return h->fail(ICNF_ERR_UNSUPPORTED, "adaptive kernel cannot be made resident on this device");
    a.dt = (r.sol && r.sol->dt > 0) ? r.sol->dt : 0.0f;
    const size_t sb = sizeof(float) * (size_t)S * r.B;
    CK(h, h->wu0.reserve(sb));
    CK(h, h->wu1.reserve(sb));
    CK(h, h->wk0.reserve(sb));
    CK(h, h->wk1.reserve(sb));
    CK(h, h->partials.reserve(sizeof(double) * 4 * 32 * (size_t)h->sm_count));
    if (r.out_loss && h->fam->fuses_loss_sum_adaptive) {
        a.out_loss = r.out_loss; a.loss_scale = r.loss_scale; a.out_lossterm = nullptr;
        r.loss_fused = true;
    }
    if (group_wants_global_norm(h) && r.global_B > 0 && h->fam->global_norm_capable(h->ws, a, mf.exact)) {
        // exact data-parallel mode: the controller's error norm is taken over the global batch
        group_fill_xchg(h, a.xg);
        a.norm_B = r.global_B;
    }
    a.wu[0] = h->wu0.as<float>(); a.wu[1] = h->wu1.as<float>();
    a.wk[0] = h->wk0.as<float>(); a.wk[1] = h->wk1.as<float>();
    a.partials = h->partials.as<double>();
    if (r.want_ckpt) {
        // room for 256 accepted steps, less for very large batches (16 GB cap): a solve that needs more reports
        // ICNF_ERR_MAX_STEPS instead of exhausting the device
        const size_t per_step = sizeof(float) * (size_t)r.B * D * std::max(1, h->fam->ckpt_stages);
        const int mem_cap = (int)std::max<size_t>(8, ((size_t)16 << 30) / std::max<size_t>(per_step, 1));
        // 256 accepted steps by default; a caller that sets `max_steps` (maxiters) to something between 256 and the glue's
        // default of 100 000 gets room for that many, up to the memory cap -- a stiff late-training flow can then be
        // differentiated instead of failing with ICNF_ERR_MAX_STEPS
        const bool explicit_cap = r.sol && r.sol->max_steps > 256 && r.sol->max_steps < 100000;
        a.max_ckpt_steps = std::min(explicit_cap ? a.ctl.max_steps : std::min(a.ctl.max_steps, 256), mem_cap);
        CK(h, h->ckpt.reserve(sizeof(float) * (size_t)(a.max_ckpt_steps + 1) * r.B * D * std::max(1, h->fam->ckpt_stages)));
        CK(h, h->steps.reserve(sizeof(StepRec) * (size_t)(a.max_ckpt_steps + 1)));
        a.ckpt = h->ckpt.as<float>();
        a.steps = h->steps.as<StepRec>();
    }
    if (vcabm) {
        CK(h, h->vc_hist.reserve(sizeof(float) * 2 * 13 * (size_t)S * r.B));
        a.vc_hist = h->vc_hist.as<float>();
        a.alg = ICNF_ALG_VCABM;
    }
    h->prof_begin(0, st);
    cudaError_t e = vcabm ? h->fam->solve_vcabm(h->ws, h->theta_host.data(), a, c.nvars, mf.exact, h->sm_count, st)
                          : h->fam->solve_adaptive(h->ws, h->theta_host.data(), a, c.nvars, mf.exact, h->sm_count, st);
    h->prof_end(0, st);
    if (vcabm && e == cudaErrorNotSupported) {
        (void)cudaGetLastError();
        return h->fail(ICNF_ERR_UNSUPPORTED, "alg = VCABM is served by the single-launch solves (tiny family, narrow path); this network runs on the multi-launch %s path: pass alg = Tsit5",
                       icnf_kernel_family(h));
    }
    if (e != cudaSuccess) return h->cuda_fail(e, "solve_adaptive launch");
    h->launches++;
    return ICNF_OK;
}

int finish_stats(icnf_handle* h, icnf_stats* stats, cudaStream_t st) {
    CK(h, cudaMemcpyAsync(h->stats_host, h->stats.p, sizeof(DevStats), cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    if (stats) {
        stats->naccept = h->stats_host->naccept;
        stats->nreject = h->stats_host->nreject;
        stats->nf = h->stats_host->nf;
        stats->status = h->stats_host->status;
        stats->t_final = h->stats_host->t_final;
        stats->dt_last = h->stats_host->dt_last;
    }
    const int s = h->stats_host->status;
    if (s != ICNF_OK) {
        const char* why = s == ICNF_ERR_MAX_STEPS ? "max_steps reached"
                          : s == ICNF_ERR_DT_UNDERFLOW ? "step size underflow"
                          : s == ICNF_ERR_NONFINITE ? "non-finite error estimate" : "device loop failed";
        return h->fail(s, "tsit5: %s (t = %g, accepted %d, rejected %d)", why, (double)h->stats_host->t_final,
                       h->stats_host->naccept, h->stats_host->nreject);
    }
    return ICNF_OK;
}

// _dev entry points run on the caller's stream; NULL is the legacy default stream (what
// torch.cuda.current_stream() is unless the caller changed it), so results are ordered with
// the caller's own work on that stream.
cudaStream_t pick_stream(icnf_handle*, void* stream) { return (cudaStream_t)stream; }

int upload(icnf_handle* h, DevBuf& buf, const float* src, size_t n, cudaStream_t st, const float** out) {
    *out = nullptr;
    if (!src || n == 0) return ICNF_OK;
    CK(h, buf.reserve(n * sizeof(float)));
    CK(h, cudaMemcpyAsync(buf.p, src, n * sizeof(float), cudaMemcpyHostToDevice, st));
    *out = buf.as<float>();
    return ICNF_OK;
}

}  // namespace

// ---------------------------------------------------------------- lifetime
extern "C" {

const char* icnf_version(void) { return "icnf_b200 0.1.0 (sm_100a)"; }

int icnf_backward_plan(const icnf_config* cfg, int exact, int sm_count, int64_t B, int32_t* threads, int32_t* grid,
                       int32_t* first, int32_t* n_blocks) {
    if (!cfg || !threads || !grid || !first || !n_blocks || B < 1 || sm_count < 1) return ICNF_ERR_INVALID;
    if (cfg->abi_version != ICNF_ABI_VERSION || cfg->n_layers < 1 || cfg->n_layers > ICNF_MAX_LAYERS) return ICNF_ERR_INVALID;
    NetShape shape;
    memset(&shape, 0, sizeof shape);
    shape.act = cfg->activation; shape.D = cfg->nvars + cfg->naug; shape.C = cfg->ncond; shape.NL = cfg->n_layers;
    for (int l = 0; l <= cfg->n_layers; ++l) shape.n[l] = cfg->sizes[l];
    if (cfg->precision != ICNF_FP32) return ICNF_ERR_UNSUPPORTED;
    for (const Family* f : tiny_registry())
        if (f->shape == shape) {
            if (!f->backward_plan) return ICNF_ERR_UNSUPPORTED;
            int t = 0, g = 0, nb = 0, fi[34] = {0};
            f->backward_plan(exact != 0, sm_count, (long long)B, &t, &g, fi, &nb);
            *threads = t; *grid = g; *n_blocks = nb;
            for (int b = 0; b <= nb; ++b) first[b] = fi[b];
            return ICNF_OK;
        }
    return ICNF_ERR_UNSUPPORTED;
}

// STEER end time (steer_tspan, src/core/base_icnf.jl:23-43) from the library's own counter-based stream, so that a
// caller without an RNG of its own (and every rank of a data-parallel step) draws the SAME t1 from a seed:
// r = rate * (2 u - 1), u = (w >> 8) * 2^-24, w = philox4x32_10(ctr = (0, 0, 0, "STER"), key = seed)[0]   (oracle/philox.py)
int icnf_steer_tspan(int mode, float t0, float t1, float steer_rate, uint64_t seed, float* t1_out) {
    if (!t1_out) return ICNF_ERR_INVALID;
    *t1_out = t1;
    if (mode != ICNF_TRAIN_REG || steer_rate == 0.0f) return ICNF_OK;
    uint32_t c[4] = {0u, 0u, 0u, 0x53544552u}, k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        if (r != 9) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
    }
    const float u = (float)(c[0] >> 8) * 5.9604644775390625e-08f;
    const float r = steer_rate * (2.0f * u - 1.0f);
    *t1_out = fmaf(fabsf(t1 - t0), r, t1);   // muladd(dt, r, t1), base_icnf.jl:38
    return ICNF_OK;
}

int icnf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int icnf_create(const icnf_config* cfg, icnf_handle** out) {
    if (out) *out = nullptr;
    if (!cfg || !out) { g_create_error = "null argument"; return ICNF_ERR_INVALID; }
    if (cfg->abi_version != ICNF_ABI_VERSION) { g_create_error = "abi_version mismatch"; return ICNF_ERR_INVALID; }
    const int D = cfg->nvars + cfg->naug;
    if (cfg->nvars < 1 || cfg->naug < 0 || cfg->ncond < 0 || cfg->n_layers < 1 || cfg->n_layers > ICNF_MAX_LAYERS) {
        g_create_error = "bad dimensions";
        return ICNF_ERR_INVALID;
    }
    const int n_in = D + (cfg->autonomous ? 0 : 1) + cfg->ncond;
    if (cfg->sizes[0] != n_in || cfg->sizes[cfg->n_layers] != D) {
        g_create_error = "sizes[0] must be nvars+naug+!autonomous+ncond and sizes[n_layers] must be nvars+naug";
        return ICNF_ERR_INVALID;
    }
    for (int l = 0; l <= cfg->n_layers; ++l)
        if (cfg->sizes[l] < 1) { g_create_error = "layer size < 1"; return ICNF_ERR_INVALID; }
    if (cfg->activation < ICNF_ACT_SOFTPLUS || cfg->activation > ICNF_ACT_IDENTITY) {
        g_create_error = "unknown activation";
        return ICNF_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return ICNF_ERR_NO_DEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return ICNF_ERR_INVALID; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return ICNF_ERR_CUDA; }

    icnf_handle* h = new (std::nothrow) icnf_handle();
    if (!h) { g_create_error = "out of memory"; return ICNF_ERR_INVALID; }
    h->cfg = *cfg;
    h->device = cfg->device;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);

    NetShape shape;
    memset(&shape, 0, sizeof shape);
    shape.act = cfg->activation; shape.D = D; shape.C = cfg->ncond; shape.NL = cfg->n_layers;
    for (int l = 0; l <= cfg->n_layers; ++l) shape.n[l] = cfg->sizes[l];
    if (cfg->precision == ICNF_FP32) {
        for (const Family* f : tiny_registry())
            if (f->shape == shape) { h->fam = f; break; }
    }
    if (!h->fam) h->fam = generic_family();
    if (!h->fam) {
        g_create_error = "no kernel family serves this network shape in this build";
        delete h;
        return ICNF_ERR_UNSUPPORTED;
    }
    if (h->fam->ws_create) h->ws = h->fam->ws_create(cfg);
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMallocHost((void**)&h->stats_host, sizeof(DevStats))) != cudaSuccess ||
        (e = cudaMallocHost((void**)&h->scalar_host, 4 * sizeof(float))) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        icnf_destroy(h);
        return ICNF_ERR_CUDA;
    }
    *out = h;
    return ICNF_OK;
}

static void group_free_fwd(icnf_handle* h);
void icnf_destroy(icnf_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    group_free_fwd(h);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->ws && h->fam && h->fam->ws_destroy) h->fam->ws_destroy(h->ws);
    DevBuf* bufs[] = {&h->theta_dev, &h->in, &h->eps, &h->ys, &h->out0, &h->out1, &h->out2, &h->wu0, &h->wu1, &h->wk0,
                      &h->wk1, &h->partials, &h->ckpt, &h->steps, &h->stats, &h->gpartial, &h->lossterm, &h->scalar,
                      &h->dtheta, &h->dxs, &h->dpbuf};
    for (DevBuf* b : bufs) b->release();
    if (h->stats_host) cudaFreeHost(h->stats_host);
    if (h->scalar_host) cudaFreeHost(h->scalar_host);
    if (h->grad_host) cudaFreeHost(h->grad_host);
    if (h->dev_done) cudaEventDestroy(h->dev_done);
    for (auto& pr : h->prof_ev)
        for (cudaEvent_t ev : pr)
            if (ev) cudaEventDestroy(ev);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* icnf_last_error(const icnf_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t icnf_n_params(const icnf_handle* h) {
    if (!h) return 0;
    int64_t n = 0;
    for (int l = 0; l < h->cfg.n_layers; ++l) n += (int64_t)h->cfg.sizes[l] * h->cfg.sizes[l + 1] + h->cfg.sizes[l + 1];
    return n;
}
int32_t icnf_n_state(const icnf_handle* h) { return h ? h->S() : 0; }
const char* icnf_kernel_family(const icnf_handle* h) {
    if (!h || !h->fam) return "";
    if (h->cfg.precision == ICNF_BF16_TC || h->cfg.precision == ICNF_BF16X3_TC) return "tc";
    return h->fam->name;
}
const char* icnf_solve_path(const icnf_handle* h, int mode) {
    if (!h || !h->fam) return "";
    if (std::string(h->fam->name) == "generic" && h->fam->global_norm_capable) {
        SolveArgs a;
        memset(&a, 0, sizeof a);
        a.mode = mode;
        if (h->fam->global_norm_capable(h->ws, a, mode == ICNF_TEST)) return "narrow";
    }
    return icnf_kernel_family(h);
}
int64_t icnf_launch_count(const icnf_handle* h) { return h ? h->launches : 0; }

int icnf_set_profiling(icnf_handle* h, int enabled) {
    if (!h) return ICNF_ERR_INVALID;
    h->profiling = enabled != 0;
    for (bool& u : h->prof_used) u = false;
    return ICNF_OK;
}

int icnf_kernel_times(icnf_handle* h, float* ms4) {
    if (!h || !ms4) return ICNF_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    for (int s = 0; s < 4; ++s) {
        ms4[s] = -1.0f;
        if (!h->prof_used[s]) continue;
        CK(h, cudaEventSynchronize(h->prof_ev[s][1]));
        CK(h, cudaEventElapsedTime(&ms4[s], h->prof_ev[s][0], h->prof_ev[s][1]));
    }
    return ICNF_OK;
}

int icnf_adam_step_dev(float* theta, const float* grad, float* m, float* v, int64_t n, int64_t step, float eta, float beta1,
                       float beta2, float epsilon, float lambda, void* stream) {
    if (!theta || !grad || !m || !v || n < 0 || step < 1) return ICNF_ERR_INVALID;
    if (n == 0) return ICNF_OK;
    const float bp1 = powf(beta1, (float)step), bp2 = powf(beta2, (float)step);
    adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(theta, grad, m, v, n, eta, beta1, beta2,
                                                                                  epsilon, lambda, bp1, bp2);
    return cudaGetLastError() == cudaSuccess ? ICNF_OK : ICNF_ERR_CUDA;
}

int icnf_measure_fp32_peak(int device, float* tflops) {
    if (!tflops) return ICNF_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return ICNF_ERR_NO_DEVICE;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float* d = nullptr;
    if (cudaMalloc(&d, 16) != cudaSuccess) return ICNF_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, grid = sms * 8;
    float best = 0.f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<grid, 256>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return ICNF_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 16 * 8 * (double)iters * 256.0 * grid;
        if (rep > 0) best = fmaxf(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return ICNF_OK;
}

int icnf_set_params(icnf_handle* h, const float* theta, int64_t n) {
    if (!h || !theta) return ICNF_ERR_INVALID;
    if (n != icnf_n_params(h)) return h->fail(ICNF_ERR_INVALID, "theta has %lld entries, network has %lld", (long long)n, (long long)icnf_n_params(h));
    CK(h, cudaSetDevice(h->device));
    { int rcj = join_dev(h); if (rcj) return rcj; }
    h->theta_host.assign(theta, theta + n);
    CK(h, h->theta_dev.reserve(sizeof(float) * n));
    CK(h, cudaMemcpyAsync(h->theta_dev.p, h->theta_host.data(), sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    if (h->fam->on_params) CK(h, h->fam->on_params(h->ws, h->theta_dev.as<float>(), h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->have_params = true;
    return ICNF_OK;
}

int icnf_set_params_dev(icnf_handle* h, const float* theta, int64_t n, void* stream) {
    if (!h || !theta) return ICNF_ERR_INVALID;
    if (n != icnf_n_params(h)) return h->fail(ICNF_ERR_INVALID, "theta has %lld entries, network has %lld", (long long)n, (long long)icnf_n_params(h));
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream);
    h->theta_host.resize(n);
    CK(h, h->theta_dev.reserve(sizeof(float) * n));
    CK(h, cudaMemcpyAsync(h->theta_dev.p, theta, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    // the tiny family passes the weights as a kernel parameter, so it needs them on the host
    CK(h, cudaMemcpyAsync(h->theta_host.data(), theta, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (h->fam->on_params) CK(h, h->fam->on_params(h->ws, h->theta_dev.as<float>(), st));
    CK(h, cudaStreamSynchronize(st));
    h->have_params = true;
    return ICNF_OK;
}

// ---------------------------------------------------------------- S1
int icnf_rhs_dev(icnf_handle* h, int mode, float t, const float* u, const float* eps, const float* ys, float* du,
                 int64_t B, void* stream) {
    NvtxRange nvtx_("icnf_rhs_dev");
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if (!u || !du) return h->fail(ICNF_ERR_INVALID, "null u/du");
    if (mode != ICNF_TEST && !eps) return h->fail(ICNF_ERR_INVALID, "TrainMode RHS needs eps");
    if (h->cfg.ncond && !ys) return h->fail(ICNF_ERR_INVALID, "conditioned flow needs ys");
    if (B == 0) return ICNF_OK;
    const ModeFlags mf = mode_flags(h->cfg, mode);
    RhsArgs a;
    memset(&a, 0, sizeof a);
    a.theta = h->theta_dev.as<float>();
    a.u = u; a.eps = eps; a.ys = ys; a.du = du; a.B = B;
    a.mode = mode; a.reg_e = mf.reg_e; a.reg_n = mf.reg_n; a.squared = h->cfg.reg_squared; a.t = t;
    cudaError_t e = h->fam->rhs(h->ws, h->theta_host.data(), a, mf.exact, h->sm_count, pick_stream(h, stream));
    if (e != cudaSuccess) return h->cuda_fail(e, "rhs launch");
    h->launches++;
    return mark_dev(h, pick_stream(h, stream));
}

int icnf_rhs(icnf_handle* h, int mode, float t, const float* u, const float* eps, const float* ys, float* du, int64_t B) {
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if ((rc = join_dev(h))) return rc;
    if (!u || !du) return h->fail(ICNF_ERR_INVALID, "null u/du");
    const int D = h->D(), S = h->S();
    const float *du_, *de_, *dy_;
    if ((rc = upload(h, h->in, u, (size_t)S * B, h->stream, &du_))) return rc;
    if ((rc = upload(h, h->eps, mode == ICNF_TEST ? nullptr : eps, (size_t)D * B, h->stream, &de_))) return rc;
    if ((rc = upload(h, h->ys, ys, (size_t)h->cfg.ncond * B, h->stream, &dy_))) return rc;
    CK(h, h->out0.reserve(sizeof(float) * (size_t)S * B + 4));
    if ((rc = icnf_rhs_dev(h, mode, t, du_, de_, dy_, h->out0.as<float>(), B, h->stream))) return rc;
    if (B) CK(h, cudaMemcpyAsync(du, h->out0.p, sizeof(float) * (size_t)S * B, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return ICNF_OK;
}

// ---------------------------------------------------------------- S2
int icnf_solve_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* u0,
                   const icnf_noise* noise, const float* eps, const float* ys, float* u_final, icnf_stats* stats,
                   int64_t B, void* stream) {
    NvtxRange nvtx_("icnf_solve_dev");
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if (!u0 || !u_final) return h->fail(ICNF_ERR_INVALID, "null u0/u_final");
    cudaStream_t st = pick_stream(h, stream);
    SolveRequest r{mode, sol, t0, t1, u0, IN_U0, noise, eps, ys, u_final, nullptr, nullptr, nullptr, nullptr, false, B};
    if ((rc = enqueue_solve(h, r, st))) return rc;
    if (stats) CK(h, cudaMemcpyAsync(stats, h->stats.p, sizeof(DevStats), cudaMemcpyDeviceToDevice, st));
    return mark_dev(h, st);
}

int icnf_solve(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* u0,
               const icnf_noise* noise, const float* eps, const float* ys, float* u_final, icnf_stats* stats, int64_t B) {
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if ((rc = join_dev(h))) return rc;
    if (!u0 || !u_final) return h->fail(ICNF_ERR_INVALID, "null u0/u_final");
    const int D = h->D(), S = h->S();
    const float *d_in, *d_eps, *d_ys;
    if ((rc = upload(h, h->in, u0, (size_t)S * B, h->stream, &d_in))) return rc;
    if ((rc = upload(h, h->eps, mode == ICNF_TEST ? nullptr : eps, (size_t)D * B, h->stream, &d_eps))) return rc;
    if ((rc = upload(h, h->ys, ys, (size_t)h->cfg.ncond * B, h->stream, &d_ys))) return rc;
    CK(h, h->out0.reserve(sizeof(float) * (size_t)S * B + 4));
    SolveRequest r{mode, sol, t0, t1, B ? d_in : u0, IN_U0, noise, d_eps, d_ys, h->out0.as<float>(), nullptr, nullptr,
                   nullptr, nullptr, false, B};
    if ((rc = enqueue_solve(h, r, h->stream))) return rc;
    if (B) CK(h, cudaMemcpyAsync(u_final, h->out0.p, sizeof(float) * (size_t)S * B, cudaMemcpyDeviceToHost, h->stream));
    return finish_stats(h, stats, h->stream);
}

int icnf_inference_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
                       const icnf_noise* noise, const float* eps, const float* ys, float* logp, float* regs,
                       icnf_stats* stats, int64_t B, void* stream) {
    NvtxRange nvtx_("icnf_inference_dev");
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if (!xs || !logp) return h->fail(ICNF_ERR_INVALID, "null xs/logp");
    cudaStream_t st = pick_stream(h, stream);
    SolveRequest r{mode, sol, t0, t1, xs, IN_XS, noise, eps, ys, nullptr, logp, regs, nullptr, nullptr, false, B};
    if ((rc = enqueue_solve(h, r, st))) return rc;
    if (stats) CK(h, cudaMemcpyAsync(stats, h->stats.p, sizeof(DevStats), cudaMemcpyDeviceToDevice, st));
    return mark_dev(h, st);
}

int icnf_inference(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
                   const icnf_noise* noise, const float* eps, const float* ys, float* logp, float* regs,
                   icnf_stats* stats, int64_t B) {
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if ((rc = join_dev(h))) return rc;
    if (!xs || !logp) return h->fail(ICNF_ERR_INVALID, "null xs/logp");
    const int D = h->D();
    const float *d_in, *d_eps, *d_ys;
    if ((rc = upload(h, h->in, xs, (size_t)h->cfg.nvars * B, h->stream, &d_in))) return rc;
    if ((rc = upload(h, h->eps, mode == ICNF_TEST ? nullptr : eps, (size_t)D * B, h->stream, &d_eps))) return rc;
    if ((rc = upload(h, h->ys, ys, (size_t)h->cfg.ncond * B, h->stream, &d_ys))) return rc;
    CK(h, h->out0.reserve(sizeof(float) * (size_t)B + 4));
    CK(h, h->out1.reserve(sizeof(float) * 3 * (size_t)B + 4));
    SolveRequest r{mode, sol, t0, t1, B ? d_in : xs, IN_XS, noise, d_eps, d_ys, nullptr, h->out0.as<float>(),
                   regs ? h->out1.as<float>() : nullptr, nullptr, nullptr, false, B};
    if ((rc = enqueue_solve(h, r, h->stream))) return rc;
    if (B) {
        CK(h, cudaMemcpyAsync(logp, h->out0.p, sizeof(float) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
        if (regs) CK(h, cudaMemcpyAsync(regs, h->out1.p, sizeof(float) * 3 * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    }
    return finish_stats(h, stats, h->stream);
}

int icnf_generate_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* z0,
                      const icnf_noise* noise, const float* eps, const float* ys, float* xs_out, icnf_stats* stats,
                      int64_t n, void* stream) {
    NvtxRange nvtx_("icnf_generate_dev");
    int rc = validate_common(h, mode, n);
    if (rc) return rc;
    if (!xs_out) return h->fail(ICNF_ERR_INVALID, "null xs_out");
    if (!z0 && !noise) return h->fail(ICNF_ERR_INVALID, "generate needs z0 or a noise seed to draw it");
    cudaStream_t st = pick_stream(h, stream);
    // reverse(steer_tspan(...)), base_icnf.jl:372: integrate t1 -> t0
    SolveRequest r{mode, sol, t1, t0, z0, z0 ? IN_Z0 : IN_Z0_DRAW, noise, eps, ys, nullptr, nullptr, nullptr, xs_out,
                   nullptr, false, n};
    if ((rc = enqueue_solve(h, r, st))) return rc;
    if (stats) CK(h, cudaMemcpyAsync(stats, h->stats.p, sizeof(DevStats), cudaMemcpyDeviceToDevice, st));
    return mark_dev(h, st);
}

int icnf_generate(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* z0,
                  const icnf_noise* noise, const float* eps, const float* ys, float* xs_out, icnf_stats* stats, int64_t n) {
    int rc = validate_common(h, mode, n);
    if (rc) return rc;
    if ((rc = join_dev(h))) return rc;
    if (!xs_out) return h->fail(ICNF_ERR_INVALID, "null xs_out");
    if (!z0 && !noise) return h->fail(ICNF_ERR_INVALID, "generate needs z0 or a noise seed to draw it");
    const int D = h->D();
    const float *d_in, *d_eps, *d_ys;
    if ((rc = upload(h, h->in, z0, (size_t)D * n, h->stream, &d_in))) return rc;
    if ((rc = upload(h, h->eps, mode == ICNF_TEST ? nullptr : eps, (size_t)D * n, h->stream, &d_eps))) return rc;
    if ((rc = upload(h, h->ys, ys, (size_t)h->cfg.ncond * n, h->stream, &d_ys))) return rc;
    CK(h, h->out0.reserve(sizeof(float) * (size_t)h->cfg.nvars * n + 4));
    SolveRequest r{mode, sol, t1, t0, d_in, (z0 && n) ? IN_Z0 : IN_Z0_DRAW, noise, d_eps, d_ys, nullptr, nullptr, nullptr,
                   h->out0.as<float>(), nullptr, false, n};
    if ((rc = enqueue_solve(h, r, h->stream))) return rc;
    if (n) CK(h, cudaMemcpyAsync(xs_out, h->out0.p, sizeof(float) * (size_t)h->cfg.nvars * n, cudaMemcpyDeviceToHost, h->stream));
    return finish_stats(h, stats, h->stream);
}

// ---------------------------------------------------------------- S3
// loss (and gradient) with all pointers on the device; loss/dtheta/dxs/stats device or null
static int loss_grad_device(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
                            const icnf_noise* noise, const float* eps, const float* ys, float* loss, float* dtheta,
                            float* dxs, int64_t B, int64_t global_batch, cudaStream_t st) {
    if (B <= 0) return h->fail(ICNF_ERR_INVALID, "loss needs at least one sample");
    const int64_t denom = global_batch > 0 ? global_batch : B;
    const bool want_grad = dtheta != nullptr || dxs != nullptr;
    if (want_grad && !h->fam->supports_backward)
        return h->fail(ICNF_ERR_UNSUPPORTED, "the %s kernel family has no backward kernel yet (network too wide for the tiny family)", h->fam->name);
    CK(h, h->lossterm.reserve(sizeof(float) * (size_t)B));
    SolveRequest r{mode, sol, t0, t1, xs, IN_XS, noise, eps, ys, nullptr, nullptr, nullptr, nullptr,
                   h->lossterm.as<float>(), want_grad, B};
    r.out_loss = loss;
    r.loss_scale = 1.0f / (float)denom;
    r.global_B = h->dp_defer_reduce || h->dp_active ? denom : 0;
    int rc = enqueue_solve(h, r, st);
    if (rc) return rc;
    if (loss && !r.loss_fused) {
        h->prof_begin(1, st);
        sum_kernel<<<1, 1024, 0, st>>>(h->lossterm.as<float>(), (long long)B, 1.0f / (float)denom, loss);
        h->prof_end(1, st);
        CK(h, cudaGetLastError());
        h->launches++;
    }
    if (!want_grad) return ICNF_OK;
    const ModeFlags mf = mode_flags(h->cfg, mode);
    const int np = (int)icnf_n_params(h);
    const int grid = h->fam->backward_grid(mf.exact, h->sm_count, B);
    const int nrows = grid * h->fam->backward_partials_per_block;
    CK(h, h->gpartial.reserve(sizeof(float) * (size_t)nrows * np));
    BackwardArgs b;
    memset(&b, 0, sizeof b);
    b.theta = h->theta_dev.as<float>();
    b.ckpt = h->ckpt.as<float>();
    b.steps = h->steps.as<StepRec>();
    b.stats = h->stats.as<DevStats>();
    b.eps = eps; b.ys = ys;
    b.gpartial = h->gpartial.as<float>();
    b.dxs = dxs;
    b.B = B;
    b.sample_offset = noise ? noise->sample_offset : 0;
    b.seed = noise ? noise->seed : 0;
    b.eps_kind = noise ? noise->kind : ICNF_EPS_SUPPLIED;
    b.mode = mode; b.reg_e = mf.reg_e; b.reg_n = mf.reg_n; b.reg_a = mf.reg_a; b.squared = h->cfg.reg_squared;
    b.lam1 = h->cfg.lambda1; b.lam2 = h->cfg.lambda2; b.lam3 = h->cfg.lambda3;
    b.inv_denominator = 1.0f / (float)denom;
    b.nvars = h->cfg.nvars;
    h->prof_begin(2, st);
    cudaError_t e = h->fam->backward(h->ws, h->theta_host.data(), b, mf.exact, grid, st);
    h->prof_end(2, st);
    if (e != cudaSuccess) return h->cuda_fail(e, "backward launch");
    h->launches++;
    h->dp_nrows = nrows;
    if (dtheta && !h->dp_defer_reduce) {
        h->prof_begin(3, st);
        reduce_grad_kernel<<<(np + 31) / 32, 256, 0, st>>>(h->gpartial.as<float>(), nrows, np, dtheta);
        h->prof_end(3, st);
        CK(h, cudaGetLastError());
        h->launches++;
    }
    return ICNF_OK;
}

int icnf_loss_grad_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
                       const icnf_noise* noise, const float* eps, const float* ys, float* loss, float* dtheta, float* dxs,
                       icnf_stats* stats, int64_t B, int64_t global_batch, void* stream) {
    NvtxRange nvtx_("icnf_loss_grad_dev");
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if (!xs) return h->fail(ICNF_ERR_INVALID, "null xs");
    cudaStream_t st = pick_stream(h, stream);
    if ((rc = loss_grad_device(h, mode, sol, t0, t1, xs, noise, eps, ys, loss, dtheta, dxs, B, global_batch, st))) return rc;
    if (stats) CK(h, cudaMemcpyAsync(stats, h->stats.p, sizeof(DevStats), cudaMemcpyDeviceToDevice, st));
    return mark_dev(h, st);
}

static int loss_grad_host(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
                          const icnf_noise* noise, const float* eps, const float* ys, float* loss, float* dtheta,
                          float* dxs, icnf_stats* stats, int64_t B, int64_t global_batch) {
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if ((rc = join_dev(h))) return rc;
    if (!xs || !loss) return h->fail(ICNF_ERR_INVALID, "null xs/loss");
    const int D = h->D();
    const int np = (int)icnf_n_params(h);
    const float *d_in, *d_eps, *d_ys;
    if ((rc = upload(h, h->in, xs, (size_t)h->cfg.nvars * B, h->stream, &d_in))) return rc;
    if ((rc = upload(h, h->eps, mode == ICNF_TEST ? nullptr : eps, (size_t)D * B, h->stream, &d_eps))) return rc;
    if ((rc = upload(h, h->ys, ys, (size_t)h->cfg.ncond * B, h->stream, &d_ys))) return rc;
    CK(h, h->scalar.reserve(4 * sizeof(float)));
    if (dtheta) CK(h, h->dtheta.reserve(sizeof(float) * (size_t)np));
    if (dxs) CK(h, h->dxs.reserve(sizeof(float) * (size_t)h->cfg.nvars * B));
    if ((rc = loss_grad_device(h, mode, sol, t0, t1, d_in, noise, d_eps, d_ys, h->scalar.as<float>(),
                               dtheta ? h->dtheta.as<float>() : nullptr, dxs ? h->dxs.as<float>() : nullptr, B,
                               global_batch, h->stream)))
        return rc;
    CK(h, cudaMemcpyAsync(h->scalar_host, h->scalar.p, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (dtheta) {
        if (h->grad_host_cap < (size_t)np) {
            if (h->grad_host) cudaFreeHost(h->grad_host);
            h->grad_host = nullptr;
            h->grad_host_cap = 0;
            CK(h, cudaMallocHost((void**)&h->grad_host, sizeof(float) * (size_t)np));
            h->grad_host_cap = (size_t)np;
        }
        CK(h, cudaMemcpyAsync(h->grad_host, h->dtheta.p, sizeof(float) * (size_t)np, cudaMemcpyDeviceToHost, h->stream));
    }
    if (dxs) CK(h, cudaMemcpyAsync(dxs, h->dxs.p, sizeof(float) * (size_t)h->cfg.nvars * B, cudaMemcpyDeviceToHost, h->stream));
    rc = finish_stats(h, stats, h->stream);   // one synchronisation for loss, gradient and statistics
    *loss = h->scalar_host[0];
    if (dtheta) memcpy(dtheta, h->grad_host, sizeof(float) * (size_t)np);
    return rc;
}

int icnf_loss(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
              const icnf_noise* noise, const float* eps, const float* ys, float* loss, icnf_stats* stats, int64_t B,
              int64_t global_batch) {
    return loss_grad_host(h, mode, sol, t0, t1, xs, noise, eps, ys, loss, nullptr, nullptr, stats, B, global_batch);
}

int icnf_loss_grad(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
                   const icnf_noise* noise, const float* eps, const float* ys, float* loss, float* dtheta, float* dxs,
                   icnf_stats* stats, int64_t B, int64_t global_batch) {
    if (h && !dtheta) return h->fail(ICNF_ERR_INVALID, "null dtheta");
    return loss_grad_host(h, mode, sol, t0, t1, xs, noise, eps, ys, loss, dtheta, dxs, stats, B, global_batch);
}

}  // extern "C"

// ---------------------------------------------------------------- data-parallel group
// NCCL is bound at run time (dlopen): the library loads without it, and inside a process that already holds an NCCL
// (PyTorch's) the same copy is picked up by its SONAME.
namespace {
struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, icnf_group_id, int) = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
    bool ok = false;
};
Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (!n.lib) { n.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* s) { void* p = dlsym(n.lib, s); if (!p) n.err = std::string("libnccl lacks ") + s; return p; };
        n.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
        n.CommInitRank = (int (*)(void**, int, icnf_group_id, int))sym("ncclCommInitRank");
        n.CommInitAll = (int (*)(void**, int, const int*))sym("ncclCommInitAll");
        n.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
        n.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
        n.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
        n.GroupStart = (int (*)())sym("ncclGroupStart");
        n.GroupEnd = (int (*)())sym("ncclGroupEnd");
        n.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        n.ok = n.err.empty();
    });
    return n;
}
constexpr int NCCL_FLOAT = 7, NCCL_UINT8 = 1, NCCL_SUM = 0;
}  // namespace

struct Group {
    void* comm = nullptr;
    int nranks = 1, rank = 0;
    // NVLink peer-memory exchange for small payloads
    unsigned char* xbuf = nullptr;
    icnf::PeerTable peers{};
    bool peer_ok = false, ipc = false;
    unsigned epoch = 0;
    int* timeout_flag = nullptr;   // device
    bool global_norm = false;      // adaptive solves of the data-parallel step use the error norm of the GLOBAL batch
};

namespace {
bool group_wants_global_norm(const icnf_handle* h) {
    const Group* g = h->grp;
    return g && g->global_norm && g->peer_ok && g->nranks > 1 && h->fam->global_norm_capable;
}
void group_fill_xchg(const icnf_handle* h, NormXchg& x) {
    x.peers = h->grp->peers;
    x.nranks = h->grp->nranks;
    x.rank = h->grp->rank;
}
}  // namespace

namespace {
#define NCK(h, call)                                                                                         \
    do {                                                                                                     \
        int r__ = (call);                                                                                    \
        if (r__ != 0) return (h)->fail(ICNF_ERR_CUDA, "%s: %s", #call, nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error"); \
    } while (0)

void group_free(icnf_handle* h) {
    Group* g = h->grp;
    if (!g) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (g->ipc)
        for (int r = 0; r < g->nranks; ++r)
            if (r != g->rank && g->peers.p[r]) cudaIpcCloseMemHandle(g->peers.p[r]);
    if (g->comm && nccl().ok) nccl().CommDestroy(g->comm);
    if (g->xbuf) cudaFree(g->xbuf);
    if (g->timeout_flag) cudaFree(g->timeout_flag);
    delete g;
    h->grp = nullptr;
}

int group_alloc_xbuf(icnf_handle* h, Group* g) {
    CK(h, cudaMalloc((void**)&g->xbuf, XBUF_BYTES));
    CK(h, cudaMemset(g->xbuf, 0, XBUF_BYTES));
    CK(h, cudaMalloc((void**)&g->timeout_flag, sizeof(int)));
    CK(h, cudaMemset(g->timeout_flag, 0, sizeof(int)));
    return ICNF_OK;
}
}  // namespace

static void group_free_fwd(icnf_handle* h) { group_free(h); }

extern "C" {

int icnf_group_unique_id(icnf_group_id* id) {
    if (!id) return ICNF_ERR_INVALID;
    if (!nccl().ok) { g_create_error = nccl().err; return ICNF_ERR_UNSUPPORTED; }
    return nccl().GetUniqueId(id) == 0 ? ICNF_OK : ICNF_ERR_CUDA;
}

int icnf_group_join(icnf_handle* h, const icnf_group_id* id, int32_t n_ranks, int32_t rank) {
    if (!h || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return ICNF_ERR_INVALID;
    if (n_ranks > XG_MAXR) return h->fail(ICNF_ERR_UNSUPPORTED, "at most %d ranks per group", XG_MAXR);
    if (!nccl().ok) return h->fail(ICNF_ERR_UNSUPPORTED, "%s", nccl().err.c_str());
    CK(h, cudaSetDevice(h->device));
    group_free(h);
    Group* g = new Group();
    g->nranks = n_ranks; g->rank = rank;
    h->grp = g;
    NCK(h, nccl().CommInitRank(&g->comm, n_ranks, *id, rank));
    int rc = group_alloc_xbuf(h, g);
    if (rc) return rc;
    // exchange the CUDA IPC handles of the exchange buffers with one all-gather, then map every peer's buffer
    cudaIpcMemHandle_t mine;
    std::vector<cudaIpcMemHandle_t> all(n_ranks);
    bool have = cudaIpcGetMemHandle(&mine, g->xbuf) == cudaSuccess;
    if (!have) { memset(&mine, 0, sizeof mine); cudaGetLastError(); }
    unsigned char* stage = nullptr;
    const size_t hb = sizeof(cudaIpcMemHandle_t) + 8;   // handle + "valid" byte, padded
    CK(h, cudaMalloc((void**)&stage, hb * (size_t)(n_ranks + 1)));
    std::vector<unsigned char> rec(hb, 0);
    memcpy(rec.data(), &mine, sizeof mine);
    rec[sizeof mine] = have ? 1 : 0;
    CK(h, cudaMemcpy(stage, rec.data(), hb, cudaMemcpyHostToDevice));
    NCK(h, nccl().AllGather(stage, stage + hb, hb, NCCL_UINT8, g->comm, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    std::vector<unsigned char> got(hb * (size_t)n_ranks);
    CK(h, cudaMemcpy(got.data(), stage + hb, got.size(), cudaMemcpyDeviceToHost));
    cudaFree(stage);
    bool all_ok = true;
    for (int r = 0; r < n_ranks; ++r) {
        if (!got[(size_t)r * hb + sizeof(cudaIpcMemHandle_t)]) all_ok = false;
        memcpy(&all[r], &got[(size_t)r * hb], sizeof(cudaIpcMemHandle_t));
    }
    g->peers.p[rank] = g->xbuf;
    if (all_ok) {
        g->ipc = true;
        for (int r = 0; r < n_ranks && all_ok; ++r) {
            if (r == rank) continue;
            void* pp = nullptr;
            if (cudaIpcOpenMemHandle(&pp, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all_ok = false; cudaGetLastError(); }
            g->peers.p[r] = (unsigned char*)pp;
        }
    }
    // every rank must take the same path: agree on peer_ok with a one-float all-reduce
    float* flag = nullptr;
    CK(h, cudaMalloc((void**)&flag, sizeof(float)));
    const float mineok = all_ok ? 0.f : 1.f;
    CK(h, cudaMemcpy(flag, &mineok, sizeof(float), cudaMemcpyHostToDevice));
    NCK(h, nccl().AllReduce(flag, flag, 1, NCCL_FLOAT, NCCL_SUM, g->comm, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    float bad = 1.f;
    CK(h, cudaMemcpy(&bad, flag, sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(flag);
    g->peer_ok = (bad == 0.f);
    return ICNF_OK;
}

int icnf_create_group(icnf_handle** handles, int32_t n) {
    if (!handles || n < 1) return ICNF_ERR_INVALID;
    for (int i = 0; i < n; ++i)
        if (!handles[i]) return ICNF_ERR_INVALID;
    icnf_handle* h0 = handles[0];
    if (n > XG_MAXR) return h0->fail(ICNF_ERR_UNSUPPORTED, "at most %d ranks per group", XG_MAXR);
    if (!nccl().ok) return h0->fail(ICNF_ERR_UNSUPPORTED, "%s", nccl().err.c_str());
    std::vector<int> devs(n);
    std::vector<void*> comms(n, nullptr);
    for (int i = 0; i < n; ++i) {
        devs[i] = handles[i]->device;
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return h0->fail(ICNF_ERR_INVALID, "icnf_create_group: two handles on device %d", devs[i]);
    }
    NCK(h0, nccl().CommInitAll(comms.data(), n, devs.data()));
    bool peer_ok = true;
    for (int i = 0; i < n; ++i) {
        icnf_handle* h = handles[i];
        CK(h, cudaSetDevice(h->device));
        group_free(h);
        Group* g = new Group();
        g->nranks = n; g->rank = i; g->comm = comms[i];
        h->grp = g;
        int rc = group_alloc_xbuf(h, g);
        if (rc) return rc;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
            if (!can) { peer_ok = false; continue; }
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer_ok = false;
            cudaGetLastError();
        }
    }
    for (int i = 0; i < n; ++i) {
        Group* g = handles[i]->grp;
        for (int j = 0; j < n; ++j) g->peers.p[j] = handles[j]->grp->xbuf;
        g->peer_ok = peer_ok;
    }
    return ICNF_OK;
}

int icnf_group_leave(icnf_handle* h) {
    if (!h) return ICNF_ERR_INVALID;
    group_free(h);
    return ICNF_OK;
}

int icnf_group_info(const icnf_handle* h, int32_t* n_ranks, int32_t* rank, int32_t* peer_memory) {
    if (!h) return ICNF_ERR_INVALID;
    if (n_ranks) *n_ranks = h->grp ? h->grp->nranks : 1;
    if (rank) *rank = h->grp ? h->grp->rank : 0;
    if (peer_memory) *peer_memory = (h->grp && h->grp->peer_ok) ? 1 : 0;
    return ICNF_OK;
}

int icnf_group_set_global_norm(icnf_handle* h, int enabled) {
    if (!h) return ICNF_ERR_INVALID;
    if (!h->grp) return h->fail(ICNF_ERR_INVALID, "handle is not in a group");
    if (enabled && !(h->grp->peer_ok && h->fam->global_norm_capable))
        return h->fail(ICNF_ERR_UNSUPPORTED, "the global error norm needs NVLink peer memory and a single-launch solve (tiny family, narrow path)");
    h->grp->global_norm = enabled != 0;
    return ICNF_OK;
}

int icnf_group_start(void) { return (nccl().ok && nccl().GroupStart() == 0) ? ICNF_OK : ICNF_ERR_UNSUPPORTED; }
int icnf_group_end(void) { return (nccl().ok && nccl().GroupEnd() == 0) ? ICNF_OK : ICNF_ERR_UNSUPPORTED; }

// Data-parallel training step: loss and gradient of this rank's columns, then ONE sum over the group of
// [dtheta; loss], all enqueued on `stream`; every rank receives the gradient of the whole batch.  Small vectors
// (the narrow-MLP family: <= 4095 parameters) are exchanged inside the gradient-reduction kernel through NVLink peer
// memory; larger ones go through ncclAllReduce.
int icnf_loss_grad_dp_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
                          const icnf_noise* noise, const float* eps, const float* ys, float* loss, float* dtheta, float* dxs,
                          icnf_stats* stats, int64_t B, int64_t global_batch, void* stream) {
    NvtxRange nvtx_("icnf_loss_grad_dp_dev");
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if (!xs || !dtheta) return h->fail(ICNF_ERR_INVALID, "null xs/dtheta");
    Group* g = h->grp;
    if (!g || g->nranks == 1)
        return icnf_loss_grad_dev(h, mode, sol, t0, t1, xs, noise, eps, ys, loss, dtheta, dxs, stats, B, global_batch, stream);
    cudaStream_t st = pick_stream(h, stream);
    const int np = (int)icnf_n_params(h);
    CK(h, h->dpbuf.reserve(sizeof(float) * (size_t)(np + 1)));
    float* buf = h->dpbuf.as<float>();
    const bool fused = g->peer_ok && (np + 1 <= XG_FLOATS) && h->fam->backward_partials_per_block == 1 &&
                       std::string(h->fam->name) == "tiny";
    h->dp_defer_reduce = fused;
    h->dp_active = true;
    rc = loss_grad_device(h, mode, sol, t0, t1, xs, noise, eps, ys, buf + np, buf, dxs, B, global_batch, st);
    h->dp_defer_reduce = false;
    h->dp_active = false;
    if (rc) return rc;
    if (fused) {
        g->epoch++;
        reduce_grad_xchg_kernel<<<(np + 1 + 31) / 32, 256, 0, st>>>(h->gpartial.as<float>(), h->dp_nrows, np, buf + np, dtheta, loss,
                                                                   g->peers, g->nranks, g->rank, g->epoch, g->timeout_flag);
        CK(h, cudaGetLastError());
        h->launches++;
    } else {
        NCK(h, nccl().AllReduce(buf, buf, (size_t)np + 1, NCCL_FLOAT, NCCL_SUM, g->comm, st));
        CK(h, cudaMemcpyAsync(dtheta, buf, sizeof(float) * (size_t)np, cudaMemcpyDeviceToDevice, st));
        if (loss) CK(h, cudaMemcpyAsync(loss, buf + np, sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (stats) CK(h, cudaMemcpyAsync(stats, h->stats.p, sizeof(DevStats), cudaMemcpyDeviceToDevice, st));
    return mark_dev(h, st);
}

int icnf_loss_grad_dp(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1, const float* xs,
                      const icnf_noise* noise, const float* eps, const float* ys, float* loss, float* dtheta, float* dxs,
                      icnf_stats* stats, int64_t B, int64_t global_batch) {
    int rc = validate_common(h, mode, B);
    if (rc) return rc;
    if (!xs || !loss || !dtheta) return h->fail(ICNF_ERR_INVALID, "null xs/loss/dtheta");
    if ((rc = join_dev(h))) return rc;
    const int D = h->D();
    const int np = (int)icnf_n_params(h);
    const float *d_in, *d_eps, *d_ys;
    if ((rc = upload(h, h->in, xs, (size_t)h->cfg.nvars * B, h->stream, &d_in))) return rc;
    if ((rc = upload(h, h->eps, mode == ICNF_TEST ? nullptr : eps, (size_t)D * B, h->stream, &d_eps))) return rc;
    if ((rc = upload(h, h->ys, ys, (size_t)h->cfg.ncond * B, h->stream, &d_ys))) return rc;
    CK(h, h->scalar.reserve(4 * sizeof(float)));
    CK(h, h->dtheta.reserve(sizeof(float) * (size_t)np));
    if (dxs) CK(h, h->dxs.reserve(sizeof(float) * (size_t)h->cfg.nvars * B));
    if ((rc = icnf_loss_grad_dp_dev(h, mode, sol, t0, t1, d_in, noise, d_eps, d_ys, h->scalar.as<float>(), h->dtheta.as<float>(),
                                    dxs ? h->dxs.as<float>() : nullptr, nullptr, B, global_batch, h->stream)))
        return rc;
    CK(h, cudaMemcpyAsync(h->scalar_host, h->scalar.p, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (h->grad_host_cap < (size_t)np) {
        if (h->grad_host) cudaFreeHost(h->grad_host);
        h->grad_host = nullptr;
        h->grad_host_cap = 0;
        CK(h, cudaMallocHost((void**)&h->grad_host, sizeof(float) * (size_t)np));
        h->grad_host_cap = (size_t)np;
    }
    CK(h, cudaMemcpyAsync(h->grad_host, h->dtheta.p, sizeof(float) * (size_t)np, cudaMemcpyDeviceToHost, h->stream));
    if (dxs) CK(h, cudaMemcpyAsync(dxs, h->dxs.p, sizeof(float) * (size_t)h->cfg.nvars * B, cudaMemcpyDeviceToHost, h->stream));
    rc = finish_stats(h, stats, h->stream);
    *loss = h->scalar_host[0];
    memcpy(dtheta, h->grad_host, sizeof(float) * (size_t)np);
    return rc;
}

}  // extern "C"
