/*
 * icnf_b200.h -- C ABI of libicnf_b200.so
 *
 * A B200-native (sm_100a) implementation of the data-parallel hot path of
 * ContinuousNormalizingFlows.jl: the batched augmented ODE right-hand side of
 * ICNF / RNODE / FFJORD / CondICNF, the Tsit5 loop that integrates it, the
 * log-density / sample / loss readouts and the training gradient.
 *
 * The reference has no FFI for this path (it is pure Julia; the extension
 * mechanism is multiple dispatch on the ICNF type parameters).  Each entry point
 * below names the reference method it stands behind; a Julia maintainer binds it
 * with `ccall` from a method specialised on a new `B200MatrixMode <: MatrixMode`
 * (see INTEGRATION.md and continuousnormalizingflows.jl_b200/julia/B200Mode.jl).
 * All citations are relative to the reference checkout's root.
 *
 * Data layout (exactly Julia's): Float32, column-major.  An `R x B` matrix is B
 * contiguous records of R floats (one record per sample).  `theta` is
 * `ComponentArrays.getdata(ps)`: [vec(W1); b1; vec(W2); b2; ...] with W_l
 * column-major n_l x n_{l-1} (src/exts/mlj_ext/core_icnf.jl:35).
 *
 * Ownership: the caller owns every buffer passed in and keeps it alive for the
 * duration of the call; the library owns all device memory behind the opaque
 * handle.  Entry points without the `_dev` suffix take HOST pointers and copy
 * in/out internally (they return when the result is in the caller's buffer);
 * `_dev` entry points take DEVICE pointers on the handle's device, enqueue on
 * the given CUDA stream (a `cudaStream_t` passed as `void*`, NULL = the legacy
 * default stream) and do not synchronise; their `loss`, `dtheta`, `dxs` and
 * `stats` arguments are DEVICE pointers too (`stats` receives an icnf_stats
 * record asynchronously; its `status` field carries the device loop's verdict).
 *
 * Errors: every function returns an `icnf_status`; no C++ exception crosses the
 * ABI; `icnf_last_error` returns a message.  There is NO CPU fallback: without a
 * usable CUDA device `icnf_create` returns ICNF_ERR_NO_DEVICE.
 *
 * Threading: a handle is used by one thread at a time; distinct handles are
 * independent.  One process per GPU for multi-GPU use; the batch is sharded by
 * columns; `icnf_noise.sample_offset` and `global_batch` tell the library where
 * this shard sits in the global batch.
 */
#ifndef ICNF_B200_H
#define ICNF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ICNF_API __attribute__((visibility("default")))
#else
#define ICNF_API
#endif

#define ICNF_MAX_LAYERS 8
#define ICNF_ABI_VERSION 1

typedef enum icnf_status {
    ICNF_OK = 0,
    ICNF_ERR_INVALID = 1,       /* bad argument / inconsistent config            */
    ICNF_ERR_CUDA = 2,          /* CUDA runtime error (message has the detail)   */
    ICNF_ERR_MAX_STEPS = 3,     /* adaptive loop hit max_steps                   */
    ICNF_ERR_DT_UNDERFLOW = 4,  /* step size underflow                           */
    ICNF_ERR_NONFINITE = 5,     /* NaN/Inf in the error estimate or state        */
    ICNF_ERR_NO_DEVICE = 6,     /* no CUDA device: there is no CPU fallback      */
    ICNF_ERR_UNSUPPORTED = 7    /* valid request this build has no kernel for    */
} icnf_status;

/* src/core/types.jl:1-7 -- TestMode, TrainMode{true}, TrainMode{false}.
 * TEST uses the exact Jacobian trace (src/core/icnf.jl:297-316); both TRAIN modes
 * use one Hutchinson probe (src/core/icnf.jl:517-536); the regularisers E, n, A
 * are non-zero only for TRAIN_REG (src/core/icnf.jl:184-251, base_icnf.jl:106-132). */
typedef enum icnf_mode { ICNF_TEST = 0, ICNF_TRAIN_REG = 1, ICNF_TRAIN_NOREG = 2 } icnf_mode;

typedef enum icnf_activation {
    ICNF_ACT_SOFTPLUS = 0,      /* NNlib.softplus, the reference default (icnf.jl:68-69) */
    ICNF_ACT_TANH = 1,
    ICNF_ACT_SIGMOID = 2,
    ICNF_ACT_IDENTITY = 3
} icnf_activation;

typedef enum icnf_eps_kind {
    ICNF_EPS_SUPPLIED = 0,      /* caller passes the D' x B probe matrix (parity mode)   */
    ICNF_EPS_GAUSSIAN = 1,      /* N(0,I): the reference default epsdist (icnf.jl:80-83) */
    ICNF_EPS_RADEMACHER = 2     /* +-1, drawn in-kernel with Philox4x32-10               */
} icnf_eps_kind;

typedef enum icnf_precision {
    ICNF_FP32 = 0,              /* fp32 FMA path (parity path, every width)              */
    ICNF_BF16_TC = 1,           /* bf16 tcgen05 tensor-core RHS for wide MLPs (about 1e-2 accurate) */
    ICNF_BF16X3_TC = 2          /* split bf16 (hi + lo operands, 3 tcgen05 MMAs per K step): tensor
                                 * cores at about 1e-5 relative accuracy, for the 1e-4 parity bar */
} icnf_precision;

/* Mirror of the `ICNF` struct and keyword constructor (src/core/icnf.jl:16-141).
 * Booleans the reference lifts into type parameters (icnf.jl:105-115) are derived
 * here: CONDITIONED = ncond != 0, AUGMENTED = naug != 0, NORM_Z = lambda1 != 0,
 * NORM_J = lambda2 != 0, NORM_Z_AUG = lambda3 != 0.  The network is the Lux
 * Chain of Dense layers of icnf.jl:67-71: sizes[0] = n_in = nvars + naug +
 * !autonomous + ncond (icnf.jl:64), sizes[n_layers] = nvars + naug; every layer
 * but the last applies `activation`. */
typedef struct icnf_config {
    int32_t abi_version;        /* ICNF_ABI_VERSION */
    int32_t nvars;
    int32_t naug;
    int32_t ncond;
    int32_t autonomous;         /* 0: time is appended to the network input (icnf.jl:147-153) */
    int32_t n_layers;
    int32_t sizes[ICNF_MAX_LAYERS + 1];
    int32_t activation;         /* icnf_activation */
    float lambda1, lambda2, lambda3;   /* icnf.jl:73-75 */
    int32_t reg_squared;        /* 0 = un-squared norms as the reference (icnf.jl:198,244); 1 = RNODE-paper squares */
    int32_t precision;          /* icnf_precision */
    int32_t device;             /* CUDA ordinal */
} icnf_config;

/* The part of `sol_kwargs` (src/core/icnf.jl:84-102) this path consumes, for the
 * Tsit5 stepper.  Zero in a controller field selects OrdinaryDiffEq's default. */
typedef struct icnf_solver {
    int32_t adaptive;           /* 1: PI-controlled steps, error norm over the whole S x B state */
    float dt;                   /* fixed step, or initial step (0 = automatic) when adaptive */
    float reltol, abstol;       /* icnf.jl:87-88 (1e-4) */
    int32_t max_steps;          /* 0 = 100000; the reference's maxiters is typemax(Int) (icnf.jl:86).  Training solves keep
                                 * checkpoints for 256 accepted steps unless 256 < max_steps < 100000 asks for more */
    float beta1, beta2, gamma, qmin, qmax, qsteady_min, qsteady_max, qoldinit;
    int32_t alg;                /* icnf_alg: 0 = Tsit5 (BASELINE.json north_star), 1 = VCABM (the reference's default,
                                 * src/core/icnf.jl:89): adaptive only; served for solve / inference / generate / loss by the
                                 * single-launch solves (tiny family and the narrow path: two hidden layers up to ~80 wide).  icnf_loss_grad differentiates discrete Runge-Kutta steps
                                 * and therefore integrates with Tsit5 whatever `alg` says (the reference's gradient is a
                                 * continuous adjoint: neither is tied to the forward steps). */
} icnf_solver;
typedef enum icnf_alg { ICNF_ALG_TSIT5 = 0, ICNF_ALG_VCABM = 1 } icnf_alg;

/* Hutchinson probe / base sample source.  `sample_offset` is the global column
 * index of this shard's first sample so that a sharded batch draws the same
 * numbers as the unsharded one. */
typedef struct icnf_noise {
    int32_t kind;               /* icnf_eps_kind */
    uint64_t seed;
    int64_t sample_offset;
} icnf_noise;

typedef struct icnf_stats {
    int32_t naccept, nreject, nf;   /* accepted / rejected steps, RHS evaluations */
    int32_t status;                 /* icnf_status of the device loop */
    float t_final, dt_last;
} icnf_stats;

typedef struct icnf_handle icnf_handle;

/* -- lifetime --------------------------------------------------------------- */
ICNF_API const char* icnf_version(void);
ICNF_API int icnf_device_count(void);
/* `ICNF(; ...)` constructor, src/core/icnf.jl:53-141 */
ICNF_API int icnf_create(const icnf_config* cfg, icnf_handle** out);
ICNF_API void icnf_destroy(icnf_handle* h);
/* message of the last failure on this handle (or of the last failed icnf_create when h == NULL) */
ICNF_API const char* icnf_last_error(const icnf_handle* h);
ICNF_API int64_t icnf_n_params(const icnf_handle* h);
ICNF_API int32_t icnf_n_state(const icnf_handle* h);        /* S = nvars + naug + 3 (icnf.jl:143-145) */
/* name of the kernel family that serves this config ("tiny", "generic", "tc") */
ICNF_API const char* icnf_kernel_family(const icnf_handle* h);
/* Where a solve of `mode` (icnf_mode) of this handle runs: "tiny", "narrow" (single-launch path of the generic family,
 * forward solves only: its reverse sweep runs on the generic SGEMMs), "generic" or "tc". */
ICNF_API const char* icnf_solve_path(const icnf_handle* h, int mode);

/* `ps` of every reference call (ComponentArray data, host or device pointer) */
ICNF_API int icnf_set_params(icnf_handle* h, const float* theta, int64_t n);
ICNF_API int icnf_set_params_dev(icnf_handle* h, const float* theta, int64_t n, void* stream);

/* -- S1: one RHS evaluation ------------------------------------------------- */
/* `augmented_f(du, u, p, t, icnf, mode, nn, st, eps)`, src/core/icnf.jl:297-339
 * (TestMode) and :517-559 (TrainMode, VJP).  u, du: S x B; eps: D' x B (ignored
 * for ICNF_TEST); ys: ncond x B or NULL. */
ICNF_API int icnf_rhs(icnf_handle* h, int mode, float t, const float* u, const float* eps,
             const float* ys, float* du, int64_t B);
ICNF_API int icnf_rhs_dev(icnf_handle* h, int mode, float t, const float* u, const float* eps,
                 const float* ys, float* du, int64_t B, void* stream);

/* -- S2: one whole solve ---------------------------------------------------- */
/* `base_sol(icnf, prob)`, src/core/base_icnf.jl:134-140: integrates u0 from t0
 * to t1 (t1 < t0 runs backwards, as generate does) and returns u(t1) only
 * (save_everystep = false, icnf.jl:85).  eps may be NULL when noise->kind is not
 * SUPPLIED or mode is ICNF_TEST. */
ICNF_API int icnf_solve(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
               const float* u0, const icnf_noise* noise, const float* eps, const float* ys,
               float* u_final, icnf_stats* stats, int64_t B);
ICNF_API int icnf_solve_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                   const float* u0, const icnf_noise* noise, const float* eps, const float* ys,
                   float* u_final, icnf_stats* stats, int64_t B, void* stream);

/* `inference(icnf, mode, xs, [ys,] ps, st)`, src/core/base_icnf.jl:406-424 =
 * inference_prob (:247-296) + base_sol + inference_sol (:158-172) fused: builds
 * u0 = [xs; 0], solves t0 -> t1, returns logp (B) and regs = [E; n; A] (3 x B).
 * regs may be NULL. */
ICNF_API int icnf_inference(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                   const float* xs, const icnf_noise* noise, const float* eps, const float* ys,
                   float* logp, float* regs, icnf_stats* stats, int64_t B);
ICNF_API int icnf_inference_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                       const float* xs, const icnf_noise* noise, const float* eps, const float* ys,
                       float* logp, float* regs, icnf_stats* stats, int64_t B, void* stream);

/* `generate(icnf, mode, [ys,] ps, st, n)`, src/core/base_icnf.jl:351-404 +
 * generate_sol (:185-194): z0 ~ basedist (D' x n; pass NULL to draw N(0,I)
 * in-kernel from noise->seed), integrates t1 -> t0, returns rows 1:nvars
 * (nvars x n).  Pass the span in inference order (t0 < t1); the reversal is done
 * here as the reference does (`reverse(steer_tspan(...))`, :372). */
ICNF_API int icnf_generate(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                  const float* z0, const icnf_noise* noise, const float* eps, const float* ys,
                  float* xs_out, icnf_stats* stats, int64_t n);
ICNF_API int icnf_generate_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                      const float* z0, const icnf_noise* noise, const float* eps, const float* ys,
                      float* xs_out, icnf_stats* stats, int64_t n, void* stream);

/* -- S3: loss and its gradient ---------------------------------------------- */
/* `loss(icnf, mode, xs, [ys,] ps, st)`, src/core/icnf.jl:628-649:
 * mean_b(-logp + l1 E + l2 n + l3 A).  `global_batch` is the mean's denominator
 * (0 = B); a shard of a larger batch passes the global size so that the shards'
 * losses and gradients SUM to the unsharded result. */
ICNF_API int icnf_loss(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
              const float* xs, const icnf_noise* noise, const float* eps, const float* ys,
              float* loss, icnf_stats* stats, int64_t B, int64_t global_batch);
/* Gradient of the above w.r.t. theta (and xs when dxs != NULL): what Zygote +
 * SciMLSensitivity produce for the reference (sensealg at src/core/icnf.jl:90-99;
 * exercised at test/ci_tests/smoke_tests.jl:132-133).  Computed by reverse-mode
 * differentiation of the discrete Tsit5 steps actually taken (step sizes held
 * constant), in one fused backward kernel. */
ICNF_API int icnf_loss_grad(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                   const float* xs, const icnf_noise* noise, const float* eps, const float* ys,
                   float* loss, float* dtheta, float* dxs, icnf_stats* stats,
                   int64_t B, int64_t global_batch);
ICNF_API int icnf_loss_grad_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                       const float* xs, const icnf_noise* noise, const float* eps, const float* ys,
                       float* loss, float* dtheta, float* dxs, icnf_stats* stats,
                       int64_t B, int64_t global_batch, void* stream);

/* `steer_tspan(icnf, mode)`, src/core/base_icnf.jl:23-43: t1 + |t1 - t0| r, r ~ U(-steer_rate, steer_rate), in
 * TrainMode{true} only; drawn from the library's counter-based Philox stream "STER" keyed on `seed`, so that every
 * rank of a data-parallel step (and a caller without an RNG) gets the same t1.  Host function, no device work. */
ICNF_API int icnf_steer_tspan(int mode, float t0, float t1, float steer_rate, uint64_t seed, float* t1_out);

/* -- multi-GPU ---------------------------------------------------
 * The batch is sharded by columns, parameters are replicated, one handle per GPU.  Inference and generate need no
 * communication.  Training has exactly one exchange: the sum over the group of [dtheta; loss].  It happens INSIDE
 * icnf_loss_grad_dp*: for small parameter vectors (the narrow-MLP family) in the gradient-reduction kernel itself,
 * through NVLink peer memory (every rank stores its slice into every peer's exchange buffer and adds the slices in
 * rank order: bit-identical results on all ranks, no extra launch); for large ones with ncclAllReduce on the same
 * stream.  NCCL is loaded at run time (libnccl.so.2); without it these entry points return ICNF_ERR_UNSUPPORTED.
 *
 * One process per GPU: rank 0 calls icnf_group_unique_id and ships the 128 bytes to the other ranks by any means;
 * every rank then calls icnf_group_join (collective).  One process, several GPUs: icnf_create_group on all handles
 * at once; calls that drive several handles from ONE thread must sit between icnf_group_start / icnf_group_end. */
typedef struct icnf_group_id { char internal[128]; } icnf_group_id;
ICNF_API int icnf_group_unique_id(icnf_group_id* id);
ICNF_API int icnf_group_join(icnf_handle* h, const icnf_group_id* id, int32_t n_ranks, int32_t rank);
ICNF_API int icnf_create_group(icnf_handle** handles, int32_t n);
ICNF_API int icnf_group_leave(icnf_handle* h);
/* peer_memory = 1 when the NVLink peer-memory exchange is available to this group */
ICNF_API int icnf_group_info(const icnf_handle* h, int32_t* n_ranks, int32_t* rank, int32_t* peer_memory);
/* Exact data-parallel mode: adaptive solves inside icnf_loss_grad_dp* take the controller's RMS
 * error norm over the GLOBAL batch -- one pair of doubles per step attempt exchanged through NVLink peer memory
 * inside the device loop -- so N shards accept and reject exactly the steps of the unsharded solve.  Off by
 * default (each shard then controls its own dt; results agree to solver tolerance).  Collective setting: every rank
 * of the group must choose the same value. */
ICNF_API int icnf_group_set_global_norm(icnf_handle* h, int enabled);
ICNF_API int icnf_group_start(void);
ICNF_API int icnf_group_end(void);
/* `loss` + gradient of the GLOBAL batch on every rank: arguments as icnf_loss_grad[_dev]; `global_batch` is the
 * unsharded batch size and noise->sample_offset the first global column of this rank's shard.  dxs stays local. */
ICNF_API int icnf_loss_grad_dp(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                      const float* xs, const icnf_noise* noise, const float* eps, const float* ys,
                      float* loss, float* dtheta, float* dxs, icnf_stats* stats,
                      int64_t B, int64_t global_batch);
ICNF_API int icnf_loss_grad_dp_dev(icnf_handle* h, int mode, const icnf_solver* sol, float t0, float t1,
                          const float* xs, const icnf_noise* noise, const float* eps, const float* ys,
                          float* loss, float* dtheta, float* dxs, icnf_stats* stats,
                          int64_t B, int64_t global_batch, void* stream);

/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
ICNF_API int64_t icnf_launch_count(const icnf_handle* h);


/* -- optimiser step ------------------------------------------ */
/* One step of `OptimiserChain(WeightDecay(lambda), Adam(eta, (beta1, beta2), epsilon))`, the
 * MLJ adapter's default (src/exts/mlj_ext/core_icnf.jl:17-24), on device-resident vectors:
 * g = grad + lambda * theta; m, v updated in place; theta -= eta * mhat / (sqrt(vhat) + epsilon).
 * `step` counts from 1.  All pointers are device pointers of length n; enqueued on `stream`. */
ICNF_API int icnf_adam_step_dev(float* theta, const float* grad, float* m, float* v, int64_t n, int64_t step, float eta,
                                float beta1, float beta2, float epsilon, float lambda, void* stream);

/* -- measurement support (bench.py) ------------------------------------------ */
/* When enabled, CUDA events are recorded on the launching stream around every
 * kernel of the next calls; icnf_kernel_times returns the device time in ms of the
 * last {forward solve or rhs, loss sum, backward, gradient reduce} launches
 * (-1 where a kernel did not run).  Off by default: no events, no overhead. */
ICNF_API int icnf_set_profiling(icnf_handle* h, int enabled);
ICNF_API int icnf_kernel_times(icnf_handle* h, float* ms4);
/* FP32 FMA-pipe throughput of `device` from an FFMA-chain microbenchmark, in
 * TFLOP/s: the roofline denominator of the narrow-MLP kernels. */
ICNF_API int icnf_measure_fp32_peak(int device, float* tflops);
/* Self-test of the tcgen05 GEMM behind precision = ICNF_BF16_TC: D (N x M, stored
 * D[n * M + m]) = A (M x K, row-major) * B (N x K, row-major)' with inputs rounded to
 * bf16 (split = 0) or to bf16 hi + lo pairs (split = 1, ICNF_BF16X3_TC) and fp32
 * accumulation; host buffers; runs on the current device. */
/* Host-only (no device is touched): the launch plan of the tiny family's backward kernel for a batch of B samples on a
 * device with `sm_count` SMs -- threads per CTA (= samples per tile), grid size, and first[0 .. n_blocks]: the first
 * thread of every weight-gradient block (block b owns threads [first[b], first[b + 1]); `first` needs 34 entries).
 * ICNF_ERR_UNSUPPORTED when the shape is not served by that kernel.  Used by the CPU test-suite. */
ICNF_API int icnf_backward_plan(const icnf_config* cfg, int exact, int sm_count, int64_t B, int32_t* threads, int32_t* grid,
                                int32_t* first, int32_t* n_blocks);

ICNF_API int icnf_tc_gemm_selftest(int M, int N, int K, const float* A, const float* B, float* D, int split);
/* Self-test of the weight-gradient form of the same kernel (K = the sample index, two operand pairs accumulated
 * into one tile, `nslices` split-K slices summed in slice order): D[n * M + m] = sum_k A[m][k] B[n][k] +
 * sum_k A2[m][k] B2[n][k], A, A2: M x K, B: N x K, B2: N2 x K (N2 <= N, 0 = no second pair), row-major host buffers. */
ICNF_API int icnf_tc_wgrad_selftest(int M, int N, int N2, int K, const float* A, const float* B, const float* A2,
                                    const float* B2, float* D, int split, int nslices);

#ifdef __cplusplus
}
#endif
#endif /* ICNF_B200_H */
