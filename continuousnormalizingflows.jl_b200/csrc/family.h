// family.h -- host-side interface between the C ABI (api.cu) and a kernel family.
// A family is a set of launchers specialised for one network shape ("tiny":
// compile-time shapes, one thread per sample) or for any shape ("generic").
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "common.cuh"

namespace icnf {

struct NetShape {
    int act, D, C, NL;
    int n[ICNF_MAX_LAYERS + 1];
    bool operator==(const NetShape& o) const {
        if (act != o.act || D != o.D || C != o.C || NL != o.NL) return false;
        for (int l = 0; l <= NL; ++l)
            if (n[l] != o.n[l]) return false;
        return true;
    }
};

struct Family {
    const char* name;
    NetShape shape;          // for tiny entries: the exact shape served
    int n_params;
    // `theta_host` is the HOST copy of the parameters (tiny kernels take the weights as a
    // kernel parameter; the generic family reads a.theta on the device and ignores it).
    // launchers return the CUDA error of the launch; `exact` selects the TestMode trace
    // `ws` is the family's per-handle workspace (null for families that need none)
    void* (*ws_create)(const icnf_config* cfg);
    void (*ws_destroy)(void* ws);
    cudaError_t (*on_params)(void* ws, const float* theta_dev, cudaStream_t st);   // after every icnf_set_params
    cudaError_t (*rhs)(void* ws, const float* theta_host, const RhsArgs& a, bool exact, int sm_count, cudaStream_t st);
    cudaError_t (*solve_fixed)(void* ws, const float* theta_host, const SolveArgs& a, int nvars, bool exact, int sm_count, cudaStream_t st);
    cudaError_t (*solve_adaptive)(void* ws, const float* theta_host, const SolveArgs& a, int nvars, bool exact, int grid, cudaStream_t st);
    // largest cooperative grid for the adaptive kernel on this device (0 = unsupported)
    int (*adaptive_max_grid)(bool exact, int sm_count);
    cudaError_t (*backward)(void* ws, const float* theta_host, const BackwardArgs& a, bool exact, int grid, cudaStream_t st);
    int (*backward_grid)(bool exact, int sm_count, long long B);
    int backward_partials_per_block;  // gradient partial rows written per CTA
    int supports_backward;
    int fuses_loss_sum_adaptive;     // the adaptive solve writes the scalar loss itself (SolveArgs::out_loss)
    int adaptive_threads;            // CTA size of the cooperative adaptive kernel (0: not cooperative)
    // host-only launch plan of the backward kernel (no device needed): CTA size, grid, first thread of every dW block
    // (`first` has room for 34 entries; entries [0, n_blocks] are filled), null when the family has no such plan
    void (*backward_plan)(bool exact, int sm_count, long long B, int* threads, int* grid, int* first, int* n_blocks);
    int ckpt_stages;                 // training checkpoints per step and sample, in units of D' floats (tiny: 6 stage inputs)
    // VCABM (variable-order Adams PECE, the reference's default alg); null: the family integrates with Tsit5 only
    cudaError_t (*solve_vcabm)(void* ws, const float* theta_host, const SolveArgs& a, int nvars, bool exact, int sm_count, cudaStream_t st);
    // does this adaptive solve run in a persistent kernel that can exchange the error-norm sums with the other ranks of a
    // group inside its device loop (exact data-parallel mode)?  null: no
    bool (*global_norm_capable)(void* ws, const SolveArgs& a, bool exact);
};

std::vector<const Family*>& tiny_registry();
struct TinyRegistrar {
    explicit TinyRegistrar(const Family* f) { tiny_registry().push_back(f); }
};

}  // namespace icnf
