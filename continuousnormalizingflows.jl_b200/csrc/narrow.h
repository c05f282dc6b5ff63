// narrow.h -- single-launch solve (exact trace or Hutchinson) for narrow MLPs with two hidden layers (narrow.cu); the
// generic family routes inference-side solves through it when the shape qualifies.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace icnf {
namespace narrow {
bool supported(const icnf_config& cfg, bool exact, const SolveArgs& a);
size_t smem_bytes(const icnf_config& cfg, bool exact);
// `amat`: the exact-trace matrix of generic.cu's on_params (device); a.wu / a.wk: two S x B float buffers each
cudaError_t solve(const icnf_config& cfg, const float* amat, const SolveArgs& a, int nvars, bool exact, bool adaptive, int sm_count, cudaStream_t st);
}  // namespace narrow
}  // namespace icnf
