// tiny-family instantiations, part A: the BASELINE.json configs 1 and 2.
#include "tiny_launch.cuh"

// config 2: two-moons RNODE, nvars 2, naug 0, default width 4 n_in: 3 -> 12 -> 12 -> 2 softplus
ICNF_REGISTER_TINY(ICNF_ACT_SOFTPLUS, 2, 0, 3, 3, 12, 12, 2)
// config 1 (examples/usage.jl, benchmark/benchmarks.jl): nvars 1, naug 2: 4 -> 16 -> 16 -> 3 softplus
ICNF_REGISTER_TINY(ICNF_ACT_SOFTPLUS, 3, 0, 3, 4, 16, 16, 3)
