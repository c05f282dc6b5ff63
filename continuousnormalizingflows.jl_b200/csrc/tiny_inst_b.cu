// tiny-family instantiations, part B: the shapes of the reference's smoke tests
// (test/ci_tests/smoke_tests.jl: nvars 2, default ICNF -> D' = 5, n_hidden = 4 n_in).
#include "tiny_launch.cuh"

// unconditioned: n_in = 6, 6 -> 24 -> 24 -> 5
ICNF_REGISTER_TINY(ICNF_ACT_SOFTPLUS, 5, 0, 3, 6, 24, 24, 5)
