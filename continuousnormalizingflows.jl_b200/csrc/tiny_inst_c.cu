// tiny-family instantiations, part C: small shapes that exercise every template
// axis in the parity tests (conditioning input, autonomous flow, tanh / sigmoid,
// one and three hidden layers).
#include "tiny_launch.cuh"

// conditioned, nvars 2, naug 1, ncond 2: n_in = 3 + 1 + 2 = 6 -> 8 -> 8 -> 3
ICNF_REGISTER_TINY(ICNF_ACT_SOFTPLUS, 3, 2, 3, 6, 8, 8, 3)
// autonomous tanh, one hidden layer, nvars 3, naug 1: 4 -> 8 -> 4
ICNF_REGISTER_TINY(ICNF_ACT_TANH, 4, 0, 2, 4, 8, 4)
// sigmoid, three hidden layers, nvars 2, naug 0: 3 -> 7 -> 9 -> 5 -> 2
ICNF_REGISTER_TINY(ICNF_ACT_SIGMOID, 2, 0, 4, 3, 7, 9, 5, 2)
// linear autonomous field z -> A z (closed-form log-density test), nvars 2: 2 -> 2
ICNF_REGISTER_TINY(ICNF_ACT_IDENTITY, 2, 0, 1, 2, 2)
