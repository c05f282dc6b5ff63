#include "narrow_kernel.cuh"
namespace icnf {
namespace narrow {
ICNF_NARROW_INSTANCE(launch_o4_softplus_exact, 4, ICNF_ACT_SOFTPLUS, true)
}
}
