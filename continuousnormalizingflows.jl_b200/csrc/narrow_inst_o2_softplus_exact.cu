#include "narrow_kernel.cuh"
namespace icnf {
namespace narrow {
ICNF_NARROW_INSTANCE(launch_o2_softplus_exact, 2, ICNF_ACT_SOFTPLUS, true)
}
}
