#include "narrow_kernel.cuh"
namespace icnf {
namespace narrow {
ICNF_NARROW_INSTANCE_VCABM(launch_vcabm_o2_exact, 2, true)
}
}
