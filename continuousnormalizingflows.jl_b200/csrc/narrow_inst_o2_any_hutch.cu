#include "narrow_kernel.cuh"
namespace icnf {
namespace narrow {
ICNF_NARROW_INSTANCE(launch_o2_any_hutch, 2, -1, false)
}
}
