#include "narrow_kernel.cuh"
namespace icnf {
namespace narrow {
ICNF_NARROW_INSTANCE(launch_o4_any_hutch, 4, -1, false)
}
}
