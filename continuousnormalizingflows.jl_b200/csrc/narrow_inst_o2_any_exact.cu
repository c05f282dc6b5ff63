#include "narrow_kernel.cuh"
namespace icnf {
namespace narrow {
ICNF_NARROW_INSTANCE(launch_o2_any_exact, 2, -1, true)
}
}
