#include "narrow_kernel.cuh"
namespace icnf {
namespace narrow {
ICNF_NARROW_INSTANCE(launch_o4_any, 4, -1)
}
}
