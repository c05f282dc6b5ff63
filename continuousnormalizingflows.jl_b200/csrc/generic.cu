// generic.cu -- kernel family for arbitrary network shapes (placeholder until the
// shared-memory tiled fp32 family lands; shapes without a tiny instantiation are
// reported as unsupported rather than computed anywhere else).
#include "family.h"

namespace icnf {
const Family* generic_family() { return nullptr; }
}  // namespace icnf
