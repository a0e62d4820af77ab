// Sweep-kernel instantiations for one reconstruction scheme (see dispatch.cuh).
#include "dispatch.cuh"
namespace astrea {
ASTREA_DEFINE_SCHEME(weno3, SCH_WENO3)
}
