// Sweep-kernel instantiations for one reconstruction scheme (see dispatch.cuh).
#include "dispatch.cuh"
namespace astrea {
ASTREA_DEFINE_SCHEME(pcm, SCH_PCM)
}
