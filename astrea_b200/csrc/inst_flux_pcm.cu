// FluxStage instantiations: piecewise-constant faces (see dispatch.cuh).
#include "dispatch.cuh"
namespace astrea {
ASTREA_DEFINE_FLUX(pcm, 0)
}
