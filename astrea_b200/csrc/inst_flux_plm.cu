// FluxStage instantiations: pointwise face conversion (PLM) (see dispatch.cuh).
#include "dispatch.cuh"
namespace astrea {
ASTREA_DEFINE_FLUX(plm, 1)
}
