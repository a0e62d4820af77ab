// FluxStage instantiations: 4th-order face conversion (PPM / WENO) (see dispatch.cuh).
#include "dispatch.cuh"
namespace astrea {
ASTREA_DEFINE_FLUX(ho, 2)
}
