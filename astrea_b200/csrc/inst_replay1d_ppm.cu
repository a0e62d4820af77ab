// Persistent 1D replay kernels of one reconstruction scheme, four Riemann solvers (see replay1d.cuh).
#include "replay1d.cuh"
namespace astrea {
ASTREA_DEFINE_REPLAY(ppm, SCH_PPM)
}
