// inst_d1.cu -- instantiates the fused sweep kernels (sweep.cuh) for a group of registered models.
#include "sweep.cuh"

namespace abcdez {
ABCDEZ_DEFINE_MODEL(ops_gauss1d, Gauss1D)
ABCDEZ_DEFINE_MODEL(ops_gauss1d_blob, Gauss1DBlob)
ABCDEZ_DEFINE_MODEL(ops_dirac, Dirac)
ABCDEZ_DEFINE_MODEL(ops_mixture, Mixture)
}  // namespace abcdez
