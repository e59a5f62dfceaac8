// inst_corr10.cu -- instantiates the fused sweep kernels (sweep.cuh) for a group of registered models.
#include "sweep.cuh"

namespace abcdez {
ABCDEZ_DEFINE_MODEL(ops_gauss_corr10, GaussCorr10)
}  // namespace abcdez
