// inst_d2.cu -- instantiates the fused sweep kernels (sweep.cuh) for a group of registered models.
#include "sweep.cuh"

namespace abcdez {
ABCDEZ_DEFINE_MODEL(ops_normdu, NormDU)
ABCDEZ_DEFINE_MODEL(ops_twod, TwoD<false>)
ABCDEZ_DEFINE_MODEL(ops_twod_inf, TwoD<true>)
}  // namespace abcdez
