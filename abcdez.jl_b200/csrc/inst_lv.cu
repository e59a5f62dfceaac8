// inst_lv.cu -- instantiates the fused sweep kernels (sweep.cuh) for a group of registered models.
#include "sweep.cuh"

namespace abcdez {
ABCDEZ_DEFINE_MODEL(ops_lotka_volterra, LotkaVolterra)
ABCDEZ_DEFINE_MODEL(ops_lotka_volterra_lin, LotkaVolterraLin)
}  // namespace abcdez
