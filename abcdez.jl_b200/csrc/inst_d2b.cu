// inst_d2b.cu -- instantiates the fused sweep kernels (sweep.cuh) for a group of registered models.
#include "sweep.cuh"

namespace abcdez {
ABCDEZ_DEFINE_MODEL(ops_wiener, Wiener)
ABCDEZ_DEFINE_MODEL(ops_socks, Socks)
ABCDEZ_DEFINE_MODEL(ops_birth_death, BirthDeath)
}  // namespace abcdez
