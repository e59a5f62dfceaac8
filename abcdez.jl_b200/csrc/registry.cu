// registry.cu -- the static model registry behind abcdez_model_lookup / abcdez_model_bind and the
// dense prior stage kernels (abcdez_prior_sample / _logpdf / _push).
#include "internal.h"

namespace abcdez {

const ModelOps* ops_gauss1d(); const ModelOps* ops_gauss1d_blob(); const ModelOps* ops_gauss_corr10();
const ModelOps* ops_dirac(); const ModelOps* ops_normdu(); const ModelOps* ops_twod(); const ModelOps* ops_twod_inf();
const ModelOps* ops_mixture(); const ModelOps* ops_wiener(); const ModelOps* ops_lotka_volterra();
const ModelOps* ops_birth_death(); const ModelOps* ops_socks(); const ModelOps* ops_gk(); const ModelOps* ops_gk_f32();
const ModelOps* ops_lotka_volterra_lin();


const ModelOps* model_ops(int id)
{
    switch (id) {
    case M_GAUSS1D: return ops_gauss1d();
    case M_GAUSS1D_BLOB: return ops_gauss1d_blob();
    case M_GAUSS_CORR10: return ops_gauss_corr10();
    case M_DIRAC: return ops_dirac();
    case M_NORMDU: return ops_normdu();
    case M_TWOD: return ops_twod();
    case M_TWOD_INF: return ops_twod_inf();
    case M_MIXTURE: return ops_mixture();
    case M_WIENER: return ops_wiener();
    case M_LOTKA_VOLTERRA: return ops_lotka_volterra();
    case M_BIRTH_DEATH: return ops_birth_death();
    case M_GK: return ops_gk();            // CTA-cooperative simulator (gk.cu)
    case M_SOCKS: return ops_socks();
    case M_GK_F32: return ops_gk_f32();    // relaxed-precision mode of the same simulator
    case M_LOTKA_VOLTERRA_LIN: return ops_lotka_volterra_lin();
    }
    return rtc_model_ops(id);          // runtime-compiled models (rtc.cu) follow the static ones
}
int model_count() { return M_COUNT + rtc_model_count(); }

static inline unsigned grid_for(int64_t N, int threads) { return (unsigned)((N + threads - 1) / threads); }

// ---------------------------------------------------------------------------------------
// prior stage calls on dense N x d arrays (abcdez_prior_sample / _logpdf / _push)
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void prior_op_kernel(const __grid_constant__ PriorDev pr, int64_t N, int op, const double* __restrict__ in,
                                double* __restrict__ out, const __grid_constant__ PhiloxKeys seed, uint32_t epoch, uint32_t id0)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double th[D], x[D];
    if (op == PRIOR_OP_SAMPLE) {
        prior_sample<D>(pr, seed, id0 + (uint32_t)i, epoch, th);
#pragma unroll
        for (int k = 0; k < D; ++k) out[i * D + k] = th[k];
        return;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) th[k] = in[i * D + k];
    push_p<D>(pr, th, x);
    if (op == PRIOR_OP_PUSH) {
#pragma unroll
        for (int k = 0; k < D; ++k) out[i * D + k] = x[k];
    } else {
        out[i] = prior_logpdf<D>(pr, x);
    }
}

template <int D>
static void l_prior(cudaStream_t st, const PriorDev& pr, int64_t N, int op, const double* in, double* out,
                    uint64_t seed, uint32_t epoch, uint32_t id0)
{
    prior_op_kernel<D><<<grid_for(N, 128), 128, 0, st>>>(pr, N, op, in, out, philox_keys(seed), epoch, id0);
}

void launch_prior_op(cudaStream_t st, int d, const PriorDev& pr, int64_t N, int op, const double* in, double* out,
                     uint64_t seed, uint32_t epoch, uint32_t id0)
{
    switch (d) {
#define C(DD) case DD: l_prior<DD>(st, pr, N, op, in, out, seed, epoch, id0); break;
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16)
#undef C
    }
}

}  // namespace abcdez
