// ctrl.cuh -- device-side schedule logic shared by the sweep and bookkeeping kernels
// (src/abcdez_smc.jl:284-292, :301, :357-376 of the reference).
#pragma once
#include "internal.h"
#include "seqsum.cuh"

namespace abcdez {

// ---------------------------------------------------------------------------------------
// history record, src/abcdez_smc.jl:284-292,362-370
// ---------------------------------------------------------------------------------------
__device__ inline void push_hist(const PopDev& P, Ctrl* c, double eps, double ess, double facc, int K)
{
    if (c->hist_len < c->hist_cap && P.hist) {
        double* h = P.hist + (size_t)c->hist_len * 8;
        h[0] = eps; h[1] = c->dmin; h[2] = c->dmax; h[3] = c->logZ; h[4] = ess; h[5] = facc;
        h[6] = c->gamma0; h[7] = (double)K;
        c->hist_len += 1;
        c->hist_iter = 1;                      // the most recent record is in the buffer (patch_extrema may complete it)
    } else {
        c->hist_iter = 0;                      // dropped: the buffer is full (reported as hist_overflow) or absent
        if (P.hist) c->hist_overflow += 1;
    }
}

// prepare the radix select of the next quantile call: Statistics.quantile type 7 at
// src/abcdez_smc.jl:301: aleph = fma(n, p, 1-p); j = clamp(trunc(aleph), 1, n-1); g = clamp(aleph-j, 0, 1)
__device__ inline void select_setup(Ctrl* c)
{
    unsigned long long n = c->n_alive_g;
    double p = c->alpha, m = 1.0 - p;
    double aleph = fma((double)n, p, m);
    long long j = (long long)trunc(aleph);
    if (j > (long long)n - 1) j = (long long)n - 1;
    if (j < 1) j = 1;
    double g = aleph - (double)j;
    g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
    c->sel_j = (unsigned long long)j;
    c->sel_rank = (unsigned long long)(j - 1);
    c->sel_prefix = 0ull;
    c->q_gamma = g;
    c->acc.cnt_le = 0ull;
    c->acc.min_gt_key = ~0ull;
}

// end of an SMC iteration, src/abcdez_smc.jl:357-376 (called by the last CTA of the last sweep)
static __device__ __noinline__ void ctrl_end_iter(const PopDev& P, Ctrl* c)
{
    c->facc = (double)c->naccs_iter / ((double)c->n_alive_g * (double)c->Ki);   // :357
    c->eps_k = c->eps;                                                         // :360
    c->iters += 1;
    push_hist(P, c, c->eps, c->ess, c->facc, c->Ki);                           // :362-370
    if (c->n_alive_g == 0) { c->status = ABCDEZ_ERR_NO_ALIVE; c->stop = 1; }     // :375
    else if (c->eps <= c->eps_target || c->nsims_total >= c->nsims_max || c->facc < c->facc_stop)
        c->stop = 1;                                                           // :376
    if (c->max_iters > 0 && c->iters >= c->max_iters) c->stop = 1;
    if (c->err) c->stop = 1;
    select_setup(c);
}

// ---------------------------------------------------------------------------------------
// helpers shared by bookkeeping.cu and head.cu
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void load4_f64(const double* __restrict__ p, size_t i0, uint32_t N, double v[4])
{
    if (i0 + 3 < N) {
        double2 a = *reinterpret_cast<const double2*>(p + i0), b = *reinterpret_cast<const double2*>(p + i0 + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < N) ? p[i0 + k] : 0.0;
    }
}

__device__ __forceinline__ void store4_f64(double* __restrict__ p, size_t i0, uint32_t N, const double v[4])
{
    if (i0 + 3 < N) {
        *reinterpret_cast<double2*>(p + i0) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(p + i0 + 2) = make_double2(v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i0 + k < N) p[i0 + k] = v[k];
    }
}

__device__ __forceinline__ uint32_t load4_u8(const uint8_t* __restrict__ p, size_t i0, uint32_t N)
{
    if (i0 + 3 < N) return *reinterpret_cast<const uint32_t*>(p + i0);
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i0 + k < N) r |= (uint32_t)p[i0 + k] << (8 * k);
    return r;
}

// extrema(delta) of the generation that the previous iteration left behind: the sweeps no longer touch
// every particle, so ranges_eps (src/abcdez_smc.jl:363) is taken from the first select pass of the next
// iteration (same delta array) and written into the history record that iteration pushed
static __device__ __noinline__ void patch_extrema(const PopDev& P, Ctrl* c)
{
    c->dmin = key_f64(__ldcg(&c->acc.dmin_key)); c->dmax = key_f64(__ldcg(&c->acc.dmax_key));
    c->acc.dmin_key = ~0ull; c->acc.dmax_key = 0ull;
    if (P.hist && c->hist_iter && c->hist_len > 0 && c->hist_len <= c->hist_cap) {   // only the record this generation pushed
        double* h = P.hist + (size_t)(c->hist_len - 1) * 8;
        h[1] = c->dmin; h[2] = c->dmax;
    }
}

// reweight pass B + the decisions of :318-324.  Builds the sequential-sum tables for the
// closed-form resampling when it will be needed.
// n_alive: this rank's alive count; n_alive_g: the whole population's (equal on one GPU)
static __device__ __noinline__ void ctrl_after_reweight(const PopDev& P, Ctrl* c, double sumsq, unsigned n_alive, unsigned n_alive_g)
{
    c->n_alive = n_alive; c->n_alive_g = n_alive_g;
    c->ess = 1.0 / sumsq;                                                   // :8,:323
    c->naccs_iter = 0ull; c->Ki = c->Kmcmc;                                 // :318-319
    c->sweep_idx = 0; c->sweeps_done = 0;
    if (c->facc < c->facc_min) c->gamma0 *= c->facc_tune;                   // :320
    c->do_resample = (c->ess < c->ess_min) ? 1 : 0;                         // :324
    if (c->do_resample && abck_is_indicator(c->kind)) {
        seqtab_build(&P.tabs[0], __ldcg(&c->acc.w_alive), (unsigned long long)n_alive_g);
        seqtab_build(&P.tabs[1], 1.0 / (double)P.Ng, (unsigned long long)P.Ng);
    }
}


}  // namespace abcdez
