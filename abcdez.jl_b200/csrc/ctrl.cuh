// ctrl.cuh -- device-side schedule logic shared by the sweep and bookkeeping kernels
// (src/abcdez_smc.jl:284-292, :301, :357-376 of the reference).
#pragma once
#include "internal.h"

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
    }
}

// prepare the radix select of the next quantile call: Statistics.quantile type 7 at
// src/abcdez_smc.jl:301: aleph = fma(n, p, 1-p); j = clamp(trunc(aleph), 1, n-1); g = clamp(aleph-j, 0, 1)
__device__ inline void select_setup(Ctrl* c)
{
    unsigned long long n = c->n_alive;
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
__device__ inline void ctrl_end_iter(const PopDev& P, Ctrl* c)
{
    c->facc = (double)c->naccs_iter / ((double)c->n_alive * (double)c->Ki);   // :357
    c->eps_k = c->eps;                                                         // :360
    c->iters += 1;
    push_hist(P, c, c->eps, c->ess, c->facc, c->Ki);                           // :362-370
    if (c->n_alive == 0) { c->status = ABCDEZ_ERR_NO_ALIVE; c->stop = 1; }     // :375
    else if (c->eps <= c->eps_target || c->nsims_total >= c->nsims_max || c->facc < c->facc_stop)
        c->stop = 1;                                                           // :376
    if (c->max_iters > 0 && c->iters >= c->max_iters) c->stop = 1;
    if (c->err) c->stop = 1;
    select_setup(c);
}


}  // namespace abcdez
