// sweep.cuh -- the per-particle sweeps of the hot path as fused sm_100a kernels, one
// instantiation per registered model:
//   init_kernel       abcde_init!        src/abcdez_init.jl:2-22  (+ prior draws, src/abcdez_smc.jl:242-243)
//   smc_sweep_kernel  abcdesmc_swarm!    src/abcdez_smc.jl:106-153
//   mc_sweep_kernel   abcdemc_swarm!     src/abcdez_mc.jl:5-61
// Thread-per-particle; theta rows are read with 16-byte vector loads (own row: streaming,
// partner rows: random gathers that touch ceil(8d/32) sectors thanks to the row layout);
// generation g is read-only and generation g+1 is written for EVERY particle (accepted ->
// proposal, otherwise copy-through), which is the reference's Jacobi double buffer
// (src/abcdez_smc.jl:337-350) without its four full-array copies per sweep.
#pragma once
#include "internal.h"
#include "ctrl.cuh"

namespace abcdez {

// ---------------------------------------------------------------------------------------
// control logic run by the last CTA of a sweep (src/abcdez_smc.jl:347-352)
// ---------------------------------------------------------------------------------------
// move the per-sweep accumulators (L2-resident, bypass L1) into the control block and reset them
__device__ __forceinline__ void sweep_collect(Ctrl* c)
{
    c->last_nsims = __ldcg(&c->acc.sweep_nsims); c->last_naccs = __ldcg(&c->acc.sweep_naccs);
    c->dmin = key_f64(__ldcg(&c->acc.dmin_key)); c->dmax = key_f64(__ldcg(&c->acc.dmax_key));
    int e = __ldcg(&c->acc.err);
    if (e && !c->err) c->err = e;
    c->acc.sweep_nsims = 0ull; c->acc.sweep_naccs = 0ull;
    c->acc.dmin_key = ~0ull; c->acc.dmax_key = 0ull;
}

__device__ inline void ctrl_after_smc_sweep(const PopDev& P, Ctrl* c)
{
    sweep_collect(c);
    c->nsims_total += (long long)c->last_nsims;
    c->naccs_iter += c->last_naccs;
    c->cur ^= 1;                                   // swap buffers, :347-350
    c->sweep_epoch += 1;
    c->sweep_idx += 1;
    c->n_sweeps += 1;
    if ((double)c->naccs_iter / (double)c->n_alive >= c->Kmcmc_min) {   // :352
        c->Ki = c->sweep_idx; c->sweeps_done = 1;
    } else if (c->sweep_idx >= c->Kmcmc) {
        c->sweeps_done = 1;
    }
    if (c->sweeps_done) ctrl_end_iter(P, c);       // :357-376
}

__device__ inline void ctrl_after_mc_sweep(Ctrl* c)
{
    sweep_collect(c);
    c->nsims_total += (long long)c->last_nsims;
    c->cur ^= 1;
    c->sweep_epoch += 1;
    c->n_sweeps += 1;
}

// block-level accumulation of the sweep counters + extrema(delta); integer atomics only, so
// the totals are independent of scheduling order
__device__ __forceinline__ void sweep_block_reduce(Ctrl* c, unsigned nsim, unsigned nacc,
                                                   unsigned long long kmin, unsigned long long kmax, int err)
{
    __shared__ unsigned s_sim[32], s_acc[32];
    __shared__ unsigned long long s_min[32], s_max[32];
    __shared__ int s_err;
    if (threadIdx.x == 0) s_err = 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    nsim = warp_sum_u(nsim); nacc = warp_sum_u(nacc);
    kmin = warp_min_u64(kmin); kmax = warp_max_u64(kmax);
    __syncthreads();
    if (err) atomicOr(&s_err, err);
    if (lane == 0) { s_sim[w] = nsim; s_acc[w] = nacc; s_min[w] = kmin; s_max[w] = kmax; }
    __syncthreads();
    if (w == 0) {
        unsigned a = lane < nw ? s_sim[lane] : 0u, b = lane < nw ? s_acc[lane] : 0u;
        unsigned long long mn = lane < nw ? s_min[lane] : ~0ull, mx = lane < nw ? s_max[lane] : 0ull;
        a = warp_sum_u(a); b = warp_sum_u(b); mn = warp_min_u64(mn); mx = warp_max_u64(mx);
        if (lane == 0) {
            if (a) atomicAdd(&c->acc.sweep_nsims, (unsigned long long)a);
            if (b) atomicAdd(&c->acc.sweep_naccs, (unsigned long long)b);
            atomicMin(&c->acc.dmin_key, mn);
            atomicMax(&c->acc.dmax_key, mx);
            if (s_err) atomicMax(&c->acc.err, s_err);
        }
    }
}

// StatsBase.wsample(rng, 1:N, alive) (src/abcdez_smc.jl:121,125) in O(1): t = u * n_alive, the
// ceil(t)-th alive particle; index 0 when t == 0 (StatsBase returns 1 then, alive or not).
__device__ __forceinline__ uint32_t wsample_alive(const uint32_t* __restrict__ alive_list, uint32_t n_alive,
                                                  uint32_t N, double u)
{
    double t = u * (double)n_alive;
    long long k = (long long)ceil(t);
    if (k <= 0) return 0u;
    if (k > (long long)n_alive) k = n_alive;
    return (n_alive == N) ? (uint32_t)(k - 1) : alive_list[k - 1];
}

constexpr int PARTNER_MAX_ATTEMPTS = 100000;
constexpr int INIT_MAX_ATTEMPTS = 100000;

// ---------------------------------------------------------------------------------------
// abcde_init!  src/abcdez_init.jl:2-22
// ---------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS)
init_kernel(PopDev P, PriorDev pr, ModelData md, uint64_t seed, int draw_prior)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    const int cur = c->cur;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned redraws = 0; int err = 0;
    unsigned long long kmin = ~0ull, kmax = 0ull;
    if (i < P.N) {
        const uint32_t pid = P.id0 + i;
        double th[D], x[D], blob[NB > 0 ? NB : 1];
        double lp;
        if (draw_prior) {                                  // src/abcdez_smc.jl:242-243
            prior_sample<D>(pr, seed, pid, 0u, th);
            push_p<D>(pr, th, x);
            lp = prior_logpdf<D>(pr, x);
        } else {
            load_row<D>(P.theta[cur], i, th);
            lp = P.logpi[cur][i];
        }
        double dl = NAN;
        if (isfinite(lp)) {                                // init.jl:9-13
            SimRng r(seed, pid, 0u, TAG_INIT_MODEL);
            push_p<D>(pr, th, x);
            dl = M::run(x, md.v, r, blob);
        }
        uint32_t attempt = 0;
        while (!isfinite(dl) || !isfinite(lp)) {           // init.jl:14-20
            if (++attempt >= (uint32_t)INIT_MAX_ATTEMPTS) { err = ABCDEZ_ERR_INIT_RETRY; break; }
            prior_sample<D>(pr, seed, pid, attempt, th);
            push_p<D>(pr, th, x);
            lp = prior_logpdf<D>(pr, x);
            SimRng r(seed, pid, attempt, TAG_INIT_MODEL);
            dl = M::run(x, md.v, r, blob);
            redraws++;
        }
        store_row<D>(P.theta[cur], i, th);
        P.logpi[cur][i] = lp;
        P.delta[cur][i] = dl;
#pragma unroll
        for (int k = 0; k < NB; ++k) P.blob[cur][(size_t)i * NB + k] = blob[k];
        kmin = kmax = f64_key(dl);
    }
    // counters: reuse the sweep accumulators (nsims slot carries the redraw count)
    sweep_block_reduce(c, redraws, 0u, kmin, kmax, err);
    if (last_block(&c->acc.ticket[0], gridDim.x)) {
        if (threadIdx.x == 0) {
            sweep_collect(c);
            c->redraws += (long long)c->last_nsims;
        }
    }
}

// ---------------------------------------------------------------------------------------
// abcdesmc_swarm!  src/abcdez_smc.jl:106-153
// ---------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS)
smc_sweep_kernel(PopDev P, PriorDev pr, ModelData md, SweepInj inj)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    if (c->stop | c->sweeps_done) return;                  // skipped sweep (early exit :352 / stop :376)
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t N = P.N, n_alive = c->n_alive;
    const double eps = c->eps, gamma0 = c->gamma0, gsig = c->gsig;
    const int kind = c->kind;
    const uint64_t seed = c->seed;
    const uint32_t epoch = c->sweep_epoch;
    const double* __restrict__ th = P.theta[cur];

    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nsim = 0, nacc = 0; int err = 0;
    unsigned long long kmin = ~0ull, kmax = 0ull;
    if (i < N) {
        double ti[D], bl[NB > 0 ? NB : 1];
        load_row<D>(th, i, ti);
        double lpi = P.logpi[cur][i], dli = P.delta[cur][i];
#pragma unroll
        for (int k = 0; k < NB; ++k) bl[k] = P.blob[cur][(size_t)i * NB + k];
        uint8_t flag = 0;
        if (P.alive[i]) {                                                  // :114
            const uint32_t pid = P.id0 + i;
            uint32_t a, b;
            if (inj.a) { a = (uint32_t)inj.a[i]; b = (uint32_t)inj.b[i]; }
            else {
                Stream ps(seed, pid, epoch, TAG_PARTNER);
                double u1, u2; uint32_t att = 0;
                a = i;
                while (a == i) {                                           // :119-122
                    if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                    ps.u2(att++, u1, u2);
                    a = wsample_alive(P.alive_list, n_alive, N, u1);
                }
                att = 0; b = a;
                while (b == a || b == i) {                                 // :123-126
                    if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                    ps.u2(att++, u1, u2);
                    b = wsample_alive(P.alive_list, n_alive, N, u2);
                }
            }
            if (!err) {
                double ta[D], tb[D], thp[D], x[D];
                load_row<D>(th, a, ta);
                load_row<D>(th, b, tb);
                Stream ms(seed, pid, epoch, TAG_MOVE);
                double z, z2;
                if (inj.z) z = inj.z[i]; else ms.n2(0u, z, z2);
                const double g = gamma0 * (1.0 + z * gsig);                // :128
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    double diff = ta[k] - tb[k];
                    double sc = diff * g;
                    thp[k] = ti[k] + sc;
                }
                push_p<D>(pr, thp, x);
                double lp = prior_logpdf<D>(pr, x);                        // :134
                if (!(lp < 0.0 && isinf(lp))) {                            // :135
                    SimRng r(seed, pid, epoch, TAG_MODEL);
                    double blp[NB > 0 ? NB : 1];
                    double dp = M::run(x, md.v, r, blp);                   // :137
                    nsim = 1; flag |= ABCDEZ_FLAG_SIM;                     // :138
                    double w = lp - lpi;                                   // :140-141, left to right
                    w = w + abck_logpdf(kind, eps, dp);
                    w = w - abck_logpdf(kind, eps, dli);
                    bool acc = (0.0 <= w);
                    if (!acc) {                                            // :145, uniform only when w < 0
                        double u, u2;
                        if (inj.u) u = inj.u[i]; else ms.u2(1u, u, u2);
                        acc = (plog(u) < w);
                    }
                    if (acc) {                                             // :146-150
                        dli = dp; lpi = lp;
#pragma unroll
                        for (int k = 0; k < D; ++k) ti[k] = thp[k];
#pragma unroll
                        for (int k = 0; k < NB; ++k) bl[k] = blp[k];
                        nacc = 1; flag |= ABCDEZ_FLAG_ACC;
                    }
                }
            }
        }
        store_row<D>(P.theta[nxt], i, ti);
        P.logpi[nxt][i] = lpi;
        P.delta[nxt][i] = dli;
#pragma unroll
        for (int k = 0; k < NB; ++k) P.blob[nxt][(size_t)i * NB + k] = bl[k];
        if (inj.flags) inj.flags[i] = flag;
        kmin = kmax = f64_key(dli);
    }
    sweep_block_reduce(c, nsim, nacc, kmin, kmax, err);
    if (last_block(&c->acc.ticket[0], gridDim.x)) {
        if (threadIdx.x == 0) ctrl_after_smc_sweep(P, c);
    }
}

// ---------------------------------------------------------------------------------------
// abcdemc_swarm!  src/abcdez_mc.jl:5-61
// ---------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS)
mc_sweep_kernel(PopDev P, PriorDev pr, ModelData md, SweepInj inj, McArgs mc)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t N = P.N;
    const double gamma0 = c->gamma0, gsig = c->gsig;
    const uint64_t seed = c->seed;
    const uint32_t epoch = c->sweep_epoch;
    const double* __restrict__ th = P.theta[cur];

    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nsim = 0, nacc = 0; int err = 0;
    unsigned long long kmin = ~0ull, kmax = 0ull;
    if (i < N) {
        double ti[D], bl[NB > 0 ? NB : 1];
        load_row<D>(th, i, ti);
        double lpi = P.logpi[cur][i], dli = P.delta[cur][i];
#pragma unroll
        for (int k = 0; k < NB; ++k) bl[k] = P.blob[cur][(size_t)i * NB + k];
        uint8_t flag = 0;
        const uint32_t pid = P.id0 + i;
        uint32_t s = i;                                                    // :18
        const double eps = (dli <= mc.eps_target) ? mc.eps_target : mc.eps_pop;   // :19
        if (dli > eps) {                                                   // :20-24
            if (inj.s) s = (uint32_t)inj.s[i];
            else {
                // cnt = #{delta <= delta_i}: upper bound in the sorted distances
                uint32_t lo = 0, hi = N;
                while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (mc.sorted_delta[m] <= dli) lo = m + 1; else hi = m; }
                Stream cs(seed, pid, epoch, TAG_MC);
                double u1, u2; cs.u2(0u, u1, u2);
                long long k = (long long)floor(u1 * (double)lo);
                if (k >= (long long)lo) k = (long long)lo - 1;
                s = mc.order[k];
            }
        }
        uint32_t a, b;
        if (inj.a) { a = (uint32_t)inj.a[i]; b = (uint32_t)inj.b[i]; }
        else {
            Stream ps(seed, pid, epoch, TAG_PARTNER);
            double u1, u2; uint32_t att = 0;
            a = s;
            while (a == s) {                                               // :25-28
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                long long k = (long long)floor(u1 * (double)N); if (k >= (long long)N) k = (long long)N - 1;
                a = (uint32_t)k;
            }
            att = 0; b = a;
            while (b == a || b == s) {                                     // :29-32
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                long long k = (long long)floor(u2 * (double)N); if (k >= (long long)N) k = (long long)N - 1;
                b = (uint32_t)k;
            }
        }
        if (!err) {
            double tsr[D], ta[D], tb[D], thp[D], x[D];
            load_row<D>(th, s, tsr);
            load_row<D>(th, a, ta);
            load_row<D>(th, b, tb);
            Stream ms(seed, pid, epoch, TAG_MOVE);
            double z, z2;
            if (inj.z) z = inj.z[i]; else ms.n2(0u, z, z2);
            const double g = gamma0 * (1.0 + z * gsig);                    // :34
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double diff = ta[k] - tb[k];
                double sc = diff * g;
                thp[k] = tsr[k] + sc;
            }
            push_p<D>(pr, thp, x);
            double lp = prior_logpdf<D>(pr, x);                            // :41
            double w_prior = lp - lpi;                                     // :42 (logpi[i], not [s])
            double u, u2;
            if (inj.u) u = inj.u[i]; else ms.u2(1u, u, u2);                // :43, always drawn
            if (!(plog(u) > fmin(0.0, w_prior))) {
                nsim = 1; flag |= ABCDEZ_FLAG_SIM;                         // :44
                SimRng r(seed, pid, epoch, TAG_MODEL);
                double blp[NB > 0 ? NB : 1];
                double dp = M::run(x, md.v, r, blp);                       // :45
                if (dp <= fmax(eps, dli)) {                                // :54-59
                    dli = dp; lpi = lp;
#pragma unroll
                    for (int k = 0; k < D; ++k) ti[k] = thp[k];
#pragma unroll
                    for (int k = 0; k < NB; ++k) bl[k] = blp[k];
                    nacc = 1; flag |= ABCDEZ_FLAG_ACC;
                }
            }
        }
        store_row<D>(P.theta[nxt], i, ti);
        P.logpi[nxt][i] = lpi;
        P.delta[nxt][i] = dli;
#pragma unroll
        for (int k = 0; k < NB; ++k) P.blob[nxt][(size_t)i * NB + k] = bl[k];
        if (inj.flags) inj.flags[i] = flag;
        kmin = kmax = f64_key(dli);
    }
    sweep_block_reduce(c, nsim, nacc, kmin, kmax, err);
    if (last_block(&c->acc.ticket[0], gridDim.x)) {
        if (threadIdx.x == 0) ctrl_after_mc_sweep(c);
    }
}

// one dist! evaluation per row (stage-level model parity); dense N x D input
template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS)
simulate_kernel(ModelData md, int64_t N, const double* __restrict__ theta_pushed, uint64_t seed,
                uint32_t epoch, uint32_t tag, uint32_t id0, double* __restrict__ dist, double* __restrict__ blobs)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double x[D], blob[NB > 0 ? NB : 1];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = theta_pushed[i * D + k];
    SimRng r(seed, id0 + (uint32_t)i, epoch, tag);
    dist[i] = M::run(x, md.v, r, blob);
    if (blobs) {
#pragma unroll
        for (int k = 0; k < NB; ++k) blobs[i * NB + k] = blob[k];
    }
}

// ---------------------------------------------------------------------------------------
// launchers + registry
// ---------------------------------------------------------------------------------------
static inline unsigned grid_for(int64_t N, int threads) { return (unsigned)((N + threads - 1) / threads); }

template <class M>
static void l_init(cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, uint64_t seed, int dp)
{
    init_kernel<M><<<grid_for(P.N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(P, pr, md, seed, dp);
}
template <class M>
static void l_smc(cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj)
{
    smc_sweep_kernel<M><<<grid_for(P.N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
}
template <class M>
static void l_mc(cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj,
                 const McArgs& mc)
{
    mc_sweep_kernel<M><<<grid_for(P.N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(P, pr, md, inj, mc);
}
template <class M>
static void l_sim(cudaStream_t st, const PriorDev*, const ModelData& md, int64_t N, const double* th, uint64_t seed,
                  uint32_t epoch, uint32_t tag, uint32_t id0, double* dist, double* blobs)
{
    simulate_kernel<M><<<grid_for(N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(md, N, th, seed, epoch, tag, id0, dist, blobs);
}


template <class M>
static ModelOps make_ops()
{
    ModelOps o;
    o.name = M::name; o.d = M::D; o.blob = M::BLOB;
    o.init = &l_init<M>; o.smc_sweep = &l_smc<M>; o.mc_sweep = &l_mc<M>; o.simulate = &l_sim<M>;
    return o;
}

// one accessor per model, defined in the inst_*.cu translation units (compiled in parallel)
#define ABCDEZ_DEFINE_MODEL(fn, M) const ModelOps* fn() { static const ModelOps o = make_ops<M>(); return &o; }

}  // namespace abcdez
