// sweep.cuh -- the per-particle sweeps of the hot path as fused sm_100a kernels, one
// instantiation per registered model:
//   init_kernel       abcde_init!        src/abcdez_init.jl:2-22  (+ prior draws, src/abcdez_smc.jl:242-243)
//   smc_sweep_kernel  abcdesmc_swarm!    src/abcdez_smc.jl:106-153
//   mc_sweep_kernel   abcdemc_swarm!     src/abcdez_mc.jl:5-61
// and, for heavy simulators (M::SPLIT), the same three steps as propose -> queue-driven simulate -> accept launches
// (smc_propose / mc_propose / init_draw, simulate_queue_kernel, smc_accept / mc_accept / init_finish; second half of this file).
// Relaxed-parity template switches of the fused kernels: SEG (warp-coherent partner segments), F32 (float theta rows).
//
// Jacobi double buffer without the copies.  The reference copies all four state arrays before every
// sweep (identity.(), src/abcdez_smc.jl:337-340) so that generation g stays read-only while g+1 is
// written.  Here the two device buffers are kept in sync *lazily*: `moved[i]` records that particle i
// was accepted in the previous executed sweep, i.e. its row in the other buffer is stale.  A sweep
// (1) repairs exactly those rows (read from g, written to g+1), (2) writes accepted proposals to g+1
// and (3) updates `moved`.  Rejected and dead particles cost no store at all, and a dead particle
// whose row is in sync is not even read.  Resampling and (re)initialisation mark every row stale.
//
// Thread j works on particle list[j] (alive first, then dead; bookkeeping.cu compact_kernel), so warps
// are homogeneous.  theta rows are read with 16-byte vector loads; partner rows are random gathers that
// touch ceil(8d/32) sectors thanks to the particle-major row layout.
#pragma once
#include "internal.h"
#include "ctrl.cuh"
#include "comm.cuh"

namespace abcdez {

// ---------------------------------------------------------------------------------------
// control logic run by the last CTA of a sweep (src/abcdez_smc.jl:347-352)
// ---------------------------------------------------------------------------------------
// move the per-sweep accumulators (L2-resident, bypass L1) into the control block and reset them
__device__ __forceinline__ void sweep_collect(Ctrl* c, bool with_extrema)
{
    c->last_nsims = __ldcg(&c->acc.sweep_nsims); c->last_naccs = __ldcg(&c->acc.sweep_naccs);
    if (with_extrema) {
        c->dmin = key_f64(__ldcg(&c->acc.dmin_key)); c->dmax = key_f64(__ldcg(&c->acc.dmax_key));
        c->acc.dmin_key = ~0ull; c->acc.dmax_key = 0ull;
    }
    int e = __ldcg(&c->acc.err);
    if (e && !c->err) c->err = e;
    c->acc.sweep_nsims = 0ull; c->acc.sweep_naccs = 0ull;
}

// sharded runs: the warp that finished the grid replaces this rank's sweep counters by the sums over all ranks
// (one low-latency in-kernel exchange, comm.cuh); integers, so every rank holds the same totals and takes the
// same early-exit / stop decisions.  All 32 lanes call; lane 0 writes the control block.
static __device__ __noinline__ void sweep_exchange_warp(const PopDev& P, Ctrl* c, bool with_extrema)
{
    const int lane = threadIdx.x & 31;
    unsigned long long rec[5] = { 0ull, 0ull, 0ull, 0ull, 0ull }, got[5];
    if (lane == 0) {
        sweep_collect(c, with_extrema);
        rec[0] = c->last_nsims; rec[1] = c->last_naccs; rec[2] = (unsigned long long)c->err;
        rec[3] = f64_key(c->dmin); rec[4] = f64_key(c->dmax);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) rec[k] = __shfl_sync(0xffffffffu, rec[k], 0);
    const bool valid = xchg_ll_warp<5>(P.x, c, rec, got);
    unsigned long long ns = valid ? got[0] : 0ull, na = valid ? got[1] : 0ull, e = valid ? got[2] : 0ull;
    unsigned long long mn = valid ? got[3] : ~0ull, mx = valid ? got[4] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ns += __shfl_xor_sync(0xffffffffu, ns, o); na += __shfl_xor_sync(0xffffffffu, na, o);
        unsigned long long t = __shfl_xor_sync(0xffffffffu, e, o); e = t > e ? t : e;
    }
    mn = warp_min_u64(mn); mx = warp_max_u64(mx);
    if (lane == 0) {
        c->last_nsims = ns; c->last_naccs = na;
        if (e && !c->err) c->err = (int)e;
        if (with_extrema) { c->dmin = key_f64(mn); c->dmax = key_f64(mx); }
    }
    __syncwarp();
}

__device__ inline void ctrl_after_smc_sweep(const PopDev& P, Ctrl* c)
{
    if (P.x.world == 1) sweep_collect(c, false);   // (sharded: sweep_exchange_warp has collected and reduced)
    c->nsims_total += (long long)c->last_nsims;
    c->naccs_iter += c->last_naccs;
    c->cur ^= 1;                                   // swap buffers, :347-350
    c->sweep_epoch += 1;
    c->sweep_idx += 1;
    c->n_sweeps += 1;
    if ((double)c->naccs_iter / (double)c->n_alive_g >= c->Kmcmc_min) {   // :352
        c->Ki = c->sweep_idx; c->sweeps_done = 1;
    } else if (c->sweep_idx >= c->Kmcmc) {
        c->sweeps_done = 1;
    }
    if (c->sweeps_done) ctrl_end_iter(P, c);       // :357-376
}

__device__ inline void ctrl_after_mc_sweep(const PopDev& P, Ctrl* c)
{
    if (P.x.world == 1) sweep_collect(c, true);    // (sharded: global extrema(delta), src/abcdez_mc.jl:146, already reduced)
    c->nsims_total += (long long)c->last_nsims;
    c->cur ^= 1;
    c->sweep_epoch += 1;
    c->n_sweeps += 1;
}

// Barrier-free epilogue: warps retire independently (no __syncthreads at the end, so a warp that has
// finished does not hold its registers hostage until the slowest warp of the CTA arrives).  Warp totals
// go to shared-memory atomics; the last warp of the CTA to arrive forwards the CTA totals with one set
// of global REDs (integer only: order independent) and takes the grid ticket.  Returns true in exactly
// one thread of the grid: the one that must run the control logic.
struct SweepSmem {
    unsigned long long mn, mx;
    unsigned sim, acc, arrived;
    int err;
};

__device__ __forceinline__ void sweep_smem_init(SweepSmem* s)
{
    if (threadIdx.x == 0) { s->mn = ~0ull; s->mx = 0ull; s->sim = 0u; s->acc = 0u; s->arrived = 0u; s->err = 0; }
    __syncthreads();
}

template <bool EXTREMA>
__device__ __forceinline__ bool sweep_finish(Ctrl* c, SweepSmem* s, unsigned nsim, unsigned nacc,
                                             unsigned long long kmin, unsigned long long kmax, int err)
{
    const int lane = threadIdx.x & 31;
    const unsigned nw = blockDim.x >> 5;
    nsim = warp_sum_u(nsim); nacc = warp_sum_u(nacc);
    if (EXTREMA) { kmin = warp_min_u64(kmin); kmax = warp_max_u64(kmax); }
    err = __reduce_max_sync(0xffffffffu, err);
    bool last = false;
    if (lane == 0) {
        if (nsim) atomicAdd(&s->sim, nsim);
        if (nacc) atomicAdd(&s->acc, nacc);
        if (EXTREMA) { atomicMin(&s->mn, kmin); atomicMax(&s->mx, kmax); }
        if (err) atomicMax(&s->err, err);
        __threadfence_block();
        if (atomicAdd(&s->arrived, 1u) == nw - 1) {
            __threadfence_block();
            volatile SweepSmem* v = s;
            unsigned a = v->sim, b = v->acc; int e = v->err;
            if (a) atomicAdd(&c->acc.sweep_nsims, (unsigned long long)a);
            if (b) atomicAdd(&c->acc.sweep_naccs, (unsigned long long)b);
            if (EXTREMA) { atomicMin(&c->acc.dmin_key, v->mn); atomicMax(&c->acc.dmax_key, v->mx); }
            if (e) atomicMax(&c->acc.err, e);
            __threadfence();
            if (atomicAdd(&c->acc.ticket[0], 1u) == gridDim.x - 1) {
                c->acc.ticket[0] = 0u;
                __threadfence();
                last = true;
            }
        }
    }
    return last;
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

// copy the scalar state of particle i from generation `cur` to `nxt`
template <int NB>
__device__ __forceinline__ void copy_scalars(const PopDev& P, int cur, int nxt, uint32_t i, double lpi, double dli)
{
    P.logpi[nxt][i] = lpi;
    P.delta[nxt][i] = dli;
#pragma unroll
    for (int k = 0; k < NB; ++k) P.blob[nxt][(size_t)i * NB + k] = P.blob[cur][(size_t)i * NB + k];
}

// ---------------------------------------------------------------------------------------
// abcde_init!  src/abcdez_init.jl:2-22
// ---------------------------------------------------------------------------------------
template <class M, bool F32 = false>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
init_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
            const __grid_constant__ PhiloxKeys seed, int draw_prior)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    const int cur = c->cur;
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned redraws = 0; int err = 0;
    unsigned long long kdl = 0ull;
    if (i < P.N) {
        const uint32_t pid = P.id0 + i;
        double th[D], x[D], blob[NB > 0 ? NB : 1];
        double lp;
        if (draw_prior) {                                  // src/abcdez_smc.jl:242-243
            prior_sample<D>(pr, seed, pid, 0u, th);
            round_row_f32<D>(th, F32);
            push_p<D>(pr, th, x);
            lp = prior_logpdf<D>(pr, x);
        } else {
            load_row<D>(P.theta[cur], i, th, F32);
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
            round_row_f32<D>(th, F32);
            push_p<D>(pr, th, x);
            lp = prior_logpdf<D>(pr, x);
            SimRng r(seed, pid, attempt, TAG_INIT_MODEL);
            dl = M::run(x, md.v, r, blob);
            redraws++;
        }
        store_row<D>(P.theta[cur], i, th, F32);
        P.logpi[cur][i] = lp;
        P.delta[cur][i] = dl;
#pragma unroll
        for (int k = 0; k < NB; ++k) P.blob[cur][(size_t)i * NB + k] = blob[k];
        P.moved[i] = 1;                                    // the other buffer is stale
        kdl = f64_key(dl);
    }
    // counters: reuse the sweep accumulators (nsims slot carries the redraw count)
    const bool last = sweep_finish<true>(c, &s_red, redraws, 0u, i < P.N ? kdl : ~0ull, i < P.N ? kdl : 0ull, err);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, true);   // global extrema, redraws, errors
    if (last) {
        if (P.x.world == 1) sweep_collect(c, true);
        c->redraws += (long long)c->last_nsims;
    }
}

// ---------------------------------------------------------------------------------------
// abcdesmc_swarm!  src/abcdez_smc.jl:106-153
// DISC: the prior has discrete marginals (push_p rounds a copy of the proposal); the common
// all-continuous case passes the proposal registers straight to the simulator.
// ---------------------------------------------------------------------------------------
// INJ: the launch may carry injected partners / jitter / accept uniforms / a flags buffer (stage-level parity calls);
// production launches (abcdez_smc_run) use INJ = false, which removes those tests -- and with them the basic-block
// boundaries that keep the partner draw, the jitter's Box-Muller pair and the loads from being scheduled together
// (101.8 -> 95.1 us per sweep).  Straightening further -- simulator and accept draw unconditional for all-Normal
// priors, the stale-row repair moved behind the accept test -- costs registers and was slower (113.6 us; the repair
// move alone 99.3 us, the unconditional accept draw alone neutral).
// SEG (relaxed-parity mode; opts.partner_segments): warp-coherent partner segments.  One pair
// of random bases per warp, lane l takes the l-th alive particle behind each base, so the two partner gathers of a
// warp read 32 neighbouring rows instead of 64 random ones (the gathers are what keeps the parity kernel off its
// roofline, profiles/README.md).  Every particle's partners are still uniform over the alive set and distinct from
// it and from each other (a lane whose segment entries collide redraws individually); only their joint law across
// the lanes of a warp differs from the reference's independent draws -- the Jacobi update does not care.
template <class M, bool DISC, int PK, bool INJ = true, bool SEG = false, bool F32 = false>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
smc_sweep_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
                 const __grid_constant__ SweepInj inj)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t N = P.N;
    // this thread's entry of the particle list is requested together with the control block's scalars (it is only
    // meaningful while some particles are dead, but always readable): one memory round trip less before the work starts
    const uint32_t listed = j < N ? P.alive_list[j] : 0u;
    if (c->stop | c->sweeps_done) return;                  // skipped sweep (early exit :352 / stop :376)
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t n_alive = c->n_alive;
    const double* __restrict__ th = P.theta[cur];
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);

    unsigned nsim = 0, nacc = 0; int err = 0;
    // every schedule scalar the particle work needs, loaded back to back (one L2 round trip, not six)
    const PhiloxKeys& seed = P.keys;
    const uint32_t epoch = c->sweep_epoch;
    const double gamma0 = c->gamma0, gsig = c->gsig, eps = c->eps;
    const int kind = c->kind;
    if (j < N) {
        const uint32_t i = (n_alive == N) ? j : listed;
        const uint8_t mv = P.moved[i];
        if (j >= n_alive) {                                                // dead particle (:114)
            if (mv) {                                                      // repair its stale row, once
                double row[D];
                load_row<D>(th, i, row, F32);
                store_row<D>(P.theta[nxt], i, row, F32);
                copy_scalars<NB>(P, cur, nxt, i, P.logpi[cur][i], P.delta[cur][i]);
                P.moved[i] = 0;
            }
            if (INJ && inj.flags) inj.flags[i] = 0;
        } else {
            uint8_t flag = 0;
            const uint32_t pid = P.id0 + i;
            // (1) partners: integer work only (plus list lookups while some particles are dead)
            uint32_t a, b;
            if (INJ && inj.a) { a = (uint32_t)inj.a[i]; b = (uint32_t)inj.b[i]; }
            else {
                // attempt k of either loop reads block k of the partner stream (u1 -> a, u2 -> b), so the first attempts
                // of both share ONE Philox block; the rejection loops continue from block 1
                Stream ps(seed, pid, epoch, TAG_PARTNER);
                double u1, u2, ua, ub; uint32_t att = 1;
                if (SEG) {
                    Stream ws(seed, P.id0 + (j & ~31u), epoch, TAG_SEGMENT);   // the warp's stream: every lane computes the same bases
                    ws.u2(0u, ua, ub);
                    const uint32_t lane = threadIdx.x & 31;
                    uint32_t ka = (uint32_t)(ua * (double)n_alive), kb = (uint32_t)(ub * (double)n_alive);
                    ka = ((ka >= n_alive ? n_alive - 1 : ka) + lane) % n_alive;       // positions in the alive list, wrapping
                    kb = ((kb >= n_alive ? n_alive - 1 : kb) + lane) % n_alive;
                    a = (n_alive == N) ? ka : P.alive_list[ka];
                    b = (n_alive == N) ? kb : P.alive_list[kb];
                    att = 0;                                                // a collision falls back to the particle's own draws, from block 0
                } else {
                    ps.u2(0u, ua, ub);
                    a = wsample_alive(P.alive_list, n_alive, N, ua);
                    b = wsample_alive(P.alive_list, n_alive, N, ub);
                }
                if (a == i || b == a || b == i) {                          // ~3 / n_alive: one rare branch around both loops
                    while (a == i) {                                       // :119-122
                        if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                        ps.u2(att++, u1, u2);
                        a = wsample_alive(P.alive_list, n_alive, N, u1);
                    }
                    att = SEG ? 0 : 1;
                    while (b == a || b == i) {                             // :123-126
                        if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                        ps.u2(att++, u1, u2);
                        b = wsample_alive(P.alive_list, n_alive, N, u2);
                    }
                }
            }
            SimRng r(seed, pid, epoch, TAG_MODEL);
            // (2) the gamma jitter (:128)
            Stream ms(seed, pid, epoch, TAG_MOVE);
            double z, z2;
            if (INJ && inj.z) z = inj.z[i]; else ms.n2(0u, z, z2);
            const double g = gamma0 * (1.0 + z * gsig);                    // :128
            // (3) own state; repair the stale row in g+1
            double thp[D];
            load_row<D>(th, i, thp, F32);
            const double lpi = P.logpi[cur][i], dli = P.delta[cur][i];
            if (mv) {
                store_row<D>(P.theta[nxt], i, thp, F32);
                copy_scalars<NB>(P, cur, nxt, i, lpi, dli);
            }
            if (!err) {
                de_proposal<D>(th, a, b, g, thp, F32);                     // :128 (F32: rounded to the float row that is stored)
                double xs[DISC ? D : 1];
                const double* x = thp;
                if (DISC) { push_p<D>(pr, thp, xs); x = xs; }
                double lp = prior_logpdf_k<D, PK>(pr, x);                  // :134
                if (!(lp < 0.0 && isinf(lp))) {                            // :135
                    double blp[NB > 0 ? NB : 1];
                    const double dp = M::run(x, md.v, r, blp);             // :137
                    nsim = 1; flag |= ABCDEZ_FLAG_SIM;                     // :138
                    double w = lp - lpi;                                   // :140-141, left to right
                    w = w + abck_logpdf(kind, eps, dp);
                    w = w - abck_logpdf(kind, eps, dli);
                    bool acc = (0.0 <= w);
                    if (!acc) {                                            // :145, uniform only when w < 0
                        double u, u2;
                        if (INJ && inj.u) { u = inj.u[i]; acc = (plog(u) < w); }
                        else { ms.u2(1u, u, u2); acc = ((u == 0.0 ? -INFINITY : plog_unit(u)) < w); }   // u in [0, 1)
                    }
                    if (acc) {                                             // :146-150
                        store_row<D>(P.theta[nxt], i, thp, F32);
                        P.logpi[nxt][i] = lp;
                        P.delta[nxt][i] = dp;
#pragma unroll
                        for (int k = 0; k < NB; ++k) P.blob[nxt][(size_t)i * NB + k] = blp[k];
                        nacc = 1; flag |= ABCDEZ_FLAG_ACC;
                    }
                }
            }
            if ((uint8_t)nacc != mv) P.moved[i] = (uint8_t)nacc;
            if (INJ && inj.flags) inj.flags[i] = flag;
        }
    }
    const bool last = sweep_finish<false>(c, &s_red, nsim, nacc, 0ull, 0ull, err);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, false);
    if (last) ctrl_after_smc_sweep(P, c);
}

// ---------------------------------------------------------------------------------------
// abcdemc_swarm!  src/abcdez_mc.jl:5-61
// ---------------------------------------------------------------------------------------
template <class M, bool DISC>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
mc_sweep_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
                const __grid_constant__ SweepInj inj, const __grid_constant__ McArgs mc)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    if (mc.from_ctrl && (c->stop | c->err)) return;                        // a generation behind an error
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t N = P.N;
    const double* __restrict__ th = P.theta[cur];
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);
    // abcdez_mc_run: the schedule scalars live on the device (no host round trip per generation)
    const double eps_target = mc.from_ctrl ? c->eps_target : mc.eps_target;
    const double eps_pop = mc.from_ctrl ? fmax(eps_target, c->dmin + 0.0 * (c->dmax - c->dmin)) : mc.eps_pop;   // src/abcdez_mc.jl:146-147, alpha = 0 (:107)

    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nsim = 0, nacc = 0; int err = 0;
    unsigned long long kdl = 0ull;
    if (i < N) {
        double thp[D];
        load_row<D>(th, i, thp);
        const double lpi = P.logpi[cur][i];
        double dli = P.delta[cur][i];
        const uint8_t mv = P.moved[i];
        if (mv) {                                                          // repair the stale row in g+1
            store_row<D>(P.theta[nxt], i, thp);
            copy_scalars<NB>(P, cur, nxt, i, lpi, dli);
        }
        uint8_t flag = 0;
        const uint32_t pid = P.id0 + i;
        const PhiloxKeys& seed = P.keys;
        const uint32_t epoch = c->sweep_epoch;
        uint32_t s = i;                                                    // :18
        const double eps = (dli <= eps_target) ? eps_target : eps_pop;   // :19
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
            Stream ps(seed, pid, epoch, TAG_PARTNER);                      // (first attempts share block 0, see smc_sweep_kernel)
            double u1, u2, ua, ub; uint32_t att = 1;
            auto pick = [N](double u) { long long k = (long long)floor(u * (double)N); if (k >= (long long)N) k = (long long)N - 1; return (uint32_t)k; };
            ps.u2(0u, ua, ub);
            a = pick(ua);
            while (a == s) {                                               // :25-28
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                a = pick(u1);
            }
            att = 1;
            b = pick(ub);
            while (b == a || b == s) {                                     // :29-32
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                b = pick(u2);
            }
        }
        if (!err) {
            Stream ms(seed, pid, epoch, TAG_MOVE);
            if (s != i) load_row<D>(th, s, thp);                           // base particle theta_s
            {
                double ta[D], tb[D];
                load_row<D>(th, a, ta);
                load_row<D>(th, b, tb);
                double z, z2;
                if (inj.z) z = inj.z[i]; else ms.n2(0u, z, z2);
                const double g = c->gamma0 * (1.0 + z * c->gsig);          // :34
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    double diff = ta[k] - tb[k];
                    double sc = diff * g;
                    thp[k] = thp[k] + sc;
                }
            }
            double xs[DISC ? D : 1];
            const double* x = thp;
            if (DISC) { push_p<D>(pr, thp, xs); x = xs; }
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
                    store_row<D>(P.theta[nxt], i, thp);
                    P.logpi[nxt][i] = lp;
                    P.delta[nxt][i] = dp;
#pragma unroll
                    for (int k = 0; k < NB; ++k) P.blob[nxt][(size_t)i * NB + k] = blp[k];
                    dli = dp;
                    nacc = 1; flag |= ABCDEZ_FLAG_ACC;
                }
            }
        }
        if ((uint8_t)nacc != mv) P.moved[i] = (uint8_t)nacc;
        if (inj.flags) inj.flags[i] = flag;
        kdl = f64_key(dli);                                                // extrema(delta), src/abcdez_mc.jl:146
    }
    const bool last = sweep_finish<true>(c, &s_red, nsim, nacc, i < N ? kdl : ~0ull, i < N ? kdl : 0ull, err);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, true);
    if (last) ctrl_after_mc_sweep(P, c);
}

// one dist! evaluation per row (stage-level model parity); dense N x D input
template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS)
simulate_kernel(const __grid_constant__ ModelData md, int64_t N, const double* __restrict__ theta_pushed, const __grid_constant__ PhiloxKeys seed,
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

#ifndef __CUDACC_RTC__
// ---------------------------------------------------------------------------------------
// abcdesmc_swarm! for HEAVY simulators (M::SPLIT): propose -> queue-driven simulate -> accept
// ---------------------------------------------------------------------------------------
// In the fused kernel a lane whose proposal leaves the prior's support (src/abcdez_smc.jl:135: `continue`) idles while the
// rest of its warp simulates -- with box priors that is 50-75 % of the lanes early in a run -- and a stochastic simulator
// keeps every warp busy until its LONGEST trajectory ends (birth-death SSA: 2.1 of 32 lanes active on average, ncu).
// Here the sweep is three launches over the same arithmetic and the same Philox streams (bit-identical results):
//   propose   partners, jitter, DE proposal, prior; proposals inside the support go into a queue (theta', log prior kept)
//   simulate  a persistent grid works the queue off: equal-length simulators stride over it with dense warps; STEPPED
//             simulators (begin / step / finish) run ONE merged loop per lane -- "fetch the next pending simulation" or
//             "advance mine by one step" -- so a lane is refilled the moment its trajectory ends
//   accept    MH accept against the current eps kernel, counters, and the sweep's control logic in the last CTA
template <class M, class = void> struct model_is_split { static constexpr bool value = false; };
template <class M> struct model_is_split<M, decltype((void)M::SPLIT)> { static constexpr bool value = M::SPLIT; };
template <class M, class = void> struct model_is_stepped { static constexpr bool value = false; };
template <class M> struct model_is_stepped<M, decltype((void)M::STEPPED)> { static constexpr bool value = M::STEPPED; };
// optional scheduling hint M::heavy(theta, data): such simulations go to the FRONT of the queue (longest first)
template <class M, class = void> struct model_has_heavy { static constexpr bool value = false; };
template <class M> struct model_has_heavy<M, decltype((void)&M::heavy)> { static constexpr bool value = true; };
// queue layout: heavy entries fill [0, heavy) from the front, the others [N - light, N) from the back; entry q of the
// hand-out order is queue[q] for q < heavy and queue[N - 1 - (q - heavy)] behind them
__device__ __forceinline__ uint32_t queue_entry(const PopDev& P, unsigned q, unsigned heavy)
{
    return q < heavy ? P.queue[q] : P.queue[P.N - 1u - (q - heavy)];
}

template <class M, bool DISC, int PK, bool SEG>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
smc_propose_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t N = P.N;
    const uint32_t listed = j < N ? P.alive_list[j] : 0u;
    if (c->stop | c->sweeps_done) return;
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t n_alive = c->n_alive;
    const double* __restrict__ th = P.theta[cur];
    const PhiloxKeys& seed = P.keys;
    const uint32_t epoch = c->sweep_epoch;
    const double gamma0 = c->gamma0, gsig = c->gsig;
    bool queued = false, hvy = false; uint32_t i = 0;
    if (j < N) {
        i = (n_alive == N) ? j : listed;
        const uint8_t mv = P.moved[i];
        if (j >= n_alive) {                                                // dead particle (:114): repair its stale row, once
            if (mv) {
                double row[D];
                load_row<D>(th, i, row);
                store_row<D>(P.theta[nxt], i, row);
                copy_scalars<NB>(P, cur, nxt, i, P.logpi[cur][i], P.delta[cur][i]);
                P.moved[i] = 0;
            }
        } else {
            const uint32_t pid = P.id0 + i;
            int err = 0;
            uint32_t a, b;
            Stream ps(seed, pid, epoch, TAG_PARTNER);
            double u1, u2, ua, ub; uint32_t att = 1;
            if (SEG) {
                Stream ws(seed, P.id0 + (j & ~31u), epoch, TAG_SEGMENT);
                ws.u2(0u, ua, ub);
                const uint32_t lane = threadIdx.x & 31;
                uint32_t ka = (uint32_t)(ua * (double)n_alive), kb = (uint32_t)(ub * (double)n_alive);
                ka = ((ka >= n_alive ? n_alive - 1 : ka) + lane) % n_alive;
                kb = ((kb >= n_alive ? n_alive - 1 : kb) + lane) % n_alive;
                a = (n_alive == N) ? ka : P.alive_list[ka];
                b = (n_alive == N) ? kb : P.alive_list[kb];
                att = 0;
            } else {
                ps.u2(0u, ua, ub);
                a = wsample_alive(P.alive_list, n_alive, N, ua);
                b = wsample_alive(P.alive_list, n_alive, N, ub);
            }
            if (a == i || b == a || b == i) {
                while (a == i) {                                           // :119-122
                    if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                    ps.u2(att++, u1, u2);
                    a = wsample_alive(P.alive_list, n_alive, N, u1);
                }
                att = SEG ? 0 : 1;
                while (b == a || b == i) {                                 // :123-126
                    if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                    ps.u2(att++, u1, u2);
                    b = wsample_alive(P.alive_list, n_alive, N, u2);
                }
            }
            Stream ms(seed, pid, epoch, TAG_MOVE);
            double z, z2;
            ms.n2(0u, z, z2);
            const double g = gamma0 * (1.0 + z * gsig);                    // :128
            double thp[D];
            load_row<D>(th, i, thp);
            if (mv) {                                                      // repair the stale row in g+1
                store_row<D>(P.theta[nxt], i, thp);
                copy_scalars<NB>(P, cur, nxt, i, P.logpi[cur][i], P.delta[cur][i]);
            }
            uint8_t flag = 0;
            if (err) atomicMax(&c->acc.err, err);
            else {
                de_proposal<D>(th, a, b, g, thp);                          // :128
                double xs[DISC ? D : 1];
                const double* x = thp;
                if (DISC) { push_p<D>(pr, thp, xs); x = xs; }
                const double lp = prior_logpdf_k<D, PK>(pr, x);            // :134
                if (!(lp < 0.0 && isinf(lp))) {                            // :135
                    store_row<D>(P.prop_theta, i, thp);
                    P.prop_lp[i] = lp;
                    flag = 1; queued = true;
                    if constexpr (model_has_heavy<M>::value) hvy = M::heavy(x, md.v);
                }
            }
            P.prop_flag[i] = flag;
        }
    }
    // queue slots: one atomic per warp and class (heavy entries from the front, the others from the back)
    const unsigned lane = threadIdx.x & 31;
    const unsigned mh = __ballot_sync(0xffffffffu, queued && hvy), ml = __ballot_sync(0xffffffffu, queued && !hvy);
    if (mh) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(mh) - 1)) base = atomicAdd(&c->acc.queue_heavy, (unsigned)__popc(mh));
        base = __shfl_sync(0xffffffffu, base, __ffs(mh) - 1);
        if (queued && hvy) P.queue[base + __popc(mh & ((1u << lane) - 1u))] = i;
    }
    if (ml) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(ml) - 1)) base = atomicAdd(&c->acc.queue_len, (unsigned)__popc(ml));
        base = __shfl_sync(0xffffffffu, base, __ffs(ml) - 1);
        if (queued && !hvy) P.queue[N - 1u - (base + __popc(ml & ((1u << lane) - 1u)))] = i;
    }
}

enum { QUEUE_MC = 0, QUEUE_SMC = 1, QUEUE_INIT = 2 };

template <class M, bool DISC>
__global__ void __launch_bounds__(SWEEP_THREADS)
simulate_queue_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md, int mode,
                      const __grid_constant__ PhiloxKeys init_keys)
{
    // mode: QUEUE_SMC abcdesmc_swarm!, QUEUE_MC abcdemc_swarm! (a propose kernel that returned early left the queue empty),
    //       QUEUE_INIT abcde_init! (streams of the init kernel: run seed, epoch 0 = first attempt, TAG_INIT_MODEL)
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    if (mode == QUEUE_SMC && (c->stop | c->sweeps_done)) return;
    const unsigned heavy = __ldcg(&c->acc.queue_heavy), len = heavy + __ldcg(&c->acc.queue_len);
    const PhiloxKeys& seed = mode == QUEUE_INIT ? init_keys : P.keys;
    const uint32_t epoch = mode == QUEUE_INIT ? 0u : c->sweep_epoch;
    const uint32_t TAG_SIM = mode == QUEUE_INIT ? (uint32_t)TAG_INIT_MODEL : (uint32_t)TAG_MODEL;
    if constexpr (!model_is_stepped<M>::value) {
        // equal-length simulations: a static stride over the queue keeps every warp dense
        for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < len; q += gridDim.x * blockDim.x) {
            const uint32_t i = queue_entry(P, q, heavy);
            double thp[D], xs[DISC ? D : 1], blp[NB > 0 ? NB : 1];
            load_row<D>(P.prop_theta, i, thp);
            const double* x = thp;
            if (DISC) { push_p<D>(pr, thp, xs); x = xs; }
            SimRng r(seed, P.id0 + i, epoch, TAG_SIM);
            P.prop_dp[i] = M::run(x, md.v, r, blp);                        // :137
#pragma unroll
            for (int k = 0; k < NB; ++k) P.prop_blob[(size_t)i * NB + k] = blp[k];
        }
    } else {
        // STEPPED simulators: one merged loop per lane -- fetch the next pending simulation, or advance mine by one step
        typename M::State st;
        SimRng r(seed, 0u, epoch, TAG_SIM);
        uint32_t i = 0;
        bool have = false, drained = false;
        const unsigned lane = threadIdx.x & 31;
        for (;;) {
            const unsigned need = __ballot_sync(0xffffffffu, !have && !drained);
            if (need) {                                                    // one atomic per warp and refill round
                unsigned base = 0;
                if (lane == (unsigned)(__ffs(need) - 1)) base = atomicAdd(&c->acc.queue_next, (unsigned)__popc(need));
                base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
                if (!have && !drained) {
                    const unsigned q = base + __popc(need & ((1u << lane) - 1u));
                    if (q < len) {
                        i = queue_entry(P, q, heavy);
                        double thp[D], xs[DISC ? D : 1];
                        load_row<D>(P.prop_theta, i, thp);
                        const double* x = thp;
                        if (DISC) { push_p<D>(pr, thp, xs); x = xs; }
                        r = SimRng(seed, P.id0 + i, epoch, TAG_SIM);
                        M::begin(st, x, md.v);
                        have = true;
                    } else drained = true;
                }
            }
            if (have) {
#pragma unroll 1
                for (int burst = 0; burst < 8 && have; ++burst) {          // a few steps between refill votes
                    if (!M::step(st, md.v, r)) {
                        double blp[NB > 0 ? NB : 1];
                        P.prop_dp[i] = M::finish(st, blp);
#pragma unroll
                        for (int k = 0; k < NB; ++k) P.prop_blob[(size_t)i * NB + k] = blp[k];
                        have = false;
                    }
                }
            }
            if (__all_sync(0xffffffffu, drained && !have)) break;
        }
    }
}

template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
smc_accept_kernel(const __grid_constant__ PopDev P)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t N = P.N;
    const uint32_t listed = j < N ? P.alive_list[j] : 0u;
    if (c->stop | c->sweeps_done) return;
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t n_alive = c->n_alive;
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);
    unsigned nsim = 0, nacc = 0;
    const double eps = c->eps; const int kind = c->kind;
    if (j < n_alive && j < N) {
        const uint32_t i = (n_alive == N) ? j : listed;
        const uint8_t mv = P.moved[i];
        if (P.prop_flag[i]) {
            const double lp = P.prop_lp[i], dp = P.prop_dp[i];
            const double lpi = P.logpi[cur][i], dli = P.delta[cur][i];
            nsim = 1;                                                      // :138
            double w = lp - lpi;                                           // :140-141, left to right
            w = w + abck_logpdf(kind, eps, dp);
            w = w - abck_logpdf(kind, eps, dli);
            bool acc = (0.0 <= w);
            if (!acc) {                                                    // :145, uniform only when w < 0
                Stream ms(P.keys, P.id0 + i, c->sweep_epoch, TAG_MOVE);
                double u, u2;
                ms.u2(1u, u, u2);
                acc = ((u == 0.0 ? -INFINITY : plog_unit(u)) < w);
            }
            if (acc) {                                                     // :146-150
                double thp[D];
                load_row<D>(P.prop_theta, i, thp);
                store_row<D>(P.theta[nxt], i, thp);
                P.logpi[nxt][i] = lp;
                P.delta[nxt][i] = dp;
#pragma unroll
                for (int k = 0; k < NB; ++k) P.blob[nxt][(size_t)i * NB + k] = P.prop_blob[(size_t)i * NB + k];
                nacc = 1;
            }
        }
        if ((uint8_t)nacc != mv) P.moved[i] = (uint8_t)nacc;
    }
    const bool last = sweep_finish<false>(c, &s_red, nsim, nacc, 0ull, 0ull, 0);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, false);
    if (last) {
        c->acc.queue_len = 0u; c->acc.queue_next = 0u; c->acc.queue_heavy = 0u;   // the next sweep's queue
        ctrl_after_smc_sweep(P, c);
    }
}

// ---------------------------------------------------------------------------------------
// abcde_init! for STEPPED simulators (src/abcdez_init.jl:2-22): the first simulation of every particle through the queue
// ---------------------------------------------------------------------------------------
// (equal-length simulators keep the fused init_kernel: their warps are dense anyway.)  draw -> queue-driven simulate ->
// finish; the redraw loop of init.jl:14-20 (non-finite distance or log prior) stays in the finish kernel, fused, for the few
// particles that need it.  Same streams as init_kernel: bit-identical populations.
template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
init_draw_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
                 const __grid_constant__ PhiloxKeys seed, int draw_prior)
{
    constexpr int D = M::D;
    Ctrl* c = P.ctrl;
    const int cur = c->cur;
    const uint32_t N = P.N;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool queued = false, hvy = false;
    if (i < N) {
        double th[D], x[D];
        double lp;
        if (draw_prior) {                                  // src/abcdez_smc.jl:242-243
            prior_sample<D>(pr, seed, P.id0 + i, 0u, th);
            push_p<D>(pr, th, x);
            lp = prior_logpdf<D>(pr, x);
            store_row<D>(P.theta[cur], i, th);
            P.logpi[cur][i] = lp;
        } else {
            load_row<D>(P.theta[cur], i, th);
            push_p<D>(pr, th, x);
            lp = P.logpi[cur][i];
        }
        if (isfinite(lp)) {                                // init.jl:9-13
            store_row<D>(P.prop_theta, i, th);
            queued = true;
            if constexpr (model_has_heavy<M>::value) hvy = M::heavy(x, md.v);
        }
        P.prop_flag[i] = queued ? 1 : 0;
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned mh = __ballot_sync(0xffffffffu, queued && hvy), ml = __ballot_sync(0xffffffffu, queued && !hvy);
    if (mh) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(mh) - 1)) base = atomicAdd(&c->acc.queue_heavy, (unsigned)__popc(mh));
        base = __shfl_sync(0xffffffffu, base, __ffs(mh) - 1);
        if (queued && hvy) P.queue[base + __popc(mh & ((1u << lane) - 1u))] = i;
    }
    if (ml) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(ml) - 1)) base = atomicAdd(&c->acc.queue_len, (unsigned)__popc(ml));
        base = __shfl_sync(0xffffffffu, base, __ffs(ml) - 1);
        if (queued && !hvy) P.queue[N - 1u - (base + __popc(ml & ((1u << lane) - 1u)))] = i;
    }
}

template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
init_finish_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
                   const __grid_constant__ PhiloxKeys seed)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    const int cur = c->cur;
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned redraws = 0; int err = 0;
    unsigned long long kdl = 0ull;
    if (i < P.N) {
        const uint32_t pid = P.id0 + i;
        double blob[NB > 0 ? NB : 1];
        double lp = P.logpi[cur][i];
        double dl = NAN;
        if (P.prop_flag[i]) {
            dl = P.prop_dp[i];
#pragma unroll
            for (int k = 0; k < NB; ++k) blob[k] = P.prop_blob[(size_t)i * NB + k];
        }
        uint32_t attempt = 0;
        while (!isfinite(dl) || !isfinite(lp)) {           // init.jl:14-20
            if (++attempt >= (uint32_t)INIT_MAX_ATTEMPTS) { err = ABCDEZ_ERR_INIT_RETRY; break; }
            double th[D], x[D];
            prior_sample<D>(pr, seed, pid, attempt, th);
            push_p<D>(pr, th, x);
            lp = prior_logpdf<D>(pr, x);
            SimRng r(seed, pid, attempt, TAG_INIT_MODEL);
            dl = M::run(x, md.v, r, blob);
            redraws++;
            store_row<D>(P.theta[cur], i, th);
            P.logpi[cur][i] = lp;
        }
        P.delta[cur][i] = dl;
#pragma unroll
        for (int k = 0; k < NB; ++k) P.blob[cur][(size_t)i * NB + k] = blob[k];
        P.moved[i] = 1;                                    // the other buffer is stale
        kdl = f64_key(dl);
    }
    const bool last = sweep_finish<true>(c, &s_red, redraws, 0u, i < P.N ? kdl : ~0ull, i < P.N ? kdl : 0ull, err);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, true);   // global extrema, redraws, errors
    if (last) {
        c->acc.queue_len = 0u; c->acc.queue_next = 0u; c->acc.queue_heavy = 0u;
        if (P.x.world == 1) sweep_collect(c, true);
        c->redraws += (long long)c->last_nsims;
    }
}

// ---------------------------------------------------------------------------------------
// abcdemc_swarm! for heavy simulators, the same three launches (src/abcdez_mc.jl:5-61)
// ---------------------------------------------------------------------------------------
template <class M, bool DISC>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
mc_propose_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
                  const __grid_constant__ McArgs mc)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    if (mc.from_ctrl && (c->stop | c->err)) return;
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t N = P.N;
    const double* __restrict__ th = P.theta[cur];
    const double eps_target = mc.from_ctrl ? c->eps_target : mc.eps_target;
    const double eps_pop = mc.from_ctrl ? fmax(eps_target, c->dmin + 0.0 * (c->dmax - c->dmin)) : mc.eps_pop;   // :146-147, alpha = 0 (:107)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool queued = false, hvy = false;
    if (i < N) {
        double thp[D];
        load_row<D>(th, i, thp);
        const double lpi = P.logpi[cur][i], dli = P.delta[cur][i];
        if (P.moved[i]) {                                                  // repair the stale row in g+1
            store_row<D>(P.theta[nxt], i, thp);
            copy_scalars<NB>(P, cur, nxt, i, lpi, dli);
        }
        const uint32_t pid = P.id0 + i;
        const PhiloxKeys& seed = P.keys;
        const uint32_t epoch = c->sweep_epoch;
        int err = 0;
        uint32_t s = i;                                                    // :18
        const double eps = (dli <= eps_target) ? eps_target : eps_pop;     // :19
        if (dli > eps) {                                                   // :20-24
            uint32_t lo = 0, hi = N;
            while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (mc.sorted_delta[m] <= dli) lo = m + 1; else hi = m; }
            Stream cs(seed, pid, epoch, TAG_MC);
            double u1, u2; cs.u2(0u, u1, u2);
            long long k = (long long)floor(u1 * (double)lo);
            if (k >= (long long)lo) k = (long long)lo - 1;
            s = mc.order[k];
        }
        uint32_t a, b;
        {
            Stream ps(seed, pid, epoch, TAG_PARTNER);
            double u1, u2, ua, ub; uint32_t att = 1;
            auto pick = [N](double u) { long long k = (long long)floor(u * (double)N); if (k >= (long long)N) k = (long long)N - 1; return (uint32_t)k; };
            ps.u2(0u, ua, ub);
            a = pick(ua);
            while (a == s) {                                               // :25-28
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                a = pick(u1);
            }
            att = 1;
            b = pick(ub);
            while (b == a || b == s) {                                     // :29-32
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { err = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                b = pick(u2);
            }
        }
        uint8_t flag = 0;
        if (err) atomicMax(&c->acc.err, err);
        else {
            Stream ms(seed, pid, epoch, TAG_MOVE);
            if (s != i) load_row<D>(th, s, thp);                           // base particle theta_s
            double z, z2;
            ms.n2(0u, z, z2);
            const double g = c->gamma0 * (1.0 + z * c->gsig);              // :34
            de_proposal<D>(th, a, b, g, thp);
            double xs[DISC ? D : 1];
            const double* x = thp;
            if (DISC) { push_p<D>(pr, thp, xs); x = xs; }
            const double lp = prior_logpdf<D>(pr, x);                      // :41
            const double w_prior = lp - lpi;                               // :42 (logpi[i], not [s])
            double u, u2;
            ms.u2(1u, u, u2);                                              // :43, always drawn
            if (!(plog(u) > fmin(0.0, w_prior))) {                         // :44: simulate
                store_row<D>(P.prop_theta, i, thp);
                P.prop_lp[i] = lp;
                flag = 1; queued = true;
                if constexpr (model_has_heavy<M>::value) hvy = M::heavy(x, md.v);
            }
        }
        P.prop_flag[i] = flag;
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned mh = __ballot_sync(0xffffffffu, queued && hvy), ml = __ballot_sync(0xffffffffu, queued && !hvy);
    if (mh) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(mh) - 1)) base = atomicAdd(&c->acc.queue_heavy, (unsigned)__popc(mh));
        base = __shfl_sync(0xffffffffu, base, __ffs(mh) - 1);
        if (queued && hvy) P.queue[base + __popc(mh & ((1u << lane) - 1u))] = i;
    }
    if (ml) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(ml) - 1)) base = atomicAdd(&c->acc.queue_len, (unsigned)__popc(ml));
        base = __shfl_sync(0xffffffffu, base, __ffs(ml) - 1);
        if (queued && !hvy) P.queue[N - 1u - (base + __popc(ml & ((1u << lane) - 1u)))] = i;
    }
}

template <class M>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
mc_accept_kernel(const __grid_constant__ PopDev P, const __grid_constant__ McArgs mc)
{
    constexpr int D = M::D, NB = M::BLOB / 8;
    Ctrl* c = P.ctrl;
    if (mc.from_ctrl && (c->stop | c->err)) return;
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t N = P.N;
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);
    const double eps_target = mc.from_ctrl ? c->eps_target : mc.eps_target;
    const double eps_pop = mc.from_ctrl ? fmax(eps_target, c->dmin + 0.0 * (c->dmax - c->dmin)) : mc.eps_pop;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nsim = 0, nacc = 0;
    unsigned long long kdl = 0ull;
    if (i < N) {
        double dli = P.delta[cur][i];
        const uint8_t mv = P.moved[i];
        if (P.prop_flag[i]) {
            nsim = 1;                                                      // :44
            const double dp = P.prop_dp[i];
            const double eps = (dli <= eps_target) ? eps_target : eps_pop; // :19
            if (dp <= fmax(eps, dli)) {                                    // :54-59
                double thp[D];
                load_row<D>(P.prop_theta, i, thp);
                store_row<D>(P.theta[nxt], i, thp);
                P.logpi[nxt][i] = P.prop_lp[i];
                P.delta[nxt][i] = dp;
#pragma unroll
                for (int k = 0; k < NB; ++k) P.blob[nxt][(size_t)i * NB + k] = P.prop_blob[(size_t)i * NB + k];
                dli = dp;
                nacc = 1;
            }
        }
        if ((uint8_t)nacc != mv) P.moved[i] = (uint8_t)nacc;
        kdl = f64_key(dli);                                                // extrema(delta), :146
    }
    const bool last = sweep_finish<true>(c, &s_red, nsim, nacc, i < N ? kdl : ~0ull, i < N ? kdl : 0ull, 0);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, true);
    if (last) {
        c->acc.queue_len = 0u; c->acc.queue_next = 0u; c->acc.queue_heavy = 0u;
        ctrl_after_mc_sweep(P, c);
    }
}

// ---------------------------------------------------------------------------------------
// launchers + registry
// ---------------------------------------------------------------------------------------
static inline unsigned grid_for(int64_t N, int threads) { return (unsigned)((N + threads - 1) / threads); }

// persistent grid of the queue-driven simulate kernel: SMs x resident CTAs (per device), at most one thread per particle
template <class M, bool DISC>
static inline unsigned queue_grid(unsigned g)
{
    static int cached[64] = { 0 };
    int dev = 0; cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    if (!cached[slot]) {
        int sms = 148, per = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, simulate_queue_kernel<M, DISC>, SWEEP_THREADS, 0) != cudaSuccess || per < 1) per = 1;
        cached[slot] = sms * per;
    }
    return g < (unsigned)cached[slot] ? g : (unsigned)cached[slot];
}

template <int D>
static inline bool prior_has_discrete(const PriorDev& pr)
{
    bool disc = false;
    for (int k = 0; k < D; ++k) disc = disc || fam_is_discrete(pr.family[k]);
    return disc;
}

template <class M>
static void l_init(const ModelOps&, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, uint64_t seed, int dp)
{
    if constexpr (model_is_split<M>::value && model_is_stepped<M>::value) {
        if (P.prop_theta && !(P.flags & POP_FP32_STATE)) {
            const unsigned g = grid_for(P.N, SWEEP_THREADS);
            const PhiloxKeys keys = philox_keys(seed);
            init_draw_kernel<M><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, keys, dp);
            if (prior_has_discrete<M::D>(pr)) simulate_queue_kernel<M, true><<<queue_grid<M, true>(g), SWEEP_THREADS, 0, st>>>(P, pr, md, QUEUE_INIT, keys);
            else simulate_queue_kernel<M, false><<<queue_grid<M, false>(g), SWEEP_THREADS, 0, st>>>(P, pr, md, QUEUE_INIT, keys);
            init_finish_kernel<M><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, keys);
            return;
        }
    }
    if (P.flags & POP_FP32_STATE) init_kernel<M, true><<<grid_for(P.N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(P, pr, md, philox_keys(seed), dp);
    else init_kernel<M><<<grid_for(P.N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(P, pr, md, philox_keys(seed), dp);
}
template <class M, bool DISC, int PK>
static void l_smc_split(cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md)
{
    const unsigned g = grid_for(P.N, SWEEP_THREADS);
    if (P.flags & POP_PARTNER_SEGMENTS) smc_propose_kernel<M, DISC, PK, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md);
    else smc_propose_kernel<M, DISC, PK, false><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md);
    simulate_queue_kernel<M, DISC><<<queue_grid<M, DISC>(g), SWEEP_THREADS, 0, st>>>(P, pr, md, QUEUE_SMC, P.keys);
    smc_accept_kernel<M><<<g, SWEEP_THREADS, 0, st>>>(P);
}

template <class M>
static void l_smc(const ModelOps&, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj)
{
    const unsigned g = grid_for(P.N, SWEEP_THREADS);
    bool all_normal = true, all_uniform = true;
    for (int k = 0; k < M::D; ++k) { all_normal = all_normal && pr.family[k] == ABCDEZ_NORMAL; all_uniform = all_uniform && pr.family[k] == ABCDEZ_UNIFORM; }
    const bool injected = inj.a || inj.b || inj.s || inj.z || inj.u || inj.flags;
    if constexpr (model_is_split<M>::value) {
        // heavy simulators: queue-driven sweep (three launches, same results); stage calls with injected randomness keep the fused kernel
        if (!injected && P.prop_theta && !(P.flags & POP_FP32_STATE)) {
            if (prior_has_discrete<M::D>(pr)) l_smc_split<M, true, PK_GENERIC>(st, P, pr, md);
            else if (all_uniform) l_smc_split<M, false, PK_UNIFORM>(st, P, pr, md);
            else l_smc_split<M, false, PK_GENERIC>(st, P, pr, md);
            return;
        }
    }
    if (P.flags & POP_FP32_STATE) {                                        // relaxed mode: float theta rows (fused kernel, production launches)
        if (prior_has_discrete<M::D>(pr)) smc_sweep_kernel<M, true, PK_GENERIC, false, false, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
        else if (all_normal) smc_sweep_kernel<M, false, PK_NORMAL, false, false, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
        else if (all_uniform) smc_sweep_kernel<M, false, PK_UNIFORM, false, false, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
        else smc_sweep_kernel<M, false, PK_GENERIC, false, false, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
        return;
    }
    if (prior_has_discrete<M::D>(pr)) smc_sweep_kernel<M, true, PK_GENERIC><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
    else if (all_normal) {
        if (injected) smc_sweep_kernel<M, false, PK_NORMAL, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
        else if (P.flags & POP_PARTNER_SEGMENTS) smc_sweep_kernel<M, false, PK_NORMAL, false, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
        else smc_sweep_kernel<M, false, PK_NORMAL, false><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
    } else if (all_uniform) {
        if (injected) smc_sweep_kernel<M, false, PK_UNIFORM, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
        else if (P.flags & POP_PARTNER_SEGMENTS) smc_sweep_kernel<M, false, PK_UNIFORM, false, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
        else smc_sweep_kernel<M, false, PK_UNIFORM, false><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
    } else smc_sweep_kernel<M, false, PK_GENERIC><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, inj);
}
template <class M>
static void l_mc(const ModelOps&, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj,
                 const McArgs& mc)
{
    if constexpr (model_is_split<M>::value) {
        const bool injected = inj.a || inj.b || inj.s || inj.z || inj.u || inj.flags;
#ifndef ABCDEZ_MC_FUSED_ONLY                                               // (A/B builds)
        if (!injected && P.prop_theta) {                                   // heavy simulators: propose -> queue-driven simulate -> accept
            const unsigned g = grid_for(P.N, SWEEP_THREADS);
            if (prior_has_discrete<M::D>(pr)) {
                mc_propose_kernel<M, true><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, mc);
                simulate_queue_kernel<M, true><<<queue_grid<M, true>(g), SWEEP_THREADS, 0, st>>>(P, pr, md, QUEUE_MC, P.keys);
            } else {
                mc_propose_kernel<M, false><<<g, SWEEP_THREADS, 0, st>>>(P, pr, md, mc);
                simulate_queue_kernel<M, false><<<queue_grid<M, false>(g), SWEEP_THREADS, 0, st>>>(P, pr, md, QUEUE_MC, P.keys);
            }
            mc_accept_kernel<M><<<g, SWEEP_THREADS, 0, st>>>(P, mc);
            return;
        }
#endif
    }
    if (prior_has_discrete<M::D>(pr))
        mc_sweep_kernel<M, true><<<grid_for(P.N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(P, pr, md, inj, mc);
    else
        mc_sweep_kernel<M, false><<<grid_for(P.N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(P, pr, md, inj, mc);
}
template <class M>
static void l_sim(const ModelOps&, cudaStream_t st, const PriorDev*, const ModelData& md, int64_t N, const double* th, uint64_t seed,
                  uint32_t epoch, uint32_t tag, uint32_t id0, double* dist, double* blobs)
{
    simulate_kernel<M><<<grid_for(N, SWEEP_THREADS), SWEEP_THREADS, 0, st>>>(md, N, th, philox_keys(seed), epoch, tag, id0, dist, blobs);
}

template <class M>
static ModelOps make_ops()
{
    ModelOps o;
    o.name = M::name; o.d = M::D; o.blob = M::BLOB;
    o.init = &l_init<M>; o.smc_sweep = &l_smc<M>; o.mc_sweep = &l_mc<M>; o.simulate = &l_sim<M>; o.dyn = nullptr; o.f32_state = 1; o.split = model_is_split<M>::value ? (model_is_stepped<M>::value ? 2 : 1) : 0;
    return o;
}

// one accessor per model, defined in the inst_*.cu translation units (compiled in parallel)
#define ABCDEZ_DEFINE_MODEL(fn, M) const ModelOps* fn() { static const ModelOps o = make_ops<M>(); return &o; }
#endif  // __CUDACC_RTC__

}  // namespace abcdez
