// bookkeeping.cu -- the SMC bookkeeping of abcdesmc! (src/abcdez_smc.jl:301-326,357-376) as
// HBM-streaming sm_100a kernels:
//   eps schedule   quantile(delta[alive], alpha)   :301      radix select on order-preserving keys
//   reweighting    abcdesmc_update_ws! + :305-315  :59-83    two fused passes (unnormalised, normalise)
//   ESS / logZ     get_ess :8, :315,:323                     deterministic two-level reductions
//   alive list     (replaces the O(N) wsample scans :121,125) tile scan + compaction
//   resampling     wsample_stratified! :15-56, abcdesmc_resample! :85-104
// All schedule decisions are taken on the device by the last CTA of the kernel that
// produces their inputs ("last block done" tickets), so one SMC iteration is a fixed list of
// launches with no host round trip; skipped stages read their flag and return.
#include "internal.h"
#include "seqsum.cuh"
#include "ctrl.cuh"
#include "comm.cuh"

namespace abcdez {

static inline unsigned tiles_for(int64_t N) { return (unsigned)((N + TILE - 1) / TILE); }
// grid of the resampling gathers (grid-stride kernels): enough CTAs to fill a B200, few enough that a skipped launch is cheap
static inline unsigned resample_grid(int64_t N) { unsigned g = (unsigned)((N + BK_THREADS - 1) / BK_THREADS); return g < 1184u ? g : 1184u; }

__global__ void end_iter_kernel(PopDev P)
{
    Ctrl* c = P.ctrl;
    if (c->stop) return;
    ctrl_end_iter(P, c);
}

// pre-loop state, src/abcdez_smc.jl:255-292
__global__ void begin_run_kernel(PopDev P)
{
    Ctrl* c = P.ctrl;
    double N = (double)P.Ng;
    double w = 1.0 / N;
    push_hist(P, c, c->eps, 1.0 / (N * (w * w)), c->facc, c->Kmcmc);
    select_setup(c);
}

// ---------------------------------------------------------------------------------------
// radix select of v[j] among the alive distances: 11-bit digits, MSB first
// ---------------------------------------------------------------------------------------
// pick the bin holding rank sel_rank; all BK_THREADS threads of the calling CTA participate
__device__ void select_pick(const PopDev& P, Ctrl* c, int shift, int nbins, bool final_pass)
{
    __shared__ unsigned s_part[BK_THREADS];
    __shared__ unsigned s_found_bin, s_found_before;
    const int per = SEL_BINS / BK_THREADS;     // 8 bins per thread
    unsigned loc[per], tot = 0;
#pragma unroll
    for (int k = 0; k < per; ++k) {
        int b = threadIdx.x * per + k;
        loc[k] = (b < nbins) ? __ldcg(&P.sel_hist[b]) : 0u;
        tot += loc[k];
    }
    s_part[threadIdx.x] = tot;
    if (threadIdx.x == 0) { s_found_bin = 0xffffffffu; s_found_before = 0; }
    __syncthreads();
    // exclusive prefix over the 256 partials (short serial loop per thread is fine: 256 adds)
    unsigned before = 0;
    for (int t = 0; t < (int)threadIdx.x; ++t) before += s_part[t];
    unsigned long long rank = c->sel_rank;
    unsigned cum = before;
#pragma unroll
    for (int k = 0; k < per; ++k) {
        if (loc[k] && rank >= cum && rank < (unsigned long long)cum + loc[k]) {
            s_found_bin = threadIdx.x * per + k; s_found_before = cum;
        }
        cum += loc[k];
    }
    __syncthreads();
    // clear the histogram for the next pass
#pragma unroll
    for (int k = 0; k < per; ++k) P.sel_hist[threadIdx.x * per + k] = 0u;
    if (threadIdx.x == 0) {
        if (s_found_bin == 0xffffffffu) {
            if (!c->err) c->err = ABCDEZ_ERR_NO_ALIVE;      // empty alive set
        } else {
            c->sel_prefix |= ((unsigned long long)s_found_bin) << shift;
            c->sel_rank = rank - s_found_before;
            if (final_pass) c->q_a = key_f64(c->sel_prefix);
        }
    }
}

__global__ void __launch_bounds__(BK_THREADS)
select_hist_kernel(PopDev P, int shift, int nbins, unsigned long long himask, int first, int final_pass)
{
    Ctrl* c = P.ctrl;
    if (c->stop) return;
    __shared__ unsigned sh[SEL_BINS];
    for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
    const unsigned long long prefix = c->sel_prefix;
    const double* __restrict__ dl = P.delta[c->cur];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    double v[4];
    load4_f64(dl, i0, P.N, v);
    uint32_t al = load4_u8(P.alive, i0, P.N);
    int nan_seen = 0;
    unsigned long long kmn = ~0ull, kmx = 0ull;      // extrema(delta) over ALL particles (ranges_eps), first pass only
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bool ok = (al >> (8 * k)) & 0xff;
        unsigned long long key = f64_key(v[k]);
        if (first && i0 + k < P.N) { kmn = key < kmn ? key : kmn; kmx = key > kmx ? key : kmx; }
        if (ok && first && isnan(v[k])) nan_seen = 1;
        ok = ok && ((key & himask) == prefix);
        unsigned digit = ok ? (unsigned)((key >> shift) & (unsigned long long)(nbins - 1)) : 0xffffffffu;
        // warp-aggregated shared-memory atomics: the leading digits of distances are highly
        // concentrated, a plain atomicAdd would serialise 32-way
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        if (ok && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&sh[digit], (unsigned)__popc(peers));
    }
    if (nan_seen) atomicMax(&c->acc.err, (int)ABCDEZ_ERR_NAN_DISTANCE);
    __shared__ unsigned long long s_mn, s_mx;
    if (first) {
        if (threadIdx.x == 0) { s_mn = ~0ull; s_mx = 0ull; }
        __syncthreads();
        kmn = warp_min_u64(kmn); kmx = warp_max_u64(kmx);
        if ((threadIdx.x & 31) == 0) { atomicMin(&s_mn, kmn); atomicMax(&s_mx, kmx); }
    }
    __syncthreads();
    if (first && threadIdx.x == 0) { atomicMin(&c->acc.dmin_key, s_mn); atomicMax(&c->acc.dmax_key, s_mx); }
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        unsigned cnt = sh[b];
        if (cnt) atomicAdd(&P.sel_hist[b], cnt);
    }
    if (last_block(&c->acc.ticket[1], gridDim.x)) {
        if (first && threadIdx.x == 0) patch_extrema(P, c);
        select_pick(P, c, shift, nbins, final_pass != 0);
    }
}

// v[j+1]: count of keys <= v[j] and the smallest key above it; then the type-7 interpolation
// and the clamp eps = max(min(q, eps), eps_target)  (src/abcdez_smc.jl:301)
__global__ void __launch_bounds__(BK_THREADS) select_next_kernel(PopDev P)
{
    Ctrl* c = P.ctrl;
    if (c->stop) return;
    const unsigned long long akey = c->sel_prefix;
    const double* __restrict__ dl = P.delta[c->cur];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    double v[4];
    load4_f64(dl, i0, P.N, v);
    uint32_t al = load4_u8(P.alive, i0, P.N);
    unsigned cnt = 0; unsigned long long mn = ~0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bool ok = (al >> (8 * k)) & 0xff;
        unsigned long long key = f64_key(v[k]);
        if (ok) { if (key <= akey) cnt++; else mn = key < mn ? key : mn; }
    }
    __shared__ unsigned s_cnt[32]; __shared__ unsigned long long s_mn[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    cnt = warp_sum_u(cnt); mn = warp_min_u64(mn);
    if (lane == 0) { s_cnt[w] = cnt; s_mn[w] = mn; }
    __syncthreads();
    if (w == 0) {
        cnt = lane < nw ? s_cnt[lane] : 0u; mn = lane < nw ? s_mn[lane] : ~0ull;
        cnt = warp_sum_u(cnt); mn = warp_min_u64(mn);
        if (lane == 0) {
            if (cnt) atomicAdd(&c->acc.cnt_le, (unsigned long long)cnt);
            if (mn != ~0ull) atomicMin(&c->acc.min_gt_key, mn);
        }
    }
    if (last_block(&c->acc.ticket[1], gridDim.x)) {
        if (threadIdx.x == 0) {
            double a = c->q_a, b;
            if (c->n_alive <= 1 || __ldcg(&c->acc.cnt_le) >= c->sel_j + 1) b = a;
            else b = key_f64(__ldcg(&c->acc.min_gt_key));
            c->q_b = b;
            double g = c->q_gamma;
            double q = (isfinite(a) && isfinite(b)) ? a + g * (b - a) : (1.0 - g) * a + g * b;
            c->q = q;
            c->eps = fmax(fmin(q, c->eps), c->eps_target);                   // :301
        }
    }
}

int launch_eps_quantile(cudaStream_t st, const PopDev& P)
{
    unsigned g = tiles_for(P.N);
    // 64-bit keys: digits at shifts 53,42,31,20,9 (11 bits) and 0 (9 bits)
    const int shifts[6] = { 53, 42, 31, 20, 9, 0 };
    for (int p = 0; p < 6; ++p) {
        int nbins = (p == 5) ? 512 : 2048;
        unsigned long long himask = (p == 0) ? 0ull : (~0ull << (shifts[p - 1]));
        select_hist_kernel<<<g, BK_THREADS, 0, st>>>(P, shifts[p], nbins, himask, p == 0, p == 5);
    }
    select_next_kernel<<<g, BK_THREADS, 0, st>>>(P);
    return 7;
}

// ---------------------------------------------------------------------------------------
// reweighting, src/abcdez_smc.jl:59-83 and :305-326
// ---------------------------------------------------------------------------------------
// pass A: ws (alive only), wprod = Wns .* ws kept unnormalised in W; wnorm = sum(wprod); logZ += log(wnorm)
__global__ void __launch_bounds__(BK_THREADS) reweight_a_kernel(PopDev P)
{
    Ctrl* c = P.ctrl;
    if (c->stop) return;
    __shared__ double s_red[32];
    const double eps_new = c->eps, eps_old = c->eps_k;
    const int kind = c->kind;
    const double* __restrict__ dl = P.delta[c->cur];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    double v[4], w[4];
    load4_f64(dl, i0, P.N, v);
    load4_f64(P.W, i0, P.N, w);
    uint32_t al = load4_u8(P.alive, i0, P.N);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bool ok = ((al >> (8 * k)) & 0xff) && (i0 + k < P.N);
        double ws = 0.0;
        if (ok) ws = abck_ws(kind, eps_new, eps_old, v[k]);                                       // :75
        w[k] = ok ? w[k] * ws : 0.0;                                                              // :308
        acc += w[k];
    }
    store4_f64(P.W, i0, P.N, w);
    double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) P.partial[blockIdx.x] = t;
    if (last_block(&c->acc.ticket[2], gridDim.x)) {
        double a = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) a += __ldcg(&P.partial[b]);
        a = block_sum(a, s_red);
        if (threadIdx.x == 0) {
            c->wnorm = a;                                                   // :309
            c->logZ += plog(a);                                              // :315
        }
    }
}

__global__ void __launch_bounds__(BK_THREADS) reweight_b_kernel(PopDev P)
{
    Ctrl* c = P.ctrl;
    if (c->stop) return;
    __shared__ double s_red[32];
    __shared__ unsigned s_cnt[32];
    const double wnorm = c->wnorm;
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    double w[4];
    load4_f64(P.W, i0, P.N, w);
    double acc = 0.0; unsigned cnt = 0; uint32_t al = 0; double wal = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < P.N) {
            w[k] = w[k] / wnorm;                                            // :310
            bool a = (w[k] > 0.0);                                          // :311
            if (a) { al |= 1u << (8 * k); cnt++; wal = w[k]; }
            acc += w[k] * w[k];
        }
    }
    store4_f64(P.W, i0, P.N, w);
    if (i0 + 3 < P.N) *reinterpret_cast<uint32_t*>(P.alive + i0) = al;
    else { for (int k = 0; k < 4; ++k) if (i0 + k < P.N) P.alive[i0 + k] = (al >> (8 * k)) & 0xff; }
    if (cnt) c->acc.w_alive = wal;     // indicator kernels: every alive weight is this same double
    double t = block_sum(acc, s_red);
    int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    cnt = warp_sum_u(cnt);
    if (lane == 0) s_cnt[wp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tc = 0;
        for (int q = 0; q < nw; ++q) tc += s_cnt[q];
        P.partial[blockIdx.x] = t;
        P.tile_cnt[blockIdx.x] = tc;
    }
    if (last_block(&c->acc.ticket[2], gridDim.x)) {
        double a = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) a += __ldcg(&P.partial[b]);
        a = block_sum(a, s_red);
        // exclusive scan of the tile counts -> tile offsets (chunked: each thread owns a contiguous chunk)
        __shared__ unsigned s_chunk[BK_THREADS];
        unsigned nt = gridDim.x, per = (nt + blockDim.x - 1) / blockDim.x;
        unsigned lo = threadIdx.x * per, hi = lo + per < nt ? lo + per : nt, sum = 0;
        for (unsigned b = lo; b < hi; ++b) sum += __ldcg(&P.tile_cnt[b]);
        s_chunk[threadIdx.x] = sum;
        __syncthreads();
        unsigned off = 0;
        for (int q = 0; q < (int)threadIdx.x; ++q) off += s_chunk[q];
        for (unsigned b = lo; b < hi; ++b) { unsigned tcnt = __ldcg(&P.tile_cnt[b]); P.tile_cnt[b] = off; off += tcnt; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_cnt[0] = off;   // total alive (last chunk's end)
        __syncthreads();
        if (threadIdx.x == 0) ctrl_after_reweight(P, c, a, s_cnt[0], s_cnt[0]);
    }
}

int launch_reweight(cudaStream_t st, const PopDev& P)
{
    unsigned g = tiles_for(P.N);
    reweight_a_kernel<<<g, BK_THREADS, 0, st>>>(P);
    reweight_b_kernel<<<g, BK_THREADS, 0, st>>>(P);
    return 2;
}

// ---------------------------------------------------------------------------------------
// compaction of the alive flags into alive_list (the O(1) replacement of wsample's scans)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BK_THREADS) compact_kernel(PopDev P)
{
    Ctrl* c = P.ctrl;
    if (c->stop || c->n_alive == P.N) return;       // identity list is implicit
    __shared__ unsigned s_w[32];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    uint32_t al = load4_u8(P.alive, i0, P.N);
    unsigned cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt += ((al >> (8 * k)) & 0xff) ? 1u : 0u;
    // block exclusive scan of cnt
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (int q = 0; q < w; ++q) woff += s_w[q];
    (void)nw;
    // alive particles (index order) fill list[0, n_alive); the dead ones follow in list[n_alive, N):
    // the sweep maps thread j to particle list[j], so alive and dead warps are homogeneous
    unsigned pos = P.tile_cnt[blockIdx.x] + woff + incl - cnt;          // alive before element i0
    unsigned dpos = c->n_alive + ((unsigned)i0 - pos);                    // dead before element i0
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < P.N) {
            if ((al >> (8 * k)) & 0xff) P.alive_list[pos++] = (uint32_t)(i0 + k);
            else P.alive_list[dpos++] = (uint32_t)(i0 + k);
        }
    }
}

int launch_compact(cudaStream_t st, const PopDev& P)
{
    compact_kernel<<<tiles_for(P.N), BK_THREADS, 0, st>>>(P);
    return 1;
}

// ---------------------------------------------------------------------------------------
// stratified resampling, src/abcdez_smc.jl:15-56 and :85-104
// ---------------------------------------------------------------------------------------
__device__ void ctrl_after_resample(const PopDev& P, Ctrl* c)
{
    double N = (double)P.Ng, w = 1.0 / N;
    c->cur ^= 1;                     // the gathered generation is the live one
    c->n_alive = P.N; c->n_alive_g = P.Ng;   // :102-103
    c->ess = 1.0 / (N * (w * w));    // get_ess(Wns) after the reset, :326
    c->n_resamples += 1;
}

// stratum draw r_si = rand(Uniform(unif0, unif1)) = unif0 + (unif1 - unif0) * u   (:46-47)
__device__ __forceinline__ double stratum_draw(const SeqTab* edges, double sval, uint32_t si, double u)
{
    double unif0 = seqtab_value(edges, si);         // running sum of sval, si steps (:46,:52)
    double unif1 = unif0 + sval;
    return unif0 + (unif1 - unif0) * u;
}

// one theta row from generation to generation (or from a peer's HBM); f32: the rows are floats (POP_FP32_STATE)
__device__ __forceinline__ void copy_theta_row(const double* __restrict__ sb, double* __restrict__ db, size_t src, size_t dst, int DS, bool f32)
{
    if (f32) {
        const float* s = reinterpret_cast<const float*>(sb) + src * DS;
        float* d = reinterpret_cast<float*>(db) + dst * DS;
        if (DS & 1) { for (int k = 0; k < DS; ++k) d[k] = s[k]; }
        else { for (int k = 0; k < DS; k += 2) *reinterpret_cast<float2*>(d + k) = *reinterpret_cast<const float2*>(s + k); }
    } else {
        const double* s = sb + src * DS;
        double* d = db + dst * DS;
        if (DS & 1) { for (int k = 0; k < DS; ++k) d[k] = s[k]; }
        else { for (int k = 0; k < DS; k += 2) *reinterpret_cast<double2*>(d + k) = *reinterpret_cast<const double2*>(s + k); }
    }
}

__device__ __forceinline__ void gather_particle(const PopDev& P, int cur, int DS, int NB, uint32_t dst, uint32_t src)
{
    copy_theta_row(P.theta[cur], P.theta[cur ^ 1], src, dst, DS, (P.flags & POP_FP32_STATE) != 0);
    P.logpi[cur ^ 1][dst] = P.logpi[cur][src];
    P.delta[cur ^ 1][dst] = P.delta[cur][src];
    for (int k = 0; k < NB; ++k) P.blob[cur ^ 1][(size_t)dst * NB + k] = P.blob[cur][(size_t)src * NB + k];
}

// closed-form path (indicator kernels): no scan at all
__global__ void __launch_bounds__(BK_THREADS)
resample_uniform_kernel(PopDev P, int DS, int NB, const double* __restrict__ inj_u, uint32_t epoch, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    __shared__ SeqTab s_w;        // weights table (the edges table is read through L1/L2)
    {   // cooperative copy of the small table into shared memory
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.tabs[0]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_w);
        for (unsigned q = threadIdx.x; q < sizeof(SeqTab) / 4; q += blockDim.x) dst[q] = src[q];
    }
    __syncthreads();
    const int cur = c->cur;
    const uint32_t N = P.N, n_alive = c->n_alive;
    const double sval = 1.0 / (double)N;                                     // :34
    // grid-stride over a capped grid (resample_grid): the kernel is launched every iteration and returns at the flag
    // test above in all but ~1 of 14, so the launch must be cheap when it has nothing to do
    for (uint32_t si = blockIdx.x * blockDim.x + threadIdx.x; si < N; si += gridDim.x * blockDim.x) {
        double u, u2;
        if (inj_u) u = inj_u[si];
        else { Stream rs(P.keys, (P.flags & POP_SYSTEMATIC) ? 0u : P.id0 + si, epoch, TAG_RESAMPLE); rs.u2(0u, u, u2); }   // systematic: ONE uniform for all strata
        double r = stratum_draw(&P.tabs[1], sval, si, u);
        unsigned long long k = seqtab_first_ge(&s_w, r);                     // :48-51 in closed form
        uint32_t src;
        if (k == 0) src = 0u;                                                // r == 0: the reference leaves i = 0
        else {
            if (k > n_alive) k = n_alive;
            src = (n_alive == N) ? (uint32_t)(k - 1) : P.alive_list[k - 1];
        }
        gather_particle(P, cur, DS, NB, si, src);                            // :96-99
        P.inds[si] = (int32_t)src;
        P.W[si] = sval; P.alive[si] = 1; P.moved[si] = 1;                    // :102-103; the old buffer is stale (nobody gathers from these)
    }
    if (last_block(&c->acc.ticket[3], gridDim.x)) {
        if (threadIdx.x == 0) ctrl_after_resample(P, c);
    }
}

// ---- sharded runs (exchange 3): GLOBAL stratified resampling over NVLink peer memory ------------
// Rank r resolves its own output strata [id0, id0 + N) against the whole population: the k-th alive particle
// of the concatenated per-rank alive lists is looked up in its owner's HBM (peer loads through the mapped
// population slabs, PeerTable) and its row is gathered straight into this rank's next generation -- no
// send/recv plan, no staging.  Same uniforms (Philox stream of the global stratum), same closed-form
// sequential sums and therefore the same indices as a single-GPU run over the same particles.
// Peers may read a rank's live generation only between two barriers: peer_barrier_kernel (everyone's alive
// list and particle rows are final) and the exchange in this kernel's last CTA (everyone is done reading).
__global__ void peer_barrier_kernel(PopDev P, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    xchg_small(P.x, c, nullptr, 0);
}

__global__ void __launch_bounds__(BK_THREADS)
resample_uniform_sharded_kernel(const __grid_constant__ PopDev P, int DS, int NB, uint32_t epoch, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    __shared__ SeqTab s_w;
    __shared__ unsigned s_end[XCHG_MAXR];         // inclusive prefix of the per-rank alive counts
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.tabs[0]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_w);
        for (unsigned q = threadIdx.x; q < sizeof(SeqTab) / 4; q += blockDim.x) dst[q] = src[q];
        if (threadIdx.x == 0) { unsigned a = 0; for (int r = 0; r < P.x.world; ++r) { a += c->rank_alive[r]; s_end[r] = a; } }
    }
    __syncthreads();
    const int cur = c->cur;
    const uint32_t N = P.N, n_alive_g = c->n_alive_g;
    const double sval = 1.0 / (double)P.Ng;                                  // :34
    for (uint32_t si = blockIdx.x * blockDim.x + threadIdx.x; si < N; si += gridDim.x * blockDim.x) {
        const uint32_t sg = P.id0 + si;                                      // global stratum
        double u, u2;
        Stream rs(P.keys, (P.flags & POP_SYSTEMATIC) ? 0u : sg, epoch, TAG_RESAMPLE); rs.u2(0u, u, u2);
        double r = stratum_draw(&P.tabs[1], sval, sg, u);
        unsigned long long k = seqtab_first_ge(&s_w, r);                     // :48-51 in closed form
        int q = 0; uint32_t src = 0u;                                        // r == 0: global particle 0 (the reference leaves i = 0)
        if (k != 0) {
            if (k > n_alive_g) k = n_alive_g;
            while (q < P.x.world - 1 && k > s_end[q]) ++q;                   // owner of the k-th alive particle
            const uint32_t kk = (uint32_t)k - (q ? s_end[q - 1] : 0u);       // 1-based among its alive particles
            const PeerPop& pp = P.peers->p[q];
            src = (c->rank_alive[q] == pp.N) ? kk - 1 : pp.alive_list[kk - 1];
        }
        const PeerPop& pp = P.peers->p[q];
        {
            copy_theta_row(pp.theta[cur], P.theta[cur ^ 1], src, si, DS, (P.flags & POP_FP32_STATE) != 0);
            P.logpi[cur ^ 1][si] = pp.logpi[cur][src];
            P.delta[cur ^ 1][si] = pp.delta[cur][src];
            for (int e = 0; e < NB; ++e) P.blob[cur ^ 1][(size_t)si * NB + e] = pp.blob[cur][(size_t)src * NB + e];
        }
        P.inds[si] = (int32_t)(pp.id0 + src);                                // global source index
        P.W[si] = sval; P.alive[si] = 1; P.moved[si] = 1;                    // :102-103; the old buffer is stale
    }
    if (last_block(&c->acc.ticket[3], gridDim.x)) {
        if (threadIdx.x == 0) {
            xchg_small(P.x, c, nullptr, 0);      // nobody reads this rank's old generation any more
            ctrl_after_resample(P, c);
        }
    }
}

__device__ __forceinline__ uint32_t search_cumsum(const double* __restrict__ cs, uint32_t N, double r);

// ---- sharded runs, general weights (Epanechnikov kernels) ------------------------------------------------------
// wsample_stratified! (src/abcdez_smc.jl:15-56) walks ONE running sum over the whole population, so rank r's
// cumulative weights must continue where rank r-1's end.  Exact mode (2): a chain of `world` exchange rounds --
// in round q rank q runs the reference's sequential FP64 sum over its block, starting from the end value of rank
// q-1, and posts its own end value; every rank leaves with all end values (Ctrl::rank_end).  Parallel mode (1):
// every rank scans its block in parallel, the block totals are exchanged once and added up in rank order to give
// the carry-in of every block (rounds like the single-GPU parallel scan: differs from the sequential sum in the
// last bits).  Then every rank resolves its own output strata: owner = first rank whose end value
// reaches r, source = first particle of that rank whose cumulative weight reaches r (binary search in the owner's
// HBM over NVLink), row gathered from the owner like in the uniform-weight kernel above.
__global__ void __launch_bounds__(BK_THREADS) scan_sequential_sharded_kernel(const __grid_constant__ PopDev P, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    __shared__ double buf[2048];
    const uint32_t N = P.N;
    double carry = 0.0;                                   // thread 0: end value of the previous rank
    for (int q = 0; q < P.x.world; ++q) {
        double s = carry;
        if (q == P.x.rank) {
            for (size_t base = 0; base < N; base += 2048) {
                for (int t = threadIdx.x; t < 2048; t += blockDim.x) buf[t] = (base + t < N) ? P.W[base + t] : 0.0;
                __syncthreads();
                if (threadIdx.x == 0) {
                    int lim = (N - base) < 2048 ? (int)(N - base) : 2048;
                    for (int t = 0; t < lim; ++t) { s = __dadd_rn(s, buf[t]); buf[t] = s; }
                }
                __syncthreads();
                for (int t = threadIdx.x; t < 2048; t += blockDim.x) if (base + t < N) P.cumsum[base + t] = buf[t];
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            unsigned long long rec[1] = { (unsigned long long)__double_as_longlong(s) };
            __threadfence_system();                       // this rank's cumsum is visible to its peers before the flag
            const unsigned slot = xchg_small(P.x, c, rec, 1);
            carry = __longlong_as_double((long long)xchg_word(P.x, slot, q, 0));
            c->rank_end[q] = carry;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) seqtab_build(&P.tabs[1], 1.0 / (double)P.Ng, (unsigned long long)P.Ng);
}

// parallel mode, step 2 of 3 (after scan_tile_sums_pop_kernel, before scan_apply_pop_kernel): block total ->
// exchange -> carry-in -> exclusive tile offsets
__global__ void scan_tile_offsets_sharded_kernel(const __grid_constant__ PopDev P, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double tot = 0.0;
        for (unsigned b = 0; b < P.ntiles; ++b) tot += P.partial[b];
        unsigned long long rec[1] = { (unsigned long long)__double_as_longlong(tot) };
        const unsigned slot = xchg_small(P.x, c, rec, 1);
        double s = 0.0;
        for (int q = 0; q < P.x.rank; ++q) s += __longlong_as_double((long long)xchg_word(P.x, slot, q, 0));
        for (unsigned b = 0; b < P.ntiles; ++b) { double t = P.partial[b]; P.partial[b] = s; s += t; }
        seqtab_build(&P.tabs[1], 1.0 / (double)P.Ng, (unsigned long long)P.Ng);
    }
}

// parallel mode, after scan_apply_pop_kernel: the end values of all blocks; also the barrier "every rank's cumulative
// weights and particle rows are final"
__global__ void scan_ends_sharded_kernel(const __grid_constant__ PopDev P, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    unsigned long long rec[1] = { (unsigned long long)__double_as_longlong(P.cumsum[P.N - 1]) };
    __threadfence_system();
    const unsigned slot = xchg_small(P.x, c, rec, 1);
    for (int q = 0; q < P.x.world; ++q) c->rank_end[q] = __longlong_as_double((long long)xchg_word(P.x, slot, q, 0));
}

__global__ void __launch_bounds__(BK_THREADS)
resample_general_sharded_kernel(const __grid_constant__ PopDev P, int DS, int NB, uint32_t epoch, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    __shared__ double s_end[XCHG_MAXR];
    if (threadIdx.x < XCHG_MAXR) s_end[threadIdx.x] = (int)threadIdx.x < P.x.world ? c->rank_end[threadIdx.x] : 0.0;
    __syncthreads();
    const int cur = c->cur;
    const uint32_t N = P.N;
    const double sval = 1.0 / (double)P.Ng;                                  // :34
    for (uint32_t si = blockIdx.x * blockDim.x + threadIdx.x; si < N; si += gridDim.x * blockDim.x) {
        const uint32_t sg = P.id0 + si;                                      // global stratum
        double u, u2;
        Stream rs(P.keys, (P.flags & POP_SYSTEMATIC) ? 0u : sg, epoch, TAG_RESAMPLE); rs.u2(0u, u, u2);
        const double r = stratum_draw(&P.tabs[1], sval, sg, u);
        int q = 0; uint32_t src = 0u;                                        // r <= 0: global particle 0 (the "i = 0" quirk, clamped)
        if (r > 0.0) {
            while (q < P.x.world - 1 && s_end[q] < r) ++q;                   // first block whose end value reaches r
            src = search_cumsum(P.peers->p[q].cumsum, P.peers->p[q].N, r);   // (r beyond the total: the last particle)
        }
        const PeerPop& pp = P.peers->p[q];
        {
            copy_theta_row(pp.theta[cur], P.theta[cur ^ 1], src, si, DS, (P.flags & POP_FP32_STATE) != 0);
            P.logpi[cur ^ 1][si] = pp.logpi[cur][src];
            P.delta[cur ^ 1][si] = pp.delta[cur][src];
            for (int e = 0; e < NB; ++e) P.blob[cur ^ 1][(size_t)si * NB + e] = pp.blob[cur][(size_t)src * NB + e];
        }
        P.inds[si] = (int32_t)(pp.id0 + src);                                // global source index
        P.W[si] = sval; P.alive[si] = 1; P.moved[si] = 1;                    // :102-103 (peers read cumsum and the old rows, not these)
    }
    if (last_block(&c->acc.ticket[3], gridDim.x)) {
        if (threadIdx.x == 0) {
            xchg_small(P.x, c, nullptr, 0);      // nobody reads this rank's old generation or cumulative weights any more
            ctrl_after_resample(P, c);
        }
    }
}

// general weights (Epanechnikov kernels): inclusive scan of W, then a search per stratum.
// mode 1: parallel three-phase scan (tile sums, tile offsets, tile rescan) -- rounds differently
// from the reference's sequential sum, documented in DESIGN.md; mode 2: sequential scan, exact.
__global__ void __launch_bounds__(BK_THREADS) scan_tile_sums_kernel(const double* __restrict__ W, uint32_t N, double* partial)
{
    __shared__ double s_red[32];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    double w[4];
    load4_f64(W, i0, N, w);
    double acc = ((w[0] + w[1]) + w[2]) + w[3];
    double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void scan_tile_offsets_kernel(double* partial, unsigned ntiles)
{
    // single thread: sequential exclusive scan of the tile sums (ntiles <= ~10^5)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (unsigned b = 0; b < ntiles; ++b) { double t = partial[b]; partial[b] = s; s += t; }
    }
}

__global__ void __launch_bounds__(BK_THREADS)
scan_apply_kernel(const double* __restrict__ W, uint32_t N, const double* __restrict__ partial, double* __restrict__ cumsum)
{
    __shared__ double s_w[32];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    double w[4];
    load4_f64(W, i0, N, w);
    w[1] += w[0]; w[2] += w[1]; w[3] += w[2];
    double incl = w[3];
    int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[wp] = incl;
    __syncthreads();
    double off = partial[blockIdx.x];
    for (int q = 0; q < wp; ++q) off += s_w[q];
    off += incl - w[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] += off;
    store4_f64(cumsum, i0, N, w);
}

// exact mode: the reference's sequential FP64 cumsum (:50), staged through shared memory so
// that the one adding thread never waits on HBM
__global__ void __launch_bounds__(BK_THREADS) scan_sequential_kernel(const double* __restrict__ W, uint32_t N, double* __restrict__ cumsum)
{
    __shared__ double buf[2048];
    double s = 0.0;
    for (size_t base = 0; base < N; base += 2048) {
        for (int q = threadIdx.x; q < 2048; q += blockDim.x) buf[q] = (base + q < N) ? W[base + q] : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            int lim = (N - base) < 2048 ? (int)(N - base) : 2048;
            for (int q = 0; q < lim; ++q) { s = __dadd_rn(s, buf[q]); buf[q] = s; }
        }
        __syncthreads();
        for (int q = threadIdx.x; q < 2048; q += blockDim.x) if (base + q < N) cumsum[base + q] = buf[q];
        __syncthreads();
    }
}

// first index (0-based) with cumsum[i] >= r; r <= 0 -> 0 ("i = 0" quirk, clamped); r > total -> N-1
__device__ __forceinline__ uint32_t search_cumsum(const double* __restrict__ cs, uint32_t N, double r)
{
    if (!(r > 0.0)) return 0u;
    uint32_t lo = 0, hi = N;      // find first i with cs[i] >= r
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (cs[m] < r) lo = m + 1; else hi = m; }
    return lo >= N ? N - 1 : lo;
}

__global__ void __launch_bounds__(BK_THREADS)
resample_general_kernel(PopDev P, int DS, int NB, const double* __restrict__ inj_u, uint32_t epoch, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    const int cur = c->cur;
    const uint32_t N = P.N;
    const double sval = 1.0 / (double)N;
    for (uint32_t si = blockIdx.x * blockDim.x + threadIdx.x; si < N; si += gridDim.x * blockDim.x) {
        double u, u2;
        if (inj_u) u = inj_u[si];
        else { Stream rs(P.keys, (P.flags & POP_SYSTEMATIC) ? 0u : P.id0 + si, epoch, TAG_RESAMPLE); rs.u2(0u, u, u2); }   // systematic: ONE uniform for all strata
        double r = stratum_draw(&P.tabs[1], sval, si, u);
        uint32_t src = search_cumsum(P.cumsum, N, r);
        gather_particle(P, cur, DS, NB, si, src);
        P.inds[si] = (int32_t)src;
        P.W[si] = sval; P.alive[si] = 1; P.moved[si] = 1;                    // (the search reads cumsum, not W)
    }
    if (last_block(&c->acc.ticket[3], gridDim.x)) {
        if (threadIdx.x == 0) ctrl_after_resample(P, c);
    }
}

// wrappers that respect the device-side do_resample flag for the scan kernels
__global__ void __launch_bounds__(BK_THREADS) scan_tile_sums_pop_kernel(PopDev P, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    __shared__ double s_red[32];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    double w[4];
    load4_f64(P.W, i0, P.N, w);
    double acc = ((w[0] + w[1]) + w[2]) + w[3];
    double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) P.partial[blockIdx.x] = t;
}

__global__ void scan_tile_offsets_pop_kernel(PopDev P, int force, int build_edges)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (unsigned b = 0; b < P.ntiles; ++b) { double t = P.partial[b]; P.partial[b] = s; s += t; }
        if (build_edges) seqtab_build(&P.tabs[1], 1.0 / (double)P.N, (unsigned long long)P.N);
    }
}

__global__ void __launch_bounds__(BK_THREADS) scan_apply_pop_kernel(PopDev P, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    __shared__ double s_w[32];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    double w[4];
    load4_f64(P.W, i0, P.N, w);
    w[1] += w[0]; w[2] += w[1]; w[3] += w[2];
    double incl = w[3];
    int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[wp] = incl;
    __syncthreads();
    double off = P.partial[blockIdx.x];
    for (int q = 0; q < wp; ++q) off += s_w[q];
    off += incl - w[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] += off;
    store4_f64(P.cumsum, i0, P.N, w);
}

__global__ void __launch_bounds__(BK_THREADS) scan_sequential_pop_kernel(PopDev P, int force)
{
    Ctrl* c = P.ctrl;
    if (c->stop || !(c->do_resample || force)) return;
    __shared__ double buf[2048];
    const uint32_t N = P.N;
    double s = 0.0;
    for (size_t base = 0; base < N; base += 2048) {
        for (int q = threadIdx.x; q < 2048; q += blockDim.x) buf[q] = (base + q < N) ? P.W[base + q] : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            int lim = (N - base) < 2048 ? (int)(N - base) : 2048;
            for (int q = 0; q < lim; ++q) { s = __dadd_rn(s, buf[q]); buf[q] = s; }
        }
        __syncthreads();
        for (int q = threadIdx.x; q < 2048; q += blockDim.x) if (base + q < N) P.cumsum[base + q] = buf[q];
        __syncthreads();
    }
    if (threadIdx.x == 0) seqtab_build(&P.tabs[1], 1.0 / (double)P.N, (unsigned long long)P.N);
}

// host-forced table build for the stage-level resample call (force=1 bypasses reweight)
__global__ void build_tabs_kernel(PopDev P)
{
    Ctrl* c = P.ctrl;
    seqtab_build(&P.tabs[0], __ldcg(&c->acc.w_alive), (unsigned long long)c->n_alive);
    seqtab_build(&P.tabs[1], 1.0 / (double)P.N, (unsigned long long)P.N);
}

// mode: 0 auto, 1 parallel scan, 2 sequential scan.  kind_is_indicator decided on the host
// from the control block mirror (the kernel type never changes during a run).
int launch_resample(cudaStream_t st, const PopDev& P, int DS, int NB, const double* inj_u, uint32_t epoch,
                    int mode, int force)
{
    unsigned gt = tiles_for(P.N), gp = resample_grid(P.N);
    if (P.x.world > 1) {             // sharded: global resampling over peer memory
        if (mode == 0) {
            peer_barrier_kernel<<<1, 1, 0, st>>>(P, force);
            resample_uniform_sharded_kernel<<<gp, BK_THREADS, 0, st>>>(P, DS, NB, epoch, force);
            return 2;
        }
        int n = 0;
        if (mode == 2) { scan_sequential_sharded_kernel<<<1, BK_THREADS, 0, st>>>(P, force); n = 1; }
        else {
            scan_tile_sums_pop_kernel<<<gt, BK_THREADS, 0, st>>>(P, force);
            scan_tile_offsets_sharded_kernel<<<1, 32, 0, st>>>(P, force);
            scan_apply_pop_kernel<<<gt, BK_THREADS, 0, st>>>(P, force);
            scan_ends_sharded_kernel<<<1, 1, 0, st>>>(P, force);
            n = 4;
        }
        resample_general_sharded_kernel<<<gp, BK_THREADS, 0, st>>>(P, DS, NB, epoch, force);
        return n + 1;
    }
    if (mode == 0) {
        if (force) build_tabs_kernel<<<1, 1, 0, st>>>(P);
        resample_uniform_kernel<<<gp, BK_THREADS, 0, st>>>(P, DS, NB, inj_u, epoch, force);
        return force ? 2 : 1;
    }
    if (mode == 2) {
        scan_sequential_pop_kernel<<<1, BK_THREADS, 0, st>>>(P, force);
        resample_general_kernel<<<gp, BK_THREADS, 0, st>>>(P, DS, NB, inj_u, epoch, force);
        return 2;
    }
    scan_tile_sums_pop_kernel<<<gt, BK_THREADS, 0, st>>>(P, force);
    scan_tile_offsets_pop_kernel<<<1, 32, 0, st>>>(P, force, 1);
    scan_apply_pop_kernel<<<gt, BK_THREADS, 0, st>>>(P, force);
    resample_general_kernel<<<gp, BK_THREADS, 0, st>>>(P, DS, NB, inj_u, epoch, force);
    return 4;
}

// wsample_stratified!(rng, weights, inds) on caller-supplied weights (test/runtests.jl:13-19)
__global__ void __launch_bounds__(BK_THREADS)
strat_indices_kernel(uint32_t N, const double* __restrict__ cumsum, const double* __restrict__ u,
                     const SeqTab* __restrict__ edges, long long* __restrict__ inds)
{
    uint32_t si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= N) return;
    double r = stratum_draw(edges, 1.0 / (double)N, si, u[si]);
    inds[si] = (long long)search_cumsum(cumsum, N, r) + 1;     // 1-based like the reference
}

__global__ void build_edges_kernel(SeqTab* tabs, uint32_t N)
{
    seqtab_build(&tabs[1], 1.0 / (double)N, (unsigned long long)N);
}

int launch_strat_indices(cudaStream_t st, int64_t N, const double* W, const double* u, double* cumsum,
                         double* partial, SeqTab* tabs, int mode, long long* inds)
{
    unsigned gt = tiles_for(N), gp = (unsigned)((N + BK_THREADS - 1) / BK_THREADS);
    int n = 0;
    build_edges_kernel<<<1, 1, 0, st>>>(tabs, (uint32_t)N); n++;
    if (mode == 2) { scan_sequential_kernel<<<1, BK_THREADS, 0, st>>>(W, (uint32_t)N, cumsum); n++; }
    else {
        scan_tile_sums_kernel<<<gt, BK_THREADS, 0, st>>>(W, (uint32_t)N, partial);
        scan_tile_offsets_kernel<<<1, 32, 0, st>>>(partial, gt);
        scan_apply_kernel<<<gt, BK_THREADS, 0, st>>>(W, (uint32_t)N, partial, cumsum);
        n += 3;
    }
    strat_indices_kernel<<<gp, BK_THREADS, 0, st>>>((uint32_t)N, cumsum, u, &tabs[1], inds); n++;
    return n;
}

int launch_end_iter(cudaStream_t st, const PopDev& P) { end_iter_kernel<<<1, 1, 0, st>>>(P); return 1; }
int launch_begin_run(cudaStream_t st, const PopDev& P) { begin_run_kernel<<<1, 1, 0, st>>>(P); return 1; }

// ---------------------------------------------------------------------------------------
// extrema(delta) over the live generation (abcdemc!, src/abcdez_mc.jl:146; stage calls)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BK_THREADS) minmax_kernel(PopDev P)
{
    Ctrl* c = P.ctrl;
    const double* __restrict__ dl = P.delta[c->cur];
    size_t i0 = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * 4;
    unsigned long long mn = ~0ull, mx = 0ull;
    for (int k = 0; k < 4; ++k) if (i0 + k < P.N) { unsigned long long key = f64_key(dl[i0 + k]); mn = key < mn ? key : mn; mx = key > mx ? key : mx; }
    mn = warp_min_u64(mn); mx = warp_max_u64(mx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&c->acc.dmin_key, mn); atomicMax(&c->acc.dmax_key, mx); }
    if (last_block(&c->acc.ticket[4], gridDim.x)) {
        if (threadIdx.x == 0) {
            if (P.x.world > 1 && c->err != ABCDEZ_ERR_NCCL) {     // global extrema
                unsigned long long rec[2] = { __ldcg(&c->acc.dmin_key), __ldcg(&c->acc.dmax_key) };
                const unsigned slot = xchg_small(P.x, c, rec, 2);
                unsigned long long mn = ~0ull, mx = 0ull;
                for (int r = 0; r < P.x.world; ++r) {
                    unsigned long long a = xchg_word(P.x, slot, r, 0), b = xchg_word(P.x, slot, r, 1);
                    mn = a < mn ? a : mn; mx = b > mx ? b : mx;
                }
                __stcg(&c->acc.dmin_key, mn); __stcg(&c->acc.dmax_key, mx);
            }
            patch_extrema(P, c);
        }
    }
}

int launch_minmax(cudaStream_t st, const PopDev& P)
{
    minmax_kernel<<<tiles_for(P.N), BK_THREADS, 0, st>>>(P);
    return 1;
}

// ---------------------------------------------------------------------------------------
// layout conversion and result assembly (push_p on every particle, src/abcdez_smc.jl:382)
// ---------------------------------------------------------------------------------------
__global__ void pack_rows_kernel(int D, int DS, int64_t N, const double* __restrict__ src, double* __restrict__ dst, int to_rows)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * D) return;
    int64_t i = e / D; int k = (int)(e - i * D);
    if (to_rows) dst[i * DS + k] = src[e]; else dst[e] = src[i * DS + k];
}

int launch_pack_rows(cudaStream_t st, int D, int64_t N, const double* dense, double* rows, int to_rows)
{
    int DS = row_stride(D);
    int64_t n = N * D;
    unsigned g = (unsigned)((n + 255) / 256);
    if (to_rows) pack_rows_kernel<<<g, 256, 0, st>>>(D, DS, N, dense, rows, 1);
    else pack_rows_kernel<<<g, 256, 0, st>>>(D, DS, N, rows, const_cast<double*>(dense), 0);
    return 1;
}

__global__ void push_rows_kernel(PopDev P, PriorDev pr, int D, int DS, double* __restrict__ out)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)P.N * D) return;
    int64_t i = e / D; int k = (int)(e - i * D);
    const double* th = P.theta[P.ctrl->cur];
    double v = (P.flags & POP_FP32_STATE) ? (double)reinterpret_cast<const float*>(th)[i * DS + k] : th[i * DS + k];
    out[e] = fam_is_discrete(pr.family[k]) ? rint(v) : v;
}

int launch_push_rows(cudaStream_t st, const PopDev& P, const PriorDev& pr, int D, double* out_dense)
{
    int64_t n = (int64_t)P.N * D;
    push_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, pr, D, row_stride(D), out_dense);
    return 1;
}

}  // namespace abcdez
