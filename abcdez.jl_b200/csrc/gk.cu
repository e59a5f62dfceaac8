// gk.cu -- config 3 of BASELINE.json: the g-and-k distribution, 4 parameters, quantile summaries of n
// (<= 16384, default 10^4) simulated draws.  One dist! evaluation is ~10^6 FP32 operations, so a particle is
// simulated by a whole CTA instead of one thread:
//   draws      thread t generates the Philox blocks t, t+256, ... (4 normals each, Box-Muller in FP32) and
//              pushes them through x = A + B (1 + 0.8 tanh(g z / 2)) (1 + z^2)^k z into shared memory;
//   summaries  the 7 octiles are order statistics x_(ceil(n j / 8)); instead of sorting, a radix multi-select
//              on order-preserving 32-bit keys (11 + 8 + 8 + 5 bits) resolves all 7 ranks in 4 passes over
//              shared memory;
//   distance   sqrt(mean squared octile difference) in FP64.
// The proposal / accept logic around it is abcdesmc_swarm! (src/abcdez_smc.jl:106-153) exactly as in sweep.cuh,
// evaluated redundantly by every thread of the CTA (same Philox streams -> same decisions); thread 0 stores.
// Bound: FP32 / SFU pipes.  Normative definition: DESIGN.md "Models"; the oracle restates
// it with glibc's libm, hence the 1e-4 relative parity tolerance of this model (tests say so).
#include "sweep.cuh"

namespace abcdez {

constexpr int GK_THREADS = 256;
constexpr int GK_MAXN = 16384;
constexpr int GK_NQ = 7;

struct GkSmem {
    unsigned hist[SEL_BINS];              // pass 1: 11-bit digit
    unsigned sub[GK_NQ][256];             // passes 2-4: one 8-bit histogram per target
    unsigned part[GK_THREADS];
    unsigned prefix[GK_NQ];               // resolved high bits of each target's key
    unsigned rank[GK_NQ];                 // rank of the target inside its current bucket
    double dist;
};

__device__ __forceinline__ unsigned f32_key(float x)
{
    unsigned b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_f32(unsigned k)
{
    unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// every thread of the CTA calls; returns the distance in every thread.  xs: n floats of shared memory.
__device__ double gk_simulate_cta(const double* th, const double* data, const Stream& rs, float* xs, GkSmem* s)
{
    const int tid = threadIdx.x;
    int n = (int)data[0];
    n = n < 8 ? 8 : (n > GK_MAXN ? GK_MAXN : n);
    const float A = (float)th[0], B = (float)th[1], g = (float)th[2], k = (float)th[3];
    __syncthreads();                                   // the previous particle's readers are done with xs / s
    for (int b = tid; b * 4 < n; b += GK_THREADS) {
        float u[4];
        rs.f4((uint32_t)b, u);
        float r1 = sqrtf(-2.0f * logf(1.0f - u[0])), r2 = sqrtf(-2.0f * logf(1.0f - u[2]));
        float s1, c1, s2, c2;
        sincosf(6.2831853f * u[1], &s1, &c1);
        sincosf(6.2831853f * u[3], &s2, &c2);
        float z[4] = { r1 * c1, r1 * s1, r2 * c2, r2 * s2 };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float zz = z[j];
            float x = A + B * (1.0f + 0.8f * tanhf(0.5f * g * zz)) * powf(1.0f + zz * zz, k) * zz;
            if (b * 4 + j < n) xs[b * 4 + j] = x;
        }
    }
    for (int q = tid; q < SEL_BINS; q += GK_THREADS) s->hist[q] = 0u;
    for (int q = tid; q < GK_NQ * 256; q += GK_THREADS) (&s->sub[0][0])[q] = 0u;
    __syncthreads();
    // ---- pass 1: top 11 bits of every key ------------------------------------------------------------
    for (int i = tid; i < n; i += GK_THREADS) atomicAdd(&s->hist[f32_key(xs[i]) >> 21], 1u);
    __syncthreads();
    {   // exclusive prefix over the 2048 bins: 8 bins per thread + a scan of the 256 partials
        unsigned loc[8], tot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { loc[q] = s->hist[tid * 8 + q]; tot += loc[q]; }
        s->part[tid] = tot;
        __syncthreads();
        unsigned before = 0;
        for (int t = 0; t < tid; ++t) before += s->part[t];
        unsigned cum = before;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (loc[q]) {
#pragma unroll
                for (int j = 0; j < GK_NQ; ++j) {
                    unsigned r = (unsigned)((n * (j + 1) + 7) / 8) - 1u;        // 0-based rank of octile j+1
                    if (r >= cum && r < cum + loc[q]) { s->prefix[j] = (unsigned)(tid * 8 + q) << 21; s->rank[j] = r - cum; }
                }
            }
            cum += loc[q];
        }
    }
    __syncthreads();
    // ---- passes 2-4: 8, 8 and 5 further bits, one small histogram per target ---------------------------
    const int shifts[3] = { 13, 5, 0 }, widths[3] = { 8, 8, 5 };
    unsigned himask = 0xffe00000u;
    for (int p = 0; p < 3; ++p) {
        const int shift = shifts[p];
        const unsigned dmask = (1u << widths[p]) - 1u;
        unsigned pf[GK_NQ];
#pragma unroll
        for (int j = 0; j < GK_NQ; ++j) pf[j] = s->prefix[j];
        for (int i = tid; i < n; i += GK_THREADS) {
            unsigned key = f32_key(xs[i]), hi = key & himask;
#pragma unroll
            for (int j = 0; j < GK_NQ; ++j)
                if (hi == pf[j]) atomicAdd(&s->sub[j][(key >> shift) & dmask], 1u);
        }
        __syncthreads();
        if (tid < GK_NQ) {                              // 7 threads walk their 256-bin histograms
            unsigned r = s->rank[tid], cum = 0, nb = dmask + 1u;
            for (unsigned q = 0; q < nb; ++q) {
                unsigned cnt = s->sub[tid][q];
                if (r < cum + cnt) { s->prefix[tid] |= q << shift; s->rank[tid] = r - cum; break; }
                cum += cnt;
            }
        }
        __syncthreads();
        for (int q = tid; q < GK_NQ * 256; q += GK_THREADS) (&s->sub[0][0])[q] = 0u;
        himask |= dmask << shift;
        __syncthreads();
    }
    if (tid == 0) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < GK_NQ; ++j) {
            double dq = (double)key_f32(s->prefix[j]) - data[1 + j];
            acc += dq * dq;
        }
        s->dist = sqrt(acc / 7.0);
    }
    __syncthreads();
    return s->dist;
}

struct GK {
    static constexpr int D = 4, BLOB = 0;
    static constexpr const char* name = "gk";
};

static inline size_t gk_smem_bytes() { return sizeof(GkSmem) + (size_t)GK_MAXN * sizeof(float); }

#define GK_SMEM_DECL                                                                     \
    extern __shared__ __align__(16) unsigned char gk_raw[];                              \
    GkSmem* gs = reinterpret_cast<GkSmem*>(gk_raw);                                      \
    float* xs = reinterpret_cast<float*>(gk_raw + sizeof(GkSmem));

// ---- abcde_init! (src/abcdez_init.jl:2-22), one CTA per particle ----------------------------------------
__global__ void __launch_bounds__(GK_THREADS)
gk_init_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
               const __grid_constant__ PhiloxKeys seed, int draw_prior)
{
    constexpr int D = GK::D;
    GK_SMEM_DECL
    Ctrl* c = P.ctrl;
    const int cur = c->cur;
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);
    unsigned redraws = 0; int err = 0;
    unsigned long long kmn = ~0ull, kmx = 0ull;
    for (uint32_t i = blockIdx.x; i < P.N; i += gridDim.x) {
        const uint32_t pid = P.id0 + i;
        double th[D], x[D], lp;
        if (draw_prior) { prior_sample<D>(pr, seed, pid, 0u, th); push_p<D>(pr, th, x); lp = prior_logpdf<D>(pr, x); }
        else { load_row<D>(P.theta[cur], i, th); lp = P.logpi[cur][i]; }
        double dl = NAN;
        if (isfinite(lp)) {
            Stream r(seed, pid, 0u, TAG_INIT_MODEL);
            push_p<D>(pr, th, x);
            dl = gk_simulate_cta(x, md.v, r, xs, gs);
        }
        uint32_t attempt = 0;
        while (!isfinite(dl) || !isfinite(lp)) {           // init.jl:14-20 (uniform over the CTA)
            if (++attempt >= (uint32_t)INIT_MAX_ATTEMPTS) { err = ABCDEZ_ERR_INIT_RETRY; break; }
            prior_sample<D>(pr, seed, pid, attempt, th);
            push_p<D>(pr, th, x);
            lp = prior_logpdf<D>(pr, x);
            Stream r(seed, pid, attempt, TAG_INIT_MODEL);
            dl = gk_simulate_cta(x, md.v, r, xs, gs);
            if (threadIdx.x == 0) redraws++;
        }
        if (threadIdx.x == 0) {
            store_row<D>(P.theta[cur], i, th);
            P.logpi[cur][i] = lp; P.delta[cur][i] = dl; P.moved[i] = 1;
            unsigned long long kdl = f64_key(dl);
            kmn = kdl < kmn ? kdl : kmn; kmx = kdl > kmx ? kdl : kmx;
        }
    }
    if (threadIdx.x != 0) err = 0;
    const bool last = sweep_finish<true>(c, &s_red, redraws, 0u, kmn, kmx, err);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, true);
    if (last) {
        if (P.x.world == 1) sweep_collect(c, true);
        c->redraws += (long long)c->last_nsims;
    }
}

// ---- abcdesmc_swarm! (src/abcdez_smc.jl:106-153), one CTA per listed particle ------------------------------
__global__ void __launch_bounds__(GK_THREADS)
gk_smc_sweep_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
                    const __grid_constant__ SweepInj inj)
{
    constexpr int D = GK::D;
    GK_SMEM_DECL
    Ctrl* c = P.ctrl;
    if (c->stop | c->sweeps_done) return;
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t N = P.N, n_alive = c->n_alive;
    const double* __restrict__ th = P.theta[cur];
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);
    const bool lead = threadIdx.x == 0;
    unsigned nsim = 0, nacc = 0; int err = 0;
    for (uint32_t j = blockIdx.x; j < N; j += gridDim.x) {
        const uint32_t i = (n_alive == N) ? j : P.alive_list[j];
        const uint8_t mv = P.moved[i];
        if (j >= n_alive) {                                                    // dead particle (:114)
            if (lead) {
                if (mv) {
                    double row[D];
                    load_row<D>(th, i, row);
                    store_row<D>(P.theta[nxt], i, row);
                    copy_scalars<0>(P, cur, nxt, i, P.logpi[cur][i], P.delta[cur][i]);
                    P.moved[i] = 0;
                }
                if (inj.flags) inj.flags[i] = 0;
            }
            continue;
        }
        uint8_t flag = 0; unsigned acc_now = 0;
        const uint32_t pid = P.id0 + i;
        const PhiloxKeys& seed = P.keys;
        const uint32_t epoch = c->sweep_epoch;
        uint32_t a, b; int perr = 0;
        if (inj.a) { a = (uint32_t)inj.a[i]; b = (uint32_t)inj.b[i]; }
        else {
            Stream ps(seed, pid, epoch, TAG_PARTNER);
            double u1, u2; uint32_t att = 0;
            a = i;
            while (a == i) {                                                   // :119-122
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { perr = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                a = wsample_alive(P.alive_list, n_alive, N, u1);
            }
            att = 0; b = a;
            while (b == a || b == i) {                                         // :123-126
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { perr = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                b = wsample_alive(P.alive_list, n_alive, N, u2);
            }
        }
        double thp[D];
        load_row<D>(th, i, thp);
        const double lpi = P.logpi[cur][i], dli = P.delta[cur][i];
        if (mv && lead) { store_row<D>(P.theta[nxt], i, thp); copy_scalars<0>(P, cur, nxt, i, lpi, dli); }
        Stream ms(seed, pid, epoch, TAG_MOVE);
        double z, z2;
        if (inj.z) z = inj.z[i]; else ms.n2(0u, z, z2);
        const double g = c->gamma0 * (1.0 + z * c->gsig);                      // :128
        if (perr) { if (lead) err = perr; }
        else {
            de_proposal<D>(th, a, b, g, thp);                                  // :128
            double xsd[D];
            push_p<D>(pr, thp, xsd);
            double lp = prior_logpdf<D>(pr, xsd);                              // :134
            if (!(lp < 0.0 && isinf(lp))) {                                    // :135 (uniform over the CTA)
                Stream r(seed, pid, epoch, TAG_MODEL);
                double dp = gk_simulate_cta(xsd, md.v, r, xs, gs);             // :137
                flag |= ABCDEZ_FLAG_SIM;                                       // :138
                const double eps = c->eps; const int kind = c->kind;
                double w = lp - lpi;                                           // :140-141
                w = w + abck_logpdf(kind, eps, dp);
                w = w - abck_logpdf(kind, eps, dli);
                bool acc = (0.0 <= w);
                if (!acc) {                                                    // :145
                    double u, u2;
                    if (inj.u) u = inj.u[i]; else ms.u2(1u, u, u2);
                    acc = (plog(u) < w);
                }
                if (acc) {                                                     // :146-150
                    if (lead) { store_row<D>(P.theta[nxt], i, thp); P.logpi[nxt][i] = lp; P.delta[nxt][i] = dp; }
                    acc_now = 1; flag |= ABCDEZ_FLAG_ACC;
                }
                if (lead) { nsim += 1; nacc += acc_now; }
            }
        }
        if (lead) {
            if ((uint8_t)acc_now != mv) P.moved[i] = (uint8_t)acc_now;
            if (inj.flags) inj.flags[i] = flag;
        }
    }
    const bool last = sweep_finish<false>(c, &s_red, nsim, nacc, 0ull, 0ull, err);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, false);
    if (last) ctrl_after_smc_sweep(P, c);
}

// ---- one dist! evaluation per row (stage-level model parity) ------------------------------------------------
__global__ void __launch_bounds__(GK_THREADS)
gk_simulate_kernel(const __grid_constant__ ModelData md, int64_t N, const double* __restrict__ theta_pushed, const __grid_constant__ PhiloxKeys seed,
                   uint32_t epoch, uint32_t tag, uint32_t id0, double* __restrict__ dist)
{
    GK_SMEM_DECL
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
        double x[4];
        for (int k = 0; k < 4; ++k) x[k] = theta_pushed[i * 4 + k];
        Stream r(seed, id0 + (uint32_t)i, epoch, tag);
        double d = gk_simulate_cta(x, md.v, r, xs, gs);
        if (threadIdx.x == 0) dist[i] = d;
    }
}

static int g_gk_grid = 0;
static unsigned gk_grid(int64_t N)
{
    if (!g_gk_grid) {
        size_t sm = gk_smem_bytes();
        cudaFuncSetAttribute(gk_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute(gk_smc_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute(gk_simulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        int dev = 0, sms = 148, per = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, gk_smc_sweep_kernel, GK_THREADS, sm) != cudaSuccess || per < 1) per = 1;
        g_gk_grid = sms * per;              // persistent CTAs: a multiple of the SM count
    }
    return (unsigned)(N < g_gk_grid ? N : g_gk_grid);
}

static void gk_l_init(const ModelOps&, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, uint64_t seed, int dp)
{
    gk_init_kernel<<<gk_grid(P.N), GK_THREADS, gk_smem_bytes(), st>>>(P, pr, md, philox_keys(seed), dp);
}
static void gk_l_smc(const ModelOps&, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj)
{
    gk_smc_sweep_kernel<<<gk_grid(P.N), GK_THREADS, gk_smem_bytes(), st>>>(P, pr, md, inj);
}
static void gk_l_sim(const ModelOps&, cudaStream_t st, const PriorDev*, const ModelData& md, int64_t N, const double* th, uint64_t seed,
                     uint32_t epoch, uint32_t tag, uint32_t id0, double* dist, double*)
{
    gk_simulate_kernel<<<gk_grid(N), GK_THREADS, gk_smem_bytes(), st>>>(md, N, th, philox_keys(seed), epoch, tag, id0, dist);
}

const ModelOps* ops_gk()
{
    static const ModelOps o = { GK::name, GK::D, GK::BLOB, &gk_l_init, &gk_l_smc, nullptr /* abcdemc!: not for this model */, &gk_l_sim, nullptr };
    return &o;
}

}  // namespace abcdez
