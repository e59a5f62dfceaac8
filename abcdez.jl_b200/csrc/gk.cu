// gk.cu -- config 3 of BASELINE.json: the g-and-k distribution, 4 parameters, quantile summaries of n
// (<= 16384, default 10^4) simulated draws.  One dist! evaluation is ~10^6 arithmetic operations, so a
// particle is simulated by a whole CTA instead of one thread:
//   draws      thread t generates the Philox blocks t, t + 256, ... (one Box-Muller pair each in FP64, two in
//              FP32; two blocks in flight per thread) into shared memory;
//   summaries  the 7 octiles are order statistics x_(ceil(n j / 8)) of
//                  x = Q(z) = A + B (1 + 0.8 (1 - e^{-g z}) / (1 + e^{-g z})) (1 + z^2)^k z.
//              Fast path (n >= 4096, B > 0, k >= 0: Q increasing): select among the z -- per-octile histograms
//              around Phi^-1(j/8), filled while the normals are drawn -- and push only the ~6 candidates per
//              octile through Q ("the fast path" below).  Generic path (everything else, and the fallback): all
//              draws through Q, then a multi-select in KEY space -- one 2048-bin histogram between the extrema
//              resolves all 7 ranks to a bucket each, buckets that are still large are refined 256 ways, and the
//              <= 64 keys left per octile are ranked directly by one warp.  Both return the exact order statistics;
//   distance   sqrt(mean squared octile difference) in FP64.
// Two registered models, same definition, different arithmetic:
//   "gk"      FP64 with the library's portable log / exp / sin / cos (common.cuh), i.e. the arithmetic a Julia
//             dist! would use (Float64).  The oracle restates it operation by operation: distances, accept
//             decisions and whole runs are bit-identical.
//   "gk_f32"  relaxed-precision mode: the draws in FP32 with portable FP32 log / exp /
//             sin / cos defined below (again restated in the oracle -> bit-identical to *its* oracle), twice the
//             lanes and shorter polynomials; its posterior agrees with "gk" within Monte-Carlo error.
// The proposal / accept logic around the simulation is abcdesmc_swarm! (src/abcdez_smc.jl:106-153) and
// abcdemc_swarm! (src/abcdez_mc.jl:5-61) exactly as in sweep.cuh, evaluated redundantly by every thread of the
// CTA (same Philox streams -> same decisions); thread 0 stores.  Bound: instruction issue on the Box-Muller / Philox
// chains (FP64 pipe 20 % busy), DESIGN.md section 5.
#include "sweep.cuh"

namespace abcdez {

#ifndef ABCDEZ_GK_THREADS
#define ABCDEZ_GK_THREADS 256
#endif
constexpr int GK_THREADS = ABCDEZ_GK_THREADS;     // threads of the CTA that simulates one particle
constexpr int GK_BPT = SEL_BINS / GK_THREADS;      // histogram bins per thread in the prefix step
constexpr int GK_MAXN = 16384;
#ifndef ABCDEZ_GK_ILP
#define ABCDEZ_GK_ILP 2
#endif
constexpr int GK_ILP = ABCDEZ_GK_ILP;              // Philox blocks (Box-Muller chains) in flight per thread in the fast path
constexpr int GK_NQ = 7;
constexpr int GK_LIST = 64;               // keys per octile ranked directly

struct GkSmem {
    unsigned hist[SEL_BINS];              // round 1: 2048 bins between the extrema
    unsigned sub[GK_NQ][256];             // refinement rounds: one 256-bin histogram per octile
    unsigned long long list[GK_NQ][GK_LIST];
    unsigned long long lo[GK_NQ], hi[GK_NQ], q[GK_NQ];
    unsigned rank[GK_NQ], cnt[GK_NQ], nlist[GK_NQ];
    unsigned part[GK_THREADS];
    unsigned long long kmin, kmax;
    float zl[8], zh[8], invw[8];          // fast path: the z windows of the 7 octiles
    unsigned region[8], tbin[GK_NQ], base[GK_NQ], miss[4];
    double dist;
};

// ---- portable FP32 math of the relaxed-precision mode (restated in oracle/abcdez_oracle.c) ------------------
// IEEE single operations only (the library is built -fmad=false; sqrtf and / are correctly rounded), explicit
// fmaf in the Horner steps.
__device__ __forceinline__ float gk_logf_pos(float x)          // x positive and normal
{
    const unsigned b = __float_as_uint(x);
    int e = (int)((b >> 23) & 0xffu) - 127;
    float m = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
    if (m > 0x1.6a09e6p+0f) { m = m * 0.5f; e += 1; }                   // sqrt(2)
    const float f = m - 1.0f;
    const float s = f / (2.0f + f);
    const float z = s * s;
    float p = 2.0f / 9.0f;
    p = fmaf(p, z, 2.0f / 7.0f);
    p = fmaf(p, z, 2.0f / 5.0f);
    p = fmaf(p, z, 2.0f / 3.0f);
    const float r = (s * z) * p;
    const float lm = 2.0f * s + r;
    return ((float)e * 0x1.62e4p-1f + lm) + (float)e * 0x1.7f7d1cp-20f;  // ln 2 = hi + lo
}
__device__ __forceinline__ float gk_expf(float x)
{
    if (x > 88.0f) return INFINITY;
    if (x < -87.0f) return 0.0f;
    const float k = floorf(x * 0x1.715476p+0f + 0.5f);                  // 1 / ln 2
    const float r = (x - k * 0x1.62e4p-1f) - k * 0x1.7f7d1cp-20f;
    float p = 1.0f / 5040.0f;
    p = fmaf(p, r, 1.0f / 720.0f);
    p = fmaf(p, r, 1.0f / 120.0f);
    p = fmaf(p, r, 1.0f / 24.0f);
    p = fmaf(p, r, 1.0f / 6.0f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    return p * __uint_as_float((unsigned)((int)k + 127) << 23);         // 2^k, k in [-126, 127]
}
__device__ __forceinline__ void gk_sincos2pif(float u, float* sn, float* cs)
{
    const float q = floorf(4.0f * u + 0.5f);
    const float r = u - 0.25f * q;
    const float x = r * 0x1.921fb6p+2f;                                 // 2 pi
    const float x2 = x * x;
    float ps = -1.0f / 39916800.0f;
    ps = fmaf(ps, x2, 1.0f / 362880.0f);
    ps = fmaf(ps, x2, -1.0f / 5040.0f);
    ps = fmaf(ps, x2, 1.0f / 120.0f);
    ps = fmaf(ps, x2, -1.0f / 6.0f);
    float pc = -1.0f / 3628800.0f;
    pc = fmaf(pc, x2, 1.0f / 40320.0f);
    pc = fmaf(pc, x2, -1.0f / 720.0f);
    pc = fmaf(pc, x2, 1.0f / 24.0f);
    pc = fmaf(pc, x2, -0.5f);
    const float s = x + x * (x2 * ps);
    const float c = 1.0f + x2 * pc;
    const int k = (int)q & 3;
    const bool swap = (k & 1) != 0;
    const float a = swap ? c : s, b = swap ? s : c;
    *sn = (k & 2) ? -a : a;
    *cs = ((k + 1) & 2) ? -b : b;
}

// ---- the simulator's arithmetic, per precision -------------------------------------------------------------
template <class T> struct GkMath;

template <> struct GkMath<double> {
    static constexpr int PER_BLOCK = 2;           // draws per Philox block
    static constexpr int MIN_CTAS = 2;            // resident CTAs per SM the register allocation must allow (= what shared memory allows at n = 10^4)
    static constexpr const char* name = "gk";
    __device__ static __forceinline__ void normals(const Stream& rs, uint32_t b, double* z) { rs.n2(b, z[0], z[1]); }
    __device__ static __forceinline__ double transform(double A, double B, double g, double k, double z)
    {
        const double gz = g * z;
        const double e = pexp(-fabs(gz));                       // in (0, 1]
        double t = pm_div_inrange(1.0 - e, 1.0 + e);            // == (1 - e) / (1 + e); tanh(|g z| / 2)
        t = gz < 0.0 ? -t : t;
        const double c = 1.0 + 0.8 * t;
        const double l = plog_unit(1.0 + z * z);                // == plog: the argument is >= 1
        const double p = pexp(k * l);                           // (1 + z^2)^k
        return A + ((B * c) * p) * z;
    }
    __device__ static __forceinline__ unsigned long long key(double x) { return f64_key(x); }
    __device__ static __forceinline__ double value(unsigned long long k) { return key_f64(k); }
};

template <> struct GkMath<float> {
    static constexpr int PER_BLOCK = 4;
    static constexpr int MIN_CTAS = 3;
    static constexpr const char* name = "gk_f32";
    __device__ static __forceinline__ void normals(const Stream& rs, uint32_t b, float* z)
    {
        float u[4];
        rs.f4(b, u);
        const float r1 = sqrtf(-2.0f * gk_logf_pos(1.0f - u[0])), r2 = sqrtf(-2.0f * gk_logf_pos(1.0f - u[2]));   // 1 - u in [2^-24, 1]
        float s1, c1, s2, c2;
        gk_sincos2pif(u[1], &s1, &c1);
        gk_sincos2pif(u[3], &s2, &c2);
        z[0] = r1 * c1; z[1] = r1 * s1; z[2] = r2 * c2; z[3] = r2 * s2;
    }
    __device__ static __forceinline__ float transform(float A, float B, float g, float k, float z)
    {
        const float gz = g * z;
        const float e = gk_expf(-fabsf(gz));
        float t = (1.0f - e) / (1.0f + e);
        t = gz < 0.0f ? -t : t;
        const float c = 1.0f + 0.8f * t;
        const float l = gk_logf_pos(1.0f + z * z);
        const float p = gk_expf(k * l);
        return A + ((B * c) * p) * z;
    }
    __device__ static __forceinline__ unsigned long long key(float x)
    {
        unsigned b = __float_as_uint(x);
        return (unsigned long long)((b & 0x80000000u) ? ~b : (b | 0x80000000u));
    }
    __device__ static __forceinline__ double value(unsigned long long k)
    {
        unsigned kk = (unsigned)k;
        unsigned b = (kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk;
        return (double)__uint_as_float(b);
    }
};

// ---- the generic path: multi-select of the 7 octile keys over the transformed draws in xs --------------------
// Leaves the selected keys in s->q (visible to every thread after the trailing barrier).
template <class T>
__device__ __noinline__ void gk_select_generic(int n, const T* xs, GkSmem* s)
{
    using MT = GkMath<T>;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    __syncthreads();
    for (int q = tid; q < SEL_BINS; q += GK_THREADS) s->hist[q] = 0u;
    for (int q = tid; q < GK_NQ * 256; q += GK_THREADS) (&s->sub[0][0])[q] = 0u;
    if (tid == 0) { s->kmin = ~0ull; s->kmax = 0ull; }
    if (tid < GK_NQ) s->nlist[tid] = 0u;
    __syncthreads();
    {
        // ---- round 1: 2048 bins of 2^sh keys between the sample's extrema --------------------------------------
        unsigned long long kmn = ~0ull, kmx = 0ull;
        for (int i = tid; i < n; i += GK_THREADS) { const unsigned long long key = MT::key(xs[i]); kmn = key < kmn ? key : kmn; kmx = key > kmx ? key : kmx; }
        kmn = warp_min_u64(kmn); kmx = warp_max_u64(kmx);
        if (lane == 0) { atomicMin(&s->kmin, kmn); atomicMax(&s->kmax, kmx); }
        __syncthreads();
        const unsigned long long kmin = s->kmin, kmax = s->kmax, width = kmax - kmin;
        const int sh = width ? max(0, 64 - __clzll((long long)width) - 11) : 0;       // (width >> sh) <= 2047
        for (int i = tid; i < n; i += GK_THREADS) atomicAdd(&s->hist[(unsigned)((MT::key(xs[i]) - kmin) >> sh)], 1u);
        __syncthreads();
        unsigned loc[GK_BPT], tot = 0;
#pragma unroll
        for (int q = 0; q < GK_BPT; ++q) { loc[q] = s->hist[tid * GK_BPT + q]; tot += loc[q]; }
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s->part[wp] = incl;
        __syncthreads();
        unsigned cum = incl - tot;
        for (int w = 0; w < wp; ++w) cum += s->part[w];
#pragma unroll
        for (int q = 0; q < GK_BPT; ++q) {
            if (loc[q]) {
#pragma unroll
                for (int j = 0; j < GK_NQ; ++j) {
                    const unsigned r = (unsigned)((n * (j + 1) + 7) / 8) - 1u;        // 0-based rank of octile j+1
                    if (r >= cum && r < cum + loc[q]) {
                        const unsigned long long lo = kmin + ((unsigned long long)(tid * GK_BPT + q) << sh);
                        const unsigned long long top = lo + ((1ull << sh) - 1ull);
                        s->lo[j] = lo; s->hi[j] = top < kmax ? top : kmax;
                        s->rank[j] = r - cum; s->cnt[j] = loc[q];
                    }
                }
            }
            cum += loc[q];
        }
        __syncthreads();
    }
    // ---- refinement: octiles whose bucket still holds more than GK_LIST keys, 256 ways per round -------------
    for (int it = 0; it < 10; ++it) {
        unsigned long long lo[GK_NQ], hi[GK_NQ];
        unsigned mask = 0;
#pragma unroll
        for (int j = 0; j < GK_NQ; ++j) {
            lo[j] = s->lo[j]; hi[j] = s->hi[j];
            if (s->cnt[j] > (unsigned)GK_LIST && lo[j] < hi[j]) mask |= 1u << j;
        }
        if (!mask) break;                                   // (uniform: read from shared memory)
        for (int i = tid; i < n; i += GK_THREADS) {
            const unsigned long long key = MT::key(xs[i]);
#pragma unroll
            for (int j = 0; j < GK_NQ; ++j) {
                if (((mask >> j) & 1u) && key >= lo[j] && key <= hi[j]) {
                    const int shj = max(0, 64 - __clzll((long long)(hi[j] - lo[j])) - 8);
                    atomicAdd(&s->sub[j][(unsigned)((key - lo[j]) >> shj)], 1u);
                }
            }
        }
        __syncthreads();
        if (tid < GK_NQ && ((mask >> tid) & 1u)) {           // 7 threads walk their 256-bin histograms
            const int shj = max(0, 64 - __clzll((long long)(hi[tid] - lo[tid])) - 8);
            unsigned r = s->rank[tid], cum = 0;
            for (unsigned q = 0; q < 256u; ++q) {
                const unsigned cnt = s->sub[tid][q];
                if (r < cum + cnt) {
                    const unsigned long long nlo = lo[tid] + ((unsigned long long)q << shj);
                    const unsigned long long top = nlo + ((1ull << shj) - 1ull);
                    s->lo[tid] = nlo; s->hi[tid] = top < hi[tid] ? top : hi[tid];
                    s->rank[tid] = r - cum; s->cnt[tid] = cnt;
                    break;
                }
                cum += cnt;
            }
        }
        __syncthreads();
        for (int q = tid; q < GK_NQ * 256; q += GK_THREADS) (&s->sub[0][0])[q] = 0u;
        __syncthreads();
    }
    // ---- gather the <= GK_LIST keys of every octile's bucket and rank them directly ----------------------------
    {
        unsigned long long lo[GK_NQ], hi[GK_NQ];
        unsigned small = 0;
#pragma unroll
        for (int j = 0; j < GK_NQ; ++j) { lo[j] = s->lo[j]; hi[j] = s->hi[j]; if (s->cnt[j] <= (unsigned)GK_LIST) small |= 1u << j; }
        for (int i = tid; i < n; i += GK_THREADS) {
            const unsigned long long key = MT::key(xs[i]);
#pragma unroll
            for (int j = 0; j < GK_NQ; ++j) {
                if (((small >> j) & 1u) && key >= lo[j] && key <= hi[j]) {
                    const unsigned idx = atomicAdd(&s->nlist[j], 1u);
                    if (idx < (unsigned)GK_LIST) s->list[j][idx] = key;
                }
            }
        }
    }
    __syncthreads();
    if (wp < GK_NQ) {                                       // warp j ranks octile j's keys: two per lane
        const int j = wp;
        if (s->cnt[j] > (unsigned)GK_LIST) { if (lane == 0) s->q[j] = s->lo[j]; }     // one key value fills the bucket
        else {
            const unsigned m = s->nlist[j] < (unsigned)GK_LIST ? s->nlist[j] : (unsigned)GK_LIST, r = s->rank[j];
#pragma unroll
            for (int h = 0; h < GK_LIST / 32; ++h) {
                const unsigned me = lane + 32 * h;
                if (me < m) {
                    const unsigned long long my = s->list[j][me];
                    unsigned below = 0;
                    for (unsigned o = 0; o < m; ++o) { const unsigned long long ot = s->list[j][o]; below += (ot < my || (ot == my && o < me)) ? 1u : 0u; }
                    if (below == r) s->q[j] = my;
                }
            }
        }
    }
    __syncthreads();
}

// ---- the fast path: select in z space ----------------------------------------------------------------------
// For B > 0 and k >= 0 the g-and-k quantile function x = Q(z) is increasing in z, so the order statistics
// commute with it: x_(r) = Q(z_(r)).  The octiles of n standard normals are known in advance up to sampling
// error -- z_(ceil(n p)) lies within a few sqrt(p (1 - p) / n) / phi(z_p) of z_p = Phi^-1(p) -- so each octile
// gets its own 256-bin histogram over z_p +- GK_WIN_SIGMAS standard errors, filled while the normals are drawn;
// the bucket of the wanted rank and its two neighbours (about 6 draws) are the only draws pushed through Q, and
// they are ranked by their x keys, so floating-point wobble of Q between neighbouring draws cannot change the
// answer (it would have to exceed a bucket, 8e-4 in z).  Result: the same order statistics, bit for bit, as
// transforming and selecting over all n draws (what the oracle does), at ~1/3 of the arithmetic.  Anything
// unexpected -- a rank outside its window or in an edge bucket, more than GK_LIST candidates -- falls back to the
// generic path below, which is also the path for n < GK_FAST_N and for parameters where Q need not be monotone.
constexpr int GK_FAST_N = 4096;                   // below this the windows of neighbouring octiles would touch
constexpr float GK_WIN_SIGMAS = 6.0f;
__constant__ float GK_ZP[GK_NQ] = { -1.15034938f, -0.67448975f, -0.31863936f, 0.0f, 0.31863936f, 0.67448975f, 1.15034938f };
__constant__ float GK_SE[GK_NQ] = { 1.60657f, 1.36263f, 1.27671f, 1.25331f, 1.27671f, 1.36263f, 1.60657f };   // sqrt(p (1 - p)) / phi(z_p)

template <class T> __device__ __forceinline__ bool gk_finite(T v) { return v == v && fabs((double)v) < (double)INFINITY; }

// region of a draw: cnt = number of window lower bounds <= z (0..7), by a 3-step binary search over the ascending
// bounds; window cnt - 1 holds the draw when z <= its upper bound
__device__ __forceinline__ unsigned gk_region(float zf, const float (&zl)[GK_NQ])
{
    const bool h4 = zf >= zl[3];
    const bool h2 = zf >= (h4 ? zl[5] : zl[1]);
    const float t = h4 ? (h2 ? zl[6] : zl[4]) : (h2 ? zl[2] : zl[0]);
    return (h4 ? 4u : 0u) + (h2 ? 2u : 0u) + (zf >= t ? 1u : 0u);
}
__device__ __forceinline__ int gk_bin(float zf, float zl, float invw)
{
    const int b = (int)((zf - zl) * invw);                    // monotone in zf; the same expression in both passes
    return b > 255 ? 255 : b;
}

// every thread of the CTA calls; returns the distance in every thread.  xs: n values of shared memory.
template <class T>
__device__ double gk_simulate_cta(const double* th, const double* data, const Stream& rs, T* xs, GkSmem* s)
{
    using MT = GkMath<T>;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    int n = (int)data[0];
    n = n < 8 ? 8 : (n > GK_MAXN ? GK_MAXN : n);
    const T A = (T)th[0], B = (T)th[1], g = (T)th[2], k = (T)th[3];
    const bool fast = n >= GK_FAST_N && gk_finite(A) && gk_finite(g) && B > (T)0 && gk_finite(B) && k >= (T)0 && gk_finite(k);
    __syncthreads();                                   // the previous particle's readers are done with xs / s
    bool done = false;
    if (fast) {
        for (int q = tid; q < GK_NQ * 256; q += GK_THREADS) (&s->sub[0][0])[q] = 0u;
        if (tid < 8) s->region[tid] = 0u;
        if (tid < 3) s->miss[tid] = 0u;
        if (tid < GK_NQ) {
            const float h = GK_WIN_SIGMAS * GK_SE[tid] / sqrtf((float)n);
            s->zl[tid] = GK_ZP[tid] - h; s->zh[tid] = GK_ZP[tid] + h; s->invw[tid] = 128.0f / h;
            s->nlist[tid] = 0u;
        }
        __syncthreads();
        float zl[GK_NQ];
#pragma unroll
        for (int j = 0; j < GK_NQ; ++j) zl[j] = s->zl[j];
        // ---- draws: the normals go to shared memory, window hits into their histograms ------------------------
        unsigned long long pk = 0ull;                          // 8 region counters of 8 bits (<= n / 256 <= 64 draws per thread)
        for (int b0 = tid; b0 * MT::PER_BLOCK < n; b0 += GK_THREADS * GK_ILP) {
            T z[GK_ILP][MT::PER_BLOCK];
#pragma unroll
            for (int u = 0; u < GK_ILP; ++u) MT::normals(rs, (uint32_t)(b0 + u * GK_THREADS), z[u]);   // independent chains; a block past n is dropped below
#pragma unroll
            for (int u = 0; u < GK_ILP; ++u) {
#pragma unroll
                for (int j = 0; j < MT::PER_BLOCK; ++j) {
                    const int i = (b0 + u * GK_THREADS) * MT::PER_BLOCK + j;
                    if (i < n) {
                        xs[i] = z[u][j];
                        const float zf = (float)z[u][j];
                        const unsigned cnt = gk_region(zf, zl);
                        pk += 1ull << (8u * cnt);
                        if (cnt) {
                            const unsigned w = cnt - 1u;
                            if (zf <= s->zh[w]) atomicAdd(&s->sub[w][gk_bin(zf, s->zl[w], s->invw[w])], 1u);
                        }
                    }
                }
            }
        }
        {   // region totals: 16-bit fields, 32 lanes x 64 draws fit
            unsigned long long ev = pk & 0x00ff00ff00ff00ffull, od = (pk >> 8) & 0x00ff00ff00ff00ffull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { ev += __shfl_xor_sync(0xffffffffu, ev, o); od += __shfl_xor_sync(0xffffffffu, od, o); }
            if (lane < 8) {
                const unsigned long long src = (lane & 1) ? od : ev;
                const unsigned v = (unsigned)((src >> (16 * (lane >> 1))) & 0xffffull);
                if (v) atomicAdd(&s->region[lane], v);
            }
        }
        __syncthreads();
        // ---- warp j: the bucket of octile j's rank ----------------------------------------------------------
        if (wp < GK_NQ) {
            const int j = wp;
            unsigned below = 0;
            for (int m = 0; m <= j; ++m) below += s->region[m];                       // draws left of window j
            const unsigned r = (unsigned)((n * (j + 1) + 7) / 8) - 1u;                // 0-based rank of octile j+1
            unsigned loc[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = s->sub[j][lane * 8 + q]; tot += loc[q]; }
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const unsigned inwin = __shfl_sync(0xffffffffu, incl, 31);
            if (r < below || r - below >= inwin) { if (lane == 0) s->miss[0] = 1u; }
            else {
                const unsigned rw = r - below;
                unsigned cum = incl - tot;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (rw >= cum && rw < cum + loc[q]) {
                        const int bin = lane * 8 + q;
                        if (bin == 0 || bin == 255) s->miss[0] = 1u;
                        else { s->tbin[j] = (unsigned)bin; s->base[j] = below + cum - s->sub[j][bin - 1]; s->rank[j] = r; }
                    }
                    cum += loc[q];
                }
            }
        }
        __syncthreads();
        // ---- candidates: the draws of that bucket and its neighbours -----------------------------------------
        if (!s->miss[0]) {
            for (int i = tid; i < n; i += GK_THREADS) {
                const float zf = (float)xs[i];
                const unsigned cnt = gk_region(zf, zl);
                if (cnt) {
                    const unsigned w = cnt - 1u;
                    if (zf <= s->zh[w]) {
                        const int d = gk_bin(zf, s->zl[w], s->invw[w]) - (int)s->tbin[w];
                        if (d >= -1 && d <= 1) {
                            const unsigned idx = atomicAdd(&s->nlist[w], 1u);
                            if (idx < (unsigned)GK_LIST) s->list[w][idx] = (unsigned long long)i;
                            else s->miss[1] = 1u;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- warp j pushes octile j's candidates through Q and ranks them by x --------------------------------
        if (wp < GK_NQ && !s->miss[0] && !s->miss[1]) {
            const int j = wp;
            const unsigned m = s->nlist[j], r = s->rank[j], base = s->base[j];
            unsigned long long my[GK_LIST / 32];
#pragma unroll
            for (int h = 0; h < GK_LIST / 32; ++h) {
                const unsigned me = lane + 32 * h;
                my[h] = me < m ? MT::key(MT::transform(A, B, g, k, xs[(int)s->list[j][me]])) : ~0ull;
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < GK_LIST / 32; ++h) { const unsigned me = lane + 32 * h; if (me < m) s->list[j][me] = my[h]; }
            __syncwarp();
            if (r < base || r - base >= m) { if (lane == 0) s->miss[2] = 1u; }
            else {
#pragma unroll
                for (int h = 0; h < GK_LIST / 32; ++h) {
                    const unsigned me = lane + 32 * h;
                    if (me < m) {
                        unsigned before = 0;
                        for (unsigned o = 0; o < m; ++o) { const unsigned long long ot = s->list[j][o]; before += (ot < my[h] || (ot == my[h] && o < me)) ? 1u : 0u; }
                        if (before == r - base) s->q[j] = my[h];
                    }
                }
            }
        }
        __syncthreads();
        done = !(s->miss[0] | s->miss[1] | s->miss[2]);
        if (!done) for (int i = tid; i < n; i += GK_THREADS) xs[i] = MT::transform(A, B, g, k, xs[i]);   // -> generic path
    } else {
        // ---- draws, all pushed through Q ------------------------------------------------------------------------
        for (int b = tid; b * MT::PER_BLOCK < n; b += GK_THREADS) {
            T z[MT::PER_BLOCK];
            MT::normals(rs, (uint32_t)b, z);
#pragma unroll
            for (int j = 0; j < MT::PER_BLOCK; ++j) {
                const T x = MT::transform(A, B, g, k, z[j]);
                const int i = b * MT::PER_BLOCK + j;
                if (i < n) xs[i] = x;
            }
        }
    }
    if (!done) gk_select_generic<T>(n, xs, s);
    if (tid == 0) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < GK_NQ; ++j) {
            const double dq = MT::value(s->q[j]) - data[1 + j];
            acc += dq * dq;
        }
        s->dist = sqrt(acc / 7.0);
    }
    __syncthreads();
    return s->dist;
}

template <class T> static inline size_t gk_smem_bytes(const ModelData& md)
{
    int n = (int)md.v[0];
    n = n < 8 ? 8 : (n > GK_MAXN ? GK_MAXN : n);
    return sizeof(GkSmem) + (((size_t)n * sizeof(T) + 15) & ~(size_t)15);
}

#define GK_SMEM_DECL                                                                     \
    extern __shared__ __align__(16) unsigned char gk_raw[];                              \
    GkSmem* gs = reinterpret_cast<GkSmem*>(gk_raw);                                      \
    T* xs = reinterpret_cast<T*>(gk_raw + sizeof(GkSmem));

// ---- abcde_init! (src/abcdez_init.jl:2-22), one CTA per particle ----------------------------------------
template <class T>
__global__ void __launch_bounds__(GK_THREADS, GkMath<T>::MIN_CTAS)
gk_init_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
               const __grid_constant__ PhiloxKeys seed, int draw_prior)
{
    constexpr int D = 4;
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
            dl = gk_simulate_cta<T>(x, md.v, r, xs, gs);
        }
        uint32_t attempt = 0;
        while (!isfinite(dl) || !isfinite(lp)) {           // init.jl:14-20 (uniform over the CTA)
            if (++attempt >= (uint32_t)INIT_MAX_ATTEMPTS) { err = ABCDEZ_ERR_INIT_RETRY; break; }
            prior_sample<D>(pr, seed, pid, attempt, th);
            push_p<D>(pr, th, x);
            lp = prior_logpdf<D>(pr, x);
            Stream r(seed, pid, attempt, TAG_INIT_MODEL);
            dl = gk_simulate_cta<T>(x, md.v, r, xs, gs);
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
template <class T>
__global__ void __launch_bounds__(GK_THREADS, GkMath<T>::MIN_CTAS)
gk_smc_sweep_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
                    const __grid_constant__ SweepInj inj)
{
    constexpr int D = 4;
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
                double dp = gk_simulate_cta<T>(xsd, md.v, r, xs, gs);          // :137
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

// ---- abcdemc_swarm! (src/abcdez_mc.jl:5-61), one CTA per particle -------------------------------------------
template <class T>
__global__ void __launch_bounds__(GK_THREADS, GkMath<T>::MIN_CTAS)
gk_mc_sweep_kernel(const __grid_constant__ PopDev P, const __grid_constant__ PriorDev pr, const __grid_constant__ ModelData md,
                   const __grid_constant__ SweepInj inj, const __grid_constant__ McArgs mc)
{
    constexpr int D = 4;
    GK_SMEM_DECL
    Ctrl* c = P.ctrl;
    if (mc.from_ctrl && (c->stop | c->err)) return;
    const int cur = c->cur, nxt = cur ^ 1;
    const uint32_t N = P.N;
    const double* __restrict__ th = P.theta[cur];
    __shared__ SweepSmem s_red;
    sweep_smem_init(&s_red);
    const bool lead = threadIdx.x == 0;
    const double eps_target = mc.from_ctrl ? c->eps_target : mc.eps_target;
    const double eps_pop = mc.from_ctrl ? fmax(eps_target, c->dmin + 0.0 * (c->dmax - c->dmin)) : mc.eps_pop;   // mc.jl:146-147
    unsigned nsim = 0, nacc = 0; int err = 0;
    unsigned long long kmn = ~0ull, kmx = 0ull;
    for (uint32_t i = blockIdx.x; i < N; i += gridDim.x) {
        double thp[D];
        load_row<D>(th, i, thp);
        const double lpi = P.logpi[cur][i];
        double dli = P.delta[cur][i];
        const uint8_t mv = P.moved[i];
        if (mv && lead) { store_row<D>(P.theta[nxt], i, thp); copy_scalars<0>(P, cur, nxt, i, lpi, dli); }
        uint8_t flag = 0; unsigned acc_now = 0;
        const uint32_t pid = P.id0 + i;
        const PhiloxKeys& seed = P.keys;
        const uint32_t epoch = c->sweep_epoch;
        uint32_t s = i;                                                        // :18
        const double eps = (dli <= eps_target) ? eps_target : eps_pop;         // :19
        if (dli > eps) {                                                       // :20-24
            if (inj.s) s = (uint32_t)inj.s[i];
            else {
                uint32_t lo = 0, hi = N;
                while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (mc.sorted_delta[m] <= dli) lo = m + 1; else hi = m; }
                Stream cs(seed, pid, epoch, TAG_MC);
                double u1, u2; cs.u2(0u, u1, u2);
                long long k = (long long)floor(u1 * (double)lo);
                if (k >= (long long)lo) k = (long long)lo - 1;
                s = mc.order[k];
            }
        }
        uint32_t a, b; int perr = 0;
        if (inj.a) { a = (uint32_t)inj.a[i]; b = (uint32_t)inj.b[i]; }
        else {
            Stream ps(seed, pid, epoch, TAG_PARTNER);
            double u1, u2, ua, ub; uint32_t att = 1;
            auto pick = [N](double u) { long long k = (long long)floor(u * (double)N); if (k >= (long long)N) k = (long long)N - 1; return (uint32_t)k; };
            ps.u2(0u, ua, ub);
            a = pick(ua);
            while (a == s) {                                                   // :25-28
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { perr = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                a = pick(u1);
            }
            att = 1;
            b = pick(ub);
            while (b == a || b == s) {                                         // :29-32
                if (att >= (uint32_t)PARTNER_MAX_ATTEMPTS) { perr = ABCDEZ_ERR_PARTNER_RETRY; break; }
                ps.u2(att++, u1, u2);
                b = pick(u2);
            }
        }
        if (perr) { if (lead) err = perr; }
        else {
            Stream ms(seed, pid, epoch, TAG_MOVE);
            if (s != i) load_row<D>(th, s, thp);                               // base particle theta_s
            double z, z2;
            if (inj.z) z = inj.z[i]; else ms.n2(0u, z, z2);
            const double g = c->gamma0 * (1.0 + z * c->gsig);                  // :34
            de_proposal<D>(th, a, b, g, thp);
            double xsd[D];
            push_p<D>(pr, thp, xsd);
            const double lp = prior_logpdf<D>(pr, xsd);                        // :41
            const double w_prior = lp - lpi;                                   // :42 (logpi[i], not [s])
            double u, u2;
            if (inj.u) u = inj.u[i]; else ms.u2(1u, u, u2);                    // :43, always drawn
            if (!(plog(u) > fmin(0.0, w_prior))) {                             // (uniform over the CTA)
                flag |= ABCDEZ_FLAG_SIM;                                       // :44
                Stream r(seed, pid, epoch, TAG_MODEL);
                const double dp = gk_simulate_cta<T>(xsd, md.v, r, xs, gs);    // :45
                if (dp <= fmax(eps, dli)) {                                    // :54-59
                    if (lead) { store_row<D>(P.theta[nxt], i, thp); P.logpi[nxt][i] = lp; P.delta[nxt][i] = dp; }
                    dli = dp;
                    acc_now = 1; flag |= ABCDEZ_FLAG_ACC;
                }
                if (lead) { nsim += 1; nacc += acc_now; }
            }
        }
        if (lead) {
            if ((uint8_t)acc_now != mv) P.moved[i] = (uint8_t)acc_now;
            if (inj.flags) inj.flags[i] = flag;
            const unsigned long long kdl = f64_key(dli);                       // extrema(delta), src/abcdez_mc.jl:146
            kmn = kdl < kmn ? kdl : kmn; kmx = kdl > kmx ? kdl : kmx;
        }
    }
    const bool last = sweep_finish<true>(c, &s_red, nsim, nacc, kmn, kmx, err);
    if (P.x.world > 1 && __shfl_sync(0xffffffffu, (int)last, 0)) sweep_exchange_warp(P, c, true);
    if (last) ctrl_after_mc_sweep(P, c);
}

// ---- one dist! evaluation per row (stage-level model parity) ------------------------------------------------
template <class T>
__global__ void __launch_bounds__(GK_THREADS, GkMath<T>::MIN_CTAS)
gk_simulate_kernel(const __grid_constant__ ModelData md, int64_t N, const double* __restrict__ theta_pushed, const __grid_constant__ PhiloxKeys seed,
                   uint32_t epoch, uint32_t tag, uint32_t id0, double* __restrict__ dist)
{
    GK_SMEM_DECL
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
        double x[4];
        for (int k = 0; k < 4; ++k) x[k] = theta_pushed[i * 4 + k];
        Stream r(seed, id0 + (uint32_t)i, epoch, tag);
        double d = gk_simulate_cta<T>(x, md.v, r, xs, gs);
        if (threadIdx.x == 0) dist[i] = d;
    }
}

// persistent grid: a multiple of the SM count (SMs x resident CTAs for this n)
template <class T>
static unsigned gk_grid(int64_t N, size_t smem)
{
    // per device: function attributes and occupancy belong to the device's context (abcdez_init_multi drives several)
    static size_t cached_smem[64] = { 0 }; static int cached_grid[64] = { 0 };
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    if (cached_smem[slot] != smem || !cached_grid[slot]) {
        const int mx = (int)(sizeof(GkSmem) + GK_MAXN * sizeof(T));
        cudaFuncSetAttribute(gk_init_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(gk_smc_sweep_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(gk_mc_sweep_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(gk_simulate_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        int sms = 148, per = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, gk_smc_sweep_kernel<T>, GK_THREADS, smem) != cudaSuccess || per < 1) per = 1;
        cached_grid[slot] = sms * per; cached_smem[slot] = smem;
    }
    return (unsigned)(N < cached_grid[slot] ? N : cached_grid[slot]);
}

template <class T>
static void gk_l_init(const ModelOps&, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, uint64_t seed, int dp)
{
    const size_t sm = gk_smem_bytes<T>(md);
    gk_init_kernel<T><<<gk_grid<T>(P.N, sm), GK_THREADS, sm, st>>>(P, pr, md, philox_keys(seed), dp);
}
template <class T>
static void gk_l_smc(const ModelOps&, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj)
{
    const size_t sm = gk_smem_bytes<T>(md);
    gk_smc_sweep_kernel<T><<<gk_grid<T>(P.N, sm), GK_THREADS, sm, st>>>(P, pr, md, inj);
}
template <class T>
static void gk_l_mc(const ModelOps&, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj,
                    const McArgs& mc)
{
    const size_t sm = gk_smem_bytes<T>(md);
    gk_mc_sweep_kernel<T><<<gk_grid<T>(P.N, sm), GK_THREADS, sm, st>>>(P, pr, md, inj, mc);
}
template <class T>
static void gk_l_sim(const ModelOps&, cudaStream_t st, const PriorDev*, const ModelData& md, int64_t N, const double* th, uint64_t seed,
                     uint32_t epoch, uint32_t tag, uint32_t id0, double* dist, double*)
{
    const size_t sm = gk_smem_bytes<T>(md);
    gk_simulate_kernel<T><<<gk_grid<T>(N, sm), GK_THREADS, sm, st>>>(md, N, th, philox_keys(seed), epoch, tag, id0, dist);
}

const ModelOps* ops_gk()
{
    static const ModelOps o = { GkMath<double>::name, 4, 0, &gk_l_init<double>, &gk_l_smc<double>, &gk_l_mc<double>, &gk_l_sim<double>, nullptr, 0, 0 };
    return &o;
}
const ModelOps* ops_gk_f32()
{
    static const ModelOps o = { GkMath<float>::name, 4, 0, &gk_l_init<float>, &gk_l_smc<float>, &gk_l_mc<float>, &gk_l_sim<float>, nullptr, 0, 0 };
    return &o;
}

}  // namespace abcdez
