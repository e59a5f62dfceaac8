// mcsort.cu -- abcdemc!: the (delta, index)-sorted order behind the "better-or-equal particle" draw
// (src/abcdez_mc.jl:23: s = rand((1:N)[delta .<= delta_i]) becomes a binary search in the sorted distances
// plus one lookup).  The library's own sort, no CUB:
//   N <= 4096   one CTA, bitonic sort of (key, index) pairs in shared memory -- one launch;
//   larger N    least-significant-digit radix sort of the order-preserving 64-bit keys, 8 passes of 8 bits,
//               each pass = per-tile digit counts, one scan, a stable scatter (ranks inside a tile come from
//               warp match.any votes walked in index order, so ties keep their index order).
// Every kernel reads the schedule from the device control block and returns when the generation does not need
// the order (all particles at or below eps_target, src/abcdez_mc.jl:19-24) or the run has stopped, so
// abcdez_mc_run enqueues the same launch list for every generation without a host round trip.
#include "internal.h"
#include "ctrl.cuh"

namespace abcdez {

constexpr int MCS_THREADS = 256;
constexpr int MCS_TILE = 2048;                    // keys per CTA and pass: 8 warps x 8 steps x 32 lanes
constexpr int MCS_SMALL = 4096;

// does this generation need the sorted order?  force: stage-level calls always do
__device__ __forceinline__ bool mcs_needed(const Ctrl* c, int force)
{
    if (force) return true;
    if (c->stop | c->err) return false;
    return c->dmax > c->eps_target;
}

// ---- N <= 4096: bitonic sort in shared memory -----------------------------------------------------------------
__global__ void __launch_bounds__(1024) mc_sort_small_kernel(PopDev P, double* __restrict__ sorted_delta, uint32_t* __restrict__ order, int force)
{
    const Ctrl* c = P.ctrl;
    if (!mcs_needed(c, force)) return;
    __shared__ unsigned long long k[MCS_SMALL];
    __shared__ uint32_t v[MCS_SMALL];
    const uint32_t N = P.N;
    const double* __restrict__ dl = P.delta[c->cur];
    uint32_t n2 = 1; while (n2 < N) n2 <<= 1;
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) { k[i] = i < N ? f64_key(dl[i]) : ~0ull; v[i] = i < N ? i : 0xffffffffu; }
    __syncthreads();
    for (uint32_t len = 2; len <= n2; len <<= 1) {
        for (uint32_t j = len >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
                const uint32_t p = i ^ j;
                if (p > i) {
                    const bool up = (i & len) == 0;
                    const unsigned long long ki = k[i], kp = k[p]; const uint32_t vi = v[i], vp = v[p];
                    const bool gt = ki > kp || (ki == kp && vi > vp);           // (key, index) order: ties stay in index order
                    if (gt == up) { k[i] = kp; k[p] = ki; v[i] = vp; v[p] = vi; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) { sorted_delta[i] = key_f64(k[i]); order[i] = v[i]; }
}

// ---- larger N: LSD radix sort ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(MCS_THREADS) mc_sort_keys_kernel(PopDev P, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, int force)
{
    const Ctrl* c = P.ctrl;
    if (!mcs_needed(c, force)) return;
    const double* __restrict__ dl = P.delta[c->cur];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.N; i += gridDim.x * blockDim.x) { keys[i] = f64_key(dl[i]); vals[i] = i; }
}

// tile_hist[digit * ntiles + tile] = keys of this tile with that digit
__global__ void __launch_bounds__(MCS_THREADS) mc_sort_count_kernel(PopDev P, const unsigned long long* __restrict__ keys, uint32_t* __restrict__ tile_hist,
                                                                    int shift, int force)
{
    const Ctrl* c = P.ctrl;
    if (!mcs_needed(c, force)) return;
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t N = P.N, base = blockIdx.x * MCS_TILE;
    for (uint32_t e = threadIdx.x; e < (uint32_t)MCS_TILE; e += MCS_THREADS) {
        const uint32_t i = base + e;
        if (i < N) atomicAdd(&h[(unsigned)((keys[i] >> shift) & 255ull)], 1u);
    }
    __syncthreads();
    tile_hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of the digit-major tile histogram (one CTA; n = 256 * ntiles entries)
__global__ void __launch_bounds__(1024) mc_sort_scan_kernel(PopDev P, uint32_t* __restrict__ tile_hist, uint32_t n, int force)
{
    const Ctrl* c = P.ctrl;
    if (!mcs_needed(c, force)) return;
    __shared__ uint32_t part[1024];
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x, lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += tile_hist[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {                                 // scan of the 1024 chunk totals by one warp (32 per lane)
        uint32_t loc[32], tot = 0;
        for (int q = 0; q < 32; ++q) { loc[q] = part[threadIdx.x * 32 + q]; tot += loc[q]; }
        uint32_t incl = tot;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
        uint32_t run = incl - tot;
        for (int q = 0; q < 32; ++q) { part[threadIdx.x * 32 + q] = run; run += loc[q]; }
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) { const uint32_t t = tile_hist[i]; tile_hist[i] = run; run += t; }
}

__global__ void __launch_bounds__(MCS_THREADS) mc_sort_scatter_kernel(PopDev P, const unsigned long long* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                                      unsigned long long* __restrict__ kout, uint32_t* __restrict__ vout,
                                                                      const uint32_t* __restrict__ tile_hist, int shift, int force)
{
    const Ctrl* c = P.ctrl;
    if (!mcs_needed(c, force)) return;
    __shared__ unsigned wh[MCS_THREADS / 32][256];          // per-warp digit counters -> running output positions
    const uint32_t N = P.N, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t wbase = blockIdx.x * MCS_TILE + w * (MCS_TILE / (MCS_THREADS / 32));      // this warp's 256 consecutive keys
    for (int q = threadIdx.x; q < (MCS_THREADS / 32) * 256; q += MCS_THREADS) (&wh[0][0])[q] = 0u;
    __syncthreads();
    unsigned long long key[8]; uint32_t val[8];
#pragma unroll
    for (int st = 0; st < 8; ++st) {
        const uint32_t i = wbase + st * 32 + lane;
        key[st] = i < N ? kin[i] : 0ull; val[st] = i < N ? vin[i] : 0u;
        if (i < N) atomicAdd(&wh[w][(unsigned)((key[st] >> shift) & 255ull)], 1u);
    }
    __syncthreads();
    {   // thread d: first output position of digit d for every warp of this tile (tile base + the earlier warps' counts)
        const unsigned d = threadIdx.x;
        unsigned run = tile_hist[(size_t)d * gridDim.x + blockIdx.x];
        for (int q = 0; q < MCS_THREADS / 32; ++q) { const unsigned t = wh[q][d]; wh[q][d] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int st = 0; st < 8; ++st) {                        // in index order: a stable rank inside the warp's keys
        const uint32_t i = wbase + st * 32 + lane;
        const bool ok = i < N;
        const unsigned d = ok ? (unsigned)((key[st] >> shift) & 255ull) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (ok) {
            const unsigned pos = wh[w][d] + __popc(peers & ((1u << lane) - 1u));
            kout[pos] = key[st]; vout[pos] = val[st];
        }
        __syncwarp();
        if (ok && lane == (unsigned)(__ffs(peers) - 1)) wh[w][d] += (unsigned)__popc(peers);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(MCS_THREADS) mc_sort_finish_kernel(PopDev P, const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                     double* __restrict__ sorted_delta, uint32_t* __restrict__ order, int force)
{
    const Ctrl* c = P.ctrl;
    if (!mcs_needed(c, force)) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.N; i += gridDim.x * blockDim.x) { sorted_delta[i] = key_f64(keys[i]); order[i] = vals[i]; }
}

size_t mc_sort_tmp_bytes(int64_t N)
{
    if (N <= MCS_SMALL) return 256;
    const size_t n = (size_t)N, nt = (n + MCS_TILE - 1) / MCS_TILE;
    return 2 * n * 8 + 2 * n * 4 + 256 * nt * 4 + 1024;
}

// sorted_delta / order <- the live generation's distances in (delta, index) order; returns the number of launches
int launch_mc_prepare(cudaStream_t st, const PopDev& P, double* sorted_delta, uint32_t* order, void* tmp, size_t tmp_bytes, int force)
{
    (void)tmp_bytes;
    const uint32_t N = P.N;
    if (N <= (uint32_t)MCS_SMALL) {
        mc_sort_small_kernel<<<1, 1024, 0, st>>>(P, sorted_delta, order, force);
        return 1;
    }
    const size_t n = N, nt = (n + MCS_TILE - 1) / MCS_TILE;
    char* p = reinterpret_cast<char*>(tmp);
    unsigned long long* k0 = reinterpret_cast<unsigned long long*>(p); p += n * 8;
    unsigned long long* k1 = reinterpret_cast<unsigned long long*>(p); p += n * 8;
    uint32_t* v0 = reinterpret_cast<uint32_t*>(p); p += n * 4;
    uint32_t* v1 = reinterpret_cast<uint32_t*>(p); p += n * 4;
    uint32_t* th = reinterpret_cast<uint32_t*>(p);
    const unsigned gs = (unsigned)((n + MCS_THREADS - 1) / MCS_THREADS) < 1184u ? (unsigned)((n + MCS_THREADS - 1) / MCS_THREADS) : 1184u;
    mc_sort_keys_kernel<<<gs, MCS_THREADS, 0, st>>>(P, k0, v0, force);
    int launches = 1;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 8 * pass;
        mc_sort_count_kernel<<<(unsigned)nt, MCS_THREADS, 0, st>>>(P, k0, th, shift, force);
        mc_sort_scan_kernel<<<1, 1024, 0, st>>>(P, th, (uint32_t)(256 * nt), force);
        mc_sort_scatter_kernel<<<(unsigned)nt, MCS_THREADS, 0, st>>>(P, k0, v0, k1, v1, th, shift, force);
        unsigned long long* tk = k0; k0 = k1; k1 = tk;
        uint32_t* tv = v0; v0 = v1; v1 = tv;
        launches += 3;
    }
    mc_sort_finish_kernel<<<gs, MCS_THREADS, 0, st>>>(P, k0, v0, sorted_delta, order, force);
    return launches + 1;
}

}  // namespace abcdez
