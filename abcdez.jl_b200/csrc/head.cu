// head.cu -- the "head" of one abcdesmc! iteration (src/abcdez_smc.jl:301-324 of the reference) as ONE
// persistent cooperative kernel:
//   eps schedule   quantile(delta[alive], alpha), clamp                 :301
//   reweighting    abcdesmc_update_ws!, wprod, wnorm, Wns, alive         :59-83, :305-311
//   logZ / ESS     logZ += log(wnorm); ess = 1/sum(Wns^2); resample flag :315, :323-324
//   alive list     compaction (the O(1) replacement of the wsample scans :121,125)
// The unfused path (bookkeeping.cu) needs 10 launches of ~10-20 us each for 8 MB of state at 10^6 particles:
// they are launch/latency bound, not bandwidth bound.  Here every CTA owns a contiguous range of
// 1024-particle tiles, the phases are separated by grid barriers (~2 us) and every grid-wide decision is
// recomputed redundantly by every CTA from the same integers / the same fixed-order partial sums, so no
// phase waits on a single "last block".  Reductions keep the per-tile partials and the summation order of
// the unfused kernels: both paths give bit-identical eps, W, wnorm, logZ and ESS.
//
// Radix select: two 11-bit digit passes over the alive distances (order-preserving 64-bit keys), then the
// few keys that share the 22-bit prefix of v[j] are compacted into a candidate list and the remaining 42
// bits, the tie test and v[j+1] are resolved on that list (per CTA in shared memory when it is small,
// grid-cooperatively while it is large).
#include "internal.h"
#include "seqsum.cuh"
#include "ctrl.cuh"
#include "comm.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace abcdez {

constexpr int HEAD_THREADS = BK_THREADS;
#ifndef ABCDEZ_HEAD_MIN_BLOCKS
#define ABCDEZ_HEAD_MIN_BLOCKS 2
#endif
constexpr int HEAD_MIN_BLOCKS = ABCDEZ_HEAD_MIN_BLOCKS;   // resident CTAs per SM the register budget is cut for
constexpr int CAND_SMEM = 4096;                   // candidates staged in shared memory for the per-CTA tail; longer
                                                  // lists are first refined grid-cooperatively, one digit per round

struct HeadSmem {
    unsigned hist[SEL_BINS];
    unsigned long long cand[CAND_SMEM];
    unsigned part[HEAD_THREADS];
    double red[32];
    unsigned wcnt[32];
    unsigned long long wmin[32];
    unsigned long long mn, mx, mab;
    unsigned found_bin, found_before, found_cnt, flag;
    unsigned long long xh[8];                     // sharded runs: this rank's header record of an exchange
    unsigned long long xall[XCHG_MAXR * 8];       // ... and every rank's, after it
    int xflag;
};

__device__ __forceinline__ int pass_shift(int p) { const int s[6] = { 53, 42, 31, 20, 9, 0 }; return s[p]; }
__device__ __forceinline__ int pass_bins(int p) { return p == 5 ? 512 : 2048; }
__device__ __forceinline__ unsigned long long pass_himask_after(int p) { return p == 5 ? ~0ull : (~0ull << pass_shift(p)); }

// pick the bin holding rank `rank` in a SEL_BINS histogram (global: read through L2; else shared); every
// thread of the CTA participates and gets the same answer.  bin == 0xffffffff: rank beyond the total.
__device__ __noinline__ void pick_bin(const unsigned* h, bool global, int nbins, unsigned long long rank, HeadSmem* s,
                         unsigned& bin, unsigned long long& before)
{
    const int per = SEL_BINS / HEAD_THREADS;
    unsigned loc[per], tot = 0;
#pragma unroll
    for (int k = 0; k < per; ++k) {
        int b = threadIdx.x * per + k;
        loc[k] = (b < nbins) ? (global ? __ldcg(&h[b]) : h[b]) : 0u;
        tot += loc[k];
    }
    // exclusive prefix of the per-thread totals: shuffle scan inside the warps, then the warp totals
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    unsigned incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
    __syncthreads();
    if (lane == 31) s->part[wp] = incl;
    if (threadIdx.x == 0) { s->found_bin = 0xffffffffu; s->found_before = 0; s->found_cnt = 0xffffffffu; }
    __syncthreads();
    unsigned bef = incl - tot;
    for (unsigned q = 0; q < wp; ++q) bef += s->part[q];
    unsigned cum = bef;
#pragma unroll
    for (int k = 0; k < per; ++k) {
        if (loc[k] && rank >= cum && rank < (unsigned long long)cum + loc[k]) { s->found_bin = threadIdx.x * per + k; s->found_before = cum; s->found_cnt = loc[k]; }
        cum += loc[k];
    }
    __syncthreads();
    bin = s->found_bin; before = s->found_before;
}

// warp-aggregated shared-memory histogram increment (leading digits of distances are highly concentrated)
__device__ __forceinline__ void hist_add(unsigned* sh, bool ok, unsigned digit)
{
    unsigned d = ok ? digit : 0xffffffffu;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    if (ok && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&sh[digit], (unsigned)__popc(peers));
}

__device__ __forceinline__ void hist_clear(HeadSmem* s)
{
    for (int b = threadIdx.x; b < SEL_BINS; b += HEAD_THREADS) s->hist[b] = 0u;
    __syncthreads();
}

__device__ __forceinline__ void hist_flush(HeadSmem* s, unsigned* gh, int nbins)
{
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += HEAD_THREADS) {
        unsigned cnt = s->hist[b];
        if (cnt) atomicAdd(&gh[b], cnt);
    }
}

// broadcast a block_sum result (valid in warp 0) to every thread
__device__ __forceinline__ double block_sum_all(double v, HeadSmem* s)
{
    double t = block_sum(v, s->red);
    __syncthreads();
    if (threadIdx.x == 0) s->red[0] = t;
    __syncthreads();
    t = s->red[0];
    __syncthreads();
    return t;
}

// resolve v[j] (rank `rank` among the M listed keys, all of which match prefix/himask), the tie test and
// v[j+1] with the digits from pass p0 on, entirely inside the CTA.  L: shared or global (L2) list.
__device__ __noinline__ void tail_select(const unsigned long long* L, unsigned M, int p0, unsigned long long prefix,
                            unsigned long long himask, unsigned long long rank, unsigned long long min_above,
                            HeadSmem* s, unsigned long long& akey, unsigned long long& bkey)
{
    const unsigned long long rank0 = rank;
    for (int p = p0; p < 6; ++p) {
        const int shift = pass_shift(p), nbins = pass_bins(p);
        hist_clear(s);
        for (unsigned i = threadIdx.x; i < M; i += HEAD_THREADS) {
            unsigned long long key = L[i];
            if ((key & himask) == prefix) atomicAdd(&s->hist[(unsigned)((key >> shift) & (unsigned long long)(nbins - 1))], 1u);
        }
        __syncthreads();
        unsigned bin; unsigned long long before;
        pick_bin(s->hist, false, nbins, rank, s, bin, before);
        const unsigned in_bin = s->found_cnt;       // keys in the picked bin (read like found_bin: stable until the next pick_bin)
        prefix |= (unsigned long long)bin << shift;
        himask = pass_himask_after(p);
        rank -= before;
        if (in_bin <= 32u && p < 5) {
            // few keys left (the usual case after one digit): gather them and rank them directly instead of
            // running the remaining digit passes; same key, the exact select does not depend on how it is found
            __syncthreads();
            if (threadIdx.x == 0) s->flag = 0u;
            __syncthreads();
            for (unsigned i = threadIdx.x; i < M; i += HEAD_THREADS) {
                unsigned long long key = L[i];
                if ((key & himask) == prefix) s->wmin[atomicAdd(&s->flag, 1u)] = key;
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                const unsigned n = s->flag;
                const unsigned long long my = threadIdx.x < n ? s->wmin[threadIdx.x] : ~0ull;
                unsigned r = 0;
                for (unsigned k = 0; k < n; ++k) {
                    const unsigned long long o = s->wmin[k];
                    r += (o < my || (o == my && k < threadIdx.x)) ? 1u : 0u;
                }
                if (threadIdx.x < n && (unsigned long long)r == rank) s->mn = my;
            }
            __syncthreads();
            prefix = s->mn;
            __syncthreads();
            break;
        }
    }
    akey = prefix;
    // #{keys <= v[j]} and min{key > v[j]} on the list
    unsigned cnt = 0; unsigned long long mn = ~0ull;
    for (unsigned i = threadIdx.x; i < M; i += HEAD_THREADS) {
        unsigned long long key = L[i];
        if (key <= akey) cnt++; else mn = key < mn ? key : mn;
    }
    cnt = warp_sum_u(cnt); mn = warp_min_u64(mn);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s->wcnt[threadIdx.x >> 5] = cnt; s->wmin[threadIdx.x >> 5] = mn; }
    __syncthreads();
    cnt = 0; mn = ~0ull;
    for (int w = 0; w < HEAD_THREADS / 32; ++w) { cnt += s->wcnt[w]; mn = s->wmin[w] < mn ? s->wmin[w] : mn; }
    __syncthreads();
    if ((unsigned long long)cnt >= rank0 + 2) bkey = akey;          // v[j+1] ties with v[j]
    else bkey = (mn != ~0ull) ? mn : min_above;
}

// ---------------------------------------------------------------------------------------
// sharded runs (exchange 1 and 2): CTA 0 turns this rank's partial results into the
// whole population's, in place, over NVLink peer memory (comm.cuh); the caller wraps each of these in
// grid barriers.  Integer sums and rank-ordered FP64 sums: bit-identical on every rank.
// ---------------------------------------------------------------------------------------
// H[0..nbins) += every other rank's histogram; optionally global extrema(delta) and the sticky error
__device__ __noinline__ void head_xchg_hist(const PopDev& P, Ctrl* c, unsigned* H, int nbins, bool first, HeadSmem* s)
{
    const unsigned tid = threadIdx.x;
    if (tid == 0 && first) {
        int e0 = c->err, e1 = __ldcg(&c->acc.err);
        s->xh[0] = __ldcg(&c->acc.dmin_key); s->xh[1] = __ldcg(&c->acc.dmax_key);
        s->xh[2] = (unsigned long long)(e0 > e1 ? e0 : e1);
    }
    __syncthreads();
    xchg_ll_block(P.x, c, s->xh, first ? 3 : 0, H, nbins, s->xall, &s->xflag);
    if (tid == 0 && first) {
        unsigned long long mn = ~0ull, mx = 0ull, e = 0ull;
        for (int r = 0; r < P.x.world; ++r) {
            unsigned long long a = s->xall[3 * r], b = s->xall[3 * r + 1], er = s->xall[3 * r + 2];
            mn = a < mn ? a : mn; mx = b > mx ? b : mx; e = er > e ? er : e;
        }
        __stcg(&c->acc.dmin_key, mn); __stcg(&c->acc.dmax_key, mx);
        if (e) { c->acc.err = (int)e; if (!c->err) c->err = (int)e; }
    }
}

// candidate-list bookkeeping of generation g: global count / extrema / min-above into c->acc.g_*; when the
// global list fits the per-CTA tail (and is not all-equal) this rank's candidates are posted into every
// rank's mailbox (XCHG_GCAND_OFF), where every CTA of every rank collects the same multiset for the tail
__device__ __noinline__ void head_xchg_cands(const PopDev& P, Ctrl* c, int g, HeadSmem* s)
{
    const unsigned tid = threadIdx.x;
    if (tid == 0) {
        s->xh[0] = __ldcg(&c->acc.cand_count[g]); s->xh[1] = __ldcg(&c->acc.cand_min[g]);
        s->xh[2] = __ldcg(&c->acc.cand_max[g]); s->xh[3] = __ldcg(&c->acc.min_above);
    }
    __syncthreads();
    xchg_ll_block(P.x, c, s->xh, 4, nullptr, 0, s->xall, &s->xflag);
    unsigned long long M = 0ull, off = 0ull, mn = ~0ull, mx = 0ull, mab = ~0ull;
    for (int r = 0; r < P.x.world; ++r) {
        unsigned long long m = s->xall[4 * r], a = s->xall[4 * r + 1], b = s->xall[4 * r + 2], ab = s->xall[4 * r + 3];
        if (r < P.x.rank) off += m;
        M += m; mn = a < mn ? a : mn; mx = b > mx ? b : mx; mab = ab < mab ? ab : mab;
    }
    if (tid == 0) {
        c->acc.g_cand_count = M; c->acc.g_cand_min = mn; c->acc.g_cand_max = mx; c->acc.g_min_above = mab;
        c->acc.g_cand_off = off;
    }
    if (M <= (unsigned long long)XCHG_GCAND && mn != mx)                   // uniform over ranks
        ll_post_keys(P.x, P.cand[g & 1], (unsigned)s->xh[0], (unsigned)off);
    __syncthreads();
}

// all-gather of the nw-word record the caller put into s->xh (thread 0): rank r's words at s->xall[nw*r ...]
__device__ __noinline__ void head_xchg_rec(const PopDev& P, Ctrl* c, int nw, HeadSmem* s)
{
    __syncthreads();
    xchg_ll_block(P.x, c, s->xh, nw, nullptr, 0, s->xall, &s->xflag);
}

__global__ void __launch_bounds__(HEAD_THREADS, HEAD_MIN_BLOCKS) head_kernel(const __grid_constant__ PopDev P)
{
    Ctrl* c = P.ctrl;
    if (c->stop) return;                                    // uniform: nobody reaches a grid barrier
    cg::grid_group grid = cg::this_grid();
    __shared__ HeadSmem s;
    const unsigned G = gridDim.x, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const uint32_t N = P.N, ntiles = P.ntiles;
    const unsigned t0 = (unsigned)((unsigned long long)blockIdx.x * ntiles / G);
    const unsigned t1 = (unsigned)((unsigned long long)(blockIdx.x + 1) * ntiles / G);
    // schedule scalars of the previous iteration (CTA 0 overwrites them after the third barrier)
    const int cur = c->cur, kind = c->kind;
    const double eps_prev = c->eps, eps_old = c->eps_k, eps_target = c->eps_target, q_gamma = c->q_gamma;
    const unsigned n_alive_prev = c->n_alive_g;
    const bool sharded = P.x.world > 1;
    const double* __restrict__ dl = P.delta[cur];
    unsigned* H1 = P.sel_hist; unsigned* H2 = P.sel_hist + SEL_BINS; unsigned* HW = P.sel_hist + 6 * SEL_BINS;
    unsigned long long rank = c->sel_rank;
    const unsigned long long rank_all = rank;

    // ---- first pass over every alive key (+ extrema(delta) over all particles, NaN check) --------------
    // Windowed digit: every alive distance lies in the support of the previous kernel, i.e. at or below eps_prev,
    // and the alpha-quantile sits in the top two binades below it in all but degenerate populations.  So the
    // first digit is the 22-bit key prefix (sign, exponent, 10 mantissa bits) RELATIVE to eps_prev's: bins
    // 1..2046 are the 2046 prefixes up to and including eps_prev's own, bin 0 collects everything below the
    // window and bin 2047 everything above it.  The map is monotone, so the select stays exact: when the rank
    // falls into one of the single-prefix bins, 22 bits of v[j] are known after ONE pass (the generic scheme
    // needs two: its leading 11-bit digit is the same for almost all distances); when it falls into a
    // collecting bin (or eps_prev is not finite: first iteration) the generic two passes run instead.
    const bool windowed = isfinite(eps_prev) && eps_prev > 0.0;
    const unsigned long long wbase = (f64_key(windowed ? eps_prev : 1.0) >> 42) - 2046ull;
    auto first_pass = [&](bool win, unsigned* H) {
        hist_clear(&s);
        unsigned long long kmn = ~0ull, kmx = 0ull; int nan_seen = 0;
        for (unsigned tile = t0; tile < t1; ++tile) {
            size_t i0 = (size_t)tile * TILE + (size_t)tid * 4;
            double v[4];
            load4_f64(dl, i0, N, v);
            uint32_t al = load4_u8(P.alive, i0, N);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                bool ok = (al >> (8 * k)) & 0xff;
                unsigned long long key = f64_key(v[k]);
                if (i0 + k < N) { kmn = key < kmn ? key : kmn; kmx = key > kmx ? key : kmx; }
                if (ok && isnan(v[k])) nan_seen = 1;
                unsigned digit;
                if (win) {
                    // (window digits are spread over ~10^3 bins: plain shared-memory atomics; the generic leading digit
                    // is the same for almost every key: warp-aggregated)
                    const unsigned long long d = key >> 42;
                    digit = d <= wbase ? 0u : (d - wbase >= 2047ull ? 2047u : (unsigned)(d - wbase));
                    if (ok) atomicAdd(&s.hist[digit], 1u);
                } else {
                    digit = (unsigned)(key >> 53);
                    hist_add(s.hist, ok, digit);
                }
            }
        }
        if (tid == 0) { s.mn = ~0ull; s.mx = 0ull; }
        __syncthreads();
        kmn = warp_min_u64(kmn); kmx = warp_max_u64(kmx);
        if (lane == 0) { atomicMin(&s.mn, kmn); atomicMax(&s.mx, kmx); }
        if (nan_seen) atomicMax(&c->acc.err, (int)ABCDEZ_ERR_NAN_DISTANCE);
        hist_flush(&s, H, 2048);
        if (tid == 0) { atomicMin(&c->acc.dmin_key, s.mn); atomicMax(&c->acc.dmax_key, s.mx); }
    };
    unsigned bin; unsigned long long before;
    unsigned long long prefix = 0ull, himask = 0ull;
    bool have22 = false;
    if (windowed) {
        first_pass(true, HW);
        grid.sync();
        if (sharded) { if (blockIdx.x == 0) head_xchg_hist(P, c, HW, 2048, true, &s); grid.sync(); }
        pick_bin(HW, true, 2048, rank, &s, bin, before);
        if (blockIdx.x == 0 && tid == 0) patch_extrema(P, c);   // ranges_eps of the previous record, :363
        if (bin == 0xffffffffu) {                           // empty alive set (uniform decision)
            grid.sync();                                    // every CTA has read HW
            if (blockIdx.x == 0 && tid == 0 && !c->err) c->err = ABCDEZ_ERR_NO_ALIVE;
            if (blockIdx.x == 0) for (int b = tid; b < SEL_BINS; b += HEAD_THREADS) HW[b] = 0u;
            return;
        }
        if (bin >= 1u && bin <= 2046u) {
            prefix = (wbase + (unsigned long long)bin) << 42; himask = ~0ull << 42;
            rank -= before;
            have22 = true;
        }
    }
    if (!have22) {
        // ---- generic pass 1: digit 53..63 --------------------------------------------------------------
        rank = rank_all;
        first_pass(false, H1);
        grid.sync();
        if (sharded) { if (blockIdx.x == 0) head_xchg_hist(P, c, H1, 2048, !windowed, &s); grid.sync(); }
        pick_bin(H1, true, 2048, rank, &s, bin, before);
        if (!windowed && blockIdx.x == 0 && tid == 0) patch_extrema(P, c);   // ranges_eps of the previous record, :363
        if (bin == 0xffffffffu) {                           // empty alive set (uniform decision)
            grid.sync();                                    // every CTA has read H1
            if (blockIdx.x == 0 && tid == 0 && !c->err) c->err = ABCDEZ_ERR_NO_ALIVE;
            if (blockIdx.x == 0) for (int b = tid; b < SEL_BINS; b += HEAD_THREADS) { P.sel_hist[b] = 0u; HW[b] = 0u; }
            return;
        }
        prefix = (unsigned long long)bin << 53; himask = ~0ull << 53;
        rank -= before;

        // ---- generic pass 2: digit 42..52 among the keys of that bin -----------------------------------
        hist_clear(&s);
        for (unsigned tile = t0; tile < t1; ++tile) {
            size_t i0 = (size_t)tile * TILE + (size_t)tid * 4;
            double v[4];
            load4_f64(dl, i0, N, v);
            uint32_t al = load4_u8(P.alive, i0, N);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned long long key = f64_key(v[k]);
                bool ok = ((al >> (8 * k)) & 0xff) && ((key & himask) == prefix);
                hist_add(s.hist, ok, (unsigned)((key >> 42) & 2047ull));
            }
        }
        hist_flush(&s, H2, 2048);
        grid.sync();
        if (sharded) { if (blockIdx.x == 0) head_xchg_hist(P, c, H2, 2048, false, &s); grid.sync(); }
        pick_bin(H2, true, 2048, rank, &s, bin, before);
        prefix |= (unsigned long long)bin << 42; himask = ~0ull << 42;
        rank -= before;
    }

    // ---- pass 3: compact the keys sharing the 22-bit prefix; min key above the prefix ----------------
    unsigned long long* cand = P.cand[0];
    {
        unsigned long long mab = ~0ull, cmn = ~0ull, cmx = 0ull;
        for (unsigned tile = t0; tile < t1; ++tile) {
            size_t i0 = (size_t)tile * TILE + (size_t)tid * 4;
            double v[4];
            load4_f64(dl, i0, N, v);
            uint32_t al = load4_u8(P.alive, i0, N);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned long long key = f64_key(v[k]);
                bool ok = (al >> (8 * k)) & 0xff;
                unsigned long long hi = key & himask;
                bool is_c = ok && hi == prefix;
                if (ok && hi > prefix) mab = key < mab ? key : mab;
                unsigned m = __ballot_sync(0xffffffffu, is_c);
                if (m) {
                    unsigned base = 0;
                    if (lane == 0) base = (unsigned)atomicAdd(&c->acc.cand_count[0], (unsigned long long)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (is_c) {
                        cand[base + __popc(m & ((1u << lane) - 1u))] = key;
                        cmn = key < cmn ? key : cmn; cmx = key > cmx ? key : cmx;
                    }
                }
            }
        }
        if (tid == 0) { s.mn = ~0ull; s.mx = 0ull; s.mab = ~0ull; }
        __syncthreads();
        mab = warp_min_u64(mab); cmn = warp_min_u64(cmn); cmx = warp_max_u64(cmx);
        if (lane == 0) { atomicMin(&s.mab, mab); atomicMin(&s.mn, cmn); atomicMax(&s.mx, cmx); }
        __syncthreads();
        if (tid == 0) {
            if (s.mab != ~0ull) atomicMin(&c->acc.min_above, s.mab);
            if (s.mn != ~0ull) { atomicMin(&c->acc.cand_min[0], s.mn); atomicMax(&c->acc.cand_max[0], s.mx); }
        }
    }
    grid.sync();
    if (sharded) { if (blockIdx.x == 0) head_xchg_cands(P, c, 0, &s); grid.sync(); }
    // candidate-list generation g: list P.cand[g & 1], accumulators cand_count/min/max[g] (reset after the run).
    // M: this rank's candidates; Mg, cmin, cmax, min_above: the whole population's
    int p_next = 2, g = 0;
    unsigned M = (unsigned)__ldcg(&c->acc.cand_count[0]);
    unsigned long long Mg = sharded ? __ldcg(&c->acc.g_cand_count) : (unsigned long long)M;
    unsigned long long cmin = sharded ? __ldcg(&c->acc.g_cand_min) : __ldcg(&c->acc.cand_min[0]);
    unsigned long long cmax = sharded ? __ldcg(&c->acc.g_cand_max) : __ldcg(&c->acc.cand_max[0]);
    unsigned long long min_above = sharded ? __ldcg(&c->acc.g_min_above) : __ldcg(&c->acc.min_above);

    // ---- large candidate lists: refine grid-cooperatively, one digit per round -----------------------
    while (Mg > (unsigned long long)CAND_SMEM && cmin != cmax && p_next < 6) {
        const int shift = pass_shift(p_next), nbins = pass_bins(p_next);
        unsigned* H = P.sel_hist + (size_t)p_next * SEL_BINS;
        const unsigned long long* src = P.cand[g & 1]; unsigned long long* dst = P.cand[(g + 1) & 1];
        hist_clear(&s);
        for (size_t i = (size_t)blockIdx.x * HEAD_THREADS + tid; i < M; i += (size_t)G * HEAD_THREADS)
            atomicAdd(&s.hist[(unsigned)((__ldcg(&src[i]) >> shift) & (unsigned long long)(nbins - 1))], 1u);
        hist_flush(&s, H, nbins);
        grid.sync();
        if (sharded) { if (blockIdx.x == 0) head_xchg_hist(P, c, H, nbins, false, &s); grid.sync(); }
        pick_bin(H, true, nbins, rank, &s, bin, before);
        prefix |= (unsigned long long)bin << shift; himask = pass_himask_after(p_next);
        rank -= before;
        {
            unsigned long long mab = ~0ull, cmn = ~0ull, cmx = 0ull;
            size_t rounds = ((size_t)M + (size_t)G * HEAD_THREADS - 1) / ((size_t)G * HEAD_THREADS);
            for (size_t r = 0; r < rounds; ++r) {
                size_t i = r * (size_t)G * HEAD_THREADS + (size_t)blockIdx.x * HEAD_THREADS + tid;
                unsigned long long key = i < M ? __ldcg(&src[i]) : 0ull;
                unsigned long long hi = key & himask;
                bool is_c = i < M && hi == prefix;
                if (i < M && hi > prefix) mab = key < mab ? key : mab;
                unsigned m = __ballot_sync(0xffffffffu, is_c);
                if (m) {
                    unsigned base = 0;
                    if (lane == 0) base = (unsigned)atomicAdd(&c->acc.cand_count[g + 1], (unsigned long long)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (is_c) {
                        dst[base + __popc(m & ((1u << lane) - 1u))] = key;
                        cmn = key < cmn ? key : cmn; cmx = key > cmx ? key : cmx;
                    }
                }
            }
            mab = warp_min_u64(mab); cmn = warp_min_u64(cmn); cmx = warp_max_u64(cmx);
            if (lane == 0) {
                if (mab != ~0ull) atomicMin(&c->acc.min_above, mab);
                if (cmn != ~0ull) { atomicMin(&c->acc.cand_min[g + 1], cmn); atomicMax(&c->acc.cand_max[g + 1], cmx); }
            }
        }
        grid.sync();
        g++; p_next++;
        if (sharded) { if (blockIdx.x == 0) head_xchg_cands(P, c, g, &s); grid.sync(); }
        M = (unsigned)__ldcg(&c->acc.cand_count[g]);
        Mg = sharded ? __ldcg(&c->acc.g_cand_count) : (unsigned long long)M;
        cmin = sharded ? __ldcg(&c->acc.g_cand_min) : __ldcg(&c->acc.cand_min[g]);
        cmax = sharded ? __ldcg(&c->acc.g_cand_max) : __ldcg(&c->acc.cand_max[g]);
        min_above = sharded ? __ldcg(&c->acc.g_min_above) : __ldcg(&c->acc.min_above);
    }

    // ---- tail: v[j], tie test, v[j+1], type-7 interpolation, clamp (every CTA, same integers) ---------
    unsigned long long akey, bkey;
    if (cmin == cmax) {                                     // all candidates equal (discrete distances, ties)
        akey = cmin;
        bkey = (rank + 1 < Mg) ? akey : min_above;
    } else {                                                // Mg <= CAND_SMEM (after p_next == 6 all keys are equal)
        if (sharded) ll_collect_keys(P.x, c, (unsigned)Mg, s.cand);        // every rank's candidates, from the local mailbox
        else { for (unsigned i = tid; i < (unsigned)Mg; i += HEAD_THREADS) s.cand[i] = __ldcg(&P.cand[g & 1][i]); }
        __syncthreads();
        tail_select(s.cand, (unsigned)Mg, p_next, prefix, himask, rank, min_above, &s, akey, bkey);
    }
    double eps;
    {
        double a = key_f64(akey);
        double b = (n_alive_prev <= 1 || bkey == ~0ull) ? a : key_f64(bkey);
        double q = (isfinite(a) && isfinite(b)) ? a + q_gamma * (b - a) : (1.0 - q_gamma) * a + q_gamma * b;
        eps = fmax(fmin(q, eps_prev), eps_target);          // :301
        if (blockIdx.x == 0 && tid == 0) { c->q_a = a; c->q_b = b; c->q = q; c->eps = eps; c->sel_prefix = akey; }
    }

    // ---- reweight pass A: ws, wprod (unnormalised, kept in W), per-tile sums -------------------------
    for (unsigned tile = t0; tile < t1; ++tile) {
        size_t i0 = (size_t)tile * TILE + (size_t)tid * 4;
        double v[4], w[4];
        load4_f64(dl, i0, N, v);
        load4_f64(P.W, i0, N, w);
        uint32_t al = load4_u8(P.alive, i0, N);
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool ok = ((al >> (8 * k)) & 0xff) && (i0 + k < N);
            double ws = 0.0;
            if (ok) ws = abck_ws(kind, eps, eps_old, v[k]);                                              // :75
            w[k] = ok ? w[k] * ws : 0.0;                                                          // :308
            acc += w[k];
        }
        store4_f64(P.W, i0, N, w);
        double t = block_sum(acc, s.red);
        if (tid == 0) P.partial[tile] = t;
    }
    grid.sync();
    double wnorm;
    {
        double a = 0.0;
        for (unsigned b = tid; b < ntiles; b += HEAD_THREADS) a += __ldcg(&P.partial[b]);
        wnorm = block_sum_all(a, &s);                                                             // :309
    }
    if (sharded) {                                          // wnorm = sum over ranks, rank order (exchange 2)
        if (blockIdx.x == 0) {
            if (tid == 0) s.xh[0] = (unsigned long long)__double_as_longlong(wnorm);
            head_xchg_rec(P, c, 1, &s);
            if (tid == 0) {
                double tot = 0.0;
                for (int r = 0; r < P.x.world; ++r) tot += __longlong_as_double((long long)s.xall[r]);
                c->acc.g_wnorm = tot;
            }
        }
        grid.sync();
        wnorm = __ldcg(&c->acc.g_wnorm);
    }
    if (blockIdx.x == 0) {
        if (tid == 0) {
            c->wnorm = wnorm; c->logZ += plog(wnorm);                                             // :315
            for (int q = 0; q < 6; ++q) { c->acc.cand_count[q] = 0ull; c->acc.cand_min[q] = ~0ull; c->acc.cand_max[q] = 0ull; }
            c->acc.min_above = ~0ull;
        }
        for (int b = tid; b < 7 * SEL_BINS; b += HEAD_THREADS) P.sel_hist[b] = 0u;              // every reader is past them
    }

    // ---- reweight pass B: Wns, alive, sum(Wns^2), alive counts per tile ------------------------------
    double* partial2 = P.partial + ntiles;
    for (unsigned tile = t0; tile < t1; ++tile) {
        size_t i0 = (size_t)tile * TILE + (size_t)tid * 4;
        double w[4];
        load4_f64(P.W, i0, N, w);
        double acc = 0.0; unsigned cnt = 0; uint32_t al = 0; double wal = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < N) {
                w[k] = w[k] / wnorm;                                                              // :310
                bool a = (w[k] > 0.0);                                                            // :311
                if (a) { al |= 1u << (8 * k); cnt++; wal = w[k]; }
                acc += w[k] * w[k];
            }
        }
        store4_f64(P.W, i0, N, w);
        if (i0 + 3 < N) *reinterpret_cast<uint32_t*>(P.alive + i0) = al;
        else { for (int k = 0; k < 4; ++k) if (i0 + k < N) P.alive[i0 + k] = (al >> (8 * k)) & 0xff; }
        if (cnt) c->acc.w_alive = wal;          // indicator kernels: every alive weight is this same double
        double t = block_sum(acc, s.red);
        cnt = warp_sum_u(cnt);
        if (lane == 0) s.wcnt[wp] = cnt;
        __syncthreads();
        if (tid == 0) {
            unsigned tc = 0;
            for (int q = 0; q < HEAD_THREADS / 32; ++q) tc += s.wcnt[q];
            partial2[tile] = t;
            P.tile_cnt[tile] = tc;
        }
        __syncthreads();
    }
    grid.sync();
    double sumsq;
    unsigned n_alive, my_off;
    {
        double a = 0.0; unsigned tot = 0, pre = 0;
        for (unsigned b = tid; b < ntiles; b += HEAD_THREADS) {
            a += __ldcg(&partial2[b]);
            unsigned tc = __ldcg(&P.tile_cnt[b]);
            tot += tc; if (b < t0) pre += tc;
        }
        sumsq = block_sum_all(a, &s);
        tot = warp_sum_u(tot); pre = warp_sum_u(pre);
        if (lane == 0) { s.wcnt[wp] = tot; s.part[wp] = pre; }
        __syncthreads();
        n_alive = 0; my_off = 0;
        for (int q = 0; q < HEAD_THREADS / 32; ++q) { n_alive += s.wcnt[q]; my_off += s.part[q]; }
        __syncthreads();
    }
    if (sharded) {                                          // sum(Wns^2), alive counts, common alive weight (exchange 2)
        if (blockIdx.x == 0) {
            if (tid == 0) {
                s.xh[0] = (unsigned long long)__double_as_longlong(sumsq); s.xh[1] = (unsigned long long)n_alive;
                s.xh[2] = (unsigned long long)__double_as_longlong(__ldcg(&c->acc.w_alive));
            }
            head_xchg_rec(P, c, 3, &s);
            if (tid == 0) {
                double tot = 0.0; unsigned ng = 0;
                for (int r = 0; r < P.x.world; ++r) {
                    tot += __longlong_as_double((long long)s.xall[3 * r]);
                    unsigned na = (unsigned)s.xall[3 * r + 1];
                    c->rank_alive[r] = na; ng += na;
                    if (na) __stcg(&c->acc.w_alive, __longlong_as_double((long long)s.xall[3 * r + 2]));   // same double on every rank
                }
                ctrl_after_reweight(P, c, tot, n_alive, ng);                                      // :318-324
            }
        }
    } else if (blockIdx.x == 0 && tid == 0) ctrl_after_reweight(P, c, sumsq, n_alive, n_alive);   // :318-324

    // ---- compaction: alive particles first (index order), the dead ones behind them ------------------
    if (n_alive != N) {
        unsigned off = my_off;
        for (unsigned tile = t0; tile < t1; ++tile) {
            size_t i0 = (size_t)tile * TILE + (size_t)tid * 4;
            uint32_t al = load4_u8(P.alive, i0, N);
            unsigned cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) cnt += ((al >> (8 * k)) & 0xff) ? 1u : 0u;
            unsigned incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            __syncthreads();
            if (lane == 31) s.wcnt[wp] = incl;
            __syncthreads();
            unsigned woff = 0, ttot = 0;
            for (int q = 0; q < HEAD_THREADS / 32; ++q) { if (q < (int)wp) woff += s.wcnt[q]; ttot += s.wcnt[q]; }
            unsigned pos = off + woff + incl - cnt;                       // alive before element i0
            unsigned dpos = n_alive + ((unsigned)i0 - pos);               // dead before element i0
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i0 + k < N) {
                    if ((al >> (8 * k)) & 0xff) P.alive_list[pos++] = (uint32_t)(i0 + k);
                    else P.alive_list[dpos++] = (uint32_t)(i0 + k);
                }
            }
            off += ttot;
        }
    }
}

static int g_head_blocks_per_sm = -1;

int launch_head(cudaStream_t st, const PopDev& P, int sm_count)
{
    if (g_head_blocks_per_sm < 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, head_kernel, HEAD_THREADS, 0) != cudaSuccess || nb < 1) nb = 1;
        g_head_blocks_per_sm = nb > 4 ? 4 : nb;
    }
    unsigned Gmax = (unsigned)(g_head_blocks_per_sm * sm_count);
    unsigned per = (P.ntiles + Gmax - 1) / Gmax;            // tiles per CTA, balanced
    unsigned G = (P.ntiles + per - 1) / per;
    if (G < 1) G = 1;
    PopDev Pc = P;
    void* args[] = { (void*)&Pc };
    cudaLaunchCooperativeKernel((const void*)head_kernel, dim3(G), dim3(HEAD_THREADS), args, 0, st);
    return 1;
}

}  // namespace abcdez
