// head.cu -- the "head" of one abcdesmc! iteration (src/abcdez_smc.jl:301-324 of the reference) as ONE
// persistent cooperative kernel:
//   eps schedule   quantile(delta[alive], alpha), clamp                 :301
//   reweighting    abcdesmc_update_ws!, wprod, wnorm, Wns, alive         :59-83, :305-311
//   logZ / ESS     logZ += log(wnorm); ess = 1/sum(Wns^2); resample flag :315, :323-324
//   alive list     compaction (the O(1) replacement of the wsample scans :121,125)
//
// Round-2 structure.  At 10^6 particles the whole working set (delta, W, alive: 17 MB) lives in L2, so the
// round-1 kernel (five passes, each a chain of load -> reduce -> barrier per 1024-particle tile, four grid
// barriers) was bound by latency, not bandwidth: 50-60 us against ~6 us of L2 traffic.  Now
//   * a CTA keeps its share of the population ON CHIP: up to KT = 4 tiles (16 distance keys and 16 weights per
//     thread) are loaded once and reused by every phase (larger shares are processed in rounds of KT tiles and
//     re-read per phase -- bandwidth bound then).  The share lives in shared memory as thread-private slots
//     ([tile][k][thread], conflict-free, no barriers needed), not in registers: a first version with register
//     arrays had to unroll every phase 16 times and came out at 26 000 SASS instructions, 37 % of whose stall
//     samples were instruction-fetch misses (profiles/README.md); loops over tiles keep the code ~5x smaller;
//   * the per-tile reductions of a round are batched (one pair of CTA barriers for KT tiles);
//   * THREE grid barriers in the common case: after the window histogram, after the candidate compaction,
//     after the per-tile weight sums.  The alive mask is decided together with wprod (wprod / wnorm > 0 <=>
//     wprod > 0 for a finite positive wnorm without underflow -- checked, ABCDEZ_ERR_BAD_ARG otherwise), so
//     the alive counts travel with the weight sums and normalisation, ESS partials and the alive-list
//     compaction need no barrier of their own; the last CTA to finish adds up the ESS partials.
//   * the eps select uses ONE adaptive window: 2047 bins of 2^s key units below key(eps_prev), with s taken
//     from the distance between the previous two thresholds in key space, so that the new quantile falls
//     near the middle of the window and its bin holds ~10^-4 of the population (30-80 candidate keys at 10^6
//     particles; the fixed 22-bit-prefix window of round 1 left 2000-5000).  The map key -> bin is monotone,
//     so the select stays exact; when the rank falls outside the window (first iteration, eps_prev = Inf, a
//     collapsing population) two generic 11-bit digit passes run instead.
//   * sharded runs: no "CTA 0 exchanges, then a second grid barrier".  After the local grid
//     barrier CTA q posts this rank's record to rank q (R CTAs post in parallel), reducer CTAs sum 256 bins
//     each into a local flag-in-word array, and EVERY CTA of every rank polls what it needs out of local
//     memory: the reduced histogram, the candidate keys of all ranks, the weight sums.  Every decision is
//     recomputed redundantly from the same integers / rank-ordered FP64 sums, so all ranks agree bit for bit.
// Reductions keep the per-tile partials and the summation order of the unfused stage kernels
// (bookkeeping.cu): both paths give bit-identical eps, W, wnorm, logZ and ESS.
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
#ifndef ABCDEZ_HEAD_KT
#define ABCDEZ_HEAD_KT 4
#endif
constexpr int HEAD_MIN_BLOCKS = ABCDEZ_HEAD_MIN_BLOCKS;   // resident CTAs per SM the register budget is cut for
constexpr int KT = ABCDEZ_HEAD_KT;                // tiles per round: KT x 4 particles per thread stay on chip
constexpr int KE = KT * 4;
constexpr int CAND_SMEM = KT >= 4 ? 2048 : 1024;                   // candidates staged in shared memory for the per-CTA tail; longer
                                                  // lists are first refined grid-cooperatively, one digit per round
constexpr int CAND_DIRECT = 256;                  // ... and at most this many are ranked directly (one per thread)
constexpr int NW = HEAD_THREADS / 32;

struct HeadSmem {
    unsigned hist[SEL_BINS];
    unsigned long long cand[CAND_SMEM];
    unsigned part[HEAD_THREADS];
    double red[KT][NW];
    double red1[32];
    unsigned wcnt[KT][NW];
    unsigned long long wmin[32];
    unsigned long long mn, mx, mab;
    unsigned found_bin, found_before, found_cnt, flag, flag2;
    int xflag;
};

__device__ __forceinline__ int pass_shift(int p) { const int s[6] = { 53, 42, 31, 20, 9, 0 }; return s[p]; }
__device__ __forceinline__ int pass_bins(int p) { return p == 5 ? 512 : 2048; }
__device__ __forceinline__ unsigned long long pass_himask_after(int p) { return p == 5 ? ~0ull : (~0ull << pass_shift(p)); }

// pick the bin holding rank `rank` in a SEL_BINS histogram of which this thread holds bins tid*8 .. tid*8+7
// in loc[]; every thread of the CTA participates and gets the same answer.  bin == 0xffffffff: rank beyond
// the total.  s->found_cnt: keys in the picked bin (stable until the next call).
__device__ __noinline__ void pick_bin(const unsigned (&loc)[SEL_BINS / HEAD_THREADS], unsigned long long rank, HeadSmem* s,
                                      unsigned& bin, unsigned long long& before)
{
    constexpr int per = SEL_BINS / HEAD_THREADS;
    unsigned tot = 0;
#pragma unroll
    for (int k = 0; k < per; ++k) tot += loc[k];
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
    // (populations are below 2^31 particles: a rank beyond 32 bits is beyond every total)
    const unsigned r32 = rank > 0xfffffffeull ? 0xffffffffu : (unsigned)rank;
    if (r32 - bef < tot && r32 >= bef) {                  // the rank falls into one of this thread's bins
        unsigned cum = bef;
#pragma unroll
        for (int k = 0; k < per; ++k) {
            if (r32 - cum < loc[k]) { s->found_bin = threadIdx.x * per + k; s->found_before = cum; s->found_cnt = loc[k]; break; }
            cum += loc[k];
        }
    }
    __syncthreads();
    bin = s->found_bin; before = s->found_before;
}

__device__ __forceinline__ void load_bins_global(const unsigned* h, int nbins, unsigned (&loc)[SEL_BINS / HEAD_THREADS])
{
#pragma unroll
    for (int k = 0; k < SEL_BINS / HEAD_THREADS; ++k) { int b = threadIdx.x * (SEL_BINS / HEAD_THREADS) + k; loc[k] = b < nbins ? __ldcg(&h[b]) : 0u; }
}
__device__ __forceinline__ void load_bins_shared(const unsigned* h, int nbins, unsigned (&loc)[SEL_BINS / HEAD_THREADS])
{
#pragma unroll
    for (int k = 0; k < SEL_BINS / HEAD_THREADS; ++k) { int b = threadIdx.x * (SEL_BINS / HEAD_THREADS) + k; loc[k] = b < nbins ? h[b] : 0u; }
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
    double t = block_sum(v, s->red1);
    __syncthreads();
    if (threadIdx.x == 0) s->red1[0] = t;
    __syncthreads();
    t = s->red1[0];
    __syncthreads();
    return t;
}

// ---------------------------------------------------------------------------------------------------
// tail: the rank-th smallest (0-based) of the M listed RELATIVE keys (all below 2^42), whether the next
// order statistic ties with it, and else the smallest listed key above it.  Entirely inside the CTA; every
// thread gets the same answer.  L: shared memory.
// ---------------------------------------------------------------------------------------------------
__device__ __noinline__ void tail_select(const unsigned long long* L, unsigned M, unsigned long long rank, HeadSmem* s,
                                         unsigned long long& a_rel, bool& tie, bool& has_b, unsigned long long& b_rel)
{
    const unsigned tid = threadIdx.x;
    const unsigned long long rank0 = rank;
    unsigned long long prefix = 0ull, himask = ~0ull << 42;
    bool found = false;
    if (M <= (unsigned)CAND_DIRECT) {
        // few keys (the usual case with the adaptive window): rank them directly, one key per thread
        __syncthreads();
        if ((tid & ~31u) < M) {                             // (warps without a key skip the M-step loop)
            const unsigned long long my = tid < M ? L[tid] : ~0ull;
            unsigned r = 0;
#pragma unroll 4
            for (unsigned k = 0; k < M; ++k) { const unsigned long long o = L[k]; r += (o < my || (o == my && k < tid)) ? 1u : 0u; }
            if (tid < M && (unsigned long long)r == rank) s->mn = my;
        }
        __syncthreads();
        prefix = s->mn;
        __syncthreads();
        found = true;
    }
    for (int p = 2; p < 6 && !found; ++p) {
        const int shift = pass_shift(p), nbins = pass_bins(p);
        hist_clear(s);
#pragma unroll 1
        for (unsigned i = tid; i < M; i += HEAD_THREADS) {
            unsigned long long key = L[i];
            if ((key & himask) == prefix) atomicAdd(&s->hist[(unsigned)((key >> shift) & (unsigned long long)(nbins - 1))], 1u);
        }
        __syncthreads();
        unsigned bin; unsigned long long before;
        unsigned loc[SEL_BINS / HEAD_THREADS];
        load_bins_shared(s->hist, nbins, loc);
        pick_bin(loc, rank, s, bin, before);
        const unsigned in_bin = s->found_cnt;
        prefix |= (unsigned long long)bin << shift;
        himask = pass_himask_after(p);
        rank -= before;
        if (in_bin <= 32u && p < 5) {
            // few keys left: gather them and rank them directly instead of running the remaining digit passes
            __syncthreads();
            if (tid == 0) s->flag = 0u;
            __syncthreads();
#pragma unroll 1
            for (unsigned i = tid; i < M; i += HEAD_THREADS) {
                unsigned long long key = L[i];
                if ((key & himask) == prefix) s->wmin[atomicAdd(&s->flag, 1u)] = key;
            }
            __syncthreads();
            if (tid < 32) {
                const unsigned n = s->flag;
                const unsigned long long my = tid < n ? s->wmin[tid] : ~0ull;
                unsigned r = 0;
                for (unsigned k = 0; k < n; ++k) {
                    const unsigned long long o = s->wmin[k];
                    r += (o < my || (o == my && k < tid)) ? 1u : 0u;
                }
                if (tid < n && (unsigned long long)r == rank) s->mn = my;
            }
            __syncthreads();
            prefix = s->mn;
            __syncthreads();
            break;
        }
    }
    a_rel = prefix;
    // #{keys <= v[j]} and min{key > v[j]} on the list
    unsigned cnt = 0; unsigned long long mn = ~0ull;
#pragma unroll 1
    for (unsigned i = tid; i < M; i += HEAD_THREADS) {
        unsigned long long key = L[i];
        if (key <= a_rel) cnt++; else mn = key < mn ? key : mn;
    }
    cnt = warp_sum_u(cnt); mn = warp_min_u64(mn);
    __syncthreads();
    if ((tid & 31) == 0) { s->part[tid >> 5] = cnt; s->wmin[tid >> 5] = mn; }
    __syncthreads();
    cnt = 0; mn = ~0ull;
    for (int w = 0; w < NW; ++w) { cnt += s->part[w]; mn = s->wmin[w] < mn ? s->wmin[w] : mn; }
    __syncthreads();
    tie = (unsigned long long)cnt >= rank0 + 2;          // v[j+1] ties with v[j]
    has_b = mn != ~0ull;
    b_rel = mn;
}

// ---------------------------------------------------------------------------------------------------
// sharded runs: flag-in-word exchanges in which every CTA takes part (comm.cuh has the word format).
// `seq` is the exchange's sequence number; every CTA of every rank counts the same sequence of exchanges
// from the value X.seq[1] held at kernel start, so nobody has to publish it inside the kernel.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long* hg_base(const XchgDev& X)
{
    return reinterpret_cast<unsigned long long*>(X.mbox[X.rank] + XCHG_HG_OFF);
}
__device__ __forceinline__ unsigned long long* gcand_region(char* mbox, int src)
{
    return reinterpret_cast<unsigned long long*>(mbox + XCHG_GCAND_OFF) + 2 * (size_t)src * XCHG_GCAND_PER;
}

// post: CTA `role` of this rank stores the record (nh 64-bit header values + an nb-bin histogram from global
// memory, either may be 0) into the mailbox of rank `role`.  Called by all threads of the CTAs whose role
// loop covers role < world.
__device__ __forceinline__ void post_record(const XchgDev& X, int to, unsigned long long seq, const unsigned long long* hdr, int nh,
                                            const unsigned* H, int nb)
{
    const unsigned slot = (unsigned)(seq & (XCHG_RING - 1)), flag = (unsigned)seq;
    unsigned long long* e = ll_entry(X.mbox[to], slot, X.rank);
    const int tid = threadIdx.x;
    if (tid < 2 * nh) st_relaxed_sys(e + tid, ll_pack((unsigned)(hdr[tid >> 1] >> (32 * (tid & 1))), flag));
    for (int k = tid; k < nb; k += HEAD_THREADS) st_relaxed_sys(e + LL_HDR + k, ll_pack(__ldcg(&H[k]), flag));
}

// collect: every thread polls the nh header values of every rank out of the local mailbox: out[r * nh + k]
// (each thread its own copy in registers would cost too many; thread t < world * nh polls value t into shared
// memory).  Ends with a CTA barrier.
__device__ __forceinline__ void collect_headers(const XchgDev& X, Ctrl* c, unsigned long long seq, int nh, unsigned long long* out, bool& ok)
{
    const unsigned slot = (unsigned)(seq & (XCHG_RING - 1)), flag = (unsigned)seq;
    const int tid = threadIdx.x;
    if (tid < X.world * nh) {
        const int r = tid / nh, k = tid - r * nh;
        const unsigned long long* e = ll_entry(X.mbox[X.rank], slot, r) + 2 * k;
        const unsigned lo = ll_poll(e, flag, ok), hi = ll_poll(e + 1, flag, ok);
        out[tid] = ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();
}

// sharded runs: all-reduce of a histogram (+ optionally the first exchange's header: extrema, error).  Every CTA
// of every rank calls; on return loc[] holds bins tid*8 .. tid*8+7 of the sum over all ranks and, when `first`,
// s_x the header records of all ranks.  No grid barrier: CTA q posts this rank's record to rank q, reducer CTAs
// sum 256 bins each into the local flag-in-word array, every CTA polls the sums from there.
static __device__ __noinline__ void all_reduce_hist(const PopDev& P, Ctrl* c, unsigned long long seq, const unsigned* H, int nbins, bool first,
                                                    unsigned (&loc)[SEL_BINS / HEAD_THREADS], unsigned long long* s_h, unsigned long long* s_x, bool& xok)
{
    const unsigned tid = threadIdx.x, G = gridDim.x;
    const int R = P.x.world;
    const unsigned slot = (unsigned)(seq & (XCHG_RING - 1)), flag = (unsigned)seq;
    if (tid == 0) {
        int e0 = c->err, e1 = __ldcg(&c->acc.err);
        s_h[0] = __ldcg(&c->acc.dmin_key); s_h[1] = __ldcg(&c->acc.dmax_key);
        s_h[2] = (unsigned long long)(e0 > e1 ? e0 : e1); s_h[3] = 0ull;
    }
    __syncthreads();
    for (int role = blockIdx.x; role < R; role += G) post_record(P.x, role, seq, s_h, first ? 4 : 0, H, nbins);
    unsigned long long* hg = hg_base(P.x);
    for (int role = blockIdx.x; role < SEL_BINS / HEAD_THREADS; role += G) {       // 8 reducer roles x 256 bins
        const int b = role * HEAD_THREADS + tid;
        if (b < nbins) {
            unsigned tot = 0;
#pragma unroll 1
            for (int r = 0; r < R; ++r) tot += ll_poll(ll_entry(P.x.mbox[P.x.rank], slot, r) + LL_HDR + b, flag, xok);
            st_relaxed_sys(hg + b, ll_pack(tot, flag));
        }
    }
    if (first) collect_headers(P.x, c, seq, 4, s_x, xok);
#pragma unroll
    for (int k = 0; k < SEL_BINS / HEAD_THREADS; ++k) {
        const int b = tid * (SEL_BINS / HEAD_THREADS) + k;
        loc[k] = b < nbins ? ll_poll(hg + b, flag, xok) : 0u;
    }
}

// The CTA's share of one round in shared memory: thread tid owns the particles tile * 1024 + tid * 4 + k of each
// of the round's tiles, as [tile][k][tid] so that its accesses are conflict-free.  Only the owning thread ever
// touches an entry, so the phases need no CTA barrier for them: this is register-like storage that a loop can index.
constexpr unsigned long long DEAD_KEY = ~0ull;            // dead or out-of-range particle (the key of one NaN pattern; NaNs are an error anyway)
struct HeadShare {
    unsigned long long key[KT][4][HEAD_THREADS];
    double w[KT][4][HEAD_THREADS];
};

// distances of the round's tiles [tb, tb + nt) -> keys; with ext != nullptr also extrema(delta) over all particles
// and the NaN check (ext[0] = min key, ext[1] = max key, ext[2] = NaN among alive)
static __device__ __noinline__ void load_keys(const PopDev& P, const double* __restrict__ dl, unsigned tb, unsigned nt, HeadShare* sh,
                                              unsigned long long* ext, bool with_w)
{
    const unsigned tid = threadIdx.x;
    const uint32_t N = P.N;
    unsigned long long kmn = ~0ull, kmx = 0ull, nan_seen = 0ull;
    // every global load of the round is issued before the first use: one L2 round trip, not one per tile
    double v[KT][4]; uint32_t al[KT]; double w[KT][4];
#pragma unroll
    for (int t = 0; t < KT; ++t) {
        const size_t i0 = (size_t)(tb + t) * TILE + (size_t)tid * 4;
        al[t] = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[t][k] = 0.0; w[t][k] = 0.0; }
        if ((unsigned)t < nt && i0 < N) {
            load4_f64(dl, i0, N, v[t]); al[t] = load4_u8(P.alive, i0, N);
            if (with_w) load4_f64(P.W, i0, N, w[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < KT; ++t) {
        if ((unsigned)t >= nt) break;
        const size_t i0 = (size_t)(tb + t) * TILE + (size_t)tid * 4;
        const int nvalid = i0 + 4 <= N ? 4 : (i0 < N ? (int)(N - i0) : 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool valid = k < nvalid, ok = valid && ((al[t] >> (8 * k)) & 0xff);
            const unsigned long long key = f64_key(v[t][k]);
            if (ext && valid) { kmn = key < kmn ? key : kmn; kmx = key > kmx ? key : kmx; if (ok && isnan(v[t][k])) nan_seen = 1ull; }
            sh->key[t][k][tid] = ok ? key : DEAD_KEY;
            if (with_w) sh->w[t][k][tid] = w[t][k];
        }
    }
    if (ext) { ext[0] = kmn; ext[1] = kmx; ext[2] = nan_seen; }
}

__global__ void __launch_bounds__(HEAD_THREADS, HEAD_MIN_BLOCKS) head_kernel(const __grid_constant__ PopDev P, const unsigned per)
{
    Ctrl* c = P.ctrl;
    if (c->stop) return;                                    // uniform: nobody reaches a grid barrier
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char head_raw[];
    HeadSmem& s = *reinterpret_cast<HeadSmem*>(head_raw);
    HeadShare* sh = reinterpret_cast<HeadShare*>(head_raw + sizeof(HeadSmem));
    __shared__ unsigned long long s_x[XCHG_MAXR * 8];       // exchanged header records: every rank's
    __shared__ unsigned long long s_h[4];                   // ... and this rank's, before it is posted
    const unsigned G = gridDim.x, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const uint32_t N = P.N, ntiles = P.ntiles;
    const unsigned t0 = blockIdx.x * per < ntiles ? blockIdx.x * per : ntiles;
    const unsigned t1 = t0 + per < ntiles ? t0 + per : ntiles;
    const unsigned rounds = (per + KT - 1) / KT;
    const bool single = rounds == 1;                        // the CTA's whole share stays in shared memory
    auto round_tiles = [&](unsigned r, unsigned& tb) -> unsigned { tb = t0 + r * KT; return tb >= t1 ? 0u : (t1 - tb < (unsigned)KT ? t1 - tb : (unsigned)KT); };
    // schedule scalars of the previous iteration (CTA 0 overwrites them after the third barrier)
    const int cur = c->cur, kind = c->kind;
    const double eps_prev = c->eps, eps_old = c->eps_k, eps_target = c->eps_target, q_gamma = c->q_gamma;
    const unsigned n_alive_prev = c->n_alive_g;
    const int win_shift_in = c->win_shift;
    const bool sharded = P.x.world > 1;
    const int R = P.x.world;
    unsigned long long seq = sharded ? __ldcg(P.x.seq + 1) : 0ull;
    bool xok = true;
    const double* __restrict__ dl = P.delta[cur];
    unsigned* H1 = P.sel_hist; unsigned* H2 = P.sel_hist + SEL_BINS; unsigned* HW = P.sel_hist + 6 * SEL_BINS;
    unsigned long long rank = c->sel_rank;
    const unsigned long long rank_all = rank;

    // ---- first pass: window histogram (+ extrema(delta) over all particles, NaN check) ---------------------
    // bin 2047 - ((kp - key) >> sft) for the 2047 * 2^sft keys at and below kp = key(eps_prev); bin 0 collects
    // everything further below, keys above kp stay outside.  The map is monotone: an exact select.
    const bool windowed = isfinite(eps_prev) && eps_prev > 0.0;
    const unsigned long long kp = f64_key(windowed ? eps_prev : 1.0);
    const int sft = (win_shift_in > 0 && win_shift_in <= 43) ? win_shift_in - 1 : 42;
    unsigned bin; unsigned long long before;
    unsigned long long lo_key = 0ull, hi_key = 0ull;        // the candidates' key range (inclusive)
    bool have_range = false;
    {
        hist_clear(&s);
        unsigned long long kmn = ~0ull, kmx = 0ull, nan_seen = 0ull;
        unsigned n_low = 0;
        for (unsigned r = 0; r < rounds; ++r) {
            unsigned tb; const unsigned nt = round_tiles(r, tb);
            unsigned long long ext[3];
            load_keys(P, dl, tb, nt, sh, ext, single);          // (single round: the weights come along, for pass A)
            kmn = ext[0] < kmn ? ext[0] : kmn; kmx = ext[1] > kmx ? ext[1] : kmx; nan_seen |= ext[2];
#pragma unroll 1
            for (unsigned t = 0; t < nt; ++t) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned long long key = sh->key[t][k][tid];
                    if (windowed) {
                        if (key <= kp) {                                       // (dead entries are above every key)
                            const unsigned long long dd = (kp - key) >> sft;
                            if (dd >= 2047ull) n_low++;                        // most keys: counted in a register
                            else atomicAdd(&s.hist[2047u - (unsigned)dd], 1u);
                        }
                    } else {
                        hist_add(s.hist, key != DEAD_KEY, (unsigned)(key >> 53));   // generic pass 1: digit 53..63
                    }
                }
            }
        }
        if (tid == 0) { s.mn = ~0ull; s.mx = 0ull; }
        __syncthreads();
        kmn = warp_min_u64(kmn); kmx = warp_max_u64(kmx);
        n_low = warp_sum_u(n_low);
        if (lane == 0) {
            atomicMin(&s.mn, kmn); atomicMax(&s.mx, kmx);
            if (n_low) atomicAdd(&s.hist[0], n_low);
        }
        if (nan_seen) atomicMax(&c->acc.err, (int)ABCDEZ_ERR_NAN_DISTANCE);
        hist_flush(&s, windowed ? HW : H1, 2048);
        if (tid == 0) { atomicMin(&c->acc.dmin_key, s.mn); atomicMax(&c->acc.dmax_key, s.mx); }
    }
    grid.sync();                                                                                   // ---- barrier 1
    // the whole population's histogram: bins tid*8 .. tid*8+7 into registers
    unsigned loc[SEL_BINS / HEAD_THREADS];
    auto first_exchange = [&](unsigned* H) {
        if (sharded) {
            all_reduce_hist(P, c, ++seq, H, 2048, true, loc, s_h, s_x, xok);
            unsigned long long mn = ~0ull, mx = 0ull, e = 0ull;
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                mn = s_x[4 * r] < mn ? s_x[4 * r] : mn; mx = s_x[4 * r + 1] > mx ? s_x[4 * r + 1] : mx;
                e = s_x[4 * r + 2] > e ? s_x[4 * r + 2] : e;
            }
            if (blockIdx.x == 0 && tid == 0) {
                __stcg(&c->acc.dmin_key, mn); __stcg(&c->acc.dmax_key, mx);
                if (e) { c->acc.err = (int)e; if (!c->err) c->err = (int)e; }
            }
        } else {
            load_bins_global(H, 2048, loc);
        }
        if (blockIdx.x == 0 && tid == 0) patch_extrema(P, c);                  // ranges_eps of the previous record, :363
    };
    unsigned long long prefix = 0ull, himask = 0ull;
    if (windowed) {
        first_exchange(HW);
        pick_bin(loc, rank, &s, bin, before);
        if (bin >= 1u && bin != 0xffffffffu) {              // inside the window
            const unsigned long long dd = 2047ull - bin;
            hi_key = kp - (dd << sft);
            const unsigned long long span = (dd + 1ull) << sft;
            lo_key = kp >= span ? kp - span + 1ull : 0ull;
            rank -= before;
            have_range = true;
        }
    }
    if (!have_range) {
        // ---- generic path: two 11-bit digit passes (first iteration, or the rank fell outside the window) -----
        rank = rank_all;
        if (windowed) {
            // the window pass has the extrema already; this pass only builds the leading-digit histogram
            hist_clear(&s);
            for (unsigned r = 0; r < rounds; ++r) {
                unsigned tb; const unsigned nt = round_tiles(r, tb);
                if (!single) load_keys(P, dl, tb, nt, sh, nullptr, false);
#pragma unroll 1
                for (unsigned t = 0; t < nt; ++t)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const unsigned long long key = sh->key[t][k][tid]; hist_add(s.hist, key != DEAD_KEY, (unsigned)(key >> 53)); }
            }
            hist_flush(&s, H1, 2048);
            grid.sync();
            if (sharded) all_reduce_hist(P, c, ++seq, H1, 2048, false, loc, s_h, s_x, xok); else load_bins_global(H1, 2048, loc);
        } else {
            first_exchange(H1);
        }
        pick_bin(loc, rank, &s, bin, before);
        if (bin == 0xffffffffu) {                           // empty alive set (uniform decision)
            grid.sync();                                    // every CTA has read the histograms
            if (blockIdx.x == 0 && tid == 0) {
                if (!c->err) c->err = ABCDEZ_ERR_NO_ALIVE;
                if (sharded) __stcg(P.x.seq + 1, seq);
            }
            if (blockIdx.x == 0) for (int b = tid; b < 7 * SEL_BINS; b += HEAD_THREADS) P.sel_hist[b] = 0u;
            return;
        }
        prefix = (unsigned long long)bin << 53; himask = ~0ull << 53;
        rank -= before;
        // generic pass 2: digit 42..52 among the keys of that bin
        hist_clear(&s);
        for (unsigned r = 0; r < rounds; ++r) {
            unsigned tb; const unsigned nt = round_tiles(r, tb);
            if (!single) load_keys(P, dl, tb, nt, sh, nullptr, false);
#pragma unroll 1
            for (unsigned t = 0; t < nt; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned long long key = sh->key[t][k][tid];
                    hist_add(s.hist, key != DEAD_KEY && (key & himask) == prefix, (unsigned)((key >> 42) & 2047ull));
                }
        }
        hist_flush(&s, H2, 2048);
        grid.sync();
        if (sharded) all_reduce_hist(P, c, ++seq, H2, 2048, false, loc, s_h, s_x, xok); else load_bins_global(H2, 2048, loc);
        pick_bin(loc, rank, &s, bin, before);
        prefix |= (unsigned long long)bin << 42;
        rank -= before;
        lo_key = prefix; hi_key = prefix | ((1ull << 42) - 1ull);
    }

    // ---- compact the candidates (keys in [lo_key, hi_key], stored relative to lo_key: < 2^42); min key above ----
    {
        unsigned long long* cand = P.cand[0];
        unsigned long long mab = ~0ull, cmn = ~0ull, cmx = 0ull;
        for (unsigned r = 0; r < rounds; ++r) {
            unsigned tb; const unsigned nt = round_tiles(r, tb);
            if (!single) load_keys(P, dl, tb, nt, sh, nullptr, false);
#pragma unroll 1
            for (unsigned t = 0; t < nt; ++t) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned long long key = sh->key[t][k][tid];
                    const bool is_c = key >= lo_key && key <= hi_key;
                    if (key > hi_key && key != DEAD_KEY) mab = key < mab ? key : mab;
                    const unsigned m = __ballot_sync(0xffffffffu, is_c);
                    if (m) {                                                   // (rare: ~10^-4 of the keys)
                        unsigned base = 0;
                        if (lane == 0) base = (unsigned)atomicAdd(&c->acc.cand_count[0], (unsigned long long)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (is_c) {
                            const unsigned long long rk = key - lo_key;
                            cand[base + __popc(m & ((1u << lane) - 1u))] = rk;
                            cmn = rk < cmn ? rk : cmn; cmx = rk > cmx ? rk : cmx;
                        }
                    }
                }
            }
        }
        if (tid == 0) { s.mn = ~0ull; s.mx = 0ull; s.mab = ~0ull; }
        __syncthreads();
        mab = warp_min_u64(mab);
        if (lane == 0 && mab != ~0ull) atomicMin(&s.mab, mab);
        if (cmn != ~0ull) { atomicMin(&s.mn, cmn); atomicMax(&s.mx, cmx); }       // (only the few threads that hold a candidate)
        __syncthreads();
        if (tid == 0) {
            if (s.mab != ~0ull) atomicMin(&c->acc.min_above, s.mab);
            if (s.mn != ~0ull) { atomicMin(&c->acc.cand_min[0], s.mn); atomicMax(&c->acc.cand_max[0], s.mx); }
        }
    }
    grid.sync();                                                                                   // ---- barrier 2
    // candidate-list generation g: list P.cand[g & 1] (relative keys), accumulators cand_count/min/max[g].
    // M: this rank's candidates; Mg, cmin, cmax, min_above: the whole population's.  Sharded: CTA q posts this
    // rank's counts (and its keys while they fit its region of the gather area) to rank q; every CTA collects.
    int g = 0, p_next = 2;
    unsigned M = 0; unsigned long long Mg = 0, cmin = ~0ull, cmax = 0ull, min_above = ~0ull;
    bool need_refine = false;                               // the candidates do not fit the per-CTA tail yet
    auto exchange_cands = [&]() {
        M = (unsigned)__ldcg(&c->acc.cand_count[g]);
        if (!sharded) {
            Mg = M; cmin = __ldcg(&c->acc.cand_min[g]); cmax = __ldcg(&c->acc.cand_max[g]); min_above = __ldcg(&c->acc.min_above);
            need_refine = Mg > (unsigned long long)CAND_SMEM;
            return;
        }
        ++seq;
        const unsigned flag = (unsigned)seq;
        if (tid == 0) {
            s_h[0] = M; s_h[1] = __ldcg(&c->acc.cand_min[g]); s_h[2] = __ldcg(&c->acc.cand_max[g]); s_h[3] = __ldcg(&c->acc.min_above);
        }
        __syncthreads();
        for (int role = blockIdx.x; role < R; role += G) {
            post_record(P.x, role, seq, s_h, 4, nullptr, 0);
            if (M <= (unsigned)XCHG_GCAND_PER) {
                unsigned long long* gr = gcand_region(P.x.mbox[role], P.x.rank);
                const unsigned long long* src = P.cand[g & 1];
#pragma unroll 1
                for (unsigned i = tid; i < M; i += HEAD_THREADS) {
                    const unsigned long long key = __ldcg(&src[i]);
                    st_relaxed_sys(gr + 2 * i, ll_pack((unsigned)key, flag));
                    st_relaxed_sys(gr + 2 * i + 1, ll_pack((unsigned)(key >> 32), flag));
                }
            }
        }
        collect_headers(P.x, c, seq, 4, s_x, xok);
        Mg = 0; cmin = ~0ull; cmax = 0ull; min_above = ~0ull;
        bool fits = true;
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const unsigned long long m = s_x[4 * r];
            Mg += m; fits = fits && m <= (unsigned long long)XCHG_GCAND_PER;
            cmin = s_x[4 * r + 1] < cmin ? s_x[4 * r + 1] : cmin; cmax = s_x[4 * r + 2] > cmax ? s_x[4 * r + 2] : cmax;
            min_above = s_x[4 * r + 3] < min_above ? s_x[4 * r + 3] : min_above;
        }
        need_refine = !(fits && Mg <= (unsigned long long)CAND_SMEM);          // (uniform over ranks)
    };
    exchange_cands();

    // ---- large candidate lists: refine grid-cooperatively, one digit per round (rare with the adaptive window) ----
    prefix = 0ull; himask = ~0ull << 42;                    // over relative keys from here on
    while (need_refine && cmin != cmax && p_next < 6) {
        const int shift = pass_shift(p_next), nbins = pass_bins(p_next);
        unsigned* H = P.sel_hist + (size_t)p_next * SEL_BINS;
        const unsigned long long* src = P.cand[g & 1]; unsigned long long* dst = P.cand[(g + 1) & 1];
        hist_clear(&s);
#pragma unroll 1
        for (size_t i = (size_t)blockIdx.x * HEAD_THREADS + tid; i < M; i += (size_t)G * HEAD_THREADS)
            atomicAdd(&s.hist[(unsigned)((__ldcg(&src[i]) >> shift) & (unsigned long long)(nbins - 1))], 1u);
        hist_flush(&s, H, nbins);
        grid.sync();
        if (sharded) all_reduce_hist(P, c, ++seq, H, nbins, false, loc, s_h, s_x, xok); else load_bins_global(H, nbins, loc);
        pick_bin(loc, rank, &s, bin, before);
        prefix |= (unsigned long long)bin << shift; himask = pass_himask_after(p_next);
        rank -= before;
        {
            unsigned long long mab = ~0ull, cmn = ~0ull, cmx = 0ull;
            const size_t rnds = ((size_t)M + (size_t)G * HEAD_THREADS - 1) / ((size_t)G * HEAD_THREADS);
#pragma unroll 1
            for (size_t r = 0; r < rnds; ++r) {
                const size_t i = r * (size_t)G * HEAD_THREADS + (size_t)blockIdx.x * HEAD_THREADS + tid;
                const unsigned long long key = i < M ? __ldcg(&src[i]) : 0ull;
                const unsigned long long hi = key & himask;
                const bool is_c = i < M && hi == prefix;
                if (i < M && hi > prefix) mab = key < mab ? key : mab;
                const unsigned m = __ballot_sync(0xffffffffu, is_c);
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
                if (mab != ~0ull) atomicMin(&c->acc.min_above, lo_key + mab);     // (absolute, like the first generation's)
                if (cmn != ~0ull) { atomicMin(&c->acc.cand_min[g + 1], cmn); atomicMax(&c->acc.cand_max[g + 1], cmx); }
            }
        }
        grid.sync();
        g++; p_next++;
        exchange_cands();
    }

    // ---- tail: v[j], tie test, v[j+1], type-7 interpolation, clamp (every CTA, same integers) ---------
    unsigned long long akey, bkey;
    if (cmin == cmax) {                                     // all candidates equal (discrete distances, ties)
        akey = lo_key + cmin;
        bkey = (rank + 1 < Mg) ? akey : min_above;
    } else {                                                // Mg <= CAND_SMEM (after p_next == 6 all keys are equal)
        if (sharded) {
            // every rank's candidates, from the local gather area: rank r's keys at its region, its count in s_x
            const unsigned flag = (unsigned)seq;
            unsigned off = 0;
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                const unsigned m = (unsigned)s_x[4 * r];
                const unsigned long long* gr = gcand_region(P.x.mbox[P.x.rank], r);
#pragma unroll 1
                for (unsigned i = tid; i < m; i += HEAD_THREADS) {
                    const unsigned lo = ll_poll(gr + 2 * i, flag, xok), hi = ll_poll(gr + 2 * i + 1, flag, xok);
                    s.cand[off + i] = ((unsigned long long)hi << 32) | lo;
                }
                off += m;
            }
        } else {
#pragma unroll 1
            for (unsigned i = tid; i < (unsigned)Mg; i += HEAD_THREADS) s.cand[i] = __ldcg(&P.cand[g & 1][i]);
        }
        __syncthreads();
        unsigned long long a_rel, b_rel; bool tie, has_b;
        // after refine rounds the list holds the keys that match (prefix, himask); the rank is relative to them
        tail_select(s.cand, (unsigned)Mg, rank, &s, a_rel, tie, has_b, b_rel);
        akey = lo_key + a_rel;
        bkey = tie ? akey : (has_b ? lo_key + b_rel : min_above);
    }
    double eps;
    {
        double a = key_f64(akey);
        double b = (n_alive_prev <= 1 || bkey == ~0ull) ? a : key_f64(bkey);
        double q = (isfinite(a) && isfinite(b)) ? a + q_gamma * (b - a) : (1.0 - q_gamma) * a + q_gamma * b;
        eps = fmax(fmin(q, eps_prev), eps_target);          // :301
        if (blockIdx.x == 0 && tid == 0) {
            c->q_a = a; c->q_b = b; c->q = q; c->eps = eps; c->sel_prefix = akey;
            // the next select's window: ~512-1023 bins between this threshold and the next quantile, if the
            // population keeps contracting at this rate in key space
            if (windowed && eps > 0.0 && eps < eps_prev) {
                const unsigned long long gap = kp - f64_key(eps);
                int sn = 63 - __clzll((long long)gap) - 9;
                sn = sn < 0 ? 0 : (sn > 42 ? 42 : sn);
                c->win_shift = sn + 1;
            }
        }
    }

    // ---- reweight pass A: ws, wprod (kept in shared memory when the share is a single round), per-tile sums and alive counts ----
    for (unsigned r = 0; r < rounds; ++r) {
        unsigned tb; const unsigned nt = round_tiles(r, tb);
        if (!single) load_keys(P, dl, tb, nt, sh, nullptr, false);
#pragma unroll 1
        for (unsigned t = 0; t < nt; ++t) {
            const size_t i0 = (size_t)(tb + t) * TILE + (size_t)tid * 4;
            double w[4] = { 0.0, 0.0, 0.0, 0.0 };
            if (single) {
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = sh->w[t][k][tid];           // prefetched with the distances
            } else if (i0 < N) load4_f64(P.W, i0, N, w);
            double acc = 0.0; unsigned cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned long long key = sh->key[t][k][tid];
                double wp_ = 0.0;
                if (key != DEAD_KEY) wp_ = w[k] * abck_ws(kind, eps, eps_old, key_f64(key));      // :75, :308
                w[k] = wp_;
                sh->w[t][k][tid] = wp_;
                acc += wp_;
                cnt += (wp_ > 0.0) ? 1u : 0u;
            }
            if (!single && i0 < N) store4_f64(P.W, i0, N, w);
            acc = warp_sum(acc); cnt = warp_sum_u(cnt);
            if (lane == 0) { s.red[t][wp] = acc; s.wcnt[t][wp] = cnt; }
        }
        __syncthreads();
        if (wp == 0) {                                      // warp 0 adds the warp totals of every tile (block_sum's second level)
#pragma unroll 1
            for (unsigned t = 0; t < nt; ++t) {
                const double tsum = warp_sum(lane < NW ? s.red[t][lane] : 0.0);
                const unsigned tc = warp_sum_u(lane < NW ? s.wcnt[t][lane] : 0u);
                if (lane == 0) { P.partial[tb + t] = tsum; P.tile_cnt[tb + t] = tc; }
            }
        }
        __syncthreads();
    }
    grid.sync();                                                                                   // ---- barrier 3
    double wnorm;
    unsigned n_alive, my_off;
    {
        double a = 0.0; unsigned tot = 0, pre = 0;
        for (unsigned b0 = tid; b0 < ntiles; b0 += 4 * HEAD_THREADS) {      // four independent loads in flight per thread; same summation order
            double pv[4]; unsigned tc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const unsigned b = b0 + j * HEAD_THREADS; pv[j] = b < ntiles ? __ldcg(&P.partial[b]) : 0.0; tc[j] = b < ntiles ? __ldcg(&P.tile_cnt[b]) : 0u; }
#pragma unroll
            for (int j = 0; j < 4; ++j) { const unsigned b = b0 + j * HEAD_THREADS; if (b < ntiles) { a += pv[j]; tot += tc[j]; if (b < t0) pre += tc[j]; } }
        }
        wnorm = block_sum_all(a, &s);                                                             // :309
        tot = warp_sum_u(tot); pre = warp_sum_u(pre);
        if (lane == 0) { s.wcnt[0][wp] = tot; s.part[wp] = pre; }
        __syncthreads();
        n_alive = 0; my_off = 0;
        for (int q = 0; q < NW; ++q) { n_alive += s.wcnt[0][q]; my_off += s.part[q]; }
        __syncthreads();
    }
    unsigned n_alive_g = n_alive;
    if (sharded) {                                          // wnorm = sum over ranks in rank order; alive counts of every rank
        ++seq;
        if (tid == 0) { s_h[0] = (unsigned long long)__double_as_longlong(wnorm); s_h[1] = (unsigned long long)n_alive; }
        __syncthreads();
        for (int role = blockIdx.x; role < R; role += G) post_record(P.x, role, seq, s_h, 2, nullptr, 0);
        collect_headers(P.x, c, seq, 2, s_x, xok);
        double tot = 0.0; unsigned ng = 0;
#pragma unroll 1
        for (int r = 0; r < R; ++r) { tot += __longlong_as_double((long long)s_x[2 * r]); ng += (unsigned)s_x[2 * r + 1]; }
        wnorm = tot; n_alive_g = ng;
        if (blockIdx.x == 0 && tid < (unsigned)R) c->rank_alive[tid] = (unsigned)s_x[2 * tid + 1];
    }
    if (blockIdx.x == 0) {
        if (tid == 0) {
            c->wnorm = wnorm; c->logZ += plog(wnorm);                                             // :315
            for (int q = 0; q < 6; ++q) { c->acc.cand_count[q] = 0ull; c->acc.cand_min[q] = ~0ull; c->acc.cand_max[q] = 0ull; }
            c->acc.min_above = ~0ull;
        }
        for (int b = tid; b < 7 * SEL_BINS; b += HEAD_THREADS) P.sel_hist[b] = 0u;              // every reader is past them
    }

    // ---- pass B: Wns, alive, sum(Wns^2) per tile, alive-first particle list -- no further grid barrier ----
    // w / wnorm through the reciprocal (pdiv_r, common.cuh: the same double as the division, 3 instructions; operands
    // outside its premises take the plain division)
    double* partial2 = P.partial + ntiles;
    const bool need_list = n_alive != N;
    const double rwn = (wnorm >= 0x1p-100 && wnorm <= 0x1p100 &&
                        ((unsigned long long)__double_as_longlong(wnorm) & 0x000fffffffffffffull) != 0x000fffffffffffffull) ? 1.0 / wnorm : 0.0;
    unsigned off = my_off;
    int mismatch = 0; double wal = 0.0; unsigned any = 0;
    for (unsigned r = 0; r < rounds; ++r) {
        unsigned tb; const unsigned nt = round_tiles(r, tb);
#pragma unroll 1
        for (unsigned t = 0; t < nt; ++t) {
            const size_t i0 = (size_t)(tb + t) * TILE + (size_t)tid * 4;
            double w[4] = { 0.0, 0.0, 0.0, 0.0 };
            if (!single && i0 < N) load4_f64(P.W, i0, N, w);
            double acc = 0.0; uint32_t nal = 0u; unsigned cnt = 0;
            const int nvalid = i0 + 4 <= N ? 4 : (i0 < N ? (int)(N - i0) : 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < nvalid) {
                    const double wp_ = single ? sh->w[t][k][tid] : w[k];
                    const double q = (wp_ == 0.0 && rwn != 0.0) ? wp_ : pdiv_r(wp_, wnorm, rwn);      // :310 (0 / wnorm without the detour)
                    const bool a = (q > 0.0);                                                     // :311
                    if (a != (wp_ > 0.0)) mismatch = 1;
                    if (a) { nal |= 1u << (8 * k); wal = q; any = 1; cnt++; }
                    acc += q * q;
                    w[k] = q;
                }
            }
            if (i0 < N) {
                store4_f64(P.W, i0, N, w);
                if (i0 + 3 < N) *reinterpret_cast<uint32_t*>(P.alive + i0) = nal;
                else { for (int k = 0; k < 4; ++k) if (i0 + k < N) P.alive[i0 + k] = (nal >> (8 * k)) & 0xff; }
            }
            // this thread's alive flags and their count stay in shared memory for the list below (its own key slots are free now)
            sh->key[t][0][tid] = ((unsigned long long)cnt << 32) | nal;
            acc = warp_sum(acc);
            unsigned incl = cnt;                                               // inclusive scan of the alive counts inside the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += u; }
            sh->key[t][1][tid] = incl;
            if (lane == 0) s.red[t][wp] = acc;
            if (lane == 31) s.wcnt[t][wp] = incl;
        }
        __syncthreads();
        if (wp == 0) {
#pragma unroll 1
            for (unsigned t = 0; t < nt; ++t) {
                const double tsum = warp_sum(lane < NW ? s.red[t][lane] : 0.0);
                if (lane == 0) partial2[tb + t] = tsum;
            }
        }
        if (need_list) {
            // alive particles first (index order), the dead ones behind them
#pragma unroll 1
            for (unsigned t = 0; t < nt; ++t) {
                const size_t i0 = (size_t)(tb + t) * TILE + (size_t)tid * 4;
                const unsigned long long pk = sh->key[t][0][tid];
                const uint32_t nal = (uint32_t)pk; const unsigned cnt = (unsigned)(pk >> 32), incl = (unsigned)sh->key[t][1][tid];
                unsigned woff = 0, ttot = 0;
#pragma unroll
                for (int q = 0; q < NW; ++q) { const unsigned wc = s.wcnt[t][q]; if (q < (int)wp) woff += wc; ttot += wc; }
                unsigned pos = off + woff + incl - cnt;                       // alive before element i0
                unsigned dpos = n_alive + ((unsigned)i0 - pos);               // dead before element i0
                const int nvalid = i0 + 4 <= N ? 4 : (i0 < N ? (int)(N - i0) : 0);
                const uint32_t i32 = (uint32_t)i0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < nvalid) {
                        // (the bounds only bite when the alive counts of pass A disagree with these flags: NaN weights,
                        // reported as an error below -- the list must still not be overrun)
                        if ((nal >> (8 * k)) & 0xff) { if (pos < N) P.alive_list[pos] = i32 + k; pos++; }
                        else { if (dpos < N) P.alive_list[dpos] = i32 + k; dpos++; }
                    }
                }
                off += ttot;
            }
        }
        __syncthreads();
    }
    if (any) c->acc.w_alive = wal;              // indicator kernels: every alive weight is this same double
    if (mismatch) atomicMax(&c->acc.alive_mismatch, 1);
    if (!xok) xchg_fail(c);

    // ---- the last CTA to finish: ESS, the decisions of :318-324 -------------------------------------------
    if (last_block(&c->acc.ticket[5], G)) {
        double a = 0.0;
        for (unsigned b0 = tid; b0 < ntiles; b0 += 4 * HEAD_THREADS) {
            double pv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const unsigned b = b0 + j * HEAD_THREADS; pv[j] = b < ntiles ? __ldcg(&partial2[b]) : 0.0; }
#pragma unroll
            for (int j = 0; j < 4; ++j) if (b0 + j * HEAD_THREADS < ntiles) a += pv[j];
        }
        double sumsq = block_sum_all(a, &s);
        if (sharded) {                                      // sum(Wns^2) in rank order, the common alive weight
            ++seq;
            if (tid < 32) {
                unsigned long long rec[2] = { (unsigned long long)__double_as_longlong(sumsq),
                                              (unsigned long long)__double_as_longlong(__ldcg(&c->acc.w_alive)) }, got[2];
                xchg_ll_warp_seq<2>(P.x, c, seq, rec, got);
                // lane r holds rank r's record: rank-ordered sum by lane 0
                double tot = 0.0; double wal_g = 0.0; bool have = false;
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
                    const unsigned long long s0 = __shfl_sync(0xffffffffu, got[0], r), s1 = __shfl_sync(0xffffffffu, got[1], r);
                    tot += __longlong_as_double((long long)s0);
                    if (!have && __ldcg(&c->rank_alive[r])) { wal_g = __longlong_as_double((long long)s1); have = true; }
                }
                if (tid == 0) {
                    if (have) __stcg(&c->acc.w_alive, wal_g);                // same double on every rank
                    sumsq = tot;
                }
            }
            if (tid == 0) __stcg(P.x.seq + 1, seq);
        }
        if (tid == 0) {
            // (wprod > 0) != (wprod / wnorm > 0) somewhere: weights far from normalised (never in a run); a NaN distance
            // (reported as such) also lands here, and the larger code wins
            if (__ldcg(&c->acc.alive_mismatch)) { c->acc.alive_mismatch = 0; atomicMax(&c->acc.err, (int)ABCDEZ_ERR_BAD_ARG); }
            ctrl_after_reweight(P, c, sumsq, n_alive, n_alive_g);                                 // :318-324
        }
    }
}

static inline size_t head_smem_bytes() { return sizeof(HeadSmem) + sizeof(HeadShare); }

// resident CTAs per SM of head_kernel, per device (cooperative launches need the whole grid resident)
static int head_blocks_per_sm(int device)
{
    static int cache[64] = { 0 };
    if (device >= 0 && device < 64 && cache[device] > 0) return cache[device];
    int nb = 0;
    cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem_bytes());
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, head_kernel, HEAD_THREADS, head_smem_bytes()) != cudaSuccess || nb < 1) nb = 1;
    nb = nb > HEAD_MIN_BLOCKS ? HEAD_MIN_BLOCKS : nb;
    if (device >= 0 && device < 64) cache[device] = nb;
    return nb;
}

// returns the number of launches (1), or -1 when the cooperative launch failed (cudaGetLastError has the reason)
int launch_head(cudaStream_t st, const PopDev& P, int sm_count)
{
    int device = 0;
    cudaGetDevice(&device);
    const unsigned Gmax = (unsigned)(head_blocks_per_sm(device) * sm_count);
    unsigned per = (P.ntiles + Gmax - 1) / Gmax;            // tiles per CTA
    if (per < 1) per = 1;
    unsigned G = (P.ntiles + per - 1) / per;
    if (G < 1) G = 1;
    PopDev Pc = P;
    void* args[] = { (void*)&Pc, (void*)&per };
    if (cudaLaunchCooperativeKernel((const void*)head_kernel, dim3(G), dim3(HEAD_THREADS), args, head_smem_bytes(), st) != cudaSuccess) return -1;
    return 1;
}

}  // namespace abcdez
