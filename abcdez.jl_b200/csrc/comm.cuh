// comm.cuh -- the cross-rank exchanges of a sharded run done INSIDE the kernels that
// produce their inputs, over NVLink peer memory: no extra launch, no host round trip.
//
// One process per GPU.  Every rank owns a mailbox in its own HBM (cudaMalloc, exported with
// cudaIpcGetMemHandle, mapped by every peer; comm.cu).  An exchange is an all-gather of a small record:
//   post   the calling CTA (or single thread) stores its record into slot (seq mod RING), entry
//          [my rank], of EVERY rank's mailbox (peer stores ride NVLink), fences at system scope and
//          then publishes the sequence number in the entry's flag word (st.release.sys);
//   wait   it then polls the R flag words of its OWN mailbox (ld.acquire.sys, local L2) until all of
//          them carry seq; the records are read back through L2 (__ldcg) and reduced in rank order,
//          so every rank computes bit-identical sums and takes identical schedule decisions.
// Every rank executes the same sequence of exchanges (the launch list and all device-side decisions
// are functions of the reduced values only), so the sequence numbers agree without negotiation.  A rank
// can be at most one exchange ahead of the slowest one (it cannot complete seq+1 before everyone has
// posted seq+1, which a rank does only after it has consumed seq), so a ring of 4 entries is never
// overwritten early.  A bounded spin (XCHG_TIMEOUT_NS) turns a lost peer into ABCDEZ_ERR_NCCL + stop
// instead of a hung GPU.
#pragma once
#include "internal.h"

namespace abcdez {

__device__ __forceinline__ unsigned long long xchg_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long* xchg_entry(char* mbox, unsigned slot, int src)
{
    return reinterpret_cast<unsigned long long*>(mbox + ((size_t)slot * XCHG_MAXR + (size_t)src) * XCHG_STRIDE);
}

// poll my own mailbox until rank `src` has published `seq` in `slot`; false on timeout
static __device__ __noinline__ bool xchg_wait_one(const XchgDev& X, unsigned slot, int src, unsigned long long seq)
{
    const unsigned long long* flag = xchg_entry(X.mbox[X.rank], slot, src);
    if (ld_acquire_sys(flag) == seq) return true;
    const unsigned long long t0 = xchg_now_ns();
    for (;;) {
#pragma unroll 1
        for (int spin = 0; spin < 64; ++spin)
            if (ld_acquire_sys(flag) == seq) return true;
        if (xchg_now_ns() - t0 > XCHG_TIMEOUT_NS) return false;
        __nanosleep(100);
    }
}

__device__ __forceinline__ void xchg_fail(Ctrl* c)
{
    c->err = ABCDEZ_ERR_NCCL; c->acc.err = ABCDEZ_ERR_NCCL; c->stop = 1;
}

// ---- single-thread exchange of up to XCHG_HDR-1 64-bit words ------------------------------------
// Called by exactly one thread of the grid (the one that runs the control logic).  Returns the
// local mailbox slot: the record of rank r starts at xchg_entry(base, slot, r) + 1.
__device__ inline unsigned xchg_small(const XchgDev& X, Ctrl* c, const unsigned long long* rec, int nwords)
{
    const unsigned long long seq = __ldcg(X.seq) + 1ull;
    const unsigned slot = (unsigned)(seq & (XCHG_RING - 1));
    for (int q = 0; q < X.world; ++q) {
        unsigned long long* e = xchg_entry(X.mbox[q], slot, X.rank);
        for (int k = 0; k < nwords; ++k) e[1 + k] = rec[k];
    }
    __threadfence_system();
    for (int q = 0; q < X.world; ++q) st_release_sys(xchg_entry(X.mbox[q], slot, X.rank), seq);
    bool ok = true;
    for (int q = 0; q < X.world; ++q) ok = xchg_wait_one(X, slot, q, seq) && ok;
    __stcg(X.seq, seq);
    if (!ok) xchg_fail(c);
    return slot;
}

__device__ __forceinline__ unsigned long long xchg_word(const XchgDev& X, unsigned slot, int src, int k)
{
    return __ldcg(xchg_entry(X.mbox[X.rank], slot, src) + 1 + k);
}

// ---- CTA-wide exchange: header words + a body of 32-bit words (a radix histogram, a key list) -----
// All threads of ONE CTA call this (blockDim.x >= XCHG_MAXR).  hdr: nh <= XCHG_HDR-1 words, body: nb <=
// XCHG_BODY/4 words (either may be null/0); hdr in shared or global memory, body in GLOBAL memory (read through L2).  s_flag: one int
// of shared memory.  Returns the slot; the record of rank r is at xchg_entry(own mailbox, slot, r):
// header words from +1, body from byte XCHG_HDR*8.
__device__ inline unsigned xchg_block(const XchgDev& X, Ctrl* c, const unsigned long long* hdr, int nh,
                                      const unsigned* body, int nb, int* s_flag)
{
    const unsigned long long seq = __ldcg(X.seq) + 1ull;    // written only by thread 0 after the final barrier
    const unsigned slot = (unsigned)(seq & (XCHG_RING - 1));
    if (threadIdx.x == 0) *s_flag = 1;
    for (int q = 0; q < X.world; ++q) {
        unsigned long long* e = xchg_entry(X.mbox[q], slot, X.rank);
        if ((int)threadIdx.x < nh) e[1 + threadIdx.x] = hdr[threadIdx.x];
        unsigned* eb = reinterpret_cast<unsigned*>(e + XCHG_HDR);
        for (int k = threadIdx.x; k < nb; k += blockDim.x) eb[k] = __ldcg(&body[k]);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < X.world) {
        st_release_sys(xchg_entry(X.mbox[threadIdx.x], slot, X.rank), seq);
        if (!xchg_wait_one(X, slot, (int)threadIdx.x, seq)) *s_flag = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __stcg(X.seq, seq);
        if (!*s_flag) xchg_fail(c);
    }
    __syncthreads();
    return slot;
}

__device__ __forceinline__ const unsigned* xchg_body(const XchgDev& X, unsigned slot, int src)
{
    return reinterpret_cast<const unsigned*>(xchg_entry(X.mbox[X.rank], slot, src) + XCHG_HDR);
}

// =====================================================================================================
// Low-latency (LL) exchanges for the per-iteration records: each 8-byte word = (seq32 << 32) | data32, written
// with one relaxed system-scope store and valid the moment its flag half matches -- no fence and no
// separate flag, so an exchange costs one NVLink crossing plus the poll (NCCL's LL protocol, done here
// inside the producing kernel).  Own sequence counter X.seq[1]; same ring argument as above.
// =====================================================================================================
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long* ll_entry(char* mbox, unsigned slot, int src)
{
    return reinterpret_cast<unsigned long long*>(mbox + XCHG_LL_OFF + ((size_t)slot * XCHG_MAXR + (size_t)src) * LL_STRIDE);
}
__device__ __forceinline__ unsigned long long ll_pack(unsigned data, unsigned flag)
{
    return ((unsigned long long)flag << 32) | (unsigned long long)data;
}
// spin until the word carries `flag`; returns its data half (ok = false on timeout)
static __device__ __noinline__ unsigned ll_poll_slow(const unsigned long long* p, unsigned flag, bool& ok);
__device__ __forceinline__ unsigned ll_poll(const unsigned long long* p, unsigned flag, bool& ok)
{
    unsigned long long v = ld_relaxed_sys(p);
    if ((unsigned)(v >> 32) == flag) return (unsigned)v;
    return ll_poll_slow(p, flag, ok);
}
static __device__ __noinline__ unsigned ll_poll_slow(const unsigned long long* p, unsigned flag, bool& ok)
{
    unsigned long long v;
    const unsigned long long t0 = xchg_now_ns();
    for (;;) {
#pragma unroll 1
        for (int spin = 0; spin < 32; ++spin) {
            v = ld_relaxed_sys(p);
            if ((unsigned)(v >> 32) == flag) return (unsigned)v;
        }
        if (xchg_now_ns() - t0 > XCHG_TIMEOUT_NS) { ok = false; return 0u; }
    }
}

// batched poll: NB independent loads are issued back to back (one L2 round trip for all of them) and
// re-issued together until every word carries `flag`
template <int NB>
__device__ __forceinline__ void ll_poll_batch(const unsigned long long* p, int stride, unsigned flag, unsigned (&data)[NB], bool& ok)
{
    unsigned long long v[NB];
    unsigned long long t0 = 0ull;
#pragma unroll 1
    for (int spin = 0;; ++spin) {
        bool all = true;
#pragma unroll
        for (int k = 0; k < NB; ++k) v[k] = ld_relaxed_sys(p + (size_t)k * stride);
#pragma unroll
        for (int k = 0; k < NB; ++k) all = all && ((unsigned)(v[k] >> 32) == flag);
        if (all) break;
        if ((spin & 31) == 31) {
            unsigned long long t = xchg_now_ns();
            if (t0 == 0ull) t0 = t;
            else if (t - t0 > XCHG_TIMEOUT_NS) { ok = false; break; }
        }
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) data[k] = (unsigned)v[k];
}

// ---- warp-wide LL all-gather of NW <= 8 64-bit words --------------------------------------------------
// All 32 lanes of ONE warp call this with the same rec[].  Lane q < world posts the record to rank q and
// then collects rank q's record from the local mailbox: on return lane q holds out[] = record of rank q
// (lanes >= world: zeros, valid = false).  The caller reduces across lanes with shuffles.
// The same all-gather with the exchange's sequence number supplied by the caller, who also publishes X.seq[1]
// when its kernel has counted several exchanges (head.cu).
template <int NW>
__device__ __forceinline__ bool xchg_ll_warp_seq(const XchgDev& X, Ctrl* c, unsigned long long seq, const unsigned long long (&rec)[NW],
                                                 unsigned long long (&out)[NW])
{
    const int lane = threadIdx.x & 31;
    const unsigned slot = (unsigned)(seq & (XCHG_RING - 1)), flag = (unsigned)seq;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NW; ++k) out[k] = 0ull;
    if (lane < X.world) {
        unsigned long long* e = ll_entry(X.mbox[lane], slot, X.rank);
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            st_relaxed_sys(e + 2 * k, ll_pack((unsigned)rec[k], flag));
            st_relaxed_sys(e + 2 * k + 1, ll_pack((unsigned)(rec[k] >> 32), flag));
        }
        unsigned half[2 * NW];
        ll_poll_batch<2 * NW>(ll_entry(X.mbox[X.rank], slot, lane), 1, flag, half, ok);
#pragma unroll
        for (int k = 0; k < NW; ++k) out[k] = ((unsigned long long)half[2 * k + 1] << 32) | half[2 * k];
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0 && !ok) xchg_fail(c);
    __syncwarp();
    return lane < X.world;
}

template <int NW>
__device__ __forceinline__ bool xchg_ll_warp(const XchgDev& X, Ctrl* c, const unsigned long long (&rec)[NW],
                                             unsigned long long (&out)[NW])
{
    const unsigned long long seq = __ldcg(X.seq + 1) + 1ull;
    const bool valid = xchg_ll_warp_seq<NW>(X, c, seq, rec, out);
    if ((threadIdx.x & 31) == 0) __stcg(X.seq + 1, seq);
    __syncwarp();
    return valid;
}

// ---- CTA-wide LL exchange: nh <= 16 header values (all-gathered) + an nb-bin histogram (all-reduced) ----
// All threads of ONE CTA (blockDim.x == 256).  hdr: shared memory, nh 64-bit words of this rank.  H: global
// memory, nb <= LL_BODY counters of this rank on entry, the sum over all ranks on return (nullptr/0: none).
// hdr_all: shared memory, [world][nh] words on return.  Ends with a CTA barrier.
__device__ inline void xchg_ll_block(const XchgDev& X, Ctrl* c, const unsigned long long* hdr, int nh,
                                     unsigned* H, int nb, unsigned long long* hdr_all, int* s_flag)
{
    const int tid = threadIdx.x;
    const unsigned long long seq = __ldcg(X.seq + 1) + 1ull;
    const unsigned slot = (unsigned)(seq & (XCHG_RING - 1)), flag = (unsigned)seq;
    if (tid == 0) *s_flag = 1;
    unsigned mine[LL_BODY / 256];                 // this thread's bins, read once
#pragma unroll
    for (int j = 0; j < LL_BODY / 256; ++j) { int k = tid + j * 256; mine[j] = (k < nb) ? __ldcg(&H[k]) : 0u; }
    for (int q = 0; q < X.world; ++q) {
        unsigned long long* e = ll_entry(X.mbox[q], slot, X.rank);
        if (tid < 2 * nh) st_relaxed_sys(e + tid, ll_pack((unsigned)(hdr[tid >> 1] >> (32 * (tid & 1))), flag));
#pragma unroll
        for (int j = 0; j < LL_BODY / 256; ++j) { int k = tid + j * 256; if (k < nb) st_relaxed_sys(e + LL_HDR + k, ll_pack(mine[j], flag)); }
    }
    bool ok = true;
    unsigned* h32 = reinterpret_cast<unsigned*>(hdr_all);
    for (int i = tid; i < X.world * 2 * nh; i += blockDim.x) {
        int r = i / (2 * nh), w = i - r * 2 * nh;
        h32[i] = ll_poll(ll_entry(X.mbox[X.rank], slot, r) + w, flag, ok);      // little-endian halves: [r][nh] 64-bit words
    }
    if (nb > 0) {
        // this thread's bins tid + 256 j of every rank; nb is a multiple of 256 (512 or 2048 bins)
        unsigned tot[LL_BODY / 256];
#pragma unroll
        for (int j = 0; j < LL_BODY / 256; ++j) tot[j] = 0u;
        for (int r = 0; r < X.world; ++r) {
            const unsigned long long* m = ll_entry(X.mbox[X.rank], slot, r) + LL_HDR + tid;
            if (nb == LL_BODY) {
                unsigned d[LL_BODY / 256];
                ll_poll_batch<LL_BODY / 256>(m, 256, flag, d, ok);
#pragma unroll
                for (int j = 0; j < LL_BODY / 256; ++j) tot[j] += d[j];
            } else {
#pragma unroll
                for (int j = 0; j < LL_BODY / 256; ++j) if (tid + j * 256 < nb) tot[j] += ll_poll(m + j * 256, flag, ok);
            }
        }
#pragma unroll
        for (int j = 0; j < LL_BODY / 256; ++j) if (tid + j * 256 < nb) H[tid + j * 256] = tot[j];
    }
    if (!ok) *s_flag = 0;
    __syncthreads();
    if (tid == 0) {
        __stcg(X.seq + 1, seq);
        if (!*s_flag) xchg_fail(c);
    }
    __syncthreads();
}

// ---- LL all-gather of key lists through every rank's XCHG_GCAND area ------------------------------------
// post (one CTA per rank): this rank's `mine` keys (global memory) land at [off, off + mine) of the gathered
// list in EVERY rank's mailbox.  collect (any number of CTAs, no grid barrier needed): poll the `total`
// gathered keys out of the local mailbox into dst (shared memory).  The flag is the LL sequence number of
// the exchange that distributed the counts (already consumed by everyone who gets here), so the area needs
// no ring: its next writer is at least three full exchanges later.
__device__ inline void ll_post_keys(const XchgDev& X, const unsigned long long* src, unsigned mine, unsigned off)
{
    const unsigned flag = (unsigned)__ldcg(X.seq + 1);
    for (int q = 0; q < X.world; ++q) {
        unsigned long long* g = reinterpret_cast<unsigned long long*>(X.mbox[q] + XCHG_GCAND_OFF) + 2 * (size_t)off;
        for (unsigned i = threadIdx.x; i < mine; i += blockDim.x) {
            unsigned long long key = __ldcg(&src[i]);
            st_relaxed_sys(g + 2 * i, ll_pack((unsigned)key, flag));
            st_relaxed_sys(g + 2 * i + 1, ll_pack((unsigned)(key >> 32), flag));
        }
    }
}

__device__ inline void ll_collect_keys(const XchgDev& X, Ctrl* c, unsigned total, unsigned long long* dst)
{
    const unsigned flag = (unsigned)__ldcg(X.seq + 1);
    bool ok = true;
    const unsigned long long* g = reinterpret_cast<const unsigned long long*>(X.mbox[X.rank] + XCHG_GCAND_OFF);
    for (unsigned i = threadIdx.x; i < total; i += blockDim.x) {
        unsigned lo = ll_poll(g + 2 * i, flag, ok), hi = ll_poll(g + 2 * i + 1, flag, ok);
        dst[i] = ((unsigned long long)hi << 32) | lo;
    }
    if (!ok) xchg_fail(c);
    __syncthreads();
}

}  // namespace abcdez
