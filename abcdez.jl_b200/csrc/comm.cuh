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
__device__ __forceinline__ bool xchg_wait_one(const XchgDev& X, unsigned slot, int src, unsigned long long seq)
{
    const unsigned long long* flag = xchg_entry(X.mbox[X.rank], slot, src);
    if (ld_acquire_sys(flag) == seq) return true;
    const unsigned long long t0 = xchg_now_ns();
    for (;;) {
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

}  // namespace abcdez
