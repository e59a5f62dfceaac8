// comm.cu -- host side of the sharded runs: one process per GPU.
//
// NCCL is the plumbing: it is dlopen-ed (torch's bundled libnccl.so.2 when the host is Python, the system
// one otherwise), bootstrapped from a caller-distributed ncclUniqueId and used for the host-level
// collectives only -- all-gathering the CUDA IPC handles of the mailboxes and population slabs, and the
// barriers around a run's set-up.  The data path of a run never calls it: the per-iteration exchanges
// are done inside the kernels over NVLink peer memory (comm.cuh), and the resampling exchange reads the
// peers' particle rows straight from their HBM through the mapped slabs (bookkeeping.cu).
#include "internal.h"
#include "comm.cuh"
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

namespace abcdez {

// ---- the few NCCL entry points we need, resolved at run time ---------------------------------------
struct NcclId { char internal[128]; };
typedef void* ncclComm_t;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT8 = 0, NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_UINT32 = 3, NCCL_INT64 = 4, NCCL_UINT64 = 5 };
enum { NCCL_SUM = 0, NCCL_PROD = 1, NCCL_MAX = 2, NCCL_MIN = 3 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string path;
};

static NcclApi g_nccl;

static bool nccl_load(std::string* why)
{
    if (g_nccl.handle) return true;
    std::vector<std::string> cands;
    if (const char* e = getenv("ABCDEZ_NCCL_LIB")) cands.push_back(e);
    cands.push_back("libnccl.so.2");          // already-loaded copy (torch) or the loader path
    cands.push_back("libnccl.so");
    cands.push_back("/usr/lib/x86_64-linux-gnu/libnccl.so.2");
    std::string tried;
    for (const std::string& p : cands) {
        void* h = dlopen(p.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (!h) { tried += p + " (" + (dlerror() ? "not loadable" : "?") + "); "; continue; }
        g_nccl.handle = h; g_nccl.path = p;
        break;
    }
    if (!g_nccl.handle) { if (why) *why = "libnccl not found: " + tried; return false; }
#define SYM(field, name)                                                                        \
    do {                                                                                        \
        *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);                                  \
        if (!g_nccl.field) { if (why) *why = std::string("libnccl lacks ") + name; g_nccl.handle = nullptr; return false; } \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather"); SYM(AllReduce, "ncclAllReduce"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
}

// ---- single-process groups (abcdez_init_multi): the ranks are threads of one process, one per GPU -------------
// Host-level collectives become an in-process rendezvous and peers are mapped with cudaDeviceEnablePeerAccess
// (plain device pointers are valid on every GPU of the process) instead of NCCL + CUDA IPC.  The kernels, the
// mailboxes and the exchange protocol are the same.
struct LocalGroup {
    int world = 1;
    std::mutex m;
    std::condition_variable cv;
    int count = 0;
    unsigned gen = 0;
    std::vector<char> buf;
};

LocalGroup* local_group_create(int world) { LocalGroup* g = new LocalGroup(); g->world = world; return g; }
void local_group_destroy(LocalGroup* g) { delete g; }

static void local_barrier(LocalGroup* g)
{
    std::unique_lock<std::mutex> lk(g->m);
    const unsigned gen = g->gen;
    if (++g->count == g->world) { g->count = 0; g->gen++; g->cv.notify_all(); }
    else g->cv.wait(lk, [&] { return g->gen != gen; });
}

static void local_allgather(LocalGroup* g, int rank, const void* in, void* out, size_t bytes)
{
    {
        std::lock_guard<std::mutex> lk(g->m);
        if (g->buf.size() < bytes * (size_t)g->world) g->buf.resize(bytes * (size_t)g->world);
    }
    local_barrier(g);                                    // the buffer has its final size before anyone writes
    memcpy(g->buf.data() + bytes * (size_t)rank, in, bytes);
    local_barrier(g);
    memcpy(out, g->buf.data(), bytes * (size_t)g->world);
    local_barrier(g);                                    // everyone has read before the next collective overwrites
}

// ---- communicator state of one context ----------------------------------------------------------------
struct Comm {
    int rank = 0, world = 1;
    LocalGroup* local = nullptr;                  // non-null: single-process group (no NCCL, no IPC)
    ncclComm_t nccl = nullptr;
    char* mbox = nullptr;                         // own mailbox (cudaMalloc, exported)
    char* peer_mbox[XCHG_MAXR] = {};              // every rank's mailbox in this process' address space
    unsigned long long* seq = nullptr;            // local exchange counter
    char* stage = nullptr; size_t stage_bytes = 0;   // device staging of the host-level all-gathers
    char* slab = nullptr; size_t slab_bytes = 0;  // shared arena: the population slab peers may read
    char* peer_slab[XCHG_MAXR] = {};
    PeerTable* d_peers = nullptr;                 // device table handed to the kernels
    std::string err;
};

static int comm_fail(Comm* cm, const std::string& msg) { cm->err = msg; return ABCDEZ_ERR_NCCL; }

#define CM_CU(call)                                                                              \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return comm_fail(cm, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define CM_NC(call)                                                                              \
    do { int r_ = (call); if (r_ != NCCL_SUCCESS) return comm_fail(cm, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); } while (0)

const char* comm_error(const Comm* cm) { return cm ? cm->err.c_str() : ""; }

// host-level all-gather of `bytes` bytes per rank (through device staging, on the context stream)
int comm_allgather(Comm* cm, cudaStream_t st, const void* in, void* out, size_t bytes)
{
    if (cm->local) {
        CM_CU(cudaStreamSynchronize(st));                 // same ordering as the NCCL path: this rank's stream has drained
        local_allgather(cm->local, cm->rank, in, out, bytes);
        return ABCDEZ_OK;
    }
    size_t need = bytes * (size_t)(cm->world + 1);
    if (cm->stage_bytes < need) {
        if (cm->stage) cudaFree(cm->stage);
        cm->stage = nullptr; cm->stage_bytes = 0;
        CM_CU(cudaMalloc((void**)&cm->stage, need));
        cm->stage_bytes = need;
    }
    char* dsend = cm->stage; char* drecv = cm->stage + bytes;
    CM_CU(cudaMemcpyAsync(dsend, in, bytes, cudaMemcpyHostToDevice, st));
    CM_NC(g_nccl.AllGather(dsend, drecv, bytes, NCCL_UINT8, cm->nccl, st));
    CM_CU(cudaMemcpyAsync(out, drecv, bytes * cm->world, cudaMemcpyDeviceToHost, st));
    CM_CU(cudaStreamSynchronize(st));
    return ABCDEZ_OK;
}

int comm_barrier(Comm* cm, cudaStream_t st)
{
    unsigned long long one = 1, all[XCHG_MAXR];
    return comm_allgather(cm, st, &one, all, sizeof one);
}

void comm_destroy(Comm* cm);
int comm_rank(const Comm* cm) { return cm->rank; }
int comm_world(const Comm* cm) { return cm->world; }

int comm_create(int rank, int world, const void* id128, cudaStream_t st, Comm** out, std::string* why)
{
    if (world > XCHG_MAXR) { *why = "at most 8 ranks (one NVSwitch domain)"; return ABCDEZ_ERR_BAD_ARG; }
    if (!nccl_load(why)) return ABCDEZ_ERR_NCCL;
    Comm* cm = new Comm();
    cm->rank = rank; cm->world = world;
    NcclId id; memcpy(&id, id128, sizeof id);
    {
        int r = g_nccl.CommInitRank(&cm->nccl, world, id, rank);
        if (r != NCCL_SUCCESS) { *why = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); delete cm; return ABCDEZ_ERR_NCCL; }
    }
    auto run = [&]() -> int {
        CM_CU(cudaMalloc((void**)&cm->mbox, XCHG_MBOX_BYTES));
        CM_CU(cudaMemsetAsync(cm->mbox, 0, XCHG_MBOX_BYTES, st));
        CM_CU(cudaMalloc((void**)&cm->seq, 64));
        CM_CU(cudaMemsetAsync(cm->seq, 0, 64, st));
        CM_CU(cudaMalloc((void**)&cm->d_peers, sizeof(PeerTable)));
        CM_CU(cudaStreamSynchronize(st));
        cudaIpcMemHandle_t mine, all[XCHG_MAXR];
        CM_CU(cudaIpcGetMemHandle(&mine, cm->mbox));
        int rc = comm_allgather(cm, st, &mine, all, sizeof mine);
        if (rc) return rc;
        for (int q = 0; q < world; ++q) {
            if (q == rank) { cm->peer_mbox[q] = cm->mbox; continue; }
            void* p = nullptr;
            CM_CU(cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess));
            cm->peer_mbox[q] = (char*)p;
        }
        return comm_barrier(cm, st);
    };
    int rc = run();
    if (rc) { *why = cm->err; comm_destroy(cm); return rc; }
    *out = cm;
    return ABCDEZ_OK;
}

static void close_peer_slabs(Comm* cm)
{
    for (int q = 0; q < cm->world; ++q) {
        if (!cm->local && q != cm->rank && cm->peer_slab[q]) cudaIpcCloseMemHandle(cm->peer_slab[q]);
        cm->peer_slab[q] = nullptr;
    }
}

// one rank (thread) of a single-process group; collective over the group's threads
int comm_create_local(int rank, int world, LocalGroup* group, int device, cudaStream_t st, Comm** out, std::string* why)
{
    if (world > XCHG_MAXR) { *why = "at most 8 GPUs (one NVSwitch domain)"; return ABCDEZ_ERR_BAD_ARG; }
    Comm* cm = new Comm();
    cm->rank = rank; cm->world = world; cm->local = group;
    auto run = [&]() -> int {
        int devs[XCHG_MAXR] = { 0 };
        local_allgather(group, rank, &device, devs, sizeof(int));
        for (int q = 0; q < world; ++q) {
            if (q == rank) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[q], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return comm_fail(cm, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        }
        CM_CU(cudaMalloc((void**)&cm->mbox, XCHG_MBOX_BYTES));
        CM_CU(cudaMemsetAsync(cm->mbox, 0, XCHG_MBOX_BYTES, st));
        CM_CU(cudaMalloc((void**)&cm->seq, 64));
        CM_CU(cudaMemsetAsync(cm->seq, 0, 64, st));
        CM_CU(cudaMalloc((void**)&cm->d_peers, sizeof(PeerTable)));
        CM_CU(cudaStreamSynchronize(st));
        char* all[XCHG_MAXR] = { nullptr };
        local_allgather(group, rank, &cm->mbox, all, sizeof(char*));
        for (int q = 0; q < world; ++q) cm->peer_mbox[q] = all[q];
        return ABCDEZ_OK;
    };
    int rc = run();
    // (every rank reaches this barrier, failed or not, so that nobody waits for a rank that gave up)
    int ok = rc == ABCDEZ_OK, oks[XCHG_MAXR];
    local_allgather(group, rank, &ok, oks, sizeof(int));
    for (int q = 0; q < world; ++q) if (!oks[q] && rc == ABCDEZ_OK) { rc = ABCDEZ_ERR_NCCL; cm->err = "a peer rank failed to set up its mailbox"; }
    if (rc) { *why = cm->err; comm_destroy(cm); return rc; }
    *out = cm;
    return ABCDEZ_OK;
}

void comm_destroy(Comm* cm)
{
    if (!cm) return;
    close_peer_slabs(cm);
    for (int q = 0; q < cm->world; ++q)
        if (!cm->local && q != cm->rank && cm->peer_mbox[q]) cudaIpcCloseMemHandle(cm->peer_mbox[q]);
    if (cm->nccl) g_nccl.CommDestroy(cm->nccl);
    if (cm->slab) cudaFree(cm->slab);
    if (cm->mbox) cudaFree(cm->mbox);
    if (cm->seq) cudaFree(cm->seq);
    if (cm->stage) cudaFree(cm->stage);
    if (cm->d_peers) cudaFree(cm->d_peers);
    delete cm;
}

// Collective.  Returns a slab of >= max-over-ranks(need) bytes that every peer has mapped; grows it
// (collectively: every rank sees the same maximum, so every rank takes the same branch) when too small.
int comm_shared_slab(Comm* cm, cudaStream_t st, size_t need, char** slab)
{
    unsigned long long mine = need, all[XCHG_MAXR];
    int rc = comm_allgather(cm, st, &mine, all, sizeof mine);
    if (rc) return rc;
    size_t mx = 0;
    for (int q = 0; q < cm->world; ++q) mx = all[q] > mx ? (size_t)all[q] : mx;
    if (cm->slab_bytes < mx) {
        close_peer_slabs(cm);
        rc = comm_barrier(cm, st); if (rc) return rc;     // nobody maps the old slab any more
        if (cm->slab) cudaFree(cm->slab);
        cm->slab = nullptr; cm->slab_bytes = 0;
        CM_CU(cudaMalloc((void**)&cm->slab, mx));
        cm->slab_bytes = mx;
        if (cm->local) {                                  // same process: the peers' device pointers are usable as they are
            char* all[XCHG_MAXR] = { nullptr };
            rc = comm_allgather(cm, st, &cm->slab, all, sizeof(char*)); if (rc) return rc;
            for (int q = 0; q < cm->world; ++q) cm->peer_slab[q] = all[q];
            *slab = cm->slab;
            return ABCDEZ_OK;
        }
        cudaIpcMemHandle_t h, hs[XCHG_MAXR];
        CM_CU(cudaIpcGetMemHandle(&h, cm->slab));
        rc = comm_allgather(cm, st, &h, hs, sizeof h); if (rc) return rc;
        for (int q = 0; q < cm->world; ++q) {
            if (q == cm->rank) { cm->peer_slab[q] = cm->slab; continue; }
            void* p = nullptr;
            CM_CU(cudaIpcOpenMemHandle(&p, hs[q], cudaIpcMemLazyEnablePeerAccess));
            cm->peer_slab[q] = (char*)p;
        }
    }
    *slab = cm->slab;
    return ABCDEZ_OK;
}

// Collective, at the start of a sharded run: publish where this rank's arrays live inside its slab, build
// the device PeerTable, and reset the mailbox ring (zeroed flags, sequence 0) between two barriers so
// that a run aborted on one rank cannot desynchronise the next one.
struct SlabLayout { unsigned long long theta[2], logpi[2], delta[2], blob[2], alive_list, cumsum; unsigned N, id0; };

int comm_begin_run(Comm* cm, cudaStream_t st, const PopDev& P, XchgDev* x, const PeerTable** d_peers)
{
    SlabLayout mine, all[XCHG_MAXR];
    auto off = [&](const void* p) { return (unsigned long long)((const char*)p - cm->slab); };
    for (int g = 0; g < 2; ++g) {
        mine.theta[g] = off(P.theta[g]); mine.logpi[g] = off(P.logpi[g]); mine.delta[g] = off(P.delta[g]);
        mine.blob[g] = off(P.blob[g]);
    }
    mine.alive_list = off(P.alive_list); mine.cumsum = off(P.cumsum); mine.N = P.N; mine.id0 = P.id0;
    int rc = comm_allgather(cm, st, &mine, all, sizeof mine);      // also a barrier: everyone left the previous run
    if (rc) return rc;
    PeerTable t; memset(&t, 0, sizeof t);
    for (int q = 0; q < cm->world; ++q) {
        char* base = cm->peer_slab[q];
        for (int g = 0; g < 2; ++g) {
            t.p[q].theta[g] = (const double*)(base + all[q].theta[g]); t.p[q].logpi[g] = (const double*)(base + all[q].logpi[g]);
            t.p[q].delta[g] = (const double*)(base + all[q].delta[g]); t.p[q].blob[g] = (const double*)(base + all[q].blob[g]);
        }
        t.p[q].alive_list = (const uint32_t*)(base + all[q].alive_list);
        t.p[q].cumsum = (const double*)(base + all[q].cumsum);
        t.p[q].N = all[q].N; t.p[q].id0 = all[q].id0;
    }
    CM_CU(cudaMemcpyAsync(cm->d_peers, &t, sizeof t, cudaMemcpyHostToDevice, st));
    CM_CU(cudaMemsetAsync(cm->mbox, 0, XCHG_MBOX_BYTES, st));
    CM_CU(cudaMemsetAsync(cm->seq, 0, 64, st));
    CM_CU(cudaStreamSynchronize(st));
    rc = comm_barrier(cm, st); if (rc) return rc;                   // every mailbox is clean before anyone posts
    x->rank = cm->rank; x->world = cm->world; x->seq = cm->seq;
    for (int q = 0; q < XCHG_MAXR; ++q) x->mbox[q] = q < cm->world ? cm->peer_mbox[q] : nullptr;
    *d_peers = cm->d_peers;
    return ABCDEZ_OK;
}

// ---- self test: rounds of the in-kernel exchanges (used by the multi-GPU tests and the latency probe) ---
// mode 0: fenced ring (xchg_block + xchg_small); mode 1: low-latency ring (xchg_ll_block + xchg_ll_warp)
__global__ void xchg_selftest_kernel(XchgDev X, Ctrl* c, int rounds, int mode, unsigned long long* out, unsigned* hist)
{
    __shared__ unsigned long long s_h[4];
    __shared__ unsigned long long s_all[XCHG_MAXR * 4];
    __shared__ int s_flag;
    unsigned long long acc = 0ull;
    for (int it = 0; it < rounds; ++it) {
        if (threadIdx.x == 0) { s_h[0] = (unsigned long long)(X.rank + 1) * 1000ull + it; s_h[1] = 0x100000000ull * (X.rank + 1) + it; }
        for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) hist[b] = (unsigned)(X.rank * 7 + b + it);
        __threadfence();
        __syncthreads();
        if (mode == 0) {
            unsigned slot = xchg_block(X, c, s_h, 2, hist, SEL_BINS, &s_flag);
            for (int r = 0; r < X.world; ++r) {
                if (threadIdx.x == 0) acc += xchg_word(X, slot, r, 0) + (xchg_word(X, slot, r, 1) >> 32);
                for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) acc += __ldcg(xchg_body(X, slot, r) + b);
            }
            if (threadIdx.x == 0) {
                unsigned long long rec[2] = { (unsigned long long)X.rank, (unsigned long long)it };
                unsigned s2 = xchg_small(X, c, rec, 2);
                for (int r = 0; r < X.world; ++r) acc += xchg_word(X, s2, r, 0) * 3ull + xchg_word(X, s2, r, 1);
            }
        } else {
            xchg_ll_block(X, c, s_h, 2, hist, SEL_BINS, s_all, &s_flag);      // hist <- sum over ranks
            for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) acc += __ldcg(&hist[b]);
            if (threadIdx.x == 0) for (int r = 0; r < X.world; ++r) acc += s_all[2 * r] + (s_all[2 * r + 1] >> 32);
            if (threadIdx.x < 32) {
                unsigned long long rec[2] = { (unsigned long long)X.rank, (unsigned long long)it }, got[2];
                bool valid = xchg_ll_warp<2>(X, c, rec, got);
                if (valid) acc += got[0] * 3ull + got[1];
            }
        }
        __syncthreads();
    }
    __shared__ unsigned long long s_acc;
    if (threadIdx.x == 0) s_acc = 0ull;
    __syncthreads();
    atomicAdd(&s_acc, acc);
    __syncthreads();
    if (threadIdx.x == 0) *out = s_acc;
}

int comm_selftest(Comm* cm, cudaStream_t st, int rounds, int mode, unsigned long long* result, double* us_per_round)
{
    Ctrl* d_ctrl = nullptr; unsigned long long* d_out = nullptr; unsigned* d_hist = nullptr;
    CM_CU(cudaMalloc((void**)&d_ctrl, sizeof(Ctrl))); CM_CU(cudaMalloc((void**)&d_out, 8)); CM_CU(cudaMalloc((void**)&d_hist, SEL_BINS * 4));
    CM_CU(cudaMemsetAsync(d_ctrl, 0, sizeof(Ctrl), st));
    CM_CU(cudaMemsetAsync(cm->mbox, 0, XCHG_MBOX_BYTES, st));
    CM_CU(cudaMemsetAsync(cm->seq, 0, 64, st));
    CM_CU(cudaStreamSynchronize(st));
    int rc = comm_barrier(cm, st); if (rc) return rc;
    XchgDev x; x.rank = cm->rank; x.world = cm->world; x.seq = cm->seq;
    for (int q = 0; q < XCHG_MAXR; ++q) x.mbox[q] = q < cm->world ? cm->peer_mbox[q] : nullptr;
    cudaEvent_t e0, e1;
    CM_CU(cudaEventCreate(&e0)); CM_CU(cudaEventCreate(&e1));
    xchg_selftest_kernel<<<1, 256, 0, st>>>(x, d_ctrl, 2, mode, d_out, d_hist);       // warm-up
    CM_CU(cudaEventRecord(e0, st));
    xchg_selftest_kernel<<<1, 256, 0, st>>>(x, d_ctrl, rounds, mode, d_out, d_hist);
    CM_CU(cudaEventRecord(e1, st));
    CM_CU(cudaGetLastError());
    Ctrl h;
    CM_CU(cudaMemcpyAsync(result, d_out, 8, cudaMemcpyDeviceToHost, st));
    CM_CU(cudaMemcpyAsync(&h, d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
    CM_CU(cudaStreamSynchronize(st));
    float ms = 0.f; CM_CU(cudaEventElapsedTime(&ms, e0, e1));
    if (us_per_round) *us_per_round = 1e3 * ms / (rounds > 0 ? rounds : 1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_ctrl); cudaFree(d_out); cudaFree(d_hist);
    rc = comm_barrier(cm, st); if (rc) return rc;
    if (h.err) return comm_fail(cm, "in-kernel exchange timed out (peer memory not reachable?)");
    return ABCDEZ_OK;
}

int nccl_unique_id(void* id128, std::string* why)
{
    if (!nccl_load(why)) return ABCDEZ_ERR_NCCL;
    NcclId id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != NCCL_SUCCESS) { *why = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return ABCDEZ_ERR_NCCL; }
    memcpy(id128, &id, sizeof id);
    return ABCDEZ_OK;
}

}  // namespace abcdez
