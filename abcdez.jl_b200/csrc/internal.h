// internal.h -- host-side objects behind the opaque handles of include/abcdez_cuda.h and the
// launcher interface between api.cu, sweep.cu and bookkeeping.cu.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#endif
#include "common.cuh"
#include "models.cuh"

namespace abcdez {

constexpr int TILE = 1024;            // particles per CTA in the bookkeeping kernels (256 thr x 4)
constexpr int BK_THREADS = 256;
#ifndef ABCDEZ_SWEEP_THREADS
#define ABCDEZ_SWEEP_THREADS 128
#endif
constexpr int SWEEP_THREADS = ABCDEZ_SWEEP_THREADS;    // thread-per-particle sweeps
#ifndef ABCDEZ_SWEEP_MIN_BLOCKS
#define ABCDEZ_SWEEP_MIN_BLOCKS 7      // 7 x 128 threads, 72 registers (28 warps per SM): measured best of 24 / 28 / 32 warps, and
                                       // 128-thread CTAs retire in finer grains than 256-thread ones (profiles/README.md)
#endif
constexpr int SWEEP_MIN_BLOCKS = ABCDEZ_SWEEP_MIN_BLOCKS;   // register cap of the fused sweep kernels
constexpr int SEL_BINS = 2048;        // 11-bit radix-select digits
constexpr int SEQTAB_MAX = 256;       // segments of the sequential-sum closed form

// closed form of the reference's sequential FP64 running sums (src/abcdez_smc.jl:46,50) for a
// constant addend: S(0)=0, S(k)=fl(S(k-1)+c).  S(kst[j] + m) = sst[j] + m*inc[j] exactly.
struct SeqTab {
    int n;
    unsigned long long kst[SEQTAB_MAX];
    double sst[SEQTAB_MAX];
    double inc[SEQTAB_MAX];
    unsigned long long kmax;      // largest valid k
};

// ---- sharded runs (comm.cuh / comm.cu): in-kernel exchanges over NVLink peer memory -----------------
constexpr int XCHG_MAXR = 8;                      // ranks of one NVSwitch domain
constexpr int XCHG_RING = 4;                      // mailbox slots (power of two)
constexpr int XCHG_HDR = 32;                      // 64-bit header words per entry: [0] flag, [1..31] small record
constexpr int XCHG_BODY = SEL_BINS * 4;           // body bytes per entry (one radix histogram)
constexpr int XCHG_STRIDE = XCHG_HDR * 8 + XCHG_BODY;
constexpr int XCHG_GCAND = 4096;                  // most candidate keys the distributed eps select gathers on every rank (= CAND_SMEM)
constexpr int XCHG_GCAND_PER = 1024;              // ... and the most one rank may contribute (its fixed region of the gather area)
// low-latency region: every 8-byte word carries 32 data bits and the 32-bit sequence number of the exchange
// that wrote it, so a word is valid as soon as it is seen -- no fence, one NVLink crossing (comm.cuh)
constexpr int LL_HDR = 32;                        // LL words of an entry's header record (16 64-bit values)
constexpr int LL_BODY = SEL_BINS;                 // LL words of an entry's body (one radix histogram)
constexpr size_t LL_STRIDE = (size_t)(LL_HDR + LL_BODY) * 8;
constexpr size_t XCHG_LL_OFF = (size_t)XCHG_RING * XCHG_MAXR * XCHG_STRIDE;
constexpr size_t XCHG_GCAND_OFF = XCHG_LL_OFF + (size_t)XCHG_RING * XCHG_MAXR * LL_STRIDE;   // 2 LL words per gathered key, one region per rank
constexpr size_t XCHG_HG_OFF = XCHG_GCAND_OFF + (size_t)XCHG_MAXR * XCHG_GCAND_PER * 16;     // the reduced histogram of an exchange (LL words, local readers)
constexpr size_t XCHG_MBOX_BYTES = XCHG_HG_OFF + (size_t)SEL_BINS * 8;
constexpr unsigned long long XCHG_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

struct XchgDev {
    int rank, world;                              // world == 1: no exchange anywhere
    unsigned long long* seq;                      // local exchange counters (device memory): [0] fenced ring, [1] LL ring
    char* mbox[XCHG_MAXR];                        // every rank's mailbox mapped into this process (own one included)
};

// what a rank needs to read a peer's particles during the global resampling (pointers valid in THIS process)
struct PeerPop {
    const double* theta[2]; const double* logpi[2]; const double* delta[2]; const double* blob[2];
    const uint32_t* alive_list;
    const double* cumsum;         // inclusive GLOBAL cumulative weights of the rank's block (general-weight resampling)
    uint32_t N, id0;
};
struct PeerTable { PeerPop p[XCHG_MAXR]; };

// device pointers of one population, passed to kernels by value
struct PopDev {
    double* theta[2];
    double* logpi[2];
    double* delta[2];
    double* blob[2];              // N x (BLOB/8) doubles
    double* W;
    uint8_t* alive;
    uint8_t* moved;               // 1: the row of this particle in the other generation buffer is stale (sweep.cuh)
    uint32_t* alive_list;         // alive particle indices (index order), then the dead ones (valid when n_alive < N)
    Ctrl* ctrl;
    double* partial;              // per-CTA reduction partials (2 x nblocks)
    uint32_t* tile_cnt;           // per-tile alive counts -> exclusive offsets
    uint32_t* sel_hist;           // 7 x SEL_BINS radix-select histograms (one per digit pass + the head's windowed first pass)
    unsigned long long* cand[2];  // head.cu candidate lists (alias cumsum / scratch)
    double* cumsum;               // N inclusive cumulative weights (general-weight resampling)
    int32_t* inds;                // N resampling indices (0-based)
    double* hist;                 // hist_cap x 8 history records
    SeqTab* tabs;                 // [0]: weights, [1]: strata edges
    uint32_t N;                   // particles of this rank
    uint32_t id0;                 // global index of this rank's first particle
    uint32_t ntiles;
    uint32_t Ng;                  // particles of the whole (sharded) population; == N on one GPU
    XchgDev x;
    const PeerTable* peers;       // device table, sharded runs only
    PhiloxKeys keys;              // round keys of the run's seed (== philox_keys(ctrl->seed)); constant-bank operands of the streams
    uint32_t flags;               // relaxed-parity modes of the run (POP_*), SURVEY.md 8f rank 4
    // queue-driven sweeps of heavy simulators (SPLIT models, sweep.cuh): the proposals that passed the prior
    double* prop_theta;           // N rows (row_stride): theta' of the sweep in flight
    double* prop_lp;              // N: log prior of theta'
    double* prop_dp;              // N: simulated distance
    double* prop_blob;            // N x (BLOB/8)
    uint8_t* prop_flag;           // N: 1 = in the queue
    uint32_t* queue;              // particle indices whose proposal must be simulated
};
enum : uint32_t { POP_PARTNER_SEGMENTS = 1u, POP_SYSTEMATIC = 2u, POP_FP32_STATE = 4u };   // relaxed-parity modes

struct SweepInj {
    const int32_t* a; const int32_t* b; const int32_t* s;
    const double* z; const double* u;
    uint8_t* flags;
};

struct McArgs {
    double eps_pop, eps_target;   // stage calls: given by the caller
    int from_ctrl;                // abcdez_mc_run: eps_target and extrema(delta) (-> eps_pop, src/abcdez_mc.jl:146-147) from the control block
    const double* sorted_delta;   // delta sorted ascending (ties in index order)
    const uint32_t* order;        // particle index per sorted position
};

#ifdef __CUDACC_RTC__
}  // namespace abcdez: runtime-compiled models (rtc.cu) see only the device-side declarations above
#else                             // ---- host side from here on ----

// per-model launchers: the static registry (inst_*.cu, gk.cu) and the runtime-compiled models (rtc.cu)
struct ModelOps {
    const char* name;
    int d, blob;
    void (*init)(const ModelOps&, cudaStream_t, const PopDev&, const PriorDev&, const ModelData&, uint64_t seed, int draw_prior);
    void (*smc_sweep)(const ModelOps&, cudaStream_t, const PopDev&, const PriorDev&, const ModelData&, const SweepInj&);
    void (*mc_sweep)(const ModelOps&, cudaStream_t, const PopDev&, const PriorDev&, const ModelData&, const SweepInj&, const McArgs&);
    void (*simulate)(const ModelOps&, cudaStream_t, const PriorDev*, const ModelData&, int64_t N, const double* theta_pushed,
                     uint64_t seed, uint32_t epoch, uint32_t tag, uint32_t id0, double* dist, double* blobs);
    void* dyn;                    // runtime-compiled models: their loaded module (rtc.cu); nullptr for the static registry
    int f32_state;                // 1: init / abcdesmc_swarm! launchers honour POP_FP32_STATE (templated kernels of the static registry)
    int split;                    // 1: sweeps run as propose -> queue-driven simulate -> accept (3 launches per sweep); 2: abcde_init! too (stepped simulators)
};
const ModelOps* model_ops(int id);
int model_count();

// runtime-compiled models (rtc.cu): CUDA source of one model struct -> NVRTC -> module -> a registry entry
int rtc_compile_model(const char* name, const char* struct_name, const char* cuda_src, int d, int blob_bytes, int load,
                      int* id, std::string* log);
const ModelOps* rtc_model_ops(int id);            // id >= M_COUNT
int rtc_model_count();

// dimension-only launchers for the prior stage calls (sweep.cu)
void launch_prior_op(cudaStream_t, int d, const PriorDev&, int64_t N, int op, const double* in, double* out,
                     uint64_t seed, uint32_t epoch, uint32_t id0);
enum { PRIOR_OP_SAMPLE = 0, PRIOR_OP_LOGPDF = 1, PRIOR_OP_PUSH = 2 };

// bookkeeping launchers (bookkeeping.cu); each returns the number of kernel launches it issued
int launch_eps_quantile(cudaStream_t, const PopDev&);                 // -> ctrl.q_a, q_b, q, eps (clamped)
int launch_reweight(cudaStream_t, const PopDev&);                     // -> W, alive, wnorm, logZ, ess, flags
int launch_compact(cudaStream_t, const PopDev&);                      // -> alive_list
int launch_head(cudaStream_t, const PopDev&, int sm_count);           // head.cu: the three above in one cooperative kernel
int launch_resample(cudaStream_t, const PopDev&, int D, int NB, const double* inj_u, uint32_t epoch,
                    int mode, int force);
int launch_end_iter(cudaStream_t, const PopDev&);
int launch_begin_run(cudaStream_t, const PopDev&);
int launch_minmax(cudaStream_t, const PopDev&);
int launch_strat_indices(cudaStream_t, int64_t N, const double* W_dev, const double* u_dev, double* cumsum_dev,
                         double* partial_dev, SeqTab* tabs_dev, int mode, long long* inds_dev);
int launch_push_rows(cudaStream_t, const PopDev&, const PriorDev&, int D, double* out_dense);
int launch_pack_rows(cudaStream_t, int D, int64_t N, const double* dense, double* rows, int to_rows);
int launch_mc_prepare(cudaStream_t, const PopDev&, double* sorted_delta, uint32_t* order, void* tmp, size_t tmp_bytes,
                      int force);          // mcsort.cu; force = 0: decided on the device (extrema(delta) vs eps_target)
size_t mc_sort_tmp_bytes(int64_t N);

// sharded runs, host side (comm.cu)
struct Comm;
struct LocalGroup;                                // the ranks of a single-process multi-GPU context (threads)
LocalGroup* local_group_create(int world);
void local_group_destroy(LocalGroup* g);
int comm_create_local(int rank, int world, LocalGroup* group, int device, cudaStream_t st, Comm** out, std::string* why);
int nccl_unique_id(void* id128, std::string* why);
int comm_create(int rank, int world, const void* id128, cudaStream_t st, Comm** out, std::string* why);
void comm_destroy(Comm* cm);
const char* comm_error(const Comm* cm);
int comm_rank(const Comm* cm);
int comm_world(const Comm* cm);
int comm_barrier(Comm* cm, cudaStream_t st);
int comm_allgather(Comm* cm, cudaStream_t st, const void* in, void* out, size_t bytes);
int comm_shared_slab(Comm* cm, cudaStream_t st, size_t need, char** slab);
int comm_begin_run(Comm* cm, cudaStream_t st, const PopDev& P, XchgDev* x, const PeerTable** d_peers);
int comm_selftest(Comm* cm, cudaStream_t st, int rounds, int mode, unsigned long long* result, double* us_per_round);

}  // namespace abcdez

struct abcdez_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int rank, world;
    abcdez::Comm* comm;           // non-null after abcdez_comm_init with world > 1
    int sm_count;
    std::vector<cudaEvent_t> ev_pool;   // reused by profile=1 runs (cudaEventCreate costs ~0.2 ms each)
    // device arena: populations are carved from one slab that lives as long as the context, so repeated
    // runs pay cudaMalloc/cudaFree (each a device-wide synchronisation) once, not 25 times per run
    char* arena; size_t arena_bytes; bool arena_busy;
    abcdez::Ctrl* h_ctrl_pool;          // pinned control-block mirror, reused like the arena
    bool h_ctrl_busy;
    abcdez::Ctrl* h_poll[2];            // pinned copies of the control block for the pipelined stop poll of abcdez_smc_run
    cudaEvent_t poll_ev[2];
    // abcdez_init_multi: this context only fans a run out to one sub-context (and host thread) per GPU
    std::vector<abcdez_ctx*> subs;
    abcdez::LocalGroup* group;
    std::vector<abcdez_ctx*> batch;     // abcdez_smc_run_batch: worker contexts (own stream + arena) on this GPU
};

struct abcdez_prior {
    abcdez::PriorDev dev;
};

struct abcdez_model {
    int id;
    const abcdez::ModelOps* ops;
    abcdez::ModelData data;
};

struct abcdez_pop {
    abcdez_ctx* ctx;
    abcdez::PriorDev prior;
    const abcdez::ModelOps* ops;
    abcdez::ModelData data;
    abcdez::PopDev dev;
    abcdez::Ctrl* h_ctrl;         // pinned mirror
    char* slab; size_t slab_bytes;   // all device arrays of the population live in this slab
    bool slab_from_arena, slab_shared, h_ctrl_from_pool;
    int64_t N;
    int D, DS, NB;
    int hist_cap;
    // scratch for injected arrays / flags
    void* scratch; size_t scratch_bytes;
    // abcdemc! sort buffers
    double* sorted_delta; uint32_t* order; void* sort_tmp; size_t sort_tmp_bytes;
    cudaEvent_t ev0, ev1;
    double last_ms; int64_t last_launches;
};

#endif  // __CUDACC_RTC__
