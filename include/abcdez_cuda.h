/*
 * abcdez_cuda.h -- C ABI of libabcdez_cuda.so, the B200 (sm_100a) implementation of
 * the ABCdeZ.jl particle hot path.
 *
 * Every entry point replaces one interface of the reference (paths relative to the
 * ABCdeZ.jl v0.6.0 tree); the Julia shim (abcdez.jl_b200/julia/ABCdeZCUDA.jl) and the
 * Python mirror (abcdez.jl_b200/host.py) bind exactly these symbols.  Plain pointers and
 * sizes only; all functions return an int status (ABCDEZ_OK == 0) and never throw.
 *
 * Layout conventions
 *   theta / P : particle-major, N rows of d doubles (== a Julia d x N Matrix{Float64})
 *   blobs     : N rows of blob_bytes bytes
 *   alive     : N bytes (0/1)
 *   indices   : 0-based int32 unless stated
 */
#ifndef ABCDEZ_CUDA_H
#define ABCDEZ_CUDA_H

#ifndef __CUDACC_RTC__          /* runtime-compiled models (abcdez_model_compile): NVRTC has no system headers */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ABCDEZ_VERSION 100
#define ABCDEZ_MAXD 16          /* max length(prior) */
#define ABCDEZ_MAXDATA 64       /* doubles of bound observed data per model */
#define ABCDEZ_MAXBLOB 64       /* max blob bytes per particle */

/* status codes (errors of src/abcdez_smc.jl:223-235, src/abcdez_mc.jl:108-110 map to BAD_ARG) */
enum {
    ABCDEZ_OK = 0,
    ABCDEZ_ERR_BAD_ARG = 1,
    ABCDEZ_ERR_CUDA = 2,
    ABCDEZ_ERR_NAN_DISTANCE = 3,   /* quantile() of a NaN distance (Statistics.quantile errors) */
    ABCDEZ_ERR_NO_ALIVE = 4,       /* "No alive particles" (src/abcdez_smc.jl:375) -- a warning there */
    ABCDEZ_ERR_INIT_RETRY = 5,     /* abcde_init! redraw limit hit (src/abcdez_init.jl:14 spins forever) */
    ABCDEZ_ERR_PARTNER_RETRY = 6,  /* partner loops spun out (src/abcdez_smc.jl:120-126 with < 3 alive) */
    ABCDEZ_ERR_NCCL = 7,
    ABCDEZ_ERR_UNSUPPORTED = 8
};

/* prior marginal families (Distributions.jl types accepted by Factored, src/abcdez_priors.jl:18-21) */
enum {
    ABCDEZ_NORMAL = 0,            /* mu, sigma */
    ABCDEZ_UNIFORM = 1,           /* a, b */
    ABCDEZ_DISCRETE_UNIFORM = 2,  /* a, b */
    ABCDEZ_LOGNORMAL = 3,         /* mu, sigma */
    ABCDEZ_EXPONENTIAL = 4,       /* scale */
    ABCDEZ_GAMMA = 5,             /* shape, scale */
    ABCDEZ_BETA = 6,              /* alpha, beta */
    ABCDEZ_NEGBIN = 7             /* r, p */
};

/* ABC kernels, src/abcdez_types.jl:26-73 */
enum {
    ABCDEZ_INDICATOR = 0,         /* Indicator0to-eps        0 <= x <= eps */
    ABCDEZ_INDICATOR_STRICT = 1,  /* IndicatorStrict0to-eps  0 <= x <  eps (default, src/abcdez_smc.jl:218) */
    ABCDEZ_EPA = 2,               /* Epa0to-eps */
    ABCDEZ_EPA_STRICT = 3         /* EpaStrict0to-eps */
};

/* per-particle flags reported by the sweeps */
#define ABCDEZ_FLAG_SIM 1         /* dist! was evaluated (nsims[i] += 1, src/abcdez_smc.jl:138) */
#define ABCDEZ_FLAG_ACC 2         /* proposal accepted (naccs[i] += 1, src/abcdez_smc.jl:150) */

typedef struct abcdez_ctx abcdez_ctx;
typedef struct abcdez_prior abcdez_prior;
typedef struct abcdez_model abcdez_model;
typedef struct abcdez_pop abcdez_pop;      /* a device-resident particle population */

/* ---- context ----------------------------------------------------------------------- */
/* One context per process and GPU.  stream: a cudaStream_t to launch on, or NULL for a
 * private stream.  Replaces the executor choice of src/abcdez_smc.jl:237. */
int abcdez_init(int device, void* stream, abcdez_ctx** out);
int abcdez_destroy(abcdez_ctx* ctx);
int abcdez_version(void);
const char* abcdez_last_error(void);       /* thread-local message of the last failure */
int abcdez_sync(abcdez_ctx* ctx);          /* cudaStreamSynchronize on the context stream */

/* Single-process multi-GPU context: `parallel=true` of src/abcdez_smc.jl:237 / src/abcdez_mc.jl:112 as "all GPUs of
 * this process".  n_gpus <= 0: every visible GPU (at most 8); device_ids NULL: 0 .. n_gpus-1.  abcdez_smc_run /
 * abcdez_mc_run on such a context run ONE population sharded over the GPUs (contiguous blocks, rank-local DE partners,
 * everything else global -- the same kernels and in-kernel NVLink exchanges as the one-process-per-GPU runs below) and
 * fill the caller's buffers with the WHOLE population's rows; one host call, no process launcher, no NCCL (peers are
 * mapped with cudaDeviceEnablePeerAccess; one internal worker thread per GPU).  Stage-level calls (abcdez_pop_*) and
 * run-state snapshots need a single-GPU context. */
int abcdez_init_multi(int n_gpus, const int* device_ids, abcdez_ctx** out);
int abcdez_ctx_gpus(const abcdez_ctx* ctx);   /* GPUs behind a context (1 for abcdez_init contexts) */

/* ---- sharded runs: one process per GPU, particles in contiguous blocks ----------------
 * Replaces the `parallel=true` executor of src/abcdez_smc.jl:237 / src/abcdez_mc.jl:112 across GPUs.
 * abcdez_nccl_unique_id: rank 0 creates a 128-byte ncclUniqueId; the caller distributes it (MPI,
 * torch.distributed, Julia Distributed ...).  abcdez_comm_init (collective): libnccl.so.2 is dlopen-ed and
 * used for the host-level plumbing only (IPC handle all-gathers, barriers); it maps every rank's mailbox
 * over CUDA IPC so that the per-iteration exchanges (eps-quantile histograms, weight sums, ESS, counters)
 * run INSIDE the kernels over NVLink peer memory and the resampling gathers read peer particles directly.
 * After it, abcdez_smc_run / abcdez_mc_run on this context are collective: opts.nparticles is the size
 * of the whole population, this rank owns the block abcdez_shard_range(nparticles, rank, world) and fills
 * its result buffers with that block's rows; scalars and histories are global and identical on all ranks.
 * DE partners are drawn inside the rank's block (island variant); everything else is global. */
int abcdez_nccl_unique_id(void* id128);
int abcdez_comm_init(abcdez_ctx* ctx, int rank, int world, const void* id128);
int abcdez_shard_range(int64_t N, int rank, int world, int64_t* lo, int64_t* hi);
/* `rounds` in-kernel exchanges (one histogram-sized and one scalar record each) through the fenced ring
 * (mode 0: barriers, resampling) or the low-latency ring (mode 1: the per-iteration records); checksum is
 * a known function of (world, rounds) -- tests/multi_gpu_worker.py -- and us_per_round their latency */
int abcdez_comm_selftest(abcdez_ctx* ctx, int rounds, int mode, uint64_t* checksum, double* us_per_round);

/* ---- prior: Factored(dists...) src/abcdez_priors.jl:18-61 ------------------------------ */
int abcdez_prior_create(abcdez_ctx* ctx, int d, const int32_t* family, const double* params /* d x 4 */,
                        abcdez_prior** out);
int abcdez_prior_destroy(abcdez_prior* p);
/* rand(rng, prior) for N particles (src/abcdez_smc.jl:242): Philox stream (seed, id0+i, epoch) */
int abcdez_prior_sample(abcdez_ctx* ctx, const abcdez_prior* p, int64_t N, uint64_t seed, uint32_t epoch,
                        int64_t id0, double* theta_out);
/* logpdf(prior, push_p(prior, theta)) (src/abcdez_smc.jl:243, src/abcdez_priors.jl:40-46) */
int abcdez_prior_logpdf(abcdez_ctx* ctx, const abcdez_prior* p, int64_t N, const double* theta,
                        double* logpi_out);
/* push_p (src/abcdez_types.jl:20-23) */
int abcdez_prior_push(abcdez_ctx* ctx, const abcdez_prior* p, int64_t N, const double* theta, double* out);

/* ---- models: the dist!(theta, ve) -> (d, blob) plugin as registered device functors ------ */
int abcdez_model_count(void);
const char* abcdez_model_name(int id);
int abcdez_model_lookup(const char* name, int* id);
int abcdez_model_info(int id, int* d, int* blob_bytes);
/* bind observed data (<= ABCDEZ_MAXDATA doubles; the `data` captured by the Julia closure) */
int abcdez_model_bind(abcdez_ctx* ctx, int id, const double* data, size_t ndata, abcdez_model** out);
int abcdez_model_destroy(abcdez_model* m);
/* Runtime-supplied model: the closest analogue of passing an arbitrary Julia `dist!` (src/abcdez_smc.jl:215).
 * cuda_src is the CUDA C++ source of one struct (placed in namespace abcdez, so Philox streams, the portable
 * math and the helpers of csrc/common.cuh are in scope):
 *     struct MyModel {
 *         static constexpr int D = 2, BLOB = 0, NOISE = 0;           // length(prior), blob bytes (multiple of 8)
 *         static constexpr const char* name = "my_model";
 *         __device__ static double run(const double* theta, const double* data, SimRng& rng, double* blob) {...}
 *     };
 * It is compiled with NVRTC for sm_100a against the library's own kernel templates (the init, abcdesmc_swarm!,
 * abcdemc_swarm! and simulate kernels are instantiated for it at run time) and registered under `name`:
 * abcdez_model_lookup / _bind and both run calls then treat it like a built-in model.  d / blob_bytes must equal
 * the struct's D / BLOB (checked at compile time).  log (optional, log_cap bytes) receives the NVRTC log.
 * ctx == NULL: compile only -- validates a model on a machine without a GPU; *id = -1. */
int abcdez_model_compile(abcdez_ctx* ctx, const char* name, const char* struct_name, const char* cuda_src,
                         int d, int blob_bytes, int* id, char* log, size_t log_cap);
/* one dist! evaluation per row of theta_pushed with stream (seed, id0+i, epoch, tag) */
int abcdez_simulate(abcdez_ctx* ctx, const abcdez_model* m, int64_t N, const double* theta_pushed,
                    uint64_t seed, uint32_t epoch, uint32_t tag, int64_t id0, double* dist_out,
                    uint8_t* blobs_out);

/* ---- ABC kernels, src/abcdez_types.jl:26-73 (host evaluation, for the known-answer tests) - */
double abcdez_kernel_pdf(int kernel, double eps, double x);
double abcdez_kernel_logpdf(int kernel, double eps, double x);

/* ---- whole runs ------------------------------------------------------------------------ */
/* kwargs of abcdesmc!, src/abcdez_smc.jl:215-220, same defaults (abcdez_smc_opts_default) */
typedef struct {
    int64_t nparticles;      /* 100 */
    double alpha;            /* 0.95 */
    double delta_ess;        /* 0.5 */
    int64_t nsims_max;       /* 10^7 */
    int32_t Kmcmc;           /* 3 */
    double Kmcmc_min;        /* 1.0 */
    int32_t kernel;          /* ABCDEZ_INDICATOR_STRICT */
    double facc_stop;        /* 0.0 */
    double facc_min;         /* 0.0 */
    double facc_tune;        /* 0.975 */
    uint64_t seed;           /* rng -> 64-bit Philox key */
    int32_t verboseout;      /* 1: fill the history arrays */
    int32_t max_iters;       /* 0 = unbounded (benchmark guard, not in the reference) */
    int32_t exact_scan;      /* 1: sequential FP64 cumsum for Epanechnikov resampling (parity mode) */
    int32_t profile;         /* 1: time the head, resampling and sweep launches of every iteration with CUDA events */
    int32_t sync_every;      /* host polls the stop flag every this many iterations (default 1) */
    int32_t fused_head;      /* 1 (default): eps quantile + reweight + ESS + alive list in one cooperative kernel;
                                0: the stage kernels one by one (same results bit for bit) */
    /* relaxed-parity performance modes (default 0 = the reference's algorithm, bit for bit) */
    int32_t systematic_resampling;  /* 1: ONE uniform for all strata (systematic resampling) instead of one per stratum
                                       (wsample_stratified!, src/abcdez_smc.jl:15-56) */
    int32_t partner_segments;       /* 1: warp-coherent DE partners -- one pair of random bases per warp, lane l takes the
                                       l-th alive particle behind each (src/abcdez_smc.jl:119-126 draws every particle's
                                       partners independently); marginally the same law, coalesced gathers */
    int32_t fp32_state;             /* 1: FP32 particle state -- the theta generations hold floats (half the bytes per row move and
                                       per random partner gather); arithmetic, log prior, distances and weights stay FP64, every
                                       proposal is rounded to float before it is scored.  Models of the static registry except
                                       g-and-k; not with run-state snapshots (ABCDEZ_ERR_UNSUPPORTED otherwise) */
    int32_t reserved1;
} abcdez_smc_opts;

typedef struct {
    /* caller-allocated outputs (host), any may be NULL */
    double* P;               /* N x d, push_p applied (src/abcdez_smc.jl:382) */
    double* Wns;             /* N */
    double* C;               /* N distances */
    uint8_t* blobs;          /* N x blob_bytes */
    /* histories, src/abcdez_smc.jl:284-292,362-370; entry 0 is the pre-loop record */
    int32_t hist_cap;
    double* h_eps; double* h_dmin; double* h_dmax; double* h_logZ; double* h_ess; double* h_facc;
    double* h_gamma0; int32_t* h_Kmcmc;
    /* scalars filled by the library */
    double eps; double logZ;
    int64_t iters; int64_t nsims;
    int32_t hist_len;
    int32_t status;          /* ABCDEZ_OK or ABCDEZ_ERR_NO_ALIVE (run still returns OK) */
    int64_t n_resamples; int64_t n_sweeps; int64_t n_launches;
    double sweep_ms;         /* profile=1: summed CUDA-event time of the sweep kernel launches */
    double total_ms;         /* CUDA-event time of the whole device loop (after init) */
    double init_ms;
    int32_t hist_dropped;    /* history records that did not fit hist_cap (the histories are truncated, not mislabelled) */
    int32_t reserved0;
    double head_ms;          /* profile=1: summed CUDA-event time of the head kernel (eps select + reweight + ESS + alive list) */
    double resample_ms;      /* profile=1: ... of the resampling launches of the iterations that did resample */
} abcdez_smc_result;

void abcdez_smc_opts_default(abcdez_smc_opts* o);
int abcdez_smc_run(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                   const abcdez_smc_opts* opts, abcdez_smc_result* res);

/* Batched runs: nruns independent abcdesmc! runs in flight together on one GPU -- replicates of one
 * model for the evidence uncertainty (docs/src/index.md:214-220) or several models for a comparison
 * (examples/minimal_example.jl:27-65).  priors / models / eps_targets / opts / results are arrays of nruns entries; every run
 * gives exactly the result of its own abcdez_smc_run call (same seeds -> same bits); status (optional) gets each run's status. */
int abcdez_smc_run_batch(abcdez_ctx* ctx, int nruns, const abcdez_prior* const* priors, const abcdez_model* const* models,
                         const double* eps_targets, const abcdez_smc_opts* opts, abcdez_smc_result* results, int* status);

/* Equally weighted posterior sample of an abcdesmc! result on the device: P[weightinds(Wns)] of test/runtests.jl:13-19,
 * 287-291 (stratified resampling with the Philox uniforms of `seed`).  P, P_out: N x d; inds_out (optional): the 1-based
 * source indices as wsample_stratified! returns them. */
int abcdez_posterior_sample(abcdez_ctx* ctx, int64_t N, int d, const double* P, const double* Wns, uint64_t seed,
                            double* P_out, int64_t* inds_out);

/* Run-state snapshots (no equivalent in the reference): abcdez_smc_run that can stop between two
 * iterations and continue later, decision by decision like the uninterrupted run.  state_out (host, capacity >=
 * abcdez_smc_state_bytes(prior, model, nparticles, result.hist_cap or 1)) receives the state the run ended in --
 * typically after opts.max_iters iterations; state_in restores one instead of drawing from the prior.  On a
 * restored run eps_target, nsims_max, facc_stop and max_iters (iterations of THIS call) are taken from the call,
 * the seed from the snapshot, and every other option must equal the snapshot's; histories cover the whole run.
 * Either pointer may be NULL (both NULL == abcdez_smc_run).  On a sharded context (abcdez_comm_init) the call is collective and
 * every rank dumps / restores the snapshot of its own block (abcdez_smc_state_bytes with the rank's particle count);
 * abcdez_init_multi contexts do not take snapshots. */
int64_t abcdez_smc_state_bytes(const abcdez_prior* prior, const abcdez_model* model, int64_t nparticles, int32_t hist_cap);
int abcdez_smc_run_state(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                         const abcdez_smc_opts* opts, abcdez_smc_result* res, const void* state_in, int64_t state_in_bytes,
                         void* state_out, int64_t state_out_cap, int64_t* state_out_bytes);

/* kwargs of abcdemc!, src/abcdez_mc.jl:102-104 */
typedef struct {
    int64_t nparticles;      /* 50 */
    int32_t generations;     /* 20 */
    uint64_t seed;
} abcdez_mc_opts;

typedef struct {
    double* P; double* C; uint8_t* blobs;     /* caller-allocated, may be NULL */
    int32_t reached_eps;                      /* maximum(delta) <= eps_target, src/abcdez_mc.jl:163 */
    int64_t nsims;
    double dmin, dmax;
    double sweep_ms, total_ms;
    int64_t n_launches;
} abcdez_mc_result;

void abcdez_mc_opts_default(abcdez_mc_opts* o);
int abcdez_mc_run(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                  const abcdez_mc_opts* opts, abcdez_mc_result* res);

/* ---- device-resident population + stage-level entry points ---------------------------------
 * These are the kernels abcdez_smc_run / abcdez_mc_run launch, exposed one by one so the
 * parity tests can feed each stage the oracle's inputs (with optional injected randomness)
 * and so bench.py can time them with state resident in HBM. */
int abcdez_pop_create(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, int64_t N,
                      int64_t id0, abcdez_pop** out);
int abcdez_pop_destroy(abcdez_pop* pop);
/* host -> device; any pointer may be NULL (left unchanged).  W/alive default to 1/N / 1. */
int abcdez_pop_upload(abcdez_pop* pop, const double* theta, const double* logpi, const double* delta,
                      const uint8_t* blobs, const double* W, const uint8_t* alive);
int abcdez_pop_download(abcdez_pop* pop, double* theta, double* logpi, double* delta, uint8_t* blobs,
                        double* W, uint8_t* alive);
/* set the schedule scalars the kernels read from the device control block */
int abcdez_pop_set(abcdez_pop* pop, double eps, double eps_kernel_prev, int32_t kernel, double gamma0,
                   double gamma_sigma, uint64_t seed, uint32_t sweep_epoch);

/* abcde_init! (src/abcdez_init.jl:2-22).  draw_prior=1: also performs the prior draws and
 * logpdf of src/abcdez_smc.jl:242-243 (attempt 0). */
int abcdez_pop_init(abcdez_pop* pop, uint64_t seed, int draw_prior, int64_t* nredraws);

/* abcdesmc_swarm! (src/abcdez_smc.jl:106-153): one Jacobi sweep; flips the ping-pong buffers.
 * inj_* NULL -> Philox contract.  flags_out (N bytes, host) optional. */
int abcdez_pop_smc_sweep(abcdez_pop* pop, const int32_t* inj_a, const int32_t* inj_b, const double* inj_z,
                         const double* inj_u, uint8_t* flags_out, int64_t* nsims, int64_t* naccs);

/* abcdemc_swarm! (src/abcdez_mc.jl:5-61) */
int abcdez_pop_mc_sweep(abcdez_pop* pop, double eps_pop, double eps_target, const int32_t* inj_s,
                        const int32_t* inj_a, const int32_t* inj_b, const double* inj_z, const double* inj_u,
                        uint8_t* flags_out, int64_t* nsims);

/* quantile(delta[alive], alpha) (src/abcdez_smc.jl:301; Statistics.quantile type 7) */
int abcdez_pop_eps_quantile(abcdez_pop* pop, double alpha, double* q, double* v_lo, double* v_hi);

/* abcdesmc_update_ws! + src/abcdez_smc.jl:305-315,323: ws, wprod, wnorm, Wns, alive, ess */
int abcdez_pop_reweight(abcdez_pop* pop, double eps_new, double* wnorm, double* ess, int64_t* n_alive);

/* src/abcdez_smc.jl:301-324 in one launch: eps = max(min(quantile(delta[alive], alpha), eps), eps_target), then the
 * reweighting, ESS and alive-list steps of the two calls above against that eps.  Bit-identical to calling them
 * one by one. */
int abcdez_pop_head(abcdez_pop* pop, double alpha, double eps_target, double* q, double* eps, double* wnorm,
                    double* ess, int64_t* n_alive);

/* abcdesmc_resample! (src/abcdez_smc.jl:85-104, wsample_stratified! :15-56): indices + gather +
 * weight reset.  uniforms (N, host) NULL -> Philox stream (seed, stratum, epoch).  inds_out
 * (N, host, 0-based int32) optional.  mode: 0 auto (closed form for indicator kernels, parallel
 * scan otherwise), 1 force parallel scan, 2 force sequential scan (bit-exact for any weights). */
int abcdez_pop_resample(abcdez_pop* pop, const double* uniforms, uint32_t epoch, int mode, int32_t* inds_out);

/* the same index computation on caller-supplied weights (test/runtests.jl:13-19 `weightinds`):
 * inds_out 1-based int64 as the reference returns them, clamped to 1..N. */
int abcdez_wsample_stratified(abcdez_ctx* ctx, int64_t N, const double* weights, const double* uniforms,
                              int mode, int64_t* inds_out);

/* CUDA-event time (ms) of the most recent stage-level call's kernels, and launch count */
int abcdez_pop_last_timing(abcdez_pop* pop, double* ms, int64_t* launches);

/* Benchmark helper: run `sweeps` back-to-back abcdesmc_swarm! launches on the resident
 * population at the current control-block settings; returns summed nsims, naccs and the
 * CUDA-event time in ms (the sweep kernel only). */
int abcdez_pop_bench_sweeps(abcdez_pop* pop, int sweeps, int64_t* nsims, int64_t* naccs, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* ABCDEZ_CUDA_H */
