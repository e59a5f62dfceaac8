// api.cu -- the C ABI of libabcdez_cuda.so (include/abcdez_cuda.h): contexts, priors, models,
// device-resident populations, the stage-level entry points used by the parity tests, and the
// two run loops abcdez_smc_run (abcdesmc!, src/abcdez_smc.jl:215-394 of the reference) and
// abcdez_mc_run (abcdemc!, src/abcdez_mc.jl:102-172).
#include "internal.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <thread>
#include <vector>

using namespace abcdez;

// ---------------------------------------------------------------------------------------
// errors: status codes + a thread-local message, never an exception across the ABI
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ABCDEZ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

#define CHECK_ARG(cond, msg)                                                                       \
    do {                                                                                           \
        if (!(cond)) return fail(ABCDEZ_ERR_BAD_ARG, msg);                                         \
    } while (0)

extern "C" const char* abcdez_last_error(void) { return g_err.c_str(); }

// stage-level calls on an abcdez_init_multi context run on its first GPU
static inline abcdez_ctx* first_gpu(abcdez_ctx* c) { return (c && !c->subs.empty()) ? c->subs[0] : c; }
extern "C" int abcdez_version(void) { return ABCDEZ_VERSION; }

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
extern "C" int abcdez_init(int device, void* stream, abcdez_ctx** out)
{
    CHECK_ARG(out != nullptr, "abcdez_init: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(ABCDEZ_ERR_CUDA, std::string("abcdez_init: no CUDA device (") + cudaGetErrorString(e) +
                    "); libabcdez_cuda has no CPU fallback");
    CHECK_ARG(device >= 0 && device < n, "abcdez_init: device index out of range");
    CU(cudaSetDevice(device));
    abcdez_ctx* c = new (std::nothrow) abcdez_ctx();
    if (!c) return fail(ABCDEZ_ERR_CUDA, "abcdez_init: out of host memory");
    c->device = device;
    c->rank = 0; c->world = 1; c->comm = nullptr; c->group = nullptr;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess) { delete c; return fail(ABCDEZ_ERR_CUDA, cudaGetErrorString(e2)); }
        c->own_stream = true;
    }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return ABCDEZ_OK;
}

extern "C" int abcdez_destroy(abcdez_ctx* ctx)
{
    if (!ctx) return ABCDEZ_OK;
    if (!ctx->subs.empty()) {                             // abcdez_init_multi: the sub-contexts own everything
        for (abcdez_ctx* sub : ctx->subs) abcdez_destroy(sub);
        if (ctx->group) local_group_destroy(ctx->group);
        delete ctx;
        return ABCDEZ_OK;
    }
    for (abcdez_ctx* w : ctx->batch) abcdez_destroy(w);
    ctx->batch.clear();
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->comm) comm_destroy(ctx->comm);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->h_ctrl_pool) cudaFreeHost(ctx->h_ctrl_pool);
    for (int b = 0; b < 2; ++b) { if (ctx->h_poll[b]) cudaFreeHost(ctx->h_poll[b]); if (ctx->poll_ev[b]) cudaEventDestroy(ctx->poll_ev[b]); }
    delete ctx;
    return ABCDEZ_OK;
}

extern "C" int abcdez_sync(abcdez_ctx* ctx)
{
    CHECK_ARG(ctx != nullptr, "abcdez_sync: ctx is NULL");
    if (!ctx->subs.empty()) { for (abcdez_ctx* sub : ctx->subs) { CU(cudaSetDevice(sub->device)); CU(cudaStreamSynchronize(sub->stream)); } return ABCDEZ_OK; }
    CU(cudaStreamSynchronize(ctx->stream));
    return ABCDEZ_OK;
}

// ---- single-process multi-GPU context (SURVEY.md 8b "Threading") -----------------------------------------
// One call from one host thread uses every GPU: the context fans a run out to one worker thread per GPU, each
// driving its own sub-context (stream, arena) exactly as one process per GPU would; peers are mapped with
// cudaDeviceEnablePeerAccess instead of CUDA IPC and the host-level collectives are an in-process rendezvous
// (comm.cu), so neither NCCL nor a process launcher is involved.  Same kernels, same in-kernel NVLink exchanges,
// same results as the one-process-per-GPU runs (and as oracle(islands = n_gpus)).
extern "C" int abcdez_init_multi(int n_gpus, const int* device_ids, abcdez_ctx** out)
{
    CHECK_ARG(out != nullptr, "abcdez_init_multi: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(ABCDEZ_ERR_CUDA, std::string("abcdez_init_multi: no CUDA device (") + cudaGetErrorString(e) + "); libabcdez_cuda has no CPU fallback");
    if (n_gpus <= 0) n_gpus = n < XCHG_MAXR ? n : XCHG_MAXR;          // all visible GPUs (one NVSwitch domain: at most 8)
    CHECK_ARG(n_gpus <= XCHG_MAXR, "abcdez_init_multi: at most 8 GPUs");
    std::vector<int> ids(n_gpus);
    for (int r = 0; r < n_gpus; ++r) {
        ids[r] = device_ids ? device_ids[r] : r;
        CHECK_ARG(ids[r] >= 0 && ids[r] < n, "abcdez_init_multi: device index out of range");
        for (int q = 0; q < r; ++q) CHECK_ARG(ids[q] != ids[r], "abcdez_init_multi: a device is listed twice");
    }
    for (int r = 0; r < n_gpus; ++r)
        for (int q = 0; q < n_gpus; ++q) {
            int can = 1;
            if (q != r) CU(cudaDeviceCanAccessPeer(&can, ids[r], ids[q]));
            if (!can) return fail(ABCDEZ_ERR_UNSUPPORTED, "abcdez_init_multi: the GPUs cannot access each other's memory (no NVLink / PCIe peer access)");
        }
    abcdez_ctx* top = new (std::nothrow) abcdez_ctx();
    if (!top) return fail(ABCDEZ_ERR_CUDA, "abcdez_init_multi: out of host memory");
    top->device = ids[0]; top->rank = 0; top->world = 1; top->comm = nullptr; top->group = nullptr;
    for (int r = 0; r < n_gpus; ++r) {
        abcdez_ctx* sub = nullptr;
        int rc = abcdez_init(ids[r], nullptr, &sub);
        if (rc) { std::string why = g_err; abcdez_destroy(top); return fail(rc, why); }
        top->subs.push_back(sub);
    }
    if (n_gpus == 1) { *out = top; return ABCDEZ_OK; }    // (a single GPU needs no communicator: runs go straight to the sub-context)
    top->group = local_group_create(n_gpus);
    std::vector<int> rcs(n_gpus, 0); std::vector<std::string> whys(n_gpus);
    std::vector<std::thread> th;
    for (int r = 0; r < n_gpus; ++r)
        th.emplace_back([&, r] {
            abcdez_ctx* sub = top->subs[r];
            cudaSetDevice(sub->device);
            rcs[r] = comm_create_local(r, n_gpus, top->group, sub->device, sub->stream, &sub->comm, &whys[r]);
            if (!rcs[r]) { sub->rank = r; sub->world = n_gpus; }
        });
    for (std::thread& t : th) t.join();
    for (int r = 0; r < n_gpus; ++r)
        if (rcs[r]) { std::string why = whys[r]; int rc = rcs[r]; abcdez_destroy(top); return fail(rc, "abcdez_init_multi: " + why); }
    cudaSetDevice(ids[0]);
    *out = top;
    return ABCDEZ_OK;
}

extern "C" int abcdez_ctx_gpus(const abcdez_ctx* ctx)
{
    if (!ctx) return 0;
    return ctx->subs.empty() ? 1 : (int)ctx->subs.size();
}

// ---- sharded runs: one process per GPU -------------------------------------------------
extern "C" int abcdez_nccl_unique_id(void* id128)
{
    CHECK_ARG(id128 != nullptr, "abcdez_nccl_unique_id: id128 is NULL");
    std::string why;
    int rc = nccl_unique_id(id128, &why);
    return rc ? fail(rc, why) : ABCDEZ_OK;
}

extern "C" int abcdez_comm_init(abcdez_ctx* ctx, int rank, int world, const void* id128)
{
    CHECK_ARG(ctx != nullptr, "abcdez_comm_init: ctx is NULL");
    CHECK_ARG(ctx->subs.empty(), "abcdez_comm_init: an abcdez_init_multi context already spans its GPUs");
    CHECK_ARG(world >= 1 && world <= XCHG_MAXR && rank >= 0 && rank < world, "abcdez_comm_init: need 0 <= rank < world <= 8");
    CHECK_ARG(ctx->comm == nullptr, "abcdez_comm_init: the context already has a communicator");
    ctx->rank = rank; ctx->world = world;
    if (world == 1) return ABCDEZ_OK;
    CHECK_ARG(id128 != nullptr, "abcdez_comm_init: id128 is NULL");
    CU(cudaSetDevice(ctx->device));
    std::string why;
    int rc = comm_create(rank, world, id128, ctx->stream, &ctx->comm, &why);
    if (rc) { ctx->rank = 0; ctx->world = 1; return fail(rc, "abcdez_comm_init: " + why); }
    return ABCDEZ_OK;
}

extern "C" int abcdez_comm_selftest(abcdez_ctx* ctx, int rounds, int mode, uint64_t* checksum, double* us_per_round)
{
    CHECK_ARG(ctx && ctx->comm && checksum, "abcdez_comm_selftest: needs a context with a communicator");
    CHECK_ARG(rounds >= 1 && rounds <= 100000, "abcdez_comm_selftest: rounds out of range");
    CHECK_ARG(mode == 0 || mode == 1, "abcdez_comm_selftest: mode must be 0 (fenced ring) or 1 (low-latency ring)");
    CU(cudaSetDevice(ctx->device));
    unsigned long long r = 0;
    int rc = comm_selftest(ctx->comm, ctx->stream, rounds, mode, &r, us_per_round);
    *checksum = r;
    return rc ? fail(rc, std::string("abcdez_comm_selftest: ") + comm_error(ctx->comm)) : ABCDEZ_OK;
}

// contiguous block of global particle indices owned by `rank`: [floor(rank N / world), floor((rank+1) N / world))
extern "C" int abcdez_shard_range(int64_t N, int rank, int world, int64_t* lo, int64_t* hi)
{
    CHECK_ARG(N >= 0 && world >= 1 && rank >= 0 && rank < world && lo && hi, "abcdez_shard_range: bad argument");
    *lo = (int64_t)(((__int128)N * rank) / world);
    *hi = (int64_t)(((__int128)N * (rank + 1)) / world);
    return ABCDEZ_OK;
}

// ---------------------------------------------------------------------------------------
// prior
// ---------------------------------------------------------------------------------------
extern "C" int abcdez_prior_create(abcdez_ctx* ctx, int d, const int32_t* family, const double* params,
                                   abcdez_prior** out)
{
    CHECK_ARG(ctx && family && params && out, "abcdez_prior_create: NULL argument");
    CHECK_ARG(d >= 1 && d <= ABCDEZ_MAXD, "abcdez_prior_create: length(prior) must be in 1..16");
    abcdez_prior* p = new (std::nothrow) abcdez_prior();
    if (!p) return fail(ABCDEZ_ERR_CUDA, "out of host memory");
    memset(&p->dev, 0, sizeof(p->dev));
    p->dev.d = d;
    for (int k = 0; k < d; ++k) {
        const double* q = params + 4 * k;
        p->dev.family[k] = family[k];
        for (int j = 0; j < 4; ++j) p->dev.p[k][j] = q[j];
        bool ok = true; double c = 0.0;
        switch (family[k]) {
        case ABCDEZ_NORMAL: ok = q[1] > 0.0; c = log(q[1]); p->dev.p[k][2] = ok ? pdiv_host_reciprocal(q[1]) : 0.0; break;
        case ABCDEZ_UNIFORM: ok = q[0] < q[1]; c = -log(q[1] - q[0]); break;
        case ABCDEZ_DISCRETE_UNIFORM: ok = q[0] <= q[1] && q[0] == rint(q[0]) && q[1] == rint(q[1]);
            c = log(1.0 / (q[1] - q[0] + 1.0)); break;
        case ABCDEZ_LOGNORMAL: ok = q[1] > 0.0; c = log(q[1]); break;
        case ABCDEZ_EXPONENTIAL: ok = q[0] > 0.0; c = log(q[0]); break;
        case ABCDEZ_GAMMA: ok = q[0] > 0.0 && q[1] > 0.0; c = lgamma(q[0]) + q[0] * log(q[1]); break;
        case ABCDEZ_BETA: ok = q[0] > 0.0 && q[1] > 0.0; c = lgamma(q[0]) + lgamma(q[1]) - lgamma(q[0] + q[1]); break;
        case ABCDEZ_NEGBIN: ok = q[0] > 0.0 && q[1] > 0.0 && q[1] <= 1.0; c = q[0] * log(q[1]) - lgamma(q[0]); break;
        default: delete p; return fail(ABCDEZ_ERR_UNSUPPORTED, "abcdez_prior_create: unsupported marginal family");
        }
        if (!ok) { delete p; return fail(ABCDEZ_ERR_BAD_ARG, "abcdez_prior_create: invalid marginal parameters"); }
        p->dev.c[k] = c;
    }
    *out = p;
    return ABCDEZ_OK;
}

extern "C" int abcdez_prior_destroy(abcdez_prior* p) { delete p; return ABCDEZ_OK; }

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
}  // namespace

static int prior_op(abcdez_ctx* ctx, const abcdez_prior* p, int64_t N, int op, const double* in, double* out,
                    uint64_t seed, uint32_t epoch, int64_t id0)
{
    ctx = first_gpu(ctx);
    CHECK_ARG(ctx && p && out, "prior op: NULL argument");
    CHECK_ARG(N >= 0, "prior op: N < 0");
    if (N == 0) return ABCDEZ_OK;
    CU(cudaSetDevice(ctx->device));
    int d = p->dev.d;
    size_t nin = (size_t)N * d * 8, nout = (op == PRIOR_OP_LOGPDF ? (size_t)N : (size_t)N * d) * 8;
    DevBuf din, dout;
    CU(din.alloc(nin)); CU(dout.alloc(nout));
    if (op != PRIOR_OP_SAMPLE) { CHECK_ARG(in != nullptr, "prior op: theta is NULL"); CU(cudaMemcpyAsync(din.p, in, nin, cudaMemcpyHostToDevice, ctx->stream)); }
    launch_prior_op(ctx->stream, d, p->dev, N, op, din.as<double>(), dout.as<double>(), seed, epoch, (uint32_t)id0);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, nout, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ABCDEZ_OK;
}

extern "C" int abcdez_prior_sample(abcdez_ctx* ctx, const abcdez_prior* p, int64_t N, uint64_t seed, uint32_t epoch,
                                   int64_t id0, double* theta_out)
{
    return prior_op(ctx, p, N, PRIOR_OP_SAMPLE, nullptr, theta_out, seed, epoch, id0);
}
extern "C" int abcdez_prior_logpdf(abcdez_ctx* ctx, const abcdez_prior* p, int64_t N, const double* theta, double* out)
{
    return prior_op(ctx, p, N, PRIOR_OP_LOGPDF, theta, out, 0, 0, 0);
}
extern "C" int abcdez_prior_push(abcdez_ctx* ctx, const abcdez_prior* p, int64_t N, const double* theta, double* out)
{
    return prior_op(ctx, p, N, PRIOR_OP_PUSH, theta, out, 0, 0, 0);
}

// ---------------------------------------------------------------------------------------
// models
// ---------------------------------------------------------------------------------------
extern "C" int abcdez_model_count(void) { return model_count(); }
extern "C" const char* abcdez_model_name(int id) { const ModelOps* o = model_ops(id); return o ? o->name : nullptr; }
extern "C" int abcdez_model_lookup(const char* name, int* id)
{
    CHECK_ARG(name && id, "abcdez_model_lookup: NULL argument");
    for (int i = 0; i < model_count(); ++i)
        if (strcmp(model_ops(i)->name, name) == 0) { *id = i; return ABCDEZ_OK; }
    return fail(ABCDEZ_ERR_BAD_ARG, std::string("abcdez_model_lookup: no registered model named '") + name + "'");
}
extern "C" int abcdez_model_info(int id, int* d, int* blob_bytes)
{
    const ModelOps* o = model_ops(id);
    CHECK_ARG(o != nullptr, "abcdez_model_info: bad model id");
    if (d) *d = o->d;
    if (blob_bytes) *blob_bytes = o->blob;
    return ABCDEZ_OK;
}
// derived constants a simulator would otherwise recompute per particle (the same IEEE operations, done once)
static void model_prepare_data(int id, double* v)
{
    if (id == M_GAUSS_CORR10) {                     // sqrt(1 - rho^2) of the AR(1) noise, models.cuh
        volatile double r2 = v[10] * v[10];         // (volatile: never contracted into an fma)
        volatile double om = 1.0 - r2;
        v[11] = sqrt(om);
    }
}

// the per-model layout of the bound data (DESIGN.md "Models"): counts embedded in the data index further entries of
// the 64-double block or bound loops inside the kernels, so they are checked here, once
static const char* model_check_data(int id, const double* v, size_t n)
{
    auto whole = [](double x, double lo, double hi) { return x == floor(x) && x >= lo && x <= hi; };
    switch (id) {
    case M_GAUSS1D: case M_GAUSS1D_BLOB: if (n < 2) return "gauss1d: data = (observation, sigma)"; break;
    case M_GAUSS_CORR10: if (n < 11 || !(fabs(v[10]) < 1.0)) return "gauss_corr10: data = y_obs[10], rho with |rho| < 1"; break;
    case M_DIRAC: case M_NORMDU: case M_MIXTURE: if (n < 1) return "this model needs data = (observation)"; break;
    case M_WIENER: if (n < 31) return "wiener: data = 31 summary values"; break;
    case M_SOCKS: if (n < 2) return "socks: data = (pairs, odds)"; break;
    case M_LOTKA_VOLTERRA: case M_LOTKA_VOLTERRA_LIN:
        if (n < 6 || !whole(v[4], 1, 29) || n < 6 + 2 * (size_t)v[4]) return "lotka_volterra: data = x0, y0, dt, substeps, nobs (1..29), sigma, obs[2 nobs]";
        if (!whole(v[3], 1, 100000) || !(v[2] > 0.0)) return "lotka_volterra: dt must be positive and substeps in 1..100000";
        break;
    case M_BIRTH_DEATH:
        if (n < 4 || !whole(v[1], 1, 60) || n < 4 + (size_t)v[1]) return "birth_death: data = n0, nobs (1..60), dt, max_events, obs[nobs]";
        if (!(v[3] >= 0.0 && v[3] <= 1e9) || !(v[2] > 0.0) || !(v[0] >= 0.0)) return "birth_death: need n0 >= 0, dt > 0 and 0 <= max_events <= 1e9";
        break;
    case M_GK: case M_GK_F32:
        if (n < 8 || !whole(v[0], 8, 16384)) return "gk: data = n (8..16384 draws), 7 observed octiles";
        break;
    }
    return nullptr;
}

extern "C" int abcdez_model_bind(abcdez_ctx* ctx, int id, const double* data, size_t ndata, abcdez_model** out)
{
    CHECK_ARG(ctx && out, "abcdez_model_bind: NULL argument");
    const ModelOps* o = model_ops(id);
    CHECK_ARG(o != nullptr, "abcdez_model_bind: bad model id");
    if (!o->smc_sweep) return fail(ABCDEZ_ERR_UNSUPPORTED, std::string("model '") + o->name + "' has no device functor in this build");
    CHECK_ARG(ndata <= ABCDEZ_MAXDATA, "abcdez_model_bind: more than ABCDEZ_MAXDATA doubles of data");
    CHECK_ARG(ndata == 0 || data != nullptr, "abcdez_model_bind: data is NULL");
    if (const char* why = model_check_data(id, data, ndata)) return fail(ABCDEZ_ERR_BAD_ARG, std::string("abcdez_model_bind: ") + why);
    abcdez_model* m = new (std::nothrow) abcdez_model();
    if (!m) return fail(ABCDEZ_ERR_CUDA, "out of host memory");
    m->id = id; m->ops = o;
    memset(&m->data, 0, sizeof(m->data));
    for (size_t i = 0; i < ndata; ++i) m->data.v[i] = data[i];
    model_prepare_data(id, m->data.v);
    *out = m;
    return ABCDEZ_OK;
}
extern "C" int abcdez_model_destroy(abcdez_model* m) { delete m; return ABCDEZ_OK; }

// runtime-supplied models (rtc.cu): compile the CUDA source of one model struct against the library's kernel
// templates and register it; ctx == NULL compiles only (no GPU needed) and leaves *id = -1
extern "C" int abcdez_model_compile(abcdez_ctx* ctx, const char* name, const char* struct_name, const char* cuda_src,
                                    int d, int blob_bytes, int* id, char* log, size_t log_cap)
{
    ctx = first_gpu(ctx);
    CHECK_ARG(name && struct_name && cuda_src && id, "abcdez_model_compile: NULL argument");
    CHECK_ARG(d >= 1 && d <= ABCDEZ_MAXD, "abcdez_model_compile: d must be in 1..16");
    CHECK_ARG(blob_bytes >= 0 && blob_bytes <= ABCDEZ_MAXBLOB && blob_bytes % 8 == 0, "abcdez_model_compile: blob_bytes must be a multiple of 8 in 0..64");
    if (ctx) {
        for (int i = 0; i < model_count(); ++i)
            if (strcmp(model_ops(i)->name, name) == 0) return fail(ABCDEZ_ERR_BAD_ARG, std::string("abcdez_model_compile: a model named '") + name + "' is already registered");
        CU(cudaSetDevice(ctx->device));
        CU(cudaFree(nullptr));                      // the primary context is current for the driver calls
    }
    std::string lg;
    int rc = rtc_compile_model(name, struct_name, cuda_src, d, blob_bytes, ctx != nullptr, id, &lg);
    if (log && log_cap) { size_t n = lg.size() < log_cap - 1 ? lg.size() : log_cap - 1; memcpy(log, lg.data(), n); log[n] = 0; }
    return rc ? fail(rc, "abcdez_model_compile: " + lg) : ABCDEZ_OK;
}

extern "C" int abcdez_simulate(abcdez_ctx* ctx, const abcdez_model* m, int64_t N, const double* theta_pushed,
                               uint64_t seed, uint32_t epoch, uint32_t tag, int64_t id0, double* dist_out,
                               uint8_t* blobs_out)
{
    ctx = first_gpu(ctx);
    CHECK_ARG(ctx && m && theta_pushed && dist_out, "abcdez_simulate: NULL argument");
    CHECK_ARG(N >= 0, "abcdez_simulate: N < 0");
    if (N == 0) return ABCDEZ_OK;
    CU(cudaSetDevice(ctx->device));
    int d = m->ops->d, B = m->ops->blob;
    DevBuf dth, dd, db;
    CU(dth.alloc((size_t)N * d * 8)); CU(dd.alloc((size_t)N * 8)); CU(db.alloc((size_t)N * (B ? B : 8)));
    CU(cudaMemcpyAsync(dth.p, theta_pushed, (size_t)N * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    m->ops->simulate(*m->ops, ctx->stream, nullptr, m->data, N, dth.as<double>(), seed, epoch, tag, (uint32_t)id0,
                     dd.as<double>(), B ? db.as<double>() : nullptr);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dist_out, dd.p, (size_t)N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (B && blobs_out) CU(cudaMemcpyAsync(blobs_out, db.p, (size_t)N * B, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ABCDEZ_OK;
}

extern "C" double abcdez_kernel_logpdf(int kernel, double eps, double x) { return abck_logpdf(kernel, eps, x); }
extern "C" double abcdez_kernel_pdf(int kernel, double eps, double x)
{
    if (!abck_insupport(kernel, eps, x)) return 0.0;
    if (abck_is_indicator(kernel)) return 1.0;
    double q = x / eps;
    return 1.0 - q * q;
}

// ---------------------------------------------------------------------------------------
// population
// ---------------------------------------------------------------------------------------
__global__ void fill_f64_kernel(double* p, int64_t n, double v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static int push_ctrl(abcdez_pop* pop)
{
    CU(cudaMemcpyAsync(pop->dev.ctrl, pop->h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, pop->ctx->stream));
    return ABCDEZ_OK;
}
static int pull_ctrl(abcdez_pop* pop)
{
    CU(cudaMemcpyAsync(pop->h_ctrl, pop->dev.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, pop->ctx->stream));
    CU(cudaStreamSynchronize(pop->ctx->stream));
    return ABCDEZ_OK;
}

static int ensure_scratch(abcdez_pop* pop, size_t bytes)
{
    if (bytes <= pop->scratch_bytes) return ABCDEZ_OK;
    return fail(ABCDEZ_ERR_BAD_ARG, "internal: population scratch too small");   // sized for every caller in pop_create_impl
}

extern "C" int abcdez_pop_destroy(abcdez_pop* pop)
{
    if (!pop) return ABCDEZ_OK;
    cudaSetDevice(pop->ctx->device);
    cudaStreamSynchronize(pop->ctx->stream);
    if (pop->slab) {
        if (pop->slab_from_arena) pop->ctx->arena_busy = false;     // the slab stays with the context
        else if (!pop->slab_shared) cudaFree(pop->slab);            // (the shared arena stays with the communicator)
    }
    if (pop->sorted_delta) cudaFree(pop->sorted_delta);
    if (pop->order) cudaFree(pop->order);
    if (pop->sort_tmp) cudaFree(pop->sort_tmp);
    if (pop->h_ctrl) {
        if (pop->h_ctrl_from_pool) pop->ctx->h_ctrl_busy = false;
        else cudaFreeHost(pop->h_ctrl);
    }
    if (pop->ev0) cudaEventDestroy(pop->ev0);
    if (pop->ev1) cudaEventDestroy(pop->ev1);
    delete pop;
    return ABCDEZ_OK;
}

// Ng > 0: this population is one shard (N particles from global index id0) of a sharded population of Ng
// particles; its slab then comes from the communicator's shared arena (mapped by every peer; collective call).
static int pop_create_impl(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, int64_t N,
                           int64_t id0, int hist_cap, abcdez_pop** out, int64_t Ng = 0)
{
    ctx = first_gpu(ctx);
    CHECK_ARG(ctx && prior && model && out, "abcdez_pop_create: NULL argument");
    CHECK_ARG(N >= 1 && N < (int64_t)0x7fffffff, "abcdez_pop_create: N must be in 1..2^31-2 per GPU");
    CHECK_ARG(prior->dev.d == model->ops->d, "abcdez_pop_create: length(prior) != model dimension");
    CHECK_ARG(id0 >= 0 && id0 + N <= (int64_t)0xffffffffll, "abcdez_pop_create: global particle ids must fit 32 bits");
    CU(cudaSetDevice(ctx->device));
    abcdez_pop* pop = new (std::nothrow) abcdez_pop();
    if (!pop) return fail(ABCDEZ_ERR_CUDA, "out of host memory");
    memset(static_cast<void*>(pop), 0, sizeof(*pop));
    pop->ctx = ctx; pop->prior = prior->dev; pop->ops = model->ops; pop->data = model->data;
    pop->N = N; pop->D = model->ops->d; pop->DS = row_stride(pop->D); pop->NB = model->ops->blob / 8;
    pop->hist_cap = hist_cap;
    PopDev& P = pop->dev;
    P.N = (uint32_t)N; P.id0 = (uint32_t)id0; P.ntiles = (uint32_t)((N + TILE - 1) / TILE);
    const bool shard = Ng > 0;
    P.Ng = (uint32_t)(shard ? Ng : N);
    P.x.rank = 0; P.x.world = 1; P.x.seq = nullptr; P.peers = nullptr;
    for (int q = 0; q < XCHG_MAXR; ++q) P.x.mbox[q] = nullptr;
    size_t n = (size_t)N;
    // one slab for every device array of the population (256-byte aligned pieces), carved from the
    // context arena when it is free, else from a private allocation
    struct Piece { void** ptr; size_t bytes; };
    const size_t scratch_need = std::max((size_t)29 * n + 64, n * (size_t)pop->D * 8);
    const bool split = model->ops->split != 0;
    Piece pieces[] = {
        { (void**)&P.theta[0], n * pop->DS * 8 }, { (void**)&P.theta[1], n * pop->DS * 8 },
        { (void**)&P.logpi[0], n * 8 }, { (void**)&P.logpi[1], n * 8 },
        { (void**)&P.delta[0], n * 8 }, { (void**)&P.delta[1], n * 8 },
        { (void**)&P.blob[0], n * pop->NB * 8 }, { (void**)&P.blob[1], n * pop->NB * 8 },
        { (void**)&P.W, n * 8 }, { (void**)&P.alive, n + 4 }, { (void**)&P.moved, n + 4 },
        { (void**)&P.alive_list, n * 4 }, { (void**)&P.ctrl, sizeof(Ctrl) },
        { (void**)&P.partial, (size_t)P.ntiles * 2 * 8 + 64 }, { (void**)&P.tile_cnt, (size_t)P.ntiles * 4 + 64 },
        { (void**)&P.sel_hist, 7 * SEL_BINS * 4 }, { (void**)&P.cumsum, n * 8 }, { (void**)&P.inds, n * 4 },
        { (void**)&P.hist, (size_t)(hist_cap > 0 ? hist_cap : 1) * 8 * 8 }, { (void**)&P.tabs, 2 * sizeof(SeqTab) },
        { (void**)&pop->scratch, scratch_need },
        // queue-driven sweeps of heavy simulators (ModelOps::split): theta', its log prior, the simulated distance / blob, the queue
        { (void**)&P.prop_theta, split ? n * pop->DS * 8 : 0 }, { (void**)&P.prop_lp, split ? n * 8 : 0 }, { (void**)&P.prop_dp, split ? n * 8 : 0 },
        { (void**)&P.prop_blob, split ? n * pop->NB * 8 : 0 }, { (void**)&P.prop_flag, split ? n + 4 : 0 }, { (void**)&P.queue, split ? n * 4 : 0 },
    };
    size_t total = 0;
    for (const Piece& pc : pieces) total += (pc.bytes + 255) & ~(size_t)255;
    if (shard) {
        int rc_ = comm_shared_slab(ctx->comm, ctx->stream, total, &pop->slab);
        if (rc_) { delete pop; return fail(rc_, std::string("abcdez: shared slab: ") + comm_error(ctx->comm)); }
        pop->slab_shared = true; pop->slab_from_arena = false;
    } else if (!ctx->arena_busy) {
        if (ctx->arena_bytes < total) {
            if (ctx->arena) cudaFree(ctx->arena);
            ctx->arena = nullptr; ctx->arena_bytes = 0;
            cudaError_t e_ = cudaMalloc((void**)&ctx->arena, total);
            if (e_ != cudaSuccess) { delete pop; return fail(ABCDEZ_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e_)); }
            ctx->arena_bytes = total;
        }
        pop->slab = ctx->arena; pop->slab_from_arena = true; ctx->arena_busy = true;
    } else {
        cudaError_t e_ = cudaMalloc((void**)&pop->slab, total);
        if (e_ != cudaSuccess) { delete pop; return fail(ABCDEZ_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e_)); }
        pop->slab_from_arena = false;
    }
    pop->slab_bytes = total;
    {
        size_t off = 0;
        for (const Piece& pc : pieces) { *pc.ptr = pop->slab + off; off += (pc.bytes + 255) & ~(size_t)255; }
        if (!split) { P.prop_theta = P.prop_lp = P.prop_dp = P.prop_blob = nullptr; P.prop_flag = nullptr; P.queue = nullptr; }
    }
    pop->scratch_bytes = scratch_need;
    P.cand[0] = reinterpret_cast<unsigned long long*>(P.cumsum);
    P.cand[1] = reinterpret_cast<unsigned long long*>(pop->scratch);
    if (!ctx->h_ctrl_busy) {
        if (!ctx->h_ctrl_pool) {
            cudaError_t e_ = cudaMallocHost((void**)&ctx->h_ctrl_pool, sizeof(Ctrl));
            if (e_ != cudaSuccess) { abcdez_pop_destroy(pop); return fail(ABCDEZ_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e_)); }
        }
        pop->h_ctrl = ctx->h_ctrl_pool; pop->h_ctrl_from_pool = true; ctx->h_ctrl_busy = true;
    } else {
        cudaError_t e_ = cudaMallocHost((void**)&pop->h_ctrl, sizeof(Ctrl));
        if (e_ != cudaSuccess) { pop->h_ctrl = nullptr; abcdez_pop_destroy(pop); return fail(ABCDEZ_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e_)); }
        pop->h_ctrl_from_pool = false;
    }
    CU(cudaEventCreateWithFlags(&pop->ev0, cudaEventDefault)); CU(cudaEventCreateWithFlags(&pop->ev1, cudaEventDefault));
    cudaStream_t st = ctx->stream;
    CU(cudaMemsetAsync(P.sel_hist, 0, 7 * SEL_BINS * 4, st));
    CU(cudaMemsetAsync(P.alive, 1, n, st));
    CU(cudaMemsetAsync(P.moved, 1, n, st));
    CU(cudaMemsetAsync(P.theta[0], 0, n * pop->DS * 8, st));
    CU(cudaMemsetAsync(P.theta[1], 0, n * pop->DS * 8, st));
    fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.W, N, 1.0 / (double)P.Ng);
    Ctrl* c = pop->h_ctrl;
    memset(c, 0, sizeof(Ctrl));
    c->eps = INFINITY; c->eps_k = INFINITY; c->eps_target = 0.0;
    c->facc = 1.0; c->gamma0 = 2.38 / sqrt(2.0 * (double)pop->D); c->gsig = 1e-5;     // src/abcdez_smc.jl:280-281
    c->alpha = 0.95; c->Kmcmc_min = 1.0; c->facc_tune = 0.975;
    c->nsims_max = (long long)10000000; c->kind = ABCDEZ_INDICATOR_STRICT; c->Kmcmc = 3; c->Ki = 3;
    c->N = (uint32_t)N; c->n_alive = (uint32_t)N; c->Ng = P.Ng; c->n_alive_g = P.Ng; c->ess_min = 0.5 * (double)P.Ng;
    c->hist_cap = hist_cap;
    c->acc.dmin_key = ~0ull; c->acc.dmax_key = 0ull; c->acc.min_gt_key = ~0ull;
    for (int q = 0; q < 6; ++q) { c->acc.cand_min[q] = ~0ull; c->acc.cand_max[q] = 0ull; }
    c->acc.min_above = ~0ull;
    c->acc.w_alive = 1.0 / (double)P.Ng;
    int rc = push_ctrl(pop);
    if (rc) { abcdez_pop_destroy(pop); return rc; }
    CU(cudaStreamSynchronize(st));
    *out = pop;
    return ABCDEZ_OK;
}

extern "C" int abcdez_pop_create(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, int64_t N,
                                 int64_t id0, abcdez_pop** out)
{
    return pop_create_impl(ctx, prior, model, N, id0, 0, out);
}

extern "C" int abcdez_pop_upload(abcdez_pop* pop, const double* theta, const double* logpi, const double* delta,
                                 const uint8_t* blobs, const double* W, const uint8_t* alive)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_upload: pop is NULL");
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    cudaStream_t st = pop->ctx->stream;
    PopDev& P = pop->dev; int cur = pop->h_ctrl->cur; size_t n = (size_t)pop->N;
    if (theta) {
        if (pop->DS == pop->D) CU(cudaMemcpyAsync(P.theta[cur], theta, n * pop->D * 8, cudaMemcpyHostToDevice, st));
        else CU(cudaMemcpy2DAsync(P.theta[cur], (size_t)pop->DS * 8, theta, (size_t)pop->D * 8, (size_t)pop->D * 8, n,
                                  cudaMemcpyHostToDevice, st));
    }
    if (logpi) CU(cudaMemcpyAsync(P.logpi[cur], logpi, n * 8, cudaMemcpyHostToDevice, st));
    if (delta) CU(cudaMemcpyAsync(P.delta[cur], delta, n * 8, cudaMemcpyHostToDevice, st));
    if (blobs && pop->NB) CU(cudaMemcpyAsync(P.blob[cur], blobs, n * pop->NB * 8, cudaMemcpyHostToDevice, st));
    if (W) CU(cudaMemcpyAsync(P.W, W, n * 8, cudaMemcpyHostToDevice, st));
    if (theta || logpi || delta || blobs) CU(cudaMemsetAsync(P.moved, 1, n, st));   // the other buffer is stale now
    if (alive) {
        CU(cudaMemcpyAsync(P.alive, alive, n, cudaMemcpyHostToDevice, st));
        // keep the control block and the compacted list consistent with the uploaded flags
        uint32_t na = 0; double wal = 0.0;
        for (size_t i = 0; i < n; ++i) if (alive[i]) { na++; if (W && wal == 0.0) wal = W[i]; }
        std::vector<uint32_t> list; list.reserve(n);       // alive first, then dead (see compact_kernel)
        for (size_t i = 0; i < n; ++i) if (alive[i]) list.push_back((uint32_t)i);
        for (size_t i = 0; i < n; ++i) if (!alive[i]) list.push_back((uint32_t)i);
        CU(cudaMemcpyAsync(P.alive_list, list.data(), n * 4, cudaMemcpyHostToDevice, st));
        pop->h_ctrl->n_alive = na; pop->h_ctrl->n_alive_g = na;
        if (W) pop->h_ctrl->acc.w_alive = wal;
        CU(cudaStreamSynchronize(st));
        rc = push_ctrl(pop); if (rc) return rc;
    }
    CU(cudaStreamSynchronize(st));
    return ABCDEZ_OK;
}

extern "C" int abcdez_pop_download(abcdez_pop* pop, double* theta, double* logpi, double* delta, uint8_t* blobs,
                                   double* W, uint8_t* alive)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_download: pop is NULL");
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    cudaStream_t st = pop->ctx->stream;
    PopDev& P = pop->dev; int cur = pop->h_ctrl->cur; size_t n = (size_t)pop->N;
    if (theta) {
        if (pop->DS == pop->D) CU(cudaMemcpyAsync(theta, P.theta[cur], n * pop->D * 8, cudaMemcpyDeviceToHost, st));
        else CU(cudaMemcpy2DAsync(theta, (size_t)pop->D * 8, P.theta[cur], (size_t)pop->DS * 8, (size_t)pop->D * 8, n,
                                  cudaMemcpyDeviceToHost, st));
    }
    if (logpi) CU(cudaMemcpyAsync(logpi, P.logpi[cur], n * 8, cudaMemcpyDeviceToHost, st));
    if (delta) CU(cudaMemcpyAsync(delta, P.delta[cur], n * 8, cudaMemcpyDeviceToHost, st));
    if (blobs && pop->NB) CU(cudaMemcpyAsync(blobs, P.blob[cur], n * pop->NB * 8, cudaMemcpyDeviceToHost, st));
    if (W) CU(cudaMemcpyAsync(W, P.W, n * 8, cudaMemcpyDeviceToHost, st));
    if (alive) CU(cudaMemcpyAsync(alive, P.alive, n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ABCDEZ_OK;
}

extern "C" int abcdez_pop_set(abcdez_pop* pop, double eps, double eps_kernel_prev, int32_t kernel, double gamma0,
                              double gamma_sigma, uint64_t seed, uint32_t sweep_epoch)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_set: pop is NULL");
    CHECK_ARG(kernel >= 0 && kernel <= 3, "abcdez_pop_set: unknown ABC kernel");
    CHECK_ARG(eps >= 0.0 && eps_kernel_prev >= 0.0, "Expected \xcf\xb5 \xe2\x89\xa5 0.0");   // src/abcdez_types.jl:30
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    Ctrl* c = pop->h_ctrl;
    c->eps = eps; c->eps_k = eps_kernel_prev; c->kind = kernel; c->gamma0 = gamma0; c->gsig = gamma_sigma;
    c->seed = seed; c->sweep_epoch = sweep_epoch;
    pop->dev.keys = philox_keys(seed);
    c->stop = 0; c->sweeps_done = 0; c->sweep_idx = 0; c->Kmcmc = 0x7fffffff; c->Ki = 1;
    c->Kmcmc_min = INFINITY; c->naccs_iter = 0; c->err = 0; c->acc.err = 0;
    rc = push_ctrl(pop); if (rc) return rc;
    CU(cudaStreamSynchronize(pop->ctx->stream));
    return ABCDEZ_OK;
}

static int check_dev_err(abcdez_pop* pop, const char* where)
{
    int e = pop->h_ctrl->err ? pop->h_ctrl->err : pop->h_ctrl->acc.err;
    if (!e) return ABCDEZ_OK;
    const char* what = e == ABCDEZ_ERR_NAN_DISTANCE ? "NaN distance among alive particles (quantile undefined)"
                     : e == ABCDEZ_ERR_NO_ALIVE ? "No alive particles"
                     : e == ABCDEZ_ERR_INIT_RETRY ? "abcde_init!: redraw limit reached without a finite distance/log-prior"
                     : e == ABCDEZ_ERR_PARTNER_RETRY ? "partner draw did not terminate (fewer than 3 alive particles?)"
                     : e == ABCDEZ_ERR_BAD_ARG ? "weights are not normalised (wprod > 0 but wprod / wnorm == 0)"
                     : e == ABCDEZ_ERR_NCCL ? "in-kernel exchange with a peer GPU timed out"
                     : "device-side error";
    char num[64]; snprintf(num, sizeof num, " [ctrl.err=%d acc.err=%d]", pop->h_ctrl->err, pop->h_ctrl->acc.err);
    return fail(e, std::string(where) + ": " + what + num);
}

#define TIME_BEGIN(pop) CU(cudaEventRecord((pop)->ev0, (pop)->ctx->stream))
#define TIME_END(pop, nl)                                                                          \
    do {                                                                                           \
        CU(cudaEventRecord((pop)->ev1, (pop)->ctx->stream));                                       \
        CU(cudaEventSynchronize((pop)->ev1));                                                      \
        float ms_ = 0.f; CU(cudaEventElapsedTime(&ms_, (pop)->ev0, (pop)->ev1));                   \
        (pop)->last_ms = ms_; (pop)->last_launches = (nl);                                         \
    } while (0)

extern "C" int abcdez_pop_last_timing(abcdez_pop* pop, double* ms, int64_t* launches)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_last_timing: pop is NULL");
    if (ms) *ms = pop->last_ms;
    if (launches) *launches = pop->last_launches;
    return ABCDEZ_OK;
}

extern "C" int abcdez_pop_init(abcdez_pop* pop, uint64_t seed, int draw_prior, int64_t* nredraws)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_init: pop is NULL");
    CU(cudaSetDevice(pop->ctx->device));
    TIME_BEGIN(pop);
    pop->ops->init(*pop->ops, pop->ctx->stream, pop->dev, pop->prior, pop->data, seed, draw_prior);
    CU(cudaGetLastError());
    TIME_END(pop, 1);
    int rc = pull_ctrl(pop); if (rc) return rc;
    if (nredraws) *nredraws = pop->h_ctrl->redraws;
    return check_dev_err(pop, "abcdez_pop_init");
}

// copy optional injected host arrays into the scratch buffer; returns device pointers
static int stage_inject(abcdez_pop* pop, const int32_t* s, const int32_t* a, const int32_t* b, const double* z,
                        const double* u, bool want_flags, SweepInj* inj)
{
    size_t n = (size_t)pop->N;
    size_t bytes = 3 * n * 4 + 2 * n * 8 + n + 64;
    int rc = ensure_scratch(pop, bytes); if (rc) return rc;
    char* base = (char*)pop->scratch;
    double* dz = (double*)base; double* du = dz + n;
    int32_t* ds = (int32_t*)(du + n); int32_t* da = ds + n; int32_t* db = da + n;
    uint8_t* df = (uint8_t*)(db + n);
    cudaStream_t st = pop->ctx->stream;
    memset(inj, 0, sizeof(*inj));
    if (z) { CU(cudaMemcpyAsync(dz, z, n * 8, cudaMemcpyHostToDevice, st)); inj->z = dz; }
    if (u) { CU(cudaMemcpyAsync(du, u, n * 8, cudaMemcpyHostToDevice, st)); inj->u = du; }
    if (s) { CU(cudaMemcpyAsync(ds, s, n * 4, cudaMemcpyHostToDevice, st)); inj->s = ds; }
    if (a && b) {
        CU(cudaMemcpyAsync(da, a, n * 4, cudaMemcpyHostToDevice, st)); inj->a = da;
        CU(cudaMemcpyAsync(db, b, n * 4, cudaMemcpyHostToDevice, st)); inj->b = db;
    }
    if (want_flags) { CU(cudaMemsetAsync(df, 0, n, st)); inj->flags = df; }
    return ABCDEZ_OK;
}

extern "C" int abcdez_pop_smc_sweep(abcdez_pop* pop, const int32_t* inj_a, const int32_t* inj_b, const double* inj_z,
                                    const double* inj_u, uint8_t* flags_out, int64_t* nsims, int64_t* naccs)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_smc_sweep: pop is NULL");
    CHECK_ARG((inj_a == nullptr) == (inj_b == nullptr), "abcdez_pop_smc_sweep: inject both partners or neither");
    CU(cudaSetDevice(pop->ctx->device));
    SweepInj inj;
    int rc = stage_inject(pop, nullptr, inj_a, inj_b, inj_z, inj_u, flags_out != nullptr, &inj); if (rc) return rc;
    TIME_BEGIN(pop);
    pop->ops->smc_sweep(*pop->ops, pop->ctx->stream, pop->dev, pop->prior, pop->data, inj);
    CU(cudaGetLastError());
    TIME_END(pop, 1);
    if (flags_out) CU(cudaMemcpyAsync(flags_out, inj.flags, (size_t)pop->N, cudaMemcpyDeviceToHost, pop->ctx->stream));
    rc = pull_ctrl(pop); if (rc) return rc;
    if (nsims) *nsims = (int64_t)pop->h_ctrl->last_nsims;
    if (naccs) *naccs = (int64_t)pop->h_ctrl->last_naccs;
    return check_dev_err(pop, "abcdez_pop_smc_sweep");
}

static int mc_prepare(abcdez_pop* pop)
{
    size_t n = (size_t)pop->N;
    if (!pop->sorted_delta) {
        CU(cudaMalloc((void**)&pop->sorted_delta, n * 8));
        CU(cudaMalloc((void**)&pop->order, n * 4));
        pop->sort_tmp_bytes = mc_sort_tmp_bytes(pop->N);
        CU(cudaMalloc(&pop->sort_tmp, pop->sort_tmp_bytes));
    }
    return ABCDEZ_OK;
}

extern "C" int abcdez_pop_mc_sweep(abcdez_pop* pop, double eps_pop, double eps_target, const int32_t* inj_s,
                                   const int32_t* inj_a, const int32_t* inj_b, const double* inj_z, const double* inj_u,
                                   uint8_t* flags_out, int64_t* nsims)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_mc_sweep: pop is NULL");
    CHECK_ARG((inj_a == nullptr) == (inj_b == nullptr), "abcdez_pop_mc_sweep: inject both partners or neither");
    if (!pop->ops->mc_sweep) return fail(ABCDEZ_ERR_UNSUPPORTED, "abcdez_pop_mc_sweep: this model has no abcdemc! sweep in this build");
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    rc = mc_prepare(pop); if (rc) return rc;
    SweepInj inj;
    rc = stage_inject(pop, inj_s, inj_a, inj_b, inj_z, inj_u, flags_out != nullptr, &inj); if (rc) return rc;
    cudaStream_t st = pop->ctx->stream;
    TIME_BEGIN(pop);
    int nl = 1;
    if (!inj_s) nl += launch_mc_prepare(st, pop->dev, pop->sorted_delta, pop->order, pop->sort_tmp, pop->sort_tmp_bytes, 1);
    McArgs mc{ eps_pop, eps_target, 0, pop->sorted_delta, pop->order };
    pop->ops->mc_sweep(*pop->ops, st, pop->dev, pop->prior, pop->data, inj, mc);
    CU(cudaGetLastError());
    TIME_END(pop, nl);
    if (flags_out) CU(cudaMemcpyAsync(flags_out, inj.flags, (size_t)pop->N, cudaMemcpyDeviceToHost, st));
    rc = pull_ctrl(pop); if (rc) return rc;
    if (nsims) *nsims = (int64_t)pop->h_ctrl->last_nsims;
    return check_dev_err(pop, "abcdez_pop_mc_sweep");
}

extern "C" int abcdez_pop_eps_quantile(abcdez_pop* pop, double alpha, double* q, double* v_lo, double* v_hi)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_eps_quantile: pop is NULL");
    CHECK_ARG(alpha >= 0.0 && alpha <= 1.0, "abcdez_pop_eps_quantile: alpha out of [0,1]");
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    Ctrl* c = pop->h_ctrl;
    CHECK_ARG(c->n_alive >= 1, "abcdez_pop_eps_quantile: no alive particles");
    double eps_keep = c->eps, tgt_keep = c->eps_target;
    c->alpha = alpha; c->stop = 0; c->err = 0; c->acc.err = 0;
    c->eps = INFINITY; c->eps_target = -INFINITY;      // the stage call returns the raw quantile
    {   // host-side select_setup (same arithmetic as ctrl.cuh)
        unsigned long long n = c->n_alive;
        double m = 1.0 - alpha, aleph = fma((double)n, alpha, m);
        long long j = (long long)trunc(aleph);
        if (j > (long long)n - 1) j = (long long)n - 1;
        if (j < 1) j = 1;
        double g = aleph - (double)j; g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
        c->sel_j = (unsigned long long)j; c->sel_rank = (unsigned long long)(j - 1); c->sel_prefix = 0; c->q_gamma = g;
        c->acc.cnt_le = 0; c->acc.min_gt_key = ~0ull;
    }
    rc = push_ctrl(pop); if (rc) return rc;
    TIME_BEGIN(pop);
    int nl = launch_eps_quantile(pop->ctx->stream, pop->dev);
    CU(cudaGetLastError());
    TIME_END(pop, nl);
    rc = pull_ctrl(pop); if (rc) return rc;
    if (q) *q = c->q;
    if (v_lo) *v_lo = c->q_a;
    if (v_hi) *v_hi = c->q_b;
    rc = check_dev_err(pop, "abcdez_pop_eps_quantile");
    c->eps = eps_keep; c->eps_target = tgt_keep;
    int rc2 = push_ctrl(pop); if (rc2) return rc2;
    CU(cudaStreamSynchronize(pop->ctx->stream));
    return rc;
}

extern "C" int abcdez_pop_reweight(abcdez_pop* pop, double eps_new, double* wnorm, double* ess, int64_t* n_alive)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_reweight: pop is NULL");
    CHECK_ARG(eps_new >= 0.0, "Expected \xcf\xb5 \xe2\x89\xa5 0.0");
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    Ctrl* c = pop->h_ctrl;
    c->eps = eps_new; c->stop = 0; c->ess_min = -1.0;       // stage call: never triggers the resample flag
    rc = push_ctrl(pop); if (rc) return rc;
    TIME_BEGIN(pop);
    int nl = launch_reweight(pop->ctx->stream, pop->dev);
    nl += launch_compact(pop->ctx->stream, pop->dev);
    CU(cudaGetLastError());
    TIME_END(pop, nl);
    rc = pull_ctrl(pop); if (rc) return rc;
    c->eps_k = eps_new;                                     // src/abcdez_smc.jl:360
    rc = push_ctrl(pop); if (rc) return rc;
    CU(cudaStreamSynchronize(pop->ctx->stream));
    if (wnorm) *wnorm = c->wnorm;
    if (ess) *ess = c->ess;
    if (n_alive) *n_alive = c->n_alive;
    return ABCDEZ_OK;
}

// the fused head of one iteration (head.cu): eps quantile + clamp, reweight, ESS, alive list
extern "C" int abcdez_pop_head(abcdez_pop* pop, double alpha, double eps_target, double* q, double* eps,
                               double* wnorm, double* ess, int64_t* n_alive)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_head: pop is NULL");
    CHECK_ARG(alpha >= 0.0 && alpha <= 1.0, "abcdez_pop_head: alpha out of [0,1]");
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    Ctrl* c = pop->h_ctrl;
    CHECK_ARG(c->n_alive >= 1, "abcdez_pop_head: no alive particles");
    c->alpha = alpha; c->eps_target = eps_target; c->stop = 0; c->err = 0; c->acc.err = 0;
    c->ess_min = -1.0;                                      // stage call: never triggers the resample flag
    {   // host-side select_setup (same arithmetic as ctrl.cuh)
        unsigned long long n = c->n_alive;
        double m = 1.0 - alpha, aleph = fma((double)n, alpha, m);
        long long j = (long long)trunc(aleph);
        if (j > (long long)n - 1) j = (long long)n - 1;
        if (j < 1) j = 1;
        double g = aleph - (double)j; g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
        c->sel_j = (unsigned long long)j; c->sel_rank = (unsigned long long)(j - 1); c->sel_prefix = 0; c->q_gamma = g;
    }
    rc = push_ctrl(pop); if (rc) return rc;
    TIME_BEGIN(pop);
    int nl = launch_head(pop->ctx->stream, pop->dev, pop->ctx->sm_count);
    if (nl < 0) return fail(ABCDEZ_ERR_CUDA, std::string("abcdez_pop_head: cooperative launch of the head kernel failed: ") + cudaGetErrorString(cudaGetLastError()));
    CU(cudaGetLastError());
    TIME_END(pop, nl);
    rc = pull_ctrl(pop); if (rc) return rc;
    if (q) *q = c->q;
    if (eps) *eps = c->eps;
    if (wnorm) *wnorm = c->wnorm;
    if (ess) *ess = c->ess;
    if (n_alive) *n_alive = c->n_alive;
    rc = check_dev_err(pop, "abcdez_pop_head");
    c->eps_k = c->eps;                                      // src/abcdez_smc.jl:360
    int rc2 = push_ctrl(pop); if (rc2) return rc2;
    CU(cudaStreamSynchronize(pop->ctx->stream));
    return rc;
}

extern "C" int abcdez_pop_resample(abcdez_pop* pop, const double* uniforms, uint32_t epoch, int mode, int32_t* inds_out)
{
    CHECK_ARG(pop != nullptr, "abcdez_pop_resample: pop is NULL");
    CHECK_ARG(mode >= 0 && mode <= 2, "abcdez_pop_resample: mode must be 0, 1 or 2");
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    Ctrl* c = pop->h_ctrl;
    CHECK_ARG(c->n_alive >= 1, "abcdez_pop_resample: no alive particles");
    int use_mode = mode;
    if (mode == 0 && !abck_is_indicator(c->kind)) use_mode = 1;
    size_t n = (size_t)pop->N;
    const double* du = nullptr;
    if (uniforms) {
        rc = ensure_scratch(pop, n * 8); if (rc) return rc;
        CU(cudaMemcpyAsync(pop->scratch, uniforms, n * 8, cudaMemcpyHostToDevice, pop->ctx->stream));
        du = (const double*)pop->scratch;
    }
    c->stop = 0;
    rc = push_ctrl(pop); if (rc) return rc;
    TIME_BEGIN(pop);
    int nl = launch_resample(pop->ctx->stream, pop->dev, pop->DS, pop->NB, du, epoch, use_mode, 1);
    CU(cudaGetLastError());
    TIME_END(pop, nl);
    if (inds_out) CU(cudaMemcpyAsync(inds_out, pop->dev.inds, n * 4, cudaMemcpyDeviceToHost, pop->ctx->stream));
    rc = pull_ctrl(pop); if (rc) return rc;
    return ABCDEZ_OK;
}

extern "C" int abcdez_wsample_stratified(abcdez_ctx* ctx, int64_t N, const double* weights, const double* uniforms,
                                         int mode, int64_t* inds_out)
{
    ctx = first_gpu(ctx);
    CHECK_ARG(ctx && weights && uniforms && inds_out, "abcdez_wsample_stratified: NULL argument");
    CHECK_ARG(N >= 1 && N < (int64_t)0x7fffffff, "abcdez_wsample_stratified: N out of range");
    CHECK_ARG(mode >= 0 && mode <= 2, "abcdez_wsample_stratified: mode must be 0, 1 or 2");
    CU(cudaSetDevice(ctx->device));
    size_t n = (size_t)N; unsigned nt = (unsigned)((N + TILE - 1) / TILE);
    DevBuf dw, du, dc, dp, dt, di;
    CU(dw.alloc(n * 8)); CU(du.alloc(n * 8)); CU(dc.alloc(n * 8)); CU(dp.alloc((size_t)nt * 8 + 64));
    CU(dt.alloc(2 * sizeof(SeqTab))); CU(di.alloc(n * 8));
    CU(cudaMemcpyAsync(dw.p, weights, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(du.p, uniforms, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    // tabs[1] holds the strata edges; strat_indices_kernel reads &tabs[0] + 1 -> pass the base
    launch_strat_indices(ctx->stream, N, dw.as<double>(), du.as<double>(), dc.as<double>(), dp.as<double>(),
                         dt.as<SeqTab>(), mode == 2 ? 2 : 1, di.as<long long>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(inds_out, di.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ABCDEZ_OK;
}

extern "C" int abcdez_pop_bench_sweeps(abcdez_pop* pop, int sweeps, int64_t* nsims, int64_t* naccs, double* ms)
{
    CHECK_ARG(pop != nullptr && sweeps >= 1, "abcdez_pop_bench_sweeps: bad argument");
    CU(cudaSetDevice(pop->ctx->device));
    int rc = pull_ctrl(pop); if (rc) return rc;
    Ctrl* c = pop->h_ctrl;
    Ctrl keep = *c;
    c->stop = 0; c->sweeps_done = 0; c->sweep_idx = 0; c->Kmcmc = 0x7fffffff; c->Kmcmc_min = INFINITY;
    long long ns0 = c->nsims_total; unsigned long long na0 = c->naccs_iter;
    rc = push_ctrl(pop); if (rc) return rc;
    SweepInj inj; memset(&inj, 0, sizeof(inj));
    TIME_BEGIN(pop);
    for (int s = 0; s < sweeps; ++s) pop->ops->smc_sweep(*pop->ops, pop->ctx->stream, pop->dev, pop->prior, pop->data, inj);
    CU(cudaGetLastError());
    TIME_END(pop, sweeps);
    rc = pull_ctrl(pop); if (rc) return rc;
    if (nsims) *nsims = c->nsims_total - ns0;
    if (naccs) *naccs = (int64_t)(c->naccs_iter - na0);
    if (ms) *ms = pop->last_ms;
    // restore the schedule scalars but keep the evolved particle state (cur, epoch)
    keep.cur = c->cur; keep.sweep_epoch = c->sweep_epoch; keep.nsims_total = c->nsims_total; keep.n_sweeps = c->n_sweeps;
    keep.dmin = c->dmin; keep.dmax = c->dmax;
    *c = keep;
    rc = push_ctrl(pop); if (rc) return rc;
    CU(cudaStreamSynchronize(pop->ctx->stream));
    return check_dev_err(pop, "abcdez_pop_bench_sweeps");
}

// ---------------------------------------------------------------------------------------
// abcdesmc!  (src/abcdez_smc.jl:215-394)
// ---------------------------------------------------------------------------------------
extern "C" void abcdez_smc_opts_default(abcdez_smc_opts* o)
{
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->nparticles = 100; o->alpha = 0.95; o->delta_ess = 0.5; o->nsims_max = 10000000; o->Kmcmc = 3;
    o->Kmcmc_min = 1.0; o->kernel = ABCDEZ_INDICATOR_STRICT; o->facc_stop = 0.0; o->facc_min = 0.0;
    o->facc_tune = 0.975; o->seed = 1; o->verboseout = 1; o->max_iters = 0; o->exact_scan = 0; o->profile = 0;
    o->sync_every = 1; o->fused_head = 1; o->systematic_resampling = 0; o->partner_segments = 0; o->fp32_state = 0;
}

// ---- run state snapshots (the reference has no equivalent) -------------------------------
// Everything abcdesmc! carries from one iteration to the next lives in the device control block, the history
// buffer and the live generation of the particle arrays, so a snapshot taken between two iterations is those
// bytes, and a restored run continues decision by decision like the uninterrupted one (the random streams are
// counter based: particle id, sweep epoch and iteration number are part of the state).
namespace {
struct SmcStateHdr {         // 96 bytes
    uint64_t magic; uint32_t version, d, ds, nb; int64_t N; int32_t hist_cap, model_id; uint64_t ctrl_bytes, total_bytes;
    uint64_t prior_hash;     // FNV-1a of the prior's families / parameters and the bound data
    char model[32];
};
static_assert(sizeof(SmcStateHdr) == 96, "snapshot header layout (DESIGN.md section 4b)");
uint64_t prior_hash(const PriorDev& pr, const ModelData& md)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    mix(&pr.d, sizeof pr.d);
    for (int k = 0; k < pr.d; ++k) { mix(&pr.family[k], sizeof(int32_t)); mix(pr.p[k], 2 * sizeof(double)); }
    mix(md.v, sizeof md.v);
    return h;
}
constexpr uint64_t SMC_STATE_MAGIC = 0x3174535a65644342ull;   // "BCdeZSt1"
size_t smc_state_size(int64_t N, int ds, int nb, int hist_cap)
{
    size_t n = (size_t)N;
    return sizeof(SmcStateHdr) + ((sizeof(Ctrl) + 7) & ~(size_t)7) + (size_t)hist_cap * 64 + n * ds * 8 + 2 * n * 8 + n * nb * 8 + n * 8
           + ((n + 7) & ~(size_t)7);
}
}  // namespace

extern "C" int64_t abcdez_smc_state_bytes(const abcdez_prior* prior, const abcdez_model* model, int64_t nparticles, int32_t hist_cap)
{
    if (!prior || !model || nparticles < 1) return -1;
    const int d = model->ops->d;
    return (int64_t)smc_state_size(nparticles, row_stride(d), model->ops->blob / 8, hist_cap > 0 ? hist_cap : 1);
}

static int smc_run_impl(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                        const abcdez_smc_opts* o, abcdez_smc_result* res, const void* state_in, size_t state_in_bytes,
                        void* state_out, size_t state_out_cap, size_t* state_out_bytes);

// abcdez_init_multi contexts: one worker thread per GPU runs the sharded implementation on its sub-context and fills its
// block of the caller's buffers; scalars and histories (identical on every rank) come from rank 0
static int smc_run_multi(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                         const abcdez_smc_opts* o, abcdez_smc_result* res)
{
    const int R = (int)ctx->subs.size();
    if (R == 1) return smc_run_impl(ctx->subs[0], prior, model, eps_target, o, res, nullptr, 0, nullptr, 0, nullptr);
    const int d = prior->dev.d, B = model->ops->blob;
    std::vector<abcdez_smc_result> rr(R, *res);
    std::vector<int> rcs(R, 0); std::vector<std::string> whys(R);
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r) {
        int64_t lo = 0, hi = 0;
        abcdez_shard_range(o->nparticles, r, R, &lo, &hi);
        if (rr[r].P) rr[r].P += lo * d;
        if (rr[r].Wns) rr[r].Wns += lo;
        if (rr[r].C) rr[r].C += lo;
        if (rr[r].blobs) rr[r].blobs += lo * B;
        if (r > 0) { rr[r].h_eps = rr[r].h_dmin = rr[r].h_dmax = rr[r].h_logZ = rr[r].h_ess = rr[r].h_facc = rr[r].h_gamma0 = nullptr; rr[r].h_Kmcmc = nullptr; }
        th.emplace_back([&, r] {
            rcs[r] = smc_run_impl(ctx->subs[r], prior, model, eps_target, o, &rr[r], nullptr, 0, nullptr, 0, nullptr);
            if (rcs[r]) whys[r] = g_err;
        });
    }
    for (std::thread& t : th) t.join();
    cudaSetDevice(ctx->device);
    for (int r = 0; r < R; ++r) if (rcs[r]) return fail(rcs[r], whys[r]);
    const abcdez_smc_result& a = rr[0];
    res->eps = a.eps; res->logZ = a.logZ; res->iters = a.iters; res->nsims = a.nsims; res->hist_len = a.hist_len; res->status = a.status;
    res->n_resamples = a.n_resamples; res->n_sweeps = a.n_sweeps; res->hist_dropped = a.hist_dropped;
    res->sweep_ms = a.sweep_ms; res->total_ms = a.total_ms; res->init_ms = a.init_ms; res->head_ms = a.head_ms; res->resample_ms = a.resample_ms;
    res->n_launches = 0;
    for (int r = 0; r < R; ++r) { res->n_launches += rr[r].n_launches; if (rr[r].total_ms > res->total_ms) res->total_ms = rr[r].total_ms; }
    return ABCDEZ_OK;
}

extern "C" int abcdez_smc_run(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                              const abcdez_smc_opts* o, abcdez_smc_result* res)
{
    CHECK_ARG(ctx && prior && model && o && res, "abcdez_smc_run: NULL argument");
    if (!ctx->subs.empty()) return smc_run_multi(ctx, prior, model, eps_target, o, res);
    return smc_run_impl(ctx, prior, model, eps_target, o, res, nullptr, 0, nullptr, 0, nullptr);
}

extern "C" int abcdez_smc_run_state(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                                    const abcdez_smc_opts* o, abcdez_smc_result* res, const void* state_in,
                                    int64_t state_in_bytes, void* state_out, int64_t state_out_cap, int64_t* state_out_bytes)
{
    CHECK_ARG(state_in_bytes >= 0 && state_out_cap >= 0, "abcdez_smc_run_state: negative size");
    CHECK_ARG(ctx != nullptr, "abcdez_smc_run_state: ctx is NULL");
    if (!ctx->subs.empty()) {
        if (ctx->subs.size() > 1 && (state_in || state_out)) return fail(ABCDEZ_ERR_UNSUPPORTED, "abcdez_smc_run_state: run-state snapshots are single-GPU in this build");
        if (ctx->subs.size() > 1) return smc_run_multi(ctx, prior, model, eps_target, o, res);
        ctx = ctx->subs[0];
    }
    CHECK_ARG(!state_out || state_out_bytes, "abcdez_smc_run_state: state_out needs state_out_bytes");
    size_t ob = 0;
    int rc = smc_run_impl(ctx, prior, model, eps_target, o, res, state_in, (size_t)state_in_bytes, state_out, (size_t)state_out_cap,
                          &ob);
    if (state_out_bytes) *state_out_bytes = (int64_t)ob;
    return rc;
}

static int smc_run_impl(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                        const abcdez_smc_opts* o, abcdez_smc_result* res, const void* state_in, size_t state_in_bytes,
                        void* state_out, size_t state_out_cap, size_t* state_out_bytes)
{
    CHECK_ARG(ctx && prior && model && o && res, "abcdez_smc_run: NULL argument");
    // argument validation with the reference's messages, src/abcdez_smc.jl:223-235
    CHECK_ARG(0.0 <= o->alpha && o->alpha < 1.0, "\xce\xb1 must be in 0 <= \xce\xb1 < 1");
    CHECK_ARG(0.0 <= o->delta_ess && o->delta_ess <= 1.0, "\xce\xb4""ess must be in 0 <= \xce\xb4""ess <= 1");
    CHECK_ARG(0.0 <= o->facc_stop && o->facc_stop <= 1.0, "facc_stop must be in 0 <= facc_stop <= 1");
    CHECK_ARG(0.0 <= o->facc_min && o->facc_min <= 1.0, "facc_min must be in 0 <= facc_min <= 1");
    CHECK_ARG(0.0 <= o->facc_tune && o->facc_tune <= 1.0, "facc_tune must be in 0 <= facc_tune <= 1");
    CHECK_ARG(0.0 <= eps_target, "\xcf\xb5_target must be non-negative");
    CHECK_ARG(1 <= o->Kmcmc, "Kmcmc must be at least 1");
    CHECK_ARG(0.0 <= o->Kmcmc_min, "Kmcmc_min must be in 0 <= Kmcmc_min <= Inf");
    CHECK_ARG(1 <= o->nsims_max, "nsims_max must be at least 1");
    CHECK_ARG(o->kernel >= 0 && o->kernel <= 3, "unknown ABC kernel");
    if (o->fp32_state) {
        if (!model->ops->f32_state) return fail(ABCDEZ_ERR_UNSUPPORTED, std::string("fp32_state: model '") + model->ops->name + "' keeps FP64 particle state in this build (g-and-k and runtime-compiled models)");
        if (state_in || state_out) return fail(ABCDEZ_ERR_UNSUPPORTED, "fp32_state: run-state snapshots hold FP64 particle state");
    }
    {
        double mn = o->alpha < o->delta_ess ? o->alpha : o->delta_ess;
        double nmin = ceil(3.0 * (double)prior->dev.d / mn);                 // :234
        if (!((double)o->nparticles >= nmin)) {
            char buf[96]; snprintf(buf, sizeof buf, "nparticles must be at least %.0f", nmin);
            return fail(ABCDEZ_ERR_BAD_ARG, buf);
        }
    }
    CU(cudaSetDevice(ctx->device));
    static const bool trace = getenv("ABCDEZ_TRACE") != nullptr;     // host-side phase timings on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    auto t_begin = now();
    double t_create = 0.0, t_loop = 0.0, t_result = 0.0;
    // sharded run (abcdez_comm_init with world > 1): nparticles is the size of the WHOLE population; this rank
    // owns the contiguous block abcdez_shard_range(nparticles, rank, world) and returns that block's rows
    const bool shard = ctx->comm != nullptr && ctx->world > 1;
    const int64_t Ng = o->nparticles;
    int64_t lo = 0, hi = Ng;
    if (shard) {
        abcdez_shard_range(Ng, ctx->rank, ctx->world, &lo, &hi);
        CHECK_ARG(Ng >= 4 * (int64_t)ctx->world, "sharded abcdesmc!: need at least 4 particles per rank");
        CHECK_ARG(Ng < (int64_t)0x7fffffff, "sharded abcdesmc!: nparticles must be below 2^31");
    }
    const int64_t N = hi - lo;
    int hist_cap = o->verboseout ? (res->hist_cap > 0 ? res->hist_cap : 0) : 0;
    int dev_hist = hist_cap > 0 ? hist_cap : 1;
    if (state_in || state_out) {
        // sharded contexts: every rank dumps / restores its own block behind its own copy of the control block (the schedule
        // scalars in it are identical on all ranks, the counts are the rank's); the mailbox sequence restarts with every run
        const size_t need = smc_state_size(N, row_stride(prior->dev.d), model->ops->blob / 8, dev_hist);
        if (state_out && state_out_cap < need) return fail(ABCDEZ_ERR_BAD_ARG, "abcdez_smc_run_state: state_out is smaller than abcdez_smc_state_bytes()");
        if (state_in) {
            const SmcStateHdr* h = (const SmcStateHdr*)state_in;
            const bool ok = state_in_bytes >= sizeof(SmcStateHdr) && h->magic == SMC_STATE_MAGIC && h->version == 2 &&
                            h->ctrl_bytes == sizeof(Ctrl) && h->N == N && (int)h->d == prior->dev.d &&
                            (int)h->nb == model->ops->blob / 8 && strncmp(h->model, model->ops->name, sizeof(h->model) - 1) == 0 &&
                            h->total_bytes == need && state_in_bytes >= need;
            if (!ok) return fail(ABCDEZ_ERR_BAD_ARG, "abcdez_smc_run_state: state_in is not a snapshot of this model / prior / nparticles / history capacity");
        }
    }
    abcdez_pop* pop = nullptr;
    int rc = pop_create_impl(ctx, prior, model, N, lo, dev_hist, &pop, shard ? Ng : 0);
    if (rc) return rc;
    if (shard) {
        rc = comm_begin_run(ctx->comm, ctx->stream, pop->dev, &pop->dev.x, &pop->dev.peers);
        if (rc) { abcdez_pop_destroy(pop); return fail(rc, std::string("sharded abcdesmc!: ") + comm_error(ctx->comm)); }
    }
    t_create = ms_since(t_begin);
    cudaStream_t st = ctx->stream;
    Ctrl* c = pop->h_ctrl;
    c->eps_target = eps_target; c->alpha = o->alpha; c->ess_min = (double)Ng * o->delta_ess;   // :259
    c->nsims_max = o->nsims_max; c->Kmcmc = o->Kmcmc; c->Ki = o->Kmcmc; c->Kmcmc_min = o->Kmcmc_min;
    c->kind = o->kernel; c->facc_stop = o->facc_stop; c->facc_min = o->facc_min; c->facc_tune = o->facc_tune;
    c->seed = o->seed; c->max_iters = o->max_iters; c->hist_cap = hist_cap;
    if (state_in) {
        // the snapshot's control block, with this call's targets: eps_target, nsims_max, facc_stop and the iteration
        // budget (max_iters counts the iterations of THIS call); every other option must be the snapshot's
        const SmcStateHdr* h = (const SmcStateHdr*)state_in;
        Ctrl s0; memcpy(&s0, (const char*)state_in + sizeof(SmcStateHdr), sizeof(Ctrl));
        const bool same = s0.alpha == o->alpha && s0.ess_min == (double)Ng * o->delta_ess && s0.Kmcmc == o->Kmcmc &&
                          s0.Kmcmc_min == o->Kmcmc_min && s0.kind == o->kernel && s0.facc_min == o->facc_min &&
                          s0.facc_tune == o->facc_tune && h->hist_cap == dev_hist && s0.hist_cap == hist_cap;
        if (!same) { abcdez_pop_destroy(pop); return fail(ABCDEZ_ERR_BAD_ARG, "abcdez_smc_run_state: options differ from the snapshot's (only eps_target, nsims_max, facc_stop and max_iters may change)"); }
        // the control block is trusted state: indices into buffers, tickets the kernels count on being zero, the prior it was drawn under
        bool sane = (s0.cur == 0 || s0.cur == 1) && s0.hist_len >= 0 && s0.hist_len <= dev_hist && s0.N == (unsigned)N && s0.Ng == (unsigned)Ng &&
                    s0.n_alive <= s0.N && s0.n_alive_g <= s0.Ng && s0.iters >= 0 && s0.sweep_idx >= 0 && s0.acc.sweep_nsims == 0 && s0.acc.sweep_naccs == 0 &&
                    h->prior_hash == prior_hash(prior->dev, model->data);
        for (int q = 0; q < 8; ++q) sane = sane && s0.acc.ticket[q] == 0u;
        for (int q = 0; q < 6; ++q) sane = sane && s0.acc.cand_count[q] == 0ull;
        if (!sane) { abcdez_pop_destroy(pop); return fail(ABCDEZ_ERR_BAD_ARG, "abcdez_smc_run_state: the snapshot's control block is inconsistent (corrupt, or taken under another prior / observed data)"); }
        *c = s0;
        c->eps_target = eps_target; c->nsims_max = o->nsims_max; c->facc_stop = o->facc_stop;
        c->max_iters = o->max_iters > 0 ? c->iters + o->max_iters : 0;
        c->status = 0;
        const bool no_alive = c->n_alive_g == 0;
        c->stop = (no_alive || c->eps <= c->eps_target || c->nsims_total >= c->nsims_max || c->facc < c->facc_stop || c->err) ? 1 : 0;   // :375-376
        if (no_alive) c->status = ABCDEZ_ERR_NO_ALIVE;
    }
    pop->dev.keys = philox_keys(c->seed);
    pop->dev.flags = (o->partner_segments ? POP_PARTNER_SEGMENTS : 0u) | (o->systematic_resampling ? POP_SYSTEMATIC : 0u) | (o->fp32_state ? POP_FP32_STATE : 0u);
    rc = push_ctrl(pop);
    std::vector<cudaEvent_t>& evs = ctx->ev_pool;
    size_t nev = 0;
    cudaEvent_t e_init0 = nullptr, e_loop0 = nullptr, e_loop1 = nullptr;
    int64_t launches = 0;
    int64_t host_iters = 0;
    const int sync_every = o->sync_every > 0 ? o->sync_every : 1;
    const int mode = abck_is_indicator(o->kernel) ? 0 : (o->exact_scan ? 2 : 1);
    SweepInj noinj; memset(&noinj, 0, sizeof(noinj));
    int pb = 0; bool have_prev = false;
    for (int b = 0; b < 2 && !rc; ++b) {
        if (!ctx->h_poll[b] && cudaMallocHost((void**)&ctx->h_poll[b], sizeof(Ctrl)) != cudaSuccess) rc = fail(ABCDEZ_ERR_CUDA, "cudaMallocHost(poll buffer) failed");
        if (!rc && !ctx->poll_ev[b] && cudaEventCreateWithFlags(&ctx->poll_ev[b], cudaEventDisableTiming) != cudaSuccess) rc = fail(ABCDEZ_ERR_CUDA, "cudaEventCreate(poll event) failed");
    }
#define RUN_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(ABCDEZ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); goto done; } } while (0)
    if (rc) goto done;
    RUN_CU(cudaEventCreate(&e_init0)); RUN_CU(cudaEventCreate(&e_loop0)); RUN_CU(cudaEventCreate(&e_loop1));
    RUN_CU(cudaEventRecord(e_init0, st));
    if (state_in) {
        // restore: control block (with this call's targets), history, live generation; the other generation is stale
        const char* p = (const char*)state_in + sizeof(SmcStateHdr) + ((sizeof(Ctrl) + 7) & ~(size_t)7);
        const size_t n = (size_t)N; const int cur = c->cur;
        RUN_CU(cudaMemcpyAsync(pop->dev.hist, p, (size_t)dev_hist * 64, cudaMemcpyHostToDevice, st)); p += (size_t)dev_hist * 64;
        RUN_CU(cudaMemcpyAsync(pop->dev.theta[cur], p, n * pop->DS * 8, cudaMemcpyHostToDevice, st)); p += n * pop->DS * 8;
        RUN_CU(cudaMemcpyAsync(pop->dev.logpi[cur], p, n * 8, cudaMemcpyHostToDevice, st)); p += n * 8;
        RUN_CU(cudaMemcpyAsync(pop->dev.delta[cur], p, n * 8, cudaMemcpyHostToDevice, st)); p += n * 8;
        if (pop->NB) RUN_CU(cudaMemcpyAsync(pop->dev.blob[cur], p, n * pop->NB * 8, cudaMemcpyHostToDevice, st));
        p += n * pop->NB * 8;
        RUN_CU(cudaMemcpyAsync(pop->dev.W, p, n * 8, cudaMemcpyHostToDevice, st)); p += n * 8;
        RUN_CU(cudaMemcpyAsync(pop->dev.alive, p, n, cudaMemcpyHostToDevice, st));
        RUN_CU(cudaMemsetAsync(pop->dev.moved, 1, n, st));
        host_iters = c->iters;
    } else {
        pop->ops->init(*pop->ops, st, pop->dev, pop->prior, pop->data, o->seed, 1);     // :242-252
        launches += (pop->ops->split == 2 ? 3 : 1) + launch_begin_run(st, pop->dev);   // :255-292 (stepped simulators: draw, simulate, finish)
    }
    RUN_CU(cudaEventRecord(e_loop0, st));
    for (;;) {                                                               // :295
        host_iters++;
        if (o->profile) {
            while (evs.size() < nev + 4) { cudaEvent_t e; RUN_CU(cudaEventCreate(&e)); evs.push_back(e); }
            RUN_CU(cudaEventRecord(evs[nev], st));
        }
        if (o->fused_head || shard) {                                        // :301-324 in one cooperative kernel
            const int nl = launch_head(st, pop->dev, ctx->sm_count);
            if (nl < 0) {    // nothing else may run on stale eps / alive state: abort the run at once
                rc = fail(ABCDEZ_ERR_CUDA, std::string("abcdez_smc_run: cooperative launch of the head kernel failed: ") + cudaGetErrorString(cudaGetLastError()));
                cudaStreamSynchronize(st);
                goto done;
            }
            launches += nl;
        } else {
            launches += launch_eps_quantile(st, pop->dev);                   // :301
            launches += launch_reweight(st, pop->dev);                       // :305-324
            launches += launch_compact(st, pop->dev);
        }
        // profile: CUDA events around the head, the resampling launches and the iteration's sweep launches
        // (skipped launches return at once)
        if (o->profile) {
            while (evs.size() < nev + 4) { cudaEvent_t e; RUN_CU(cudaEventCreate(&e)); evs.push_back(e); }
            RUN_CU(cudaEventRecord(evs[nev + 1], st));
        }
        launches += launch_resample(st, pop->dev, pop->DS, pop->NB, nullptr, (uint32_t)host_iters, mode, 0);   // :324-326
        if (o->profile) RUN_CU(cudaEventRecord(evs[nev + 2], st));
        for (int k = 0; k < o->Kmcmc; ++k) {                                 // :336-353
            pop->ops->smc_sweep(*pop->ops, st, pop->dev, pop->prior, pop->data, noinj);
            launches += pop->ops->split ? 3 : 1;                             // (heavy simulators: propose, simulate, accept)
        }
        if (o->profile) { RUN_CU(cudaEventRecord(evs[nev + 3], st)); nev += 4; }
        if (host_iters % sync_every == 0) {
            // pipelined stop poll: request a copy of the control block behind this batch, then look at the copy
            // requested behind the PREVIOUS batch -- the host stays one batch ahead of the device, so the device
            // never idles while the host waits and re-enqueues (kernels behind a stop return at once)
            RUN_CU(cudaMemcpyAsync(ctx->h_poll[pb], pop->dev.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
            RUN_CU(cudaEventRecord(ctx->poll_ev[pb], st));
            if (have_prev) {
                RUN_CU(cudaEventSynchronize(ctx->poll_ev[pb ^ 1]));
                const Ctrl* pc = ctx->h_poll[pb ^ 1];
                if (pc->stop || pc->err || pc->acc.err) break;
            }
            have_prev = true; pb ^= 1;
        }
        if (host_iters > 10000000) break;
    }
    RUN_CU(cudaEventRecord(e_loop1, st));
    if (state_out) {
        // snapshot between two iterations (before the final extrema pass touches the accumulators)
        RUN_CU(cudaMemcpyAsync(c, pop->dev.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
        RUN_CU(cudaStreamSynchronize(st));
        const size_t n = (size_t)N; const int cur = c->cur;
        SmcStateHdr h; memset(&h, 0, sizeof h);
        h.magic = SMC_STATE_MAGIC; h.version = 2; h.prior_hash = prior_hash(pop->prior, pop->data); h.d = (uint32_t)pop->D; h.ds = (uint32_t)pop->DS; h.nb = (uint32_t)pop->NB;
        h.N = N; h.hist_cap = dev_hist; h.model_id = model->id; h.ctrl_bytes = sizeof(Ctrl);
        h.total_bytes = smc_state_size(N, pop->DS, pop->NB, dev_hist);
        strncpy(h.model, model->ops->name, sizeof(h.model) - 1);
        char* p = (char*)state_out;
        memcpy(p, &h, sizeof h); p += sizeof h;
        memset(p, 0, (sizeof(Ctrl) + 7) & ~(size_t)7); memcpy(p, c, sizeof(Ctrl)); p += (sizeof(Ctrl) + 7) & ~(size_t)7;
        RUN_CU(cudaMemcpyAsync(p, pop->dev.hist, (size_t)dev_hist * 64, cudaMemcpyDeviceToHost, st)); p += (size_t)dev_hist * 64;
        RUN_CU(cudaMemcpyAsync(p, pop->dev.theta[cur], n * pop->DS * 8, cudaMemcpyDeviceToHost, st)); p += n * pop->DS * 8;
        RUN_CU(cudaMemcpyAsync(p, pop->dev.logpi[cur], n * 8, cudaMemcpyDeviceToHost, st)); p += n * 8;
        RUN_CU(cudaMemcpyAsync(p, pop->dev.delta[cur], n * 8, cudaMemcpyDeviceToHost, st)); p += n * 8;
        if (pop->NB) RUN_CU(cudaMemcpyAsync(p, pop->dev.blob[cur], n * pop->NB * 8, cudaMemcpyDeviceToHost, st));
        p += n * pop->NB * 8;
        RUN_CU(cudaMemcpyAsync(p, pop->dev.W, n * 8, cudaMemcpyDeviceToHost, st)); p += n * 8;
        memset(p, 0, (n + 7) & ~(size_t)7);
        RUN_CU(cudaMemcpyAsync(p, pop->dev.alive, n, cudaMemcpyDeviceToHost, st));
        RUN_CU(cudaStreamSynchronize(st));
        *state_out_bytes = h.total_bytes;
    }
    launches += launch_minmax(st, pop->dev);      // extrema(delta) of the final generation -> last history record
    RUN_CU(cudaGetLastError());
    RUN_CU(cudaMemcpyAsync(c, pop->dev.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
    RUN_CU(cudaStreamSynchronize(st));
    t_loop = ms_since(t_begin);
    rc = check_dev_err(pop, "abcdez_smc_run");
    if (rc == ABCDEZ_ERR_NO_ALIVE) rc = ABCDEZ_OK;       // a warning in the reference (:375); reported in status
    if (rc) goto done;
    {
        float ms = 0.f;
        RUN_CU(cudaEventElapsedTime(&ms, e_init0, e_loop0)); res->init_ms = ms;
        RUN_CU(cudaEventElapsedTime(&ms, e_loop0, e_loop1)); res->total_ms = ms;
        double sw = 0.0, hd = 0.0, rs = 0.0;
        for (size_t i = 0; i + 3 < nev; i += 4) {
            RUN_CU(cudaEventElapsedTime(&ms, evs[i], evs[i + 1])); hd += ms;
            RUN_CU(cudaEventElapsedTime(&ms, evs[i + 1], evs[i + 2])); if (ms > 0.012f) rs += ms;   // (launches that return at the flag take ~2 us)
            RUN_CU(cudaEventElapsedTime(&ms, evs[i + 2], evs[i + 3])); sw += ms;
        }
        res->sweep_ms = sw; res->head_ms = hd; res->resample_ms = rs;
    }
    res->eps = c->eps; res->logZ = c->logZ; res->iters = c->iters; res->nsims = c->nsims_total;
    res->status = c->status; res->n_resamples = c->n_resamples; res->n_sweeps = c->n_sweeps; res->n_launches = launches;
    res->hist_len = hist_cap > 0 ? c->hist_len : 0;
    res->hist_dropped = hist_cap > 0 ? c->hist_overflow : 0;
    // results, :382-393
    if (res->P) {
        rc = ensure_scratch(pop, (size_t)N * pop->D * 8); if (rc) goto done;
        launch_push_rows(st, pop->dev, pop->prior, pop->D, (double*)pop->scratch);
        RUN_CU(cudaMemcpyAsync(res->P, pop->scratch, (size_t)N * pop->D * 8, cudaMemcpyDeviceToHost, st));
    }
    if (res->Wns) RUN_CU(cudaMemcpyAsync(res->Wns, pop->dev.W, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    if (res->C) RUN_CU(cudaMemcpyAsync(res->C, pop->dev.delta[c->cur], (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    if (res->blobs && pop->NB) RUN_CU(cudaMemcpyAsync(res->blobs, pop->dev.blob[c->cur], (size_t)N * pop->NB * 8, cudaMemcpyDeviceToHost, st));
    RUN_CU(cudaStreamSynchronize(st));
    if (res->hist_len > 0) {
        std::vector<double> h((size_t)res->hist_len * 8);
        RUN_CU(cudaMemcpy(h.data(), pop->dev.hist, h.size() * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < res->hist_len; ++i) {
            const double* r = &h[(size_t)i * 8];
            if (res->h_eps) res->h_eps[i] = r[0];
            if (res->h_dmin) res->h_dmin[i] = r[1];
            if (res->h_dmax) res->h_dmax[i] = r[2];
            if (res->h_logZ) res->h_logZ[i] = r[3];
            if (res->h_ess) res->h_ess[i] = r[4];
            if (res->h_facc) res->h_facc[i] = r[5];
            if (res->h_gamma0) res->h_gamma0[i] = r[6];
            if (res->h_Kmcmc) res->h_Kmcmc[i] = (int32_t)r[7];
        }
    }
done:
    t_result = ms_since(t_begin);
    if (e_init0) cudaEventDestroy(e_init0);
    if (e_loop0) cudaEventDestroy(e_loop0);
    if (e_loop1) cudaEventDestroy(e_loop1);
    abcdez_pop_destroy(pop);
    if (trace)
        fprintf(stderr, "[abcdez] smc_run N=%lld: create %.2f ms, loop done at %.2f, results at %.2f, destroyed at %.2f (device loop %.2f ms)\n",
                (long long)N, t_create, t_loop, t_result, ms_since(t_begin), res->total_ms);
    return rc;
#undef RUN_CU
}

// ---------------------------------------------------------------------------------------
// after the run: equally weighted posterior sample, batched replicate / multi-model runs
// ---------------------------------------------------------------------------------------
__global__ void strata_uniforms_kernel(const __grid_constant__ PhiloxKeys keys, int64_t N, double* __restrict__ u)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Stream rs(keys, (uint32_t)i, 0u, TAG_RESAMPLE);
    double a, b; rs.u2(0u, a, b);
    u[i] = a;
}
__global__ void gather_rows_kernel(int64_t N, int d, const double* __restrict__ P, const long long* __restrict__ inds, double* __restrict__ out)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * d) return;
    const int64_t i = e / d; const int k = (int)(e - i * d);
    long long src = inds[i] - 1;                          // wsample_stratified! indices are 1-based; "i = 0" (r == 0) clamps to the first row
    src = src < 0 ? 0 : (src >= N ? N - 1 : src);
    out[e] = P[src * d + k];
}

// P[weightinds(Wns)] of test/runtests.jl:13-19,287-291 on the device: stratified resampling of the weighted particles into an
// equally weighted sample (Philox uniforms of `seed`); inds_out (optional) receives the 1-based source indices
extern "C" int abcdez_posterior_sample(abcdez_ctx* ctx, int64_t N, int d, const double* P, const double* Wns, uint64_t seed,
                                       double* P_out, int64_t* inds_out)
{
    ctx = first_gpu(ctx);
    CHECK_ARG(ctx && P && Wns && P_out, "abcdez_posterior_sample: NULL argument");
    CHECK_ARG(N >= 1 && N < (int64_t)0x7fffffff && d >= 1 && d <= ABCDEZ_MAXD, "abcdez_posterior_sample: N or d out of range");
    {
        double sum = 0.0;
        for (int64_t i = 0; i < N; ++i) sum += Wns[i];
        CHECK_ARG(fabs(sum - 1.0) < 1e-8, "Sum of weights expected to be 1.0 (approximately)");       // test/runtests.jl:14
    }
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)N; const unsigned nt = (unsigned)((N + TILE - 1) / TILE);
    DevBuf dw, du, dc, dp, dt, di, dP, dO;
    CU(dw.alloc(n * 8)); CU(du.alloc(n * 8)); CU(dc.alloc(n * 8)); CU(dp.alloc((size_t)nt * 8 + 64)); CU(dt.alloc(2 * sizeof(SeqTab)));
    CU(di.alloc(n * 8)); CU(dP.alloc(n * d * 8)); CU(dO.alloc(n * d * 8));
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(dw.p, Wns, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dP.p, P, n * d * 8, cudaMemcpyHostToDevice, st));
    strata_uniforms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(philox_keys(seed), N, du.as<double>());
    launch_strat_indices(st, N, dw.as<double>(), du.as<double>(), dc.as<double>(), dp.as<double>(), dt.as<SeqTab>(), 2, di.as<long long>());
    gather_rows_kernel<<<(unsigned)((n * d + 255) / 256), 256, 0, st>>>(N, d, dP.as<double>(), di.as<long long>(), dO.as<double>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(P_out, dO.p, n * d * 8, cudaMemcpyDeviceToHost, st));
    if (inds_out) CU(cudaMemcpyAsync(inds_out, di.p, n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ABCDEZ_OK;
}

// Batched runs: `nruns` independent abcdesmc! runs (replicates of one model for the evidence uncertainty, docs/src/index.md:214-220;
// several models for a comparison, examples/minimal_example.jl:27-65) in flight together.  A 1000-particle run is a chain of
// ~800 launches of a few microseconds each and leaves the GPU almost empty; here up to 16 worker threads, each with its own
// stream and arena on the same GPU, pull runs from the list, so their kernels overlap on the device.  Every run gives exactly
// the result of its own abcdez_smc_run call.  status[i] (optional) receives the status of run i; the return value is the first failure.
extern "C" int abcdez_smc_run_batch(abcdez_ctx* ctx, int nruns, const abcdez_prior* const* priors, const abcdez_model* const* models,
                                    const double* eps_targets, const abcdez_smc_opts* opts, abcdez_smc_result* results, int* status)
{
    ctx = first_gpu(ctx);
    CHECK_ARG(ctx && priors && models && eps_targets && opts && results, "abcdez_smc_run_batch: NULL argument");
    CHECK_ARG(nruns >= 1 && nruns <= 65536, "abcdez_smc_run_batch: nruns out of range");
    CHECK_ARG(ctx->comm == nullptr, "abcdez_smc_run_batch: batched runs are single-GPU");
    const int W = nruns < 16 ? nruns : 16;
    while ((int)ctx->batch.size() < W) {                  // worker contexts live as long as the parent (arenas are reused)
        abcdez_ctx* w = nullptr;
        int rc = abcdez_init(ctx->device, nullptr, &w);
        if (rc) return rc;
        ctx->batch.push_back(w);
    }
    std::vector<int> rcs(nruns, 0); std::vector<std::string> whys(nruns);
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int w = 0; w < W; ++w)
        th.emplace_back([&, w] {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= nruns) break;
                rcs[i] = smc_run_impl(ctx->batch[w], priors[i], models[i], eps_targets[i], &opts[i], &results[i], nullptr, 0, nullptr, 0, nullptr);
                if (rcs[i]) whys[i] = g_err;
            }
        });
    for (std::thread& t : th) t.join();
    cudaSetDevice(ctx->device);
    int first = ABCDEZ_OK;
    for (int i = 0; i < nruns; ++i) {
        if (status) status[i] = rcs[i];
        if (rcs[i] && !first) { first = rcs[i]; g_err = "abcdez_smc_run_batch: run " + std::to_string(i) + ": " + whys[i]; }
    }
    return first;
}

// ---------------------------------------------------------------------------------------
// abcdemc!  (src/abcdez_mc.jl:102-172)
// ---------------------------------------------------------------------------------------
extern "C" void abcdez_mc_opts_default(abcdez_mc_opts* o)
{
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->nparticles = 50; o->generations = 20; o->seed = 1;
}

static int mc_run_impl(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                       const abcdez_mc_opts* o, abcdez_mc_result* res);

extern "C" int abcdez_mc_run(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                             const abcdez_mc_opts* o, abcdez_mc_result* res)
{
    CHECK_ARG(ctx && prior && model && o && res, "abcdez_mc_run: NULL argument");
    if (ctx->subs.empty()) return mc_run_impl(ctx, prior, model, eps_target, o, res);
    const int R = (int)ctx->subs.size();
    if (R == 1) return mc_run_impl(ctx->subs[0], prior, model, eps_target, o, res);
    const int d = prior->dev.d, B = model->ops->blob;
    std::vector<abcdez_mc_result> rr(R, *res);
    std::vector<int> rcs(R, 0); std::vector<std::string> whys(R);
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r) {
        int64_t lo = 0, hi = 0;
        abcdez_shard_range(o->nparticles, r, R, &lo, &hi);
        if (rr[r].P) rr[r].P += lo * d;
        if (rr[r].C) rr[r].C += lo;
        if (rr[r].blobs) rr[r].blobs += lo * B;
        th.emplace_back([&, r] {
            rcs[r] = mc_run_impl(ctx->subs[r], prior, model, eps_target, o, &rr[r]);
            if (rcs[r]) whys[r] = g_err;
        });
    }
    for (std::thread& t : th) t.join();
    cudaSetDevice(ctx->device);
    for (int r = 0; r < R; ++r) if (rcs[r]) return fail(rcs[r], whys[r]);
    res->reached_eps = rr[0].reached_eps; res->nsims = rr[0].nsims; res->dmin = rr[0].dmin; res->dmax = rr[0].dmax;
    res->sweep_ms = rr[0].sweep_ms; res->total_ms = rr[0].total_ms; res->n_launches = 0;
    for (int r = 0; r < R; ++r) res->n_launches += rr[r].n_launches;
    return ABCDEZ_OK;
}

static int mc_run_impl(abcdez_ctx* ctx, const abcdez_prior* prior, const abcdez_model* model, double eps_target,
                       const abcdez_mc_opts* o, abcdez_mc_result* res)
{
    CHECK_ARG(ctx && prior && model && o && res, "abcdez_mc_run: NULL argument");
    CHECK_ARG(0.0 <= eps_target, "\xcf\xb5_target must be non-negative");      // src/abcdez_mc.jl:108
    CHECK_ARG(5 <= o->nparticles, "nparticles must be at least 5");          // :109
    CHECK_ARG(1 <= o->generations, "generations must be at least 1");        // :110
    if (!model->ops->mc_sweep) return fail(ABCDEZ_ERR_UNSUPPORTED, std::string("abcdemc!: model '") + model->ops->name + "' has no abcdemc! sweep in this build");
    CHECK_ARG(o->generations <= 1000000, "generations must be at most 10^6");
    CU(cudaSetDevice(ctx->device));
    // sharded run: nparticles is the whole population; base particle and partners are drawn inside the rank's
    // block, extrema(delta) (:146) and the simulation count are global (in-kernel exchange after every sweep)
    const bool shard = ctx->comm != nullptr && ctx->world > 1;
    const int64_t Ng = o->nparticles;
    int64_t lo = 0, hi = Ng;
    if (shard) {
        abcdez_shard_range(Ng, ctx->rank, ctx->world, &lo, &hi);
        CHECK_ARG(Ng >= 5 * (int64_t)ctx->world, "sharded abcdemc!: need at least 5 particles per rank");
        CHECK_ARG(Ng < (int64_t)0x7fffffff, "sharded abcdemc!: nparticles must be below 2^31");
    }
    const int64_t N = hi - lo;
    abcdez_pop* pop = nullptr;
    int rc = pop_create_impl(ctx, prior, model, N, lo, 1, &pop, shard ? Ng : 0);
    if (rc) return rc;
    if (shard) {
        rc = comm_begin_run(ctx->comm, ctx->stream, pop->dev, &pop->dev.x, &pop->dev.peers);
        if (rc) { abcdez_pop_destroy(pop); return fail(rc, std::string("sharded abcdemc!: ") + comm_error(ctx->comm)); }
    }
    cudaStream_t st = ctx->stream;
    Ctrl* c = pop->h_ctrl;
    c->seed = o->seed; c->eps_target = eps_target;
    pop->dev.keys = philox_keys(o->seed);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int64_t launches = 0;
    SweepInj noinj; memset(&noinj, 0, sizeof(noinj));
#define RUN_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(ABCDEZ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); goto done; } } while (0)
    rc = push_ctrl(pop); if (rc) goto done;
    rc = mc_prepare(pop); if (rc) goto done;
    RUN_CU(cudaEventCreate(&e0)); RUN_CU(cudaEventCreate(&e1));
    pop->ops->init(*pop->ops, st, pop->dev, pop->prior, pop->data, o->seed, 1);         // :117-125
    launches += pop->ops->split == 2 ? 3 : 1;
    RUN_CU(cudaEventRecord(e0, st));
    {
        // The generation loop (:134-161) is a fixed launch list: extrema(delta) of the live generation come from the
        // previous kernel's epilogue and stay in the device control block, eps_pop (:146-147) is derived from them
        // inside the sweep, the sort kernels return at once when no particle is above eps_target (:19-24), and every
        // kernel behind an error returns at once -- no host round trip per generation.
        McArgs mc{ 0.0, eps_target, 1, pop->sorted_delta, pop->order };
        for (int it = 0; it < o->generations; ++it) {                        // :134
            launches += launch_mc_prepare(st, pop->dev, pop->sorted_delta, pop->order, pop->sort_tmp, pop->sort_tmp_bytes, 0);
            pop->ops->mc_sweep(*pop->ops, st, pop->dev, pop->prior, pop->data, noinj, mc);   // :149
            launches += pop->ops->split ? 3 : 1;                                   // (heavy simulators: propose, simulate, accept)
        }
    }
    RUN_CU(cudaEventRecord(e1, st));
    RUN_CU(cudaGetLastError());
    RUN_CU(cudaMemcpyAsync(c, pop->dev.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
    RUN_CU(cudaStreamSynchronize(st));
    rc = check_dev_err(pop, "abcdez_mc_run"); if (rc) goto done;
    { float ms = 0.f; RUN_CU(cudaEventElapsedTime(&ms, e0, e1)); res->total_ms = ms; res->sweep_ms = ms; }
    res->reached_eps = (c->dmax <= eps_target) ? 1 : 0;                      // :163
    res->nsims = c->nsims_total; res->dmin = c->dmin; res->dmax = c->dmax; res->n_launches = launches;
    if (res->P) {
        rc = ensure_scratch(pop, (size_t)N * pop->D * 8); if (rc) goto done;
        launch_push_rows(st, pop->dev, pop->prior, pop->D, (double*)pop->scratch);   // :166
        RUN_CU(cudaMemcpyAsync(res->P, pop->scratch, (size_t)N * pop->D * 8, cudaMemcpyDeviceToHost, st));
    }
    if (res->C) RUN_CU(cudaMemcpyAsync(res->C, pop->dev.delta[c->cur], (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    if (res->blobs && pop->NB) RUN_CU(cudaMemcpyAsync(res->blobs, pop->dev.blob[c->cur], (size_t)N * pop->NB * 8, cudaMemcpyDeviceToHost, st));
    RUN_CU(cudaStreamSynchronize(st));
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    abcdez_pop_destroy(pop);
    return rc;
#undef RUN_CU
}
