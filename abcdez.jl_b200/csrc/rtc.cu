// rtc.cu -- runtime-supplied models: the closest analogue of handing abcdesmc! an arbitrary Julia `dist!`
// (src/abcdez_smc.jl:215).  The caller passes the CUDA source of ONE model struct (the same shape as the
// structs of models.cuh: D, BLOB, NOISE, name, run(theta, data, rng, blob)); it is compiled with NVRTC against
// the library's own kernel templates (sweep.cuh and friends, embedded as strings by build.py), loaded as a
// module and registered next to the static models, so init / abcdesmc_swarm! / abcdemc_swarm! / simulate for
// it are the SAME kernels, instantiated at run time.  libnvrtc and libcuda are dlopen-ed: the library still
// loads (and compiles models, e.g. for validation) on a machine without a GPU.
#include "internal.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace abcdez {

extern const char* const rtc_header_names[];
extern const char* const rtc_header_texts[];
extern const int rtc_header_count;

// ---- NVRTC / driver entry points, resolved at run time ------------------------------------------------
typedef void* nvrtcProgram;
struct RtcApi {
    void* nv = nullptr; void* cu = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    int (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*cuModuleLoadData)(void**, const void*) = nullptr;
    int (*cuModuleGetFunction)(void**, void*, const char*) = nullptr;
    int (*cuLaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    int (*cuGetErrorString)(int, const char**) = nullptr;
};
static RtcApi g_rtc;

static bool load_syms(void* h, std::string* why, std::initializer_list<std::pair<void**, const char*>> syms)
{
    for (auto& s : syms) {
        *s.first = dlsym(h, s.second);
        if (!*s.first) { *why = std::string("missing symbol ") + s.second; return false; }
    }
    return true;
}

static bool nvrtc_load(std::string* why)
{
    if (g_rtc.nv) return true;
    const char* cands[] = { getenv("ABCDEZ_NVRTC_LIB"), "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.13" };
    void* h = nullptr;
    for (const char* c : cands) if (c && (h = dlopen(c, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) { *why = "libnvrtc not found (set ABCDEZ_NVRTC_LIB)"; return false; }
    if (!load_syms(h, why, { { (void**)&g_rtc.CreateProgram, "nvrtcCreateProgram" }, { (void**)&g_rtc.DestroyProgram, "nvrtcDestroyProgram" },
                             { (void**)&g_rtc.CompileProgram, "nvrtcCompileProgram" }, { (void**)&g_rtc.GetProgramLogSize, "nvrtcGetProgramLogSize" },
                             { (void**)&g_rtc.GetProgramLog, "nvrtcGetProgramLog" }, { (void**)&g_rtc.GetCUBINSize, "nvrtcGetCUBINSize" },
                             { (void**)&g_rtc.GetCUBIN, "nvrtcGetCUBIN" }, { (void**)&g_rtc.AddNameExpression, "nvrtcAddNameExpression" },
                             { (void**)&g_rtc.GetLoweredName, "nvrtcGetLoweredName" }, { (void**)&g_rtc.GetErrorString, "nvrtcGetErrorString" } }))
        return false;
    g_rtc.nv = h;
    return true;
}

static bool driver_load(std::string* why)
{
    if (g_rtc.cu) return true;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) { *why = "libcuda.so.1 not found (no NVIDIA driver)"; return false; }
    if (!load_syms(h, why, { { (void**)&g_rtc.cuModuleLoadData, "cuModuleLoadData" }, { (void**)&g_rtc.cuModuleGetFunction, "cuModuleGetFunction" },
                             { (void**)&g_rtc.cuLaunchKernel, "cuLaunchKernel" }, { (void**)&g_rtc.cuGetErrorString, "cuGetErrorString" } }))
        return false;
    g_rtc.cu = h;
    return true;
}

// ---- a loaded runtime model ----------------------------------------------------------------------------------
enum { K_INIT = 0, K_SMC_GEN, K_SMC_DISC, K_SMC_NORMAL, K_SMC_UNIFORM, K_MC, K_MC_DISC, K_SIM, K_COUNT };
struct DynModel {
    std::string name;
    void* module = nullptr;
    void* fn[K_COUNT] = {};
    ModelOps ops;
};
static std::vector<DynModel*> g_dyn;
static std::mutex g_dyn_mu;

static inline unsigned grid_of(int64_t N, int threads) { return (unsigned)((N + threads - 1) / threads); }

static void dyn_launch(void* f, unsigned grid, cudaStream_t st, void** params)
{
    g_rtc.cuLaunchKernel(f, grid, 1, 1, (unsigned)SWEEP_THREADS, 1, 1, 0, (void*)st, params, nullptr);
}

static bool has_discrete(const PriorDev& pr)
{
    for (int k = 0; k < pr.d; ++k) if (fam_is_discrete(pr.family[k])) return true;
    return false;
}

static void dyn_init(const ModelOps& o, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, uint64_t seed, int dp)
{
    DynModel* m = (DynModel*)o.dyn;
    const PhiloxKeys keys = philox_keys(seed);
    void* params[] = { (void*)&P, (void*)&pr, (void*)&md, (void*)&keys, (void*)&dp };
    dyn_launch(m->fn[K_INIT], grid_of(P.N, SWEEP_THREADS), st, params);
}
static void dyn_smc(const ModelOps& o, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj)
{
    DynModel* m = (DynModel*)o.dyn;
    bool all_normal = true, all_uniform = true;
    for (int k = 0; k < pr.d; ++k) { all_normal = all_normal && pr.family[k] == ABCDEZ_NORMAL; all_uniform = all_uniform && pr.family[k] == ABCDEZ_UNIFORM; }
    int k = has_discrete(pr) ? K_SMC_DISC : all_normal ? K_SMC_NORMAL : all_uniform ? K_SMC_UNIFORM : K_SMC_GEN;
    void* params[] = { (void*)&P, (void*)&pr, (void*)&md, (void*)&inj };
    dyn_launch(m->fn[k], grid_of(P.N, SWEEP_THREADS), st, params);
}
static void dyn_mc(const ModelOps& o, cudaStream_t st, const PopDev& P, const PriorDev& pr, const ModelData& md, const SweepInj& inj,
                   const McArgs& mc)
{
    DynModel* m = (DynModel*)o.dyn;
    void* params[] = { (void*)&P, (void*)&pr, (void*)&md, (void*)&inj, (void*)&mc };
    dyn_launch(m->fn[has_discrete(pr) ? K_MC_DISC : K_MC], grid_of(P.N, SWEEP_THREADS), st, params);
}
static void dyn_sim(const ModelOps& o, cudaStream_t st, const PriorDev*, const ModelData& md, int64_t N, const double* th, uint64_t seed,
                    uint32_t epoch, uint32_t tag, uint32_t id0, double* dist, double* blobs)
{
    DynModel* m = (DynModel*)o.dyn;
    const PhiloxKeys keys = philox_keys(seed);
    void* params[] = { (void*)&md, (void*)&N, (void*)&th, (void*)&keys, (void*)&epoch, (void*)&tag, (void*)&id0, (void*)&dist, (void*)&blobs };
    dyn_launch(m->fn[K_SIM], grid_of(N, SWEEP_THREADS), st, params);
}

const ModelOps* rtc_model_ops(int id)
{
    std::lock_guard<std::mutex> lk(g_dyn_mu);
    int k = id - M_COUNT;
    return (k >= 0 && k < (int)g_dyn.size()) ? &g_dyn[k]->ops : nullptr;
}
int rtc_model_count() { std::lock_guard<std::mutex> lk(g_dyn_mu); return (int)g_dyn.size(); }

// load != 0: the cubin is loaded into the current CUDA context and registered (*id = new model id);
// load == 0: compile only (no GPU needed), *id = -1.
int rtc_compile_model(const char* name, const char* struct_name, const char* cuda_src, int d, int blob_bytes, int load,
                      int* id, std::string* log)
{
    std::string why;
    if (!nvrtc_load(&why)) { *log = why; return ABCDEZ_ERR_UNSUPPORTED; }
    // the translation unit: the library's kernel templates + the caller's struct + explicit instantiations
    std::string M = std::string("abcdez::") + struct_name;
    const std::string kexpr[K_COUNT] = {
        "abcdez::init_kernel<" + M + ">",
        "abcdez::smc_sweep_kernel<" + M + ", false, abcdez::PK_GENERIC>", "abcdez::smc_sweep_kernel<" + M + ", true, abcdez::PK_GENERIC>",
        "abcdez::smc_sweep_kernel<" + M + ", false, abcdez::PK_NORMAL>", "abcdez::smc_sweep_kernel<" + M + ", false, abcdez::PK_UNIFORM>",
        "abcdez::mc_sweep_kernel<" + M + ", false>", "abcdez::mc_sweep_kernel<" + M + ", true>",
        "abcdez::simulate_kernel<" + M + ">" };
    std::string tu = "#include \"sweep.cuh\"\nnamespace abcdez {\n#line 1 \"model.cu\"\n";
    tu += cuda_src;
    tu += "\n#line 1 \"abcdez_instantiate.cu\"\n";
    char buf[512];
    snprintf(buf, sizeof buf, "static_assert(%s::D == %d, \"the model's D differs from the d passed to abcdez_model_compile\");\n"
                              "static_assert(%s::BLOB == %d && %s::BLOB %% 8 == 0 && %s::BLOB <= ABCDEZ_MAXBLOB, \"the model's BLOB differs from blob_bytes (or is not a multiple of 8 <= 64)\");\n"
                              "static_assert(%s::D >= 1 && %s::D <= ABCDEZ_MAXD, \"D out of range\");\n",
             struct_name, d, struct_name, blob_bytes, struct_name, struct_name, struct_name, struct_name);
    tu += buf;
    tu += "}  // namespace abcdez\n";
    nvrtcProgram prog = nullptr;
    int rc = g_rtc.CreateProgram(&prog, tu.c_str(), "abcdez_user_model.cu", rtc_header_count, rtc_header_texts, rtc_header_names);
    if (rc) { *log = std::string("nvrtcCreateProgram: ") + g_rtc.GetErrorString(rc); return ABCDEZ_ERR_CUDA; }
    for (int k = 0; k < K_COUNT; ++k) g_rtc.AddNameExpression(prog, kexpr[k].c_str());
    char tdef[64], bdef[64];
    snprintf(tdef, sizeof tdef, "-DABCDEZ_SWEEP_THREADS=%d", SWEEP_THREADS);
    snprintf(bdef, sizeof bdef, "-DABCDEZ_SWEEP_MIN_BLOCKS=%d", SWEEP_MIN_BLOCKS);
    const char* opts[] = { "-arch=sm_100a", "-std=c++17", "-fmad=false", "-default-device", "-lineinfo", tdef, bdef };
    rc = g_rtc.CompileProgram(prog, (int)(sizeof opts / sizeof opts[0]), opts);
    {
        size_t n = 0;
        g_rtc.GetProgramLogSize(prog, &n);
        if (n > 1) { std::string l(n, '\0'); g_rtc.GetProgramLog(prog, &l[0]); l.resize(strlen(l.c_str())); *log = l; }
    }
    if (rc) { g_rtc.DestroyProgram(&prog); *log = std::string("NVRTC: ") + g_rtc.GetErrorString(rc) + "\n" + *log; return ABCDEZ_ERR_BAD_ARG; }
    size_t nbin = 0;
    g_rtc.GetCUBINSize(prog, &nbin);
    std::vector<char> cubin(nbin);
    g_rtc.GetCUBIN(prog, cubin.data());
    std::string lowered[K_COUNT];
    for (int k = 0; k < K_COUNT; ++k) {
        const char* ln = nullptr;
        rc = g_rtc.GetLoweredName(prog, kexpr[k].c_str(), &ln);
        if (rc || !ln) { g_rtc.DestroyProgram(&prog); *log = "nvrtcGetLoweredName failed for " + kexpr[k]; return ABCDEZ_ERR_CUDA; }
        lowered[k] = ln;
    }
    g_rtc.DestroyProgram(&prog);
    *id = -1;
    if (!load) return ABCDEZ_OK;
    if (!driver_load(&why)) { *log = why; return ABCDEZ_ERR_CUDA; }
    DynModel* m = new DynModel();
    m->name = name;
    int cr = g_rtc.cuModuleLoadData(&m->module, cubin.data());
    if (cr) {
        const char* es = nullptr; g_rtc.cuGetErrorString(cr, &es);
        *log = std::string("cuModuleLoadData: ") + (es ? es : "error"); delete m; return ABCDEZ_ERR_CUDA;
    }
    for (int k = 0; k < K_COUNT; ++k) {
        cr = g_rtc.cuModuleGetFunction(&m->fn[k], m->module, lowered[k].c_str());
        if (cr) { *log = "cuModuleGetFunction failed for " + kexpr[k]; delete m; return ABCDEZ_ERR_CUDA; }
    }
    m->ops.name = m->name.c_str(); m->ops.d = d; m->ops.blob = blob_bytes;
    m->ops.init = &dyn_init; m->ops.smc_sweep = &dyn_smc; m->ops.mc_sweep = &dyn_mc; m->ops.simulate = &dyn_sim; m->ops.dyn = m; m->ops.f32_state = 0; m->ops.split = 0;
    std::lock_guard<std::mutex> lk(g_dyn_mu);
    g_dyn.push_back(m);
    *id = M_COUNT + (int)g_dyn.size() - 1;
    return ABCDEZ_OK;
}

}  // namespace abcdez
