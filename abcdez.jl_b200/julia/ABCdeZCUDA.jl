# ABCdeZCUDA.jl -- thin `ccall` shim that keeps the ABCdeZ.jl user interface
# (`abcdesmc!`, `abcdemc!`, `Factored`, the ABC kernels; reference src/ABCdeZ.jl:1-20) and runs the
# particle hot path in libabcdez_cuda.so (include/abcdez_cuda.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia`.  It is a mechanical mirror of
# the ctypes binding in ../host.py (same symbols, same struct layouts, same defaults and messages);
# every executable check goes through that binding.  See INTEGRATION.md.
#
# Usage (drop-in for the reference, except that `dist!` is a registered device functor):
#
#   using ABCdeZCUDA, Distributions
#   prior = Normal(0, sqrt(10))
#   dist! = DeviceModel("gauss1d", [3.0, 1.0])        # instead of a Julia closure
#   r = abcdesmc!(prior, dist!, 0.3, nothing, nparticles=1000)
#   evidence = exp(r.logZ)
module ABCdeZCUDA

using Distributions
using Random

export Factored, abcdesmc!, abcdemc!, DeviceModel, compile_model, Context, MultiContext, nccl_unique_id, comm_init!, shard_range
export Indicator0toϵ, IndicatorStrict0toϵ, Epa0toϵ, EpaStrict0toϵ
export weightinds, posterior_sample, evidence_at, model_probabilities, evidence_uncertainty, abcdesmc_batch

const LIB = get(ENV, "ABCDEZ_LIB", joinpath(@__DIR__, "..", "libabcdez_cuda.so"))

# ---- status / errors (include/abcdez_cuda.h: every export returns an int status) -----------------
const ABCDEZ_OK = Cint(0)
const ABCDEZ_ERR_NO_ALIVE = Cint(4)

last_error() = unsafe_string(ccall((:abcdez_last_error, LIB), Cstring, ()))
check(rc::Integer) = rc == ABCDEZ_OK ? nothing : error(last_error())     # reference: error("...") -> ErrorException

# ---- Factored, reference src/abcdez_priors.jl:18-61 -----------------------------------------------
struct Factored{N} <: Distribution{Multivariate, Continuous}
    p::NTuple{N, UnivariateDistribution}
    Factored(args::UnivariateDistribution...) = new{length(args)}(args)
end
Base.length(::Factored{N}) where {N} = N
# the Distributions interface the reference extends (src/abcdez_priors.jl:27-54; exercised by test/runtests.jl:21-36),
# evaluated by the library's prior kernels (abcdez_prior_logpdf / abcdez_prior_sample) -- defined below, after Context

# (family, params) of a marginal; families of include/abcdez_cuda.h
marginal(d::Normal) = (Cint(0), (d.μ, d.σ, 0.0, 0.0))
marginal(d::Uniform) = (Cint(1), (d.a, d.b, 0.0, 0.0))
marginal(d::DiscreteUniform) = (Cint(2), (Float64(d.a), Float64(d.b), 0.0, 0.0))
marginal(d::LogNormal) = (Cint(3), (d.μ, d.σ, 0.0, 0.0))
marginal(d::Exponential) = (Cint(4), (d.θ, 0.0, 0.0, 0.0))
marginal(d::Gamma) = (Cint(5), (d.α, d.θ, 0.0, 0.0))
marginal(d::Beta) = (Cint(6), (d.α, d.β, 0.0, 0.0))
marginal(d::NegativeBinomial) = (Cint(7), (d.r, d.p, 0.0, 0.0))
marginal(d) = error("ABCdeZCUDA: unsupported prior marginal $(typeof(d))")

marginals(p::Factored) = collect(p.p)
marginals(p::UnivariateDistribution) = [p]          # bare univariate prior: d = 1, scalar θ (examples/minimal_example.jl:17)

# ---- ABC kernels, reference src/abcdez_types.jl:26-73 (type objects select the kernel enum) -------
for (name, kind, strict, epa) in ((:Indicator0toϵ, 0, false, false), (:IndicatorStrict0toϵ, 1, true, false),
                                  (:Epa0toϵ, 2, false, true), (:EpaStrict0toϵ, 3, true, true))
    @eval begin
        struct $name <: ContinuousUnivariateDistribution
            ϵ::Float64
            function $name(ϵ)
                ϵ ≥ 0.0 || error("Expected ϵ ≥ 0.0")
                new(ϵ)
            end
        end
        kernel_kind(::Type{$name}) = Cint($kind)
        Distributions.pdf(d::$name, x::Real) = ccall((:abcdez_kernel_pdf, LIB), Cdouble, (Cint, Cdouble, Cdouble), $kind, d.ϵ, x)
        Distributions.logpdf(d::$name, x::Real) = ccall((:abcdez_kernel_logpdf, LIB), Cdouble, (Cint, Cdouble, Cdouble), $kind, d.ϵ, x)
    end
end

# ---- handles ---------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    rank::Int
    world::Int
    Context(h::Ptr{Cvoid}, rank::Int, world::Int) = new(h, rank, world)
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:abcdez_init, LIB), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, r))
        c = new(r[], 0, 1)
        finalizer(x -> ccall((:abcdez_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
        c
    end
end
# ---- sharded runs: one Julia process per GPU (include/abcdez_cuda.h "sharded runs") --------------------
# Rank 0 calls nccl_unique_id() and ships the 128 bytes to the other ranks (Distributed.jl, MPI.jl, a file ...);
# every rank then calls comm_init!(ctx, rank, world, id).  After that abcdesmc!/abcdemc! on `ctx` are collective:
# `nparticles` is the whole population, each rank gets the rows of its block shard_range(nparticles, rank, world).
function nccl_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:abcdez_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    id
end
function comm_init!(ctx::Context, rank::Integer, world::Integer, id::Vector{UInt8})
    check(ccall((:abcdez_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, rank, world, id))
    ctx.rank = rank; ctx.world = world
    ctx
end
function shard_range(N::Integer, rank::Integer, world::Integer)     # 0-based [lo, hi)
    lo = Ref{Int64}(0); hi = Ref{Int64}(0)
    check(ccall((:abcdez_shard_range, LIB), Cint, (Int64, Cint, Cint, Ref{Int64}, Ref{Int64}), N, rank, world, lo, hi))
    lo[], hi[]
end
local_count(ctx::Context, N::Integer) = ctx.world == 1 ? N : ((lo, hi) = shard_range(N, ctx.rank, ctx.world); hi - lo)

const DEFAULT_CTX = Ref{Union{Nothing, Context}}(nothing)
default_context() = (DEFAULT_CTX[] === nothing && (DEFAULT_CTX[] = Context(parse(Int, get(ENV, "LOCAL_RANK", "0")))); DEFAULT_CTX[])

# ---- single-process multi-GPU context (abcdez_init_multi): `parallel=true` == every GPU of this process ----------
# abcdesmc!/abcdemc! on it run ONE population sharded over the GPUs and return the whole population.
function MultiContext(n_gpus::Integer=0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:abcdez_init_multi, LIB), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), n_gpus, C_NULL, r))
    c = Context(r[], 0, 1)
    finalizer(x -> ccall((:abcdez_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
    c
end
const MULTI_CTX = Ref{Union{Nothing, Context}}(nothing)
multi_context() = (MULTI_CTX[] === nothing && (MULTI_CTX[] = MultiContext(0)); MULTI_CTX[])
run_context(parallel::Bool) = parallel ? multi_context() : default_context()      # src/abcdez_smc.jl:237: the executor choice

# ---- Factored as a Distribution (src/abcdez_priors.jl:27-54) ----------------------------------------------------------
function Distributions.logpdf(d::Factored{N}, x) where {N}
    ctx = default_context(); ph = prior_handle(ctx, d)
    th = Float64[float(x[k]) for k in 1:N]; out = Ref{Cdouble}(0.0)
    check(ccall((:abcdez_prior_logpdf, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ref{Cdouble}), ctx.h, ph, 1, th, out))
    ccall((:abcdez_prior_destroy, LIB), Cint, (Ptr{Cvoid},), ph)
    out[]
end
Distributions.pdf(d::Factored, x) = exp(logpdf(d, x))
function Base.rand(rng::AbstractRNG, d::Factored{N}) where {N}
    ctx = default_context(); ph = prior_handle(ctx, d)
    th = zeros(Float64, N)
    check(ccall((:abcdez_prior_sample, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, UInt64, UInt32, Int64, Ptr{Cdouble}),
                ctx.h, ph, 1, rand(rng, UInt64), 0, 0, th))
    ccall((:abcdez_prior_destroy, LIB), Cint, (Ptr{Cvoid},), ph)
    ntuple(k -> d.p[k] isa DiscreteDistribution ? round(Int, th[k]) : th[k], N)      # tuple of marginal draws, :53-54
end
Base.rand(d::Factored) = rand(Random.default_rng(), d)

"`dist!` as a registered device functor bound to observed data (the data a Julia closure would capture)."
struct DeviceModel
    name::String
    data::Vector{Float64}
end

"""
    compile_model(name, struct_name, cuda_src, d; blob_bytes=0, ctx=default_context())

Runtime-supplied `dist!`: the CUDA source of one model struct (see include/abcdez_cuda.h, abcdez_model_compile) is
compiled with NVRTC against the library's kernel templates and registered; `DeviceModel(name, data)` then works
like a built-in model.  Throws with the NVRTC log on a compile error.
"""
function compile_model(name::String, struct_name::String, cuda_src::String, d::Integer; blob_bytes::Integer=0,
                       ctx::Context=default_context())
    id = Ref{Cint}(-1); log = zeros(UInt8, 1 << 16)
    check(ccall((:abcdez_model_compile, LIB), Cint,
                (Ptr{Cvoid}, Cstring, Cstring, Cstring, Cint, Cint, Ref{Cint}, Ptr{UInt8}, Csize_t),
                ctx.h, name, struct_name, cuda_src, d, blob_bytes, id, log, length(log)))
    Int(id[])
end

function prior_handle(ctx::Context, prior)
    ms = marginals(prior)
    fam = Cint[marginal(m)[1] for m in ms]
    par = zeros(Float64, 4, length(ms))                 # column-major 4 x d == C's d x 4 rows
    for (k, m) in enumerate(ms); par[:, k] .= marginal(m)[2]; end
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:abcdez_prior_create, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ref{Ptr{Cvoid}}),
                ctx.h, length(ms), fam, par, r))
    r[]
end

function model_handle(ctx::Context, m::DeviceModel)
    id = Ref{Cint}(0)
    check(ccall((:abcdez_model_lookup, LIB), Cint, (Cstring, Ref{Cint}), m.name, id))
    d = Ref{Cint}(0); b = Ref{Cint}(0)
    check(ccall((:abcdez_model_info, LIB), Cint, (Cint, Ref{Cint}, Ref{Cint}), id[], d, b))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:abcdez_model_bind, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Csize_t, Ref{Ptr{Cvoid}}),
                ctx.h, id[], m.data, length(m.data), r))
    r[], Int(d[]), Int(b[])
end

# ---- C structs (field order == include/abcdez_cuda.h) ----------------------------------------------
mutable struct SmcOpts
    nparticles::Int64; alpha::Cdouble; delta_ess::Cdouble; nsims_max::Int64; Kmcmc::Int32
    Kmcmc_min::Cdouble; kernel::Int32; facc_stop::Cdouble; facc_min::Cdouble; facc_tune::Cdouble
    seed::UInt64; verboseout::Int32; max_iters::Int32; exact_scan::Int32; profile::Int32; sync_every::Int32
    fused_head::Int32; systematic_resampling::Int32; partner_segments::Int32
    fp32_state::Int32; reserved1::Int32
    SmcOpts() = new()
end

mutable struct SmcResult
    P::Ptr{Cdouble}; Wns::Ptr{Cdouble}; C::Ptr{Cdouble}; blobs::Ptr{UInt8}
    hist_cap::Int32
    h_eps::Ptr{Cdouble}; h_dmin::Ptr{Cdouble}; h_dmax::Ptr{Cdouble}; h_logZ::Ptr{Cdouble}; h_ess::Ptr{Cdouble}
    h_facc::Ptr{Cdouble}; h_gamma0::Ptr{Cdouble}; h_Kmcmc::Ptr{Int32}
    eps::Cdouble; logZ::Cdouble; iters::Int64; nsims::Int64; hist_len::Int32; status::Int32
    n_resamples::Int64; n_sweeps::Int64; n_launches::Int64; sweep_ms::Cdouble; total_ms::Cdouble; init_ms::Cdouble
    hist_dropped::Int32; reserved0::Int32; head_ms::Cdouble; resample_ms::Cdouble
    SmcResult() = new()
end

mutable struct McOpts
    nparticles::Int64; generations::Int32; seed::UInt64
end

mutable struct McResult
    P::Ptr{Cdouble}; C::Ptr{Cdouble}; blobs::Ptr{UInt8}; reached_eps::Int32; nsims::Int64
    dmin::Cdouble; dmax::Cdouble; sweep_ms::Cdouble; total_ms::Cdouble; n_launches::Int64
    McResult() = new()
end

seed_from(rng::AbstractRNG) = rand(rng, UInt64)       # `rng` -> the 64-bit Philox key
seed_from(s::Integer) = UInt64(s)

# P as the reference returns it (src/abcdez_smc.jl:382): a Vector of tuples (Factored) or scalars (univariate),
# with discrete coordinates as Int (push_p, src/abcdez_types.jl:20-23)
function particles(prior, P::Matrix{Float64})
    ms = marginals(prior)
    conv(m, x) = m isa DiscreteDistribution ? round(Int, x) : x
    prior isa Factored ? [ntuple(k -> conv(ms[k], P[k, i]), length(ms)) for i in axes(P, 2)] :
                         [conv(ms[1], P[1, i]) for i in axes(P, 2)]
end

blobs_out(raw::Matrix{UInt8}, B::Int) = B == 0 ? fill(nothing, size(raw, 2)) : [raw[1:B, i] for i in axes(raw, 2)]

"""
    abcdesmc!(prior, dist!, ϵ_target, varexternal; kwargs...)

Same positional arguments, keyword names and defaults as the reference (src/abcdez_smc.jl:215-220).
`dist!` is a `DeviceModel`; `varexternal` is accepted and ignored (device functors keep their scratch
in registers); `parallel=true` runs the population sharded over every GPU of the process (`MultiContext`), the
counterpart of the reference's `ThreadedEx()` (src/abcdez_smc.jl:237).  Not in the reference: `max_iters`, `state`,
`return_state` (run-state snapshots) and the relaxed-parity modes `systematic_resampling`, `partner_segments`, `fp32_state`.
"""
function abcdesmc!(prior, dist!::DeviceModel, ϵ_target, varexternal;
                   nparticles::Int=100, α=0.95, δess=0.5, nsims_max::Int=10^7, Kmcmc::Int=3, Kmcmc_min=1.0,
                   ABCk=IndicatorStrict0toϵ, facc_stop=0.0, facc_min=0.0, facc_tune=0.975,
                   verbose::Bool=true, verboseout::Bool=true, rng=Random.default_rng(), parallel::Bool=false,
                   ctx::Context=run_context(parallel), hist_cap::Int=8192,
                   max_iters::Int=0, state::Union{Nothing,Vector{UInt8}}=nothing, return_state::Bool=false,
                   systematic_resampling::Bool=false, partner_segments::Bool=false, fp32_state::Bool=false)
    # max_iters / state / return_state are not in the reference: run-state snapshots (abcdez_smc_run_state); the
    # returned NamedTuple gains a `state` field when return_state=true
    Kmcmc_min > facc_min || @warn("Kmcmc_min should be larger than facc_min")         # src/abcdez_smc.jl:232
    ph = prior_handle(ctx, prior)
    mh, d, B = model_handle(ctx, dist!)
    o = SmcOpts()
    ccall((:abcdez_smc_opts_default, LIB), Cvoid, (Ref{SmcOpts},), o)
    o.nparticles = nparticles; o.alpha = α; o.delta_ess = δess; o.nsims_max = nsims_max; o.Kmcmc = Kmcmc
    o.Kmcmc_min = Kmcmc_min; o.kernel = kernel_kind(ABCk); o.facc_stop = facc_stop; o.facc_min = facc_min
    o.facc_tune = facc_tune; o.seed = seed_from(rng); o.verboseout = verboseout || verbose; o.max_iters = max_iters
    o.systematic_resampling = systematic_resampling; o.partner_segments = partner_segments; o.fp32_state = fp32_state
    verbose && @info("Preparing abcde in smc mode", nparticles, α, δess, parallel)       # src/abcdez_smc.jl:238-239
    N = local_count(ctx, nparticles)                   # sharded: the rows of this rank's block
    P = Matrix{Float64}(undef, d, N); Wns = Vector{Float64}(undef, N); C = Vector{Float64}(undef, N)
    bl = zeros(UInt8, max(B, 1), N)
    h = [zeros(Float64, hist_cap) for _ in 1:7]; hK = zeros(Int32, hist_cap)
    r = SmcResult()
    state_out = nothing
    GC.@preserve P Wns C bl h hK begin
        r.P = pointer(P); r.Wns = pointer(Wns); r.C = pointer(C); r.blobs = pointer(bl)
        r.hist_cap = (verboseout || verbose) ? hist_cap : 0
        r.h_eps, r.h_dmin, r.h_dmax, r.h_logZ, r.h_ess, r.h_facc, r.h_gamma0 = pointer.(h)
        r.h_Kmcmc = pointer(hK)
        # argument errors come back with the reference's messages (src/abcdez_smc.jl:223-235)
        if state === nothing && !return_state
            check(ccall((:abcdez_smc_run, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ref{SmcOpts}, Ref{SmcResult}),
                        ctx.h, ph, mh, ϵ_target, o, r))
        else
            need = ccall((:abcdez_smc_state_bytes, LIB), Int64, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32), ph, mh, N, r.hist_cap)
            sout = return_state ? Vector{UInt8}(undef, need) : UInt8[]
            s_in = state === nothing ? UInt8[] : state
            nout = Ref{Int64}(0)
            GC.@preserve s_in sout check(ccall((:abcdez_smc_run_state, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ref{SmcOpts}, Ref{SmcResult}, Ptr{UInt8}, Int64, Ptr{UInt8}, Int64, Ref{Int64}),
                ctx.h, ph, mh, ϵ_target, o, r, state === nothing ? C_NULL : pointer(s_in), length(s_in),
                return_state ? pointer(sout) : C_NULL, length(sout), nout))
            return_state && resize!(sout, nout[])
            state_out = return_state ? sout : nothing
        end
    end
    ccall((:abcdez_prior_destroy, LIB), Cint, (Ptr{Cvoid},), ph); ccall((:abcdez_model_destroy, LIB), Cint, (Ptr{Cvoid},), mh)
    r.status == ABCDEZ_ERR_NO_ALIVE && @warn("No alive particles")                      # src/abcdez_smc.jl:375
    if verbose                                                                          # the per-iteration @info of src/abcdez_smc.jl:372,
        for it in 2:r.hist_len                                                          # printed from the history the device recorded
            @info "Finished run:" iteration = it - 1 ϵ = h[1][it] range_ϵ = (h[2][it], h[3][it]) logZ = h[4][it] ess = h[5][it] facc = h[6][it] Kmcmc = hK[it]
        end
        r.hist_dropped > 0 && @warn("history truncated: $(r.hist_dropped) records did not fit hist_cap")
        @info "Final run:" iteration = r.iters nsim = r.nsims ϵ = r.eps logZ = r.logZ       # src/abcdez_smc.jl:379
    end
    θs = particles(prior, P); blobs = blobs_out(bl, B)
    out = if verboseout                                                                 # src/abcdez_smc.jl:388-393
        n = r.hist_len
        (P = θs, Wns = Wns, C = C, ϵ = r.eps, logZ = r.logZ, blobs = blobs,
         ϵs = h[1][1:n], ranges_ϵ = collect(zip(h[2][1:n], h[3][1:n])), logZs = h[4][1:n], esss = h[5][1:n],
         faccs = h[6][1:n], γ0s = h[7][1:n], Kmcmcs = Int.(hK[1:n]))
    else
        (P = θs, Wns = Wns, C = C, ϵ = r.eps, logZ = r.logZ, blobs = blobs)
    end
    return_state ? merge(out, (state = state_out,)) : out
end

"""
    abcdemc!(prior, dist!, ϵ_target, varexternal; nparticles=50, generations=20, verbose=true, rng, parallel=false)

Reference: src/abcdez_mc.jl:102-172.
"""
function abcdemc!(prior, dist!::DeviceModel, ϵ_target, varexternal;
                  nparticles::Int=50, generations::Int=20, verbose=true, rng=Random.default_rng(),
                  parallel::Bool=false, ctx::Context=run_context(parallel))
    ph = prior_handle(ctx, prior)
    mh, d, B = model_handle(ctx, dist!)
    o = McOpts(nparticles, generations, seed_from(rng))
    N = local_count(ctx, nparticles)
    P = Matrix{Float64}(undef, d, N); C = Vector{Float64}(undef, N); bl = zeros(UInt8, max(B, 1), N)
    r = McResult()
    GC.@preserve P C bl begin
        r.P = pointer(P); r.C = pointer(C); r.blobs = pointer(bl)
        check(ccall((:abcdez_mc_run, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ref{McOpts}, Ref{McResult}),
                    ctx.h, ph, mh, ϵ_target, o, r))
    end
    ccall((:abcdez_prior_destroy, LIB), Cint, (Ptr{Cvoid},), ph); ccall((:abcdez_model_destroy, LIB), Cint, (Ptr{Cvoid},), mh)
    verbose && (@info "End:" converged = r.reached_eps != 0 nsim = r.nsims range_ϵ = (r.dmin, r.dmax))
    (P = particles(prior, P), C = C, reached_ϵ = r.reached_eps != 0, blobs = blobs_out(bl, B))   # src/abcdez_mc.jl:171
end

# ---- after the run: what the reference's tests, example and docs do with a result -------------------------------
"`weightinds` of test/runtests.jl:13-19: stratified resampling indices of normalised weights (1-based)"
function weightinds(ws::Vector{Float64}; rng=Random.default_rng(), ctx::Context=default_context())
    abs(sum(ws) - 1.0) < 1e-8 || error("Sum of weights expected to be 1.0 (approximately)")
    clamp.(wsample_stratified!(rng, ws, zeros(Int, length(ws)); ctx=ctx), 1, length(ws))
end

"Equally weighted posterior sample `P[weightinds(Wns)]` (test/runtests.jl:287-291), resampled and gathered on the device"
function posterior_sample(r; rng=Random.default_rng(), ctx::Context=default_context())
    N = length(r.P); d = length(first(r.P))
    P = Float64[float(r.P[i][k]) for k in 1:d, i in 1:N]; out = similar(P)
    check(ccall((:abcdez_posterior_sample, LIB), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, UInt64, Ptr{Cdouble}, Ptr{Int64}),
                ctx.h, N, d, P, r.Wns, rand(rng, UInt64), out, C_NULL))
    d == 1 && !(first(r.P) isa Tuple) ? vec(out) : [ntuple(k -> out[k, i], d) for i in 1:N]
end

"log evidence of a `verboseout` run at tolerance ϵ from its ladder (docs/src/index.md:282-284)"
function evidence_at(r, ϵ::Real)
    i = findlast(>=(ϵ), r.ϵs)
    i === nothing && error("ϵ is above the whole ladder of this run")
    r.logZs[i]
end

"Posterior model probabilities (examples/minimal_example.jl:58-65) from log evidences"
function model_probabilities(logZs::AbstractVector{<:Real}; prior_probs=fill(1 / length(logZs), length(logZs)))
    w = exp.(logZs .- maximum(logZs)) .* prior_probs
    w ./ sum(w)
end

"""
    abcdesmc_batch(prior, dist!, ϵ_target, nruns; kwargs...) -> Vector{Float64} of logZ

`nruns` replicates of one `abcdesmc!` run in flight together on one GPU (abcdez_smc_run_batch): the evidence
uncertainty of docs/src/index.md:214-220 in about the time of one run.  Returns the log evidences.
"""
function abcdesmc_batch(prior, dist!::DeviceModel, ϵ_target, nruns::Integer; nparticles::Int=100, α=0.95, δess=0.5, nsims_max::Int=10^7,
                        Kmcmc::Int=3, Kmcmc_min=1.0, ABCk=IndicatorStrict0toϵ, rng=Random.default_rng(), ctx::Context=default_context())
    ph = prior_handle(ctx, prior); mh, d, B = model_handle(ctx, dist!)
    opts = Vector{SmcOpts}(undef, nruns); res = Vector{SmcResult}(undef, nruns)
    for i in 1:nruns
        o = SmcOpts(); ccall((:abcdez_smc_opts_default, LIB), Cvoid, (Ref{SmcOpts},), o)
        o.nparticles = nparticles; o.alpha = α; o.delta_ess = δess; o.nsims_max = nsims_max; o.Kmcmc = Kmcmc; o.Kmcmc_min = Kmcmc_min
        o.kernel = kernel_kind(ABCk); o.seed = rand(rng, UInt64); o.verboseout = 0
        opts[i] = o
        r = SmcResult(); r.P = C_NULL; r.Wns = C_NULL; r.C = C_NULL; r.blobs = C_NULL; r.hist_cap = 0
        r.h_eps = r.h_dmin = r.h_dmax = r.h_logZ = r.h_ess = r.h_facc = r.h_gamma0 = C_NULL; r.h_Kmcmc = C_NULL
        res[i] = r
    end
    # the C side takes arrays of structs: pack the mutable structs into contiguous buffers
    so, sr = sizeof(SmcOpts), sizeof(SmcResult)
    ob = Vector{UInt8}(undef, so * nruns); rb = Vector{UInt8}(undef, sr * nruns)
    GC.@preserve ob rb begin
        for i in 1:nruns
            unsafe_copyto!(Ptr{SmcOpts}(pointer(ob, 1 + (i - 1) * so)), Ptr{SmcOpts}(pointer_from_objref(opts[i])), 1)
            unsafe_copyto!(Ptr{SmcResult}(pointer(rb, 1 + (i - 1) * sr)), Ptr{SmcResult}(pointer_from_objref(res[i])), 1)
        end
        phs = fill(ph, nruns); mhs = fill(mh, nruns); eps = fill(Float64(ϵ_target), nruns)
        check(ccall((:abcdez_smc_run_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Cint}),
                    ctx.h, nruns, phs, mhs, eps, ob, rb, C_NULL))
        for i in 1:nruns
            unsafe_copyto!(Ptr{SmcResult}(pointer_from_objref(res[i])), Ptr{SmcResult}(pointer(rb, 1 + (i - 1) * sr)), 1)
        end
    end
    ccall((:abcdez_prior_destroy, LIB), Cint, (Ptr{Cvoid},), ph); ccall((:abcdez_model_destroy, LIB), Cint, (Ptr{Cvoid},), mh)
    [r.logZ for r in res]
end

"The docs' advice (docs/src/index.md:214-220): repeat the run, summarise the log evidences -> (mean, std, logZs)"
function evidence_uncertainty(prior, dist!::DeviceModel, ϵ_target; repeats::Int=8, kwargs...)
    lz = abcdesmc_batch(prior, dist!, ϵ_target, repeats; kwargs...)
    m = sum(lz) / length(lz)
    (m, length(lz) > 1 ? sqrt(sum(abs2, lz .- m) / (length(lz) - 1)) : 0.0, lz)
end

# `ABCdeZ.wsample_stratified!(rng, weights, inds)` (src/abcdez_smc.jl:15-56; used by test/runtests.jl:13-19)
function wsample_stratified!(rng::AbstractRNG, weights::Vector{Float64}, inds::Vector{Int}; ctx::Context=default_context())
    u = rand(rng, length(weights))
    check(ccall((:abcdez_wsample_stratified, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Int64}),
                ctx.h, length(weights), weights, u, 2, inds))
    inds
end

end # module
