# KissABCCuda.jl -- thin Julia shim over libkissabc_cuda.so (include/kissabc_cuda.h).
# NOT executed in this repository's CI (no Julia in the build image); every call it makes is mirrored 1:1 by the
# tested Python ctypes host (kissabc.jl_b200/api.py).  Drop next to KissABC.jl and `using KissABCCuda`.
#
# It adds methods to the reference's own entry points so existing scripts keep working:
#     smc(prior, cost::DeviceCost; kw...)                                  (ref src/smc.jl:92-206)
#     sample(ApproxKernelizedPosterior(prior, cost::DeviceCost, eps), AIS(N), Ns; kw...)   (ref src/KissABC.jl:35-94)
module KissABCCuda

using KissABC, Distributions, MonteCarloMeasurements
import KissABC: smc, sample, AIS, ApproxKernelizedPosterior, Factored

const LIB = get(ENV, "KISSABC_CUDA_LIB", joinpath(@__DIR__, "..", "libkissabc_cuda.so"))

# ---- PODs, same layout as the header
struct KabcPrior
    kind::Int32; _pad::Int32
    p0::Float64; p1::Float64; lo::Float64; hi::Float64
end
struct KabcModel
    kind::Int32; precision::Int32; n_draws::Int32; n_target::Int32
    target::NTuple{32,Float64}; param::NTuple{8,Float64}
end
struct KabcSmcConfig
    nparticles::Int64; alpha::Float64; mcmc_retrys::Int64; mcmc_tol::Float64; epstol::Float64
    r_epstol::Float64; min_r_ess::Float64; max_stretch::Float64; verbose::Int32; max_iterations::Int32
end
struct KabcAisConfig
    nwalkers::Int64; nsamples::Int64; ntransitions::Int64; discard_initial::Int64; thinning::Int64
    retry_sampling::Int64; scale::Float64; posterior::Int32; _pad::Int32
end
struct KabcSmcLog
    iteration::Int64; eps::Float64; n_alive::Int64; flag::Int32; resampled::Int32
    accepted::Int64; cost_evals::Int64; sweeps::Int64
end

# ---- the registered device costs that replace the `cost` closure
abstract type DeviceCost end
pad(v, n) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, n)
struct NormalMeanStd <: DeviceCost; n::Int; mean::Float64; std::Float64; weight::Float64; f32::Bool; end
NormalMeanStd(; n=1000, mean=2.0, std=0.04, weight=50.0, f32=true) = NormalMeanStd(n, mean, std, weight, f32)
pod(c::NormalMeanStd) = KabcModel(0, c.f32, c.n, 2, pad((c.mean, c.std), 32), pad((c.weight,), 8))
struct MA2 <: DeviceCost; n::Int; target::NTuple{2,Float64}; f32::Bool; end
pod(c::MA2) = KabcModel(1, c.f32, c.n, 2, pad(c.target, 32), pad((), 8))
struct GandK <: DeviceCost; n::Int; target::NTuple{7,Float64}; c::Float64; f32::Bool; end
pod(c::GandK) = KabcModel(2, c.f32, c.n, 7, pad(c.target, 32), pad((c.c,), 8))
struct LotkaVolterra <: DeviceCost; target::Vector{Float64}; x0::Float64; y0::Float64; T::Float64; max_events::Int; f32::Bool; end
pod(c::LotkaVolterra) = KabcModel(3, c.f32, 0, length(c.target), pad(c.target, 32),
                                  pad((c.x0, c.y0, c.T, length(c.target) ÷ 2, c.max_events), 8))
# the simulators of the reference's own integration tests (test/runtests.jl:34-44 and :105-112)
struct Socks <: DeviceCost; target::NTuple{2,Float64}; n_picked::Int; end
pod(c::Socks) = KabcModel(5, 0, 0, 2, pad(c.target, 32), pad((c.n_picked,), 8))
struct NoisyProduct <: DeviceCost; target::Float64; noise::Float64; end
pod(c::NoisyProduct) = KabcModel(4, 0, 0, 1, pad((c.target,), 32), pad((2.0, c.noise), 8))

pod(d::Uniform) = KabcPrior(0, 0, d.a, d.b, d.a, d.b)
pod(d::Normal) = KabcPrior(1, 0, d.μ, d.σ, -Inf, Inf)
pod(d::Truncated{<:Normal}) = KabcPrior(2, 0, d.untruncated.μ, d.untruncated.σ, d.lower, d.upper)
pod(d::Beta) = KabcPrior(3, 0, d.α, d.β, 0.0, 1.0)
pod(d::NegativeBinomial) = KabcPrior(4, 0, d.r, d.p, 0.0, Inf)      # discrete: the library applies push_p (round) itself
pod(d::DiscreteUniform) = KabcPrior(5, 0, d.a, d.b, d.a, d.b)
pods(p::Factored) = KabcPrior[pod(q) for q in p.p]
pods(p::UnivariateDistribution) = KabcPrior[pod(p)]

lasterror() = unsafe_string(ccall((:kabc_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 || error(lasterror())       # the reference raises ErrorException: same here

mutable struct Context
    h::Ptr{Cvoid}
    function Context(; device=0, seed=UInt64(0x4B49535341424300))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:kabc_ctx_create, LIB), Cint, (Cint, UInt64, Ref{Ptr{Cvoid}}), device, seed, r))
        c = new(r[]); finalizer(x -> ccall((:kabc_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c); c
    end
    # one rank of a multi-GPU job (one process per GPU): `id` = the 128 bytes of nccl_unique_id() made by rank 0 and moved by
    # the host (MPI.Bcast!, Distributed).  The context exchanges and maps its peer arena by itself.
    function Context(rank::Integer, world::Integer, id::Vector{UInt8}; device=0, seed=UInt64(0x4B49535341424300))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:kabc_ctx_create_dist, LIB), Cint, (Cint, UInt64, Cint, Cint, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
                    device, seed, rank, world, id, r))
        c = new(r[]); finalizer(x -> ccall((:kabc_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c); c
    end
end
nccl_unique_id() = (id = Vector{UInt8}(undef, 128); check(ccall((:kabc_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)
const DEFAULT = Ref{Union{Nothing,Context}}(nothing)
ctx() = (DEFAULT[] === nothing && (DEFAULT[] = Context()); DEFAULT[])

bundle(θ::Matrix{Float64}) = (P = [Particles(θ[:, k]) for k in 1:size(θ, 2)]; length(P) == 1 ? P[1] : P)

function smc(prior::Distribution, cost::DeviceCost; nparticles::Int=100, alpha=0.95, mcmc_retrys::Int=0,
             mcmc_tol=0.015, epstol=0.0, r_epstol=(1 - alpha)^1.5 / 50, min_r_ess=alpha^2, max_stretch=2.0,
             verbose::Bool=false, parallel::Bool=false, context::Context=ctx())
    pr = pods(prior); d = length(pr); N = nparticles
    cfg = KabcSmcConfig(N, alpha, mcmc_retrys, mcmc_tol, epstol, r_epstol, min_r_ess, max_stretch, verbose, 0)
    θ = Matrix{Float64}(undef, N, d); alive = Vector{UInt8}(undef, N); C = Vector{Float64}(undef, N)
    ϵ = Ref(0.0); it = Ref{Int64}(0); ev = Ref{Int64}(0); log = Vector{KabcSmcLog}(undef, 1 << 16)
    check(ccall((:kabc_smc_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{KabcPrior}, Cint, Ref{KabcModel}, Ref{KabcSmcConfig}, Ptr{Float64}, Ptr{UInt8},
                 Ptr{Float64}, Ref{Float64}, Ref{Int64}, Ref{Int64}, Ptr{KabcSmcLog}, Int64),
                context.h, pr, d, pod(cost), cfg, θ, alive, C, ϵ, it, ev, log, length(log)))
    verbose && foreach(r -> println("(iteration, ϵ, ESS) = ", (r.iteration, r.eps, r.n_alive)), log[1:it[]])
    (P = bundle(θ[alive .== 1, :]), C = C, ϵ = ϵ[])            # ref src/smc.jl:200-205
end

function sample(model::ApproxKernelizedPosterior{<:Any,<:DeviceCost}, spl::AIS, Ns::Integer; ntransitions::Int=1,
                discard_initial::Int=0, thinning::Int=1, retry_sampling::Int=100, context::Context=ctx(), kwargs...)
    pr = pods(model.prior); d = length(pr)
    cfg = KabcAisConfig(spl.nparticles, Ns, ntransitions, discard_initial, thinning, retry_sampling, model.scale, 0, 0)
    out = Matrix{Float64}(undef, Ns, d); ev = Ref{Int64}(0); acc = Ref{Int64}(0)
    check(ccall((:kabc_ais_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{KabcPrior}, Cint, Ref{KabcModel}, Ref{KabcAisConfig}, Ptr{Float64}, Ref{Int64}, Ref{Int64}),
                context.h, pr, d, pod(model.cost), cfg, out, ev, acc))
    bundle(out)                                                  # ref src/KissABC.jl:82-94
end

# hard-threshold posterior, ref src/types.jl:76-104: same entry point, posterior = 1, scale = maxcost
function sample(model::KissABC.ApproxPosterior{<:Any,<:DeviceCost}, spl::AIS, Ns::Integer; ntransitions::Int=1,
                discard_initial::Int=0, thinning::Int=1, retry_sampling::Int=100, context::Context=ctx(), kwargs...)
    pr = pods(model.prior); d = length(pr)
    cfg = KabcAisConfig(spl.nparticles, Ns, ntransitions, discard_initial, thinning, retry_sampling, model.maxcost, 1, 0)
    out = Matrix{Float64}(undef, Ns, d); ev = Ref{Int64}(0); acc = Ref{Int64}(0)
    check(ccall((:kabc_ais_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{KabcPrior}, Cint, Ref{KabcModel}, Ref{KabcAisConfig}, Ptr{Float64}, Ref{Int64}, Ref{Int64}),
                context.h, pr, d, pod(model.cost), cfg, out, ev, acc))
    bundle(out)
end

# ---- ABCDE / pfilter, ref src/smc.jl:275-428
struct KabcAbcdeConfig
    nparticles::Int64; generations::Int64; eps_target::Float64; alpha::Float64; proposal_width::Float64
    earlystop::Int32; _pad::Int32
end
struct KabcPfilterConfig
    nparticles::Int64; q::Float64; eff_tol::Float64; epstol::Float64; proposal_width::Float64; max_iters::Int64
end

function KissABC.ABCDE(prior::Distribution, cost::DeviceCost, ϵ_target; nparticles=50, generations=20, α=0, parallel=false,
                       earlystop=false, verbose=true, proposal_width=1.0, context::Context=ctx())
    pr = pods(prior); d = length(pr); N = nparticles
    cfg = KabcAbcdeConfig(N, generations, ϵ_target, α, proposal_width, earlystop, 0)
    θ = Matrix{Float64}(undef, N, d); Δ = Vector{Float64}(undef, N)
    reached = Ref{Int32}(0); nsim = Ref{Int64}(0); gens = Ref{Int64}(0)
    check(ccall((:kabc_abcde_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{KabcPrior}, Cint, Ref{KabcModel}, Ref{KabcAbcdeConfig}, Ptr{Float64}, Ptr{Float64},
                 Ref{Int32}, Ref{Int64}, Ref{Int64}),
                context.h, pr, d, pod(cost), cfg, θ, Δ, reached, nsim, gens))
    (P=bundle(θ), C=Particles(Δ), reached_ϵ=reached[] != 0)
end

function KissABC.pfilter(prior::Distribution, cost::DeviceCost, N; q=0.7, eff_tol=0.1, epstol=-Inf, max_iters=Inf,
                         proposal_width=0.75, verbose=false, parallel=false, context::Context=ctx())
    pr = pods(prior); d = length(pr)
    n = ccall((:kabc_pfilter_nparticles, LIB), Int64, (Int64, Cint, Float64), N, d, q)
    cfg = KabcPfilterConfig(N, q, eff_tol, epstol, proposal_width, isinf(max_iters) ? 0 : Int64(max_iters))
    θ = Matrix{Float64}(undef, n, d); C = Vector{Float64}(undef, n)
    ϵ = Ref(0.0); it = Ref{Int64}(0); reps = Ref{Int64}(0); ev = Ref{Int64}(0)
    check(ccall((:kabc_pfilter_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{KabcPrior}, Cint, Ref{KabcModel}, Ref{KabcPfilterConfig}, Ptr{Float64}, Ptr{Float64},
                 Ref{Float64}, Ref{Int64}, Ref{Int64}, Ref{Int64}),
                context.h, pr, d, pod(cost), cfg, θ, C, ϵ, it, reps, ev))
    (P=bundle(θ), C=Particles(C))
end

export DeviceCost, NormalMeanStd, MA2, GandK, LotkaVolterra, Socks, NoisyProduct, Context
end # module
