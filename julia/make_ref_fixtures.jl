# make_ref_fixtures.jl -- reference-held golden vectors for the hot path.
#
#     julia --project=<env with KissABC 3.0.1> julia/make_ref_fixtures.jl [path/to/KissABC.jl/src/KissABC.jl]
#
# Runs the UNMODIFIED reference (KissABC.smc; the AIS init step and `transition!` through the two AbstractMCMC.step
# methods) on a PhiloxRNG in SERIAL mode -- one Philox word stream (seed; tag 6, id 0, epoch 0) consumed in the reference's
# own order -- and writes tests/golden/ref_smc_*.json / ref_ais_*.json.  tests/test_ref_fixtures.py replays the same stream
# through the CPU oracle's serial mode (kor_smc_set_serial / kor_ais_set_serial: identical control logic to the
# per-particle-stream mode the device is bit-compared with) and requires
#   * identical integer results: iterations, ESS per iteration, alive mask, accepted counts (decisions),
#   * floats within 1e-12 relative: eps per iteration, theta, costs, log-densities.
# Floats are not bit-identical by construction: the reference's own arithmetic uses Julia's log (src/smc.jl:166,
# src/transition.jl:58), pairwise `sum` in mean/std and `hypot` (README.md:46-52), where the variate spec fixes sequential
# sums and sqrt(a^2+b^2); the test states the bound.
#
# NOT executed in this repository's CI (no Julia in the build image).  Floats are written as their UInt64 bit patterns.
using Random, Statistics, Distributions
if length(ARGS) >= 1
    include(ARGS[1])              # a checkout of the reference: .../KissABC.jl/src/KissABC.jl
    using .KissABC
else
    using KissABC
end
import AbstractMCMC
include(joinpath(@__DIR__, "PhiloxRNG.jl"))
using .PhiloxRNGs

const SEED = UInt64(0x4B49535341424300)
const OUT = joinpath(@__DIR__, "..", "tests", "golden")

bits(x::Real) = string(reinterpret(UInt64, Float64(x)))
jarr(v) = "[" * join(v, ",") * "]"
jbits(v) = jarr(bits.(v))
jstr(s) = "\"" * s * "\""
jobj(pairs) = "{" * join(["\"$(k)\":$(v)" for (k, v) in pairs], ",") * "}"

# ---- the registered simulators, written the way a KissABC user writes a cost closure (README.md:35-52), drawing from the
#      run's own rng through the spec's normal generator
function cost_normal(rng, n)                       # README.md:46-52, targets of test/runtests.jl:283-284
    function (θ)
        μ, σ = θ
        x = spec_normals(rng, n) .* σ .+ μ
        hypot(mean(x) - 2.0, (std(x) - 0.04) * 50)
    end
end
function cost_ma2(rng, n, target)                  # SURVEY.md Appendix B
    function (θ)
        t1, t2 = θ
        (t1 > -2.0 && t1 < 2.0 && t1 + t2 > -1.0 && t1 - t2 < 1.0) || return Inf
        e = spec_normals(rng, n + 2)
        y = [e[t+2] + t1 * e[t+1] + t2 * e[t] for t in 1:n]
        a1 = 0.0; for t in 2:n; a1 += y[t] * y[t-1]; end
        a2 = 0.0; for t in 3:n; a2 += y[t] * y[t-2]; end
        sqrt((a1 / n - target[1])^2 + (a2 / n - target[2])^2)
    end
end

function cost_noisyprod(rng)                       # test/runtests.jl:105-112 with the run's rng instead of the global one
    function (θ)
        n, du = θ
        abs((n * n + du) * (n + randn(rng) * 0.01) - 5.5)
    end
end

# `verbose && @show iteration, ϵ, ESS` (src/smc.jl:143) is the only per-iteration output of the reference: capture it
function capture_stdout(f)
    path, io = mktemp()
    local res
    redirect_stdout(io) do
        res = f()
    end
    close(io)
    txt = read(path, String)
    rm(path)
    res, txt
end
function parse_show(txt)
    its, eps, ess = Int[], Float64[], Int[]
    for m in eachmatch(r"\(iteration, ϵ, ESS\) = \((\d+), ([^,]+), (\d+)\)", txt)
        push!(its, parse(Int, m.captures[1])); push!(eps, parse(Float64, m.captures[2])); push!(ess, parse(Int, m.captures[3]))
    end
    its, eps, ess
end

function smc_fixture(name, prior_spec, prior, mkcost, model_spec; kw...)
    rng = PhiloxRNG(SEED)
    cost = mkcost(rng)
    res, txt = capture_stdout(() -> smc(prior, cost; rng=rng, verbose=true, parallel=false, kw...))
    its, eps, ess = parse_show(txt)
    P = res.P isa AbstractVector ? res.P : [res.P]
    open(joinpath(OUT, "ref_smc_$(name).json"), "w") do io
        write(io, jobj([
            "kind" => jstr("smc"), "name" => jstr(name), "seed" => string(SEED), "prior" => prior_spec, "model" => model_spec,
            "kwargs" => jobj([string(k) => (v isa Integer ? string(v) : bits(v)) for (k, v) in kw]),
            "int_kwargs" => jarr([jstr(string(k)) for (k, v) in kw if v isa Integer]),
            "iterations" => string(length(its)), "eps_per_iteration" => jbits(eps), "ess_per_iteration" => jarr(ess),
            "eps" => bits(res.ϵ), "C" => jbits(res.C),
            "P" => jarr([jbits(p.particles) for p in P]),
            "words_consumed" => string(words_consumed(rng)),
            "julia" => jstr(string(VERSION))]))
    end
    println("ref_smc_$(name): ", length(its), " iterations, eps = ", res.ϵ, ", alive = ", length(P[1].particles))
end

function ais_fixture(name, prior_spec, prior, mkcost, model_spec, scale, N, steps, ntransitions; posterior=0)
    rng = PhiloxRNG(SEED)
    # posterior = 1: the hard-threshold ApproxPosterior (src/types.jl:76-104), `scale` is its maxcost and the second slot of a
    # log-density is the cost
    model = posterior == 1 ? ApproxPosterior(prior, mkcost(rng), scale) : ApproxKernelizedPosterior(prior, mkcost(rng), scale)
    spl = AIS(N)
    sample0, state = AbstractMCMC.step(rng, model, spl)                       # src/KissABC.jl:35-64
    flat(ps) = [Float64(p.x[k]) for k in 1:length(prior), p in ps]           # d x N
    th_init = flat(state.sample)
    lp_init = [ld[1] for ld in state.loglikelihood]; ll_init = [ld[2] for ld in state.loglikelihood]   # (logprior, loglikelihood | cost)
    samples = [collect(Float64, sample0.x)]
    for s in 1:steps
        smp, state = AbstractMCMC.step(rng, model, spl, state; ntransitions=ntransitions)   # src/KissABC.jl:66-80
        push!(samples, collect(Float64, smp.x))
    end
    th = flat(state.sample)
    lp = [ld[1] for ld in state.loglikelihood]; ll = [ld[2] for ld in state.loglikelihood]
    open(joinpath(OUT, "ref_ais_$(name).json"), "w") do io
        write(io, jobj([
            "kind" => jstr("ais"), "name" => jstr(name), "seed" => string(SEED), "prior" => prior_spec, "model" => model_spec,
            "scale" => bits(scale), "nwalkers" => string(N), "steps" => string(steps), "ntransitions" => string(ntransitions),
            "posterior" => string(posterior),
            "theta_init" => jbits(vec(permutedims(th_init))), "lp_init" => jbits(lp_init), "ll_init" => jbits(ll_init),
            "samples" => jarr([jbits(s) for s in samples]),
            "theta" => jbits(vec(permutedims(th))), "lp" => jbits(lp), "ll" => jbits(ll),
            "words_consumed" => string(words_consumed(rng)),
            "julia" => jstr(string(VERSION))]))
    end
    println("ref_ais_$(name): ", steps, " steps x ", ntransitions, " transitions")
end

# prior specs in the oracle's notation (tests/common.py): only laws whose Distributions.jl sampler is a fixed transform of
# rand / randn / rand(a:b) (Uniform: a + (b-a) rand; Normal: mu + sigma randn; DiscreteUniform: rand(rng, a:b)) can be
# replayed; Truncated uses its own rejection scheme
const UU_NORMAL = "[[\"uniform\",1,3],[\"uniform\",0.01,0.2]]"
const UU_MA2 = "[[\"uniform\",-2,2],[\"uniform\",-1,1]]"
const MA2_T = (0.72, 0.2)

smc_fixture("normal_defaults", UU_NORMAL, Factored(Uniform(1, 3), Uniform(0.01, 0.2)), r -> cost_normal(r, 200),
            "{\"kind\":\"normal\",\"n\":200}"; nparticles=400, epstol=0.05)
smc_fixture("normal_sparse_resampling", UU_NORMAL, Factored(Uniform(1, 3), Uniform(0.01, 0.2)), r -> cost_normal(r, 100),
            "{\"kind\":\"normal\",\"n\":100}"; nparticles=300, alpha=0.8, min_r_ess=0.4, mcmc_retrys=2, mcmc_tol=0.3, epstol=0.1)
smc_fixture("ma2", UU_MA2, Factored(Uniform(-2, 2), Uniform(-1, 1)), r -> cost_ma2(r, 100, MA2_T),
            "{\"kind\":\"ma2\",\"n\":100}"; nparticles=500, alpha=0.9, epstol=0.2)
ais_fixture("normal", UU_NORMAL, Factored(Uniform(1, 3), Uniform(0.01, 0.2)), r -> cost_normal(r, 100),
            "{\"kind\":\"normal\",\"n\":100}", 0.05, 12, 60, 3)
ais_fixture("ma2", UU_MA2, Factored(Uniform(-2, 2), Uniform(-1, 1)), r -> cost_ma2(r, 100, MA2_T),
            "{\"kind\":\"ma2\",\"n\":100}", 0.2, 10, 40, 2)
# discrete component: DiscreteUniform is sampled as rand(rng, a:b) and push_p rounds it (src/types.jl:32)
smc_fixture("noisyprod_discrete", "[[\"normal\",1,0.5],[\"duniform\",1,10]]", Factored(Normal(1, 0.5), DiscreteUniform(1, 10)),
            cost_noisyprod, "{\"kind\":\"noisyprod\"}"; nparticles=200, epstol=0.02)
ais_fixture("hard_normal", UU_NORMAL, Factored(Uniform(1, 3), Uniform(0.01, 0.2)), r -> cost_normal(r, 100),
            "{\"kind\":\"normal\",\"n\":100}", 0.3, 12, 60, 3; posterior=1)
