# dump_reference.jl -- golden vectors FROM THE UNMODIFIED REFERENCE (VERDICT r1 item 10 / SURVEY 8(c)).
#
# Run on any machine with Julia and ContinuousNormalizingFlows.jl v0.31 (+ OrdinaryDiffEqTsit5):
#
#     julia --project=<env with ContinuousNormalizingFlows, OrdinaryDiffEqTsit5, Zygote, ComponentArrays, Lux>
#           tests/golden/dump_reference.jl  tests/golden/reference
#
# It runs the package's own public API (`inference`, `generate`, `loss`, Zygote gradient of `loss`) with
# `alg = Tsit5()`, fixed seeds and Float32, and writes, per case, raw little-endian Float32 files plus
# `manifest.json`.  Nothing of this repository is imported.  tests/test_reference_golden.py picks the
# directory up when it exists (it is skipped while it does not: Julia is not installed in the build
# container or on the GPU box) and holds BOTH the CPU oracle and the CUDA path to these numbers --
# that is the step that turns "parity unpinned" into pinned.
#
# Layout of every matrix: Julia's own column-major R x B (one column per sample) = what the C ABI takes.
import ContinuousNormalizingFlows as CNF
import ComponentArrays, Lux, LuxCore, Random, Zygote, OrdinaryDiffEqTsit5, SciMLSensitivity, ADTypes, Distributions

const outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "reference")
mkpath(outdir)

writef32(name, a) = open(io -> write(io, Float32.(vec(collect(a)))), joinpath(outdir, name), "w")

jstr(x::AbstractString) = "\"" * x * "\""
jstr(x::Bool) = x ? "true" : "false"
jstr(x::Real) = string(x)
jstr(x::AbstractVector) = "[" * join(jstr.(x), ", ") * "]"
jstr(d::AbstractDict) = "{" * join([jstr(string(k)) * ": " * jstr(v) for (k, v) in d], ", ") * "}"

function sol_kwargs(; adaptive::Bool, dt::Float32 = 0.0f0, tol::Float32 = 1.0f-4)
    base = (;
        save_everystep = false,
        alg = OrdinaryDiffEqTsit5.Tsit5(),
        reltol = tol,
        abstol = tol,
        sensealg = SciMLSensitivity.QuadratureAdjoint(;
            autodiff = true, autojacvec = SciMLSensitivity.ZygoteVJP(), reltol = tol, abstol = tol),
    )
    return adaptive ? merge(base, (; maxiters = typemax(Int))) : merge(base, (; adaptive = false, dt = dt))
end

# one case = one ICNF configuration + inputs; every quantity is computed with the rng re-seeded to `seed`,
# so that the Hutchinson probe (base_icnf.jl:258-259) is the SAME matrix in all of them and in `eps.f32`
function dump_case(name; nvariables, naugments, nconditions = 0, n_hidden = nothing, B, seed, adaptive, dt = 0.0f0,
                   steer_rate = 0.0f0, tol = 1.0f-4)
    rng = Random.Xoshiro(seed)
    kw = (; nvariables, naugments, nconditions, rng, steer_rate,
          compute_mode = CNF.LuxVecJacMatrixMode(ADTypes.AutoZygote()),
          sol_kwargs = sol_kwargs(; adaptive, dt, tol))
    icnf = isnothing(n_hidden) ? CNF.ICNF(; kw...) : CNF.ICNF(; kw..., n_hidden)
    ps, st = LuxCore.setup(Random.Xoshiro(seed + 1), icnf)
    ps = ComponentArrays.ComponentArray(ps)
    # non-zero biases so that the bias path is exercised (Lux initialises them to zero)
    θ = ComponentArrays.getdata(ps)
    θ .+= 0.05f0 .* randn(Random.Xoshiro(seed + 2), Float32, length(θ))
    data_rng = Random.Xoshiro(seed + 3)
    xs = randn(data_rng, Float32, nvariables, B)
    ys = nconditions > 0 ? randn(data_rng, Float32, nconditions, B) : nothing
    D = nvariables + naugments
    args(x) = isnothing(ys) ? (x,) : (x, ys)

    # the probe and the steered end time the calls below will draw (same rng state, same order)
    Random.seed!(icnf.rng, seed)
    ϵ = zeros(Float32, D, B)
    Random.rand!(icnf.rng, icnf.epsdist, ϵ)
    t0, t1 = CNF.steer_tspan(icnf, CNF.TrainMode{true}())

    Random.seed!(icnf.rng, seed)
    logp_test, _ = CNF.inference(icnf, CNF.TestMode(), args(xs)..., ps, st)
    Random.seed!(icnf.rng, seed)
    logp_train, (E, n, A) = CNF.inference(icnf, CNF.TrainMode{true}(), args(xs)..., ps, st)
    Random.seed!(icnf.rng, seed)
    l = CNF.loss(icnf, CNF.TrainMode{true}(), args(xs)..., ps, st)
    Random.seed!(icnf.rng, seed)
    gθ = only(Zygote.gradient(p -> CNF.loss(icnf, CNF.TrainMode{true}(), args(xs)..., p, st), ps))
    Random.seed!(icnf.rng, seed)
    gx = only(Zygote.gradient(x -> CNF.loss(icnf, CNF.TrainMode{true}(), args(x)..., ps, st), xs))
    # generate: base sample drawn from the re-seeded rng (base_icnf.jl:360-361), then the probe
    Random.seed!(icnf.rng, seed)
    z0 = zeros(Float32, D, B)
    Random.rand!(icnf.rng, icnf.basedist, z0)
    Random.seed!(icnf.rng, seed)
    gen = isnothing(ys) ? CNF.generate(icnf, CNF.TestMode(), ps, st, B) : CNF.generate(icnf, CNF.TestMode(), ys, ps, st, B)

    for (f, a) in (("theta", θ), ("xs", xs), ("eps", ϵ), ("logp_test", logp_test), ("logp_train", logp_train),
                   ("E", E), ("n", n), ("A", A), ("loss", [l]), ("dtheta", ComponentArrays.getdata(gθ)), ("dxs", gx),
                   ("z0", z0), ("generate", gen))
        writef32("$(name)_$(f).f32", a)
    end
    isnothing(ys) || writef32("$(name)_ys.f32", ys)
    sizes = [l.in_dims for l in icnf.nn.layers]
    push!(sizes, icnf.nn.layers[end].out_dims)
    return Dict("name" => name, "nvariables" => nvariables, "naugments" => naugments, "nconditions" => nconditions,
                "sizes" => sizes, "B" => B, "seed" => seed, "adaptive" => adaptive, "dt" => dt, "reltol" => tol,
                "abstol" => tol, "steer_rate" => steer_rate, "t0" => t0, "t1" => t1,
                "lambda" => [icnf.λ₁, icnf.λ₂, icnf.λ₃])
end

cases = [
    # fixed step: the stepper arithmetic is pinned by the tableau alone -> tight comparison (1e-5)
    dump_case("usage_fixed"; nvariables = 1, naugments = 2, B = 64, seed = 11, adaptive = false, dt = 0.125f0),
    dump_case("moons_fixed"; nvariables = 2, naugments = 0, B = 256, seed = 12, adaptive = false, dt = 0.125f0),
    dump_case("cond_fixed"; nvariables = 2, naugments = 1, nconditions = 2, n_hidden = 8, B = 64, seed = 13, adaptive = false, dt = 0.25f0),
    # adaptive at the package's default tolerances: controller constants, initial step and error norm of
    # OrdinaryDiffEq enter -> comparison at solver tolerance; the manifest records nothing about the steps
    dump_case("usage_adaptive"; nvariables = 1, naugments = 2, B = 64, seed = 21, adaptive = true),
    dump_case("moons_adaptive_steer"; nvariables = 2, naugments = 0, B = 256, seed = 22, adaptive = true, steer_rate = 0.1f0),
    dump_case("gmm16_adaptive"; nvariables = 16, naugments = 0, B = 128, seed = 23, adaptive = true),
]
open(joinpath(outdir, "manifest.json"), "w") do io
    write(io, "{\"reference\": \"ContinuousNormalizingFlows.jl\", \"alg\": \"Tsit5\", \"cases\": " * jstr(cases) * "}\n")
end
println("wrote ", length(cases), " cases to ", outdir)
