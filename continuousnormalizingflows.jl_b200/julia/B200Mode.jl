# B200Mode.jl -- Julia glue for libicnf_b200.so (NOT executed in the build container:
# Julia is not installed there; the Python ctypes mirror in ../api.py exercises the
# same C ABI).  A maintainer of ContinuousNormalizingFlows.jl adds this file to
# src/ (or a package extension) and `include`s it after core/utils.jl.
#
# It adds one compute mode, `B200MatrixMode <: MatrixMode`, and specialises the three
# seams of SURVEY.md 8(b) on it.  Everything else in the package -- `inference`,
# `generate`, `ICNFDist`, `ICNFModel`, the MLJ fit loop -- keeps calling the same
# generic functions and lands here through dispatch.

const libicnf = "libicnf_b200.so"   # on LD_LIBRARY_PATH, or an absolute path

struct B200MatrixMode{ADBack <: ADTypes.AbstractADType} <: MatrixMode{ADBack}
    adback::ADBack          # unused: gradients come from icnf_loss_grad
    precision::Int32        # icnf_precision: 0 = fp32, 1 = bf16 tensor cores (2e-2), 2 = split bf16 tensor cores (1e-4)
    device::Int32           # CUDA ordinal
end
function B200MatrixMode(adback::ADTypes.AbstractADType = ADTypes.AutoZygote(); precision::Symbol = :fp32, device::Integer = 0)
    codes = Dict(:fp32 => 0, :bf16_tc => 1, :bf16x3_tc => 2)
    haskey(codes, precision) || error("B200MatrixMode: precision must be :fp32, :bf16_tc or :bf16x3_tc")
    return B200MatrixMode(adback, Int32(codes[precision]), Int32(device))
end

# ---- C structs (include/icnf_b200.h) ---------------------------------------------
struct IcnfConfig
    abi_version::Int32
    nvars::Int32
    naug::Int32
    ncond::Int32
    autonomous::Int32
    n_layers::Int32
    sizes::NTuple{9, Int32}
    activation::Int32
    lambda1::Float32
    lambda2::Float32
    lambda3::Float32
    reg_squared::Int32
    precision::Int32
    device::Int32
end

struct IcnfSolver
    adaptive::Int32
    dt::Float32
    reltol::Float32
    abstol::Float32
    max_steps::Int32
    beta1::Float32; beta2::Float32; gamma::Float32; qmin::Float32; qmax::Float32
    qsteady_min::Float32; qsteady_max::Float32; qoldinit::Float32
    alg::Int32                      # 0 = Tsit5, 1 = VCABM (the reference's default, icnf.jl:89)
end

struct IcnfNoise
    kind::Int32
    seed::UInt64
    sample_offset::Int64
end

mutable struct IcnfStats
    naccept::Int32; nreject::Int32; nf::Int32; status::Int32
    t_final::Float32; dt_last::Float32
    IcnfStats() = new(0, 0, 0, 0, 0.0f0, 0.0f0)
end

mode_code(::TestMode) = Int32(0)
mode_code(::TrainMode{true}) = Int32(1)
mode_code(::TrainMode{false}) = Int32(2)

# ---- handle cache: never stored in ps / st / icnf, so machines stay serialisable
#      (test/ci_tests/smoke_tests.jl:142) ------------------------------------------
const HANDLES = IdDict{Any, Ptr{Cvoid}}()

# icnf_activation codes; anything else has no kernel
function activation_code(f)
    f === NNlib.softplus && return Int32(0)
    (f === tanh || f === NNlib.tanh_fast) && return Int32(1)
    (f === NNlib.sigmoid || f === NNlib.sigmoid_fast) && return Int32(2)
    f === identity && return Int32(3)
    error("B200MatrixMode: no kernel for activation $(f); use softplus, tanh, sigmoid or identity")
end

# A PlanarLayer (src/layers/planar_layer.jl: f(x) = u * act(w'x + b), parameters (u, w, b)) IS Dense(n_in => 1, act)
# followed by a bias-free Dense(1 => n_out): it reaches the kernels as that pair, with the parameters re-ordered to
# [w; b; u; 0] on the way in and the gradient re-ordered back on the way out (the output bias is not a parameter).
is_planar(nn::Lux.Chain) = length(nn.layers) == 1 && nn.layers[1] isa PlanarLayer
planar_use_bias(::PlanarLayer{USE_BIAS}) where {USE_BIAS} = USE_BIAS

function dense_sizes(nn::Lux.Chain)
    is_planar(nn) && return Int32[nn.layers[1].in_dims, 1, nn.layers[1].out_dims]
    sizes = Int32[nn.layers[1].in_dims]
    for l in nn.layers
        push!(sizes, l.out_dims)
    end
    return sizes
end

function planar_to_lib(pl::PlanarLayer, θ::Vector{Float32})
    no, ni = pl.out_dims, pl.in_dims
    b = planar_use_bias(pl) ? θ[(no + ni + 1):(no + ni + 1)] : zeros(Float32, 1)
    return vcat(θ[(no + 1):(no + ni)], b, θ[1:no], zeros(Float32, no))
end

function planar_grad_from_lib(pl::PlanarLayer, g::Vector{Float32})
    no, ni = pl.out_dims, pl.in_dims
    return planar_use_bias(pl) ? vcat(g[(ni + 2):(ni + 1 + no)], g[1:ni], g[(ni + 1):(ni + 1)]) :
           vcat(g[(ni + 2):(ni + 1 + no)], g[1:ni])
end

function handle(icnf::ICNF{Float32, <:B200MatrixMode})
    get!(HANDLES, icnf) do
        sizes = dense_sizes(icnf.nn)
        padded = ntuple(i -> i <= length(sizes) ? sizes[i] : Int32(0), 9)
        length(sizes) - 1 <= 8 || error("B200MatrixMode: at most 8 Dense layers")
        planar = is_planar(icnf.nn)
        planar || all(l -> l isa Lux.Dense, icnf.nn.layers) ||
            error("B200MatrixMode: nn must be a Lux.Chain of Dense layers or of one PlanarLayer")
        act = activation_code(planar || length(icnf.nn.layers) > 1 ? icnf.nn.layers[1].activation : identity)
        planar || all(l -> activation_code(l.activation) == act, icnf.nn.layers[1:(end - 1)]) ||
            error("B200MatrixMode: all hidden layers must share one activation")
        planar || activation_code(icnf.nn.layers[end].activation) == 3 || error("B200MatrixMode: the last Dense layer must be linear")
        autonomous = icnf isa ICNF{Float32, <:Any, <:Any, <:Any, true}
        cfg = Ref(IcnfConfig(1, icnf.nvariables, icnf.naugments,
            first(sizes) - icnf.nvariables - icnf.naugments - !autonomous,
            autonomous, length(sizes) - 1, padded, act,
            icnf.λ₁, icnf.λ₂, icnf.λ₃, 0, icnf.compute_mode.precision, icnf.compute_mode.device))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:icnf_create, libicnf), Cint, (Ref{IcnfConfig}, Ref{Ptr{Cvoid}}), cfg, out)
        rc == 0 || error(unsafe_string(ccall((:icnf_last_error, libicnf), Cstring, (Ptr{Cvoid},), C_NULL)))
        out[]
    end
end

check(h, rc) = rc == 0 || error(unsafe_string(ccall((:icnf_last_error, libicnf), Cstring, (Ptr{Cvoid},), h)))

function set_params!(h, ps, icnf = nothing)
    θ = ComponentArrays.getdata(ps)::Vector{Float32}
    (!isnothing(icnf) && is_planar(icnf.nn)) && (θ = planar_to_lib(icnf.nn.layers[1], θ))
    GC.@preserve θ check(h, ccall((:icnf_set_params, libicnf), Cint, (Ptr{Cvoid}, Ptr{Float32}, Int64), h, θ, length(θ)))
end

# The library integrates with Tsit5 or with the reference's default `alg = VCABM()` (icnf.jl:89; inference / generate /
# loss on the narrow-MLP family -- the gradient entry point differentiates discrete Tsit5 steps whatever `alg` says, see
# include/icnf_b200.h).  Any other `alg` is an error, never a silent substitution (SURVEY D1).  Fixed-step solves pass
# `adaptive = false, dt = ...`.
function tsit5_opts(icnf)
    kw = icnf.sol_kwargs
    alg = get(kw, :alg, nothing)
    algname = isnothing(alg) ? :none : nameof(typeof(alg))
    (algname === :Tsit5 || algname === :VCABM) ||
        error("B200MatrixMode integrates with Tsit5 or VCABM: construct the ICNF with sol_kwargs = (; alg = Tsit5(), ...); got $(alg)")
    adaptive = get(kw, :adaptive, true)
    maxiters = get(kw, :maxiters, 100000)
    return Ref(IcnfSolver(adaptive ? 1 : 0, Float32(get(kw, :dt, 0.0f0)), Float32(get(kw, :reltol, 1.0f-4)),
        Float32(get(kw, :abstol, 1.0f-4)), Int32(min(maxiters, typemax(Int32))), 0, 0, 0, 0, 0, 0, 0, 0,
        algname === :VCABM ? Int32(1) : Int32(0)))
end

# ---- S2: one ccall per solve (replaces base_sol, src/core/base_icnf.jl:134-140) ---
function base_sol(
    icnf::ICNF{Float32, <:B200MatrixMode, INPLACE},
    prob::SciMLBase.AbstractODEProblem{<:AbstractMatrix{<:Real}, NTuple{2, Float32}, INPLACE},
) where {INPLACE}
    h = handle(icnf)
    set_params!(h, prob.p, icnf)
    f = prob.f.f                      # the closure built by make_ode_func: carries mode, ϵ, ys
    u0 = Matrix{Float32}(prob.u0)
    ufinal = similar(u0)
    ϵ = f.ϵ
    ys = f.nn isa CondLayer ? Matrix{Float32}(f.nn.ys) : nothing
    noise = Ref(IcnfNoise(0, 0, 0))   # SUPPLIED: identical noise to the CPU path
    stats = IcnfStats()
    t0, t1 = prob.tspan
    GC.@preserve u0 ufinal ϵ ys begin
        check(h, ccall((:icnf_solve, libicnf), Cint,
            (Ptr{Cvoid}, Cint, Ref{IcnfSolver}, Float32, Float32, Ptr{Float32}, Ref{IcnfNoise}, Ptr{Float32},
             Ptr{Float32}, Ptr{Float32}, Ref{IcnfStats}, Int64),
            h, mode_code(f.mode), tsit5_opts(icnf), t0, t1, u0, noise, ϵ,
            isnothing(ys) ? C_NULL : pointer(ys), ufinal, stats, size(u0, 2)))
    end
    return ufinal
end

# ---- S3: loss and its gradient (replaces the Zygote + SciMLSensitivity path,
#      src/core/icnf.jl:90-99, :628-649) ------------------------------------------
function b200_loss_grad(icnf, mode, xs, ys, ps; want_dxs = false)
    h = handle(icnf)
    set_params!(h, ps, icnf)
    B = size(xs, 2)
    ϵ = similar(xs, icnf.nvariables + icnf.naugments, B)
    Random.rand!(icnf.rng, icnf.epsdist, ϵ)                     # base_icnf.jl:258-259
    t0, t1 = steer_tspan(icnf, mode)                            # base_icnf.jl:23-43
    # the library's parameter count (a PlanarLayer is served as a Dense pair with n_out more entries: its zero output bias)
    dθ = zeros(Float32, length(ComponentArrays.getdata(ps)) + (is_planar(icnf.nn) ? icnf.nn.layers[1].out_dims + !planar_use_bias(icnf.nn.layers[1]) : 0))
    dxs = want_dxs ? similar(xs) : nothing
    loss = Ref{Float32}(0)
    stats = IcnfStats()
    noise = Ref(IcnfNoise(0, 0, 0))
    GC.@preserve xs ys ϵ dθ dxs begin
        check(h, ccall((:icnf_loss_grad, libicnf), Cint,
            (Ptr{Cvoid}, Cint, Ref{IcnfSolver}, Float32, Float32, Ptr{Float32}, Ref{IcnfNoise}, Ptr{Float32},
             Ptr{Float32}, Ref{Float32}, Ptr{Float32}, Ptr{Float32}, Ref{IcnfStats}, Int64, Int64),
            h, mode_code(mode), tsit5_opts(icnf), t0, t1, xs, noise, ϵ,
            isnothing(ys) ? C_NULL : pointer(ys), loss, dθ, want_dxs ? pointer(dxs) : C_NULL, stats, B, 0))
    end
    is_planar(icnf.nn) && (dθ = planar_grad_from_lib(icnf.nn.layers[1], dθ))
    return loss[], dθ, dxs
end

function loss(icnf::ICNF{Float32, <:B200MatrixMode}, mode::Mode, xs::AbstractMatrix{<:Real}, ps::Any, st::NamedTuple)
    return first(b200_loss_grad(icnf, mode, Matrix{Float32}(xs), nothing, ps))
end

# conditioned signature, src/core/icnf.jl:639-649
function loss(icnf::ICNF{Float32, <:B200MatrixMode}, mode::Mode, xs::AbstractMatrix{<:Real}, ys::AbstractMatrix{<:Real},
              ps::Any, st::NamedTuple)
    return first(b200_loss_grad(icnf, mode, Matrix{Float32}(xs), Matrix{Float32}(ys), ps))
end

function ChainRulesCore.rrule(
    ::typeof(loss), icnf::ICNF{Float32, <:B200MatrixMode}, mode::Mode, xs::AbstractMatrix{<:Real}, ps::Any, st::NamedTuple,
)
    l, dθ, dxs = b200_loss_grad(icnf, mode, Matrix{Float32}(xs), nothing, ps; want_dxs = true)
    function loss_pullback(l̄)
        NT = ChainRulesCore.NoTangent()
        return NT, NT, NT, l̄ .* dxs, ComponentArrays.ComponentArray(l̄ .* dθ, ComponentArrays.getaxes(ps)), NT
    end
    return l, loss_pullback
end

function ChainRulesCore.rrule(
    ::typeof(loss), icnf::ICNF{Float32, <:B200MatrixMode}, mode::Mode, xs::AbstractMatrix{<:Real}, ys::AbstractMatrix{<:Real},
    ps::Any, st::NamedTuple,
)
    l, dθ, dxs = b200_loss_grad(icnf, mode, Matrix{Float32}(xs), Matrix{Float32}(ys), ps; want_dxs = true)
    function cond_loss_pullback(l̄)
        NT = ChainRulesCore.NoTangent()
        # the conditioning input is data: the library returns no gradient for it (neither do the reference's callers ask)
        return NT, NT, NT, l̄ .* dxs, ChainRulesCore.ZeroTangent(),
               ComponentArrays.ComponentArray(l̄ .* dθ, ComponentArrays.getaxes(ps)), NT
    end
    return l, cond_loss_pullback
end

# ---- multi-GPU (SURVEY 8(e)): one handle per device, batch columns sharded, ONE all-reduce of [dθ; loss] inside the
#      library (NCCL over NVLink; tiny gradients go through peer memory in one kernel).  Single process, several GPUs:
#
#          hs = [handle(icnf_on_device_g) for g in 0:(G - 1)]
#          check(C_NULL, ccall((:icnf_create_group, libicnf), Cint, (Ptr{Ptr{Cvoid}}, Int32), hs, G))
#
#      then call `icnf_loss_grad_dp` (same arguments as `icnf_loss_grad`, `global_batch` = the unsharded batch,
#      `noise.sample_offset` = first global column of the shard) once per handle from one task per device; every
#      shard returns the gradient of the WHOLE batch.  One process per GPU (MPI.jl / Distributed.jl): rank 0 calls
#      `icnf_group_unique_id`, ships the 128 bytes to the other ranks, every rank calls `icnf_group_join`.
