"""
RbqEngine.jl — `ccall` binding of librbq.so underneath the reference's RobotDynamics model API.

NOT EXECUTED IN THE BUILD CONTAINER (no Julia there): written against include/rbq.h and the reference call sites
    RD.discrete_dynamics(::Type{RK3}, model, astate, acontrol, time, dt)     spin13.jl:44-58 (and every spinYY.jl)
    RD.discrete_jacobian!(::Type{RK3}, ∇f, model, z)                          RobotDynamics 0.2.1 default (ForwardDiff)
    run_sim_prop(gate_count, gate_type; ...)                                  spin.jl:1553-1608
Usage from a reference script (e.g. src/spin/spin13.jl), after its own `struct Model ... end`:

    include(joinpath(ENV["RBQ_ENGINE_PATH"], "julia", "RbqEngine.jl"))
    RbqEngine.@accelerate Model RbqEngine.params_spin13()      # adds the ccall-backed methods for this Model type

`julia/selfcheck.jl` compares the accelerated methods with the un-overridden reference ones when Julia is available.
"""
module RbqEngine

using StaticArrays
using LinearAlgebra: diag
import RobotDynamics
const RD = RobotDynamics

const LIB = get(ENV, "RBQ_LIB", joinpath(@__DIR__, "..", "rbqoc_b200", "librbq.so"))

# ---- include/rbq.h mirrors -----------------------------------------------------------------------
const RBQ_MAX_SAMPLE_STATES = 4
const RBQ_HIST_BINS = 256

struct ModelParams                      # rbq_model_params (128 bytes)
    model_id::Int32
    derivative_order::Int32
    decay_aware::Int32
    time_optimal::Int32
    sample_state_count::Int32
    algo::Int32
    nominal_offset::NTuple{4,Int32}
    integrator::Int32                   # 0 = the model file's exponential step; 2 / 3 / 4 = RobotDynamics RK2 / RK3 / RK4 of RD.dynamics
    reserved::Int32
    fq::Float64
    fq_samples::NTuple{2,Float64}
    namp::Float64
    fq_cov::Float64
    alpha::Float64
    S_diag::NTuple{4,Float64}
end

struct Stats                            # rbq_stats
    count::Int64
    nonfinite::Int64
    sum::Float64
    sumsq::Float64
    min::Float64
    max::Float64
    hist::NTuple{RBQ_HIST_BINS,Int64}
end

const FQ = 1.4e-2
_params(id; DO=0, DA=false, TO=false, ssc=1, nominal=(0, 0, 0, 0), fq=FQ, fq_samples=(0.0, 0.0), namp=0.0,
        fq_cov=0.0, alpha=1.0, S=(1.0, 1.0, 1.0, 1.0), integrator=0) =
    ModelParams(id, DO, DA, TO, ssc, 0, Int32.(nominal), Int32(integrator), Int32(0), fq, fq_samples, namp, fq_cov, alpha, S)

params_spin13() = _params(13)
params_spin12(fq_cov=FQ * 1e-2) = _params(12; fq_samples=(FQ + fq_cov, FQ - fq_cov))            # spin12.jl:204-205
params_spin18(namp) = _params(18; namp=namp)                                                     # spin18.jl:50-52
params_spin15(DA::Bool, TO::Bool) = _params(15; DA=DA, TO=TO)                                     # spin15.jl:40-42
params_spin11(DO::Int) = _params(11; DO=DO)                                                       # spin11.jl:51-52
params_spin17(DO::Int) = _params(17; DO=DO)
# spin30.jl:60-66: nominal_idxs[i][1]-1 is the 0-based offset of sample state i's nominal block
# spin25.jl:67-74: wnamp_cov travels in the fq_cov slot (include/rbq.h)
params_spin25(S, nominal_idxs, namp, wnamp_cov, alpha, ssc) =
    _params(25; ssc=ssc, namp=namp, fq_cov=wnamp_cov, alpha=alpha, S=Tuple(diag(S)),
            nominal=ntuple(i -> i <= ssc ? nominal_idxs[i][1] - 1 : 0, 4))
params_spin30(S, nominal_idxs, fq_cov, alpha, ssc) =
    _params(30; ssc=ssc, fq_cov=fq_cov, alpha=alpha, S=Tuple(diag(S)),
            nominal=ntuple(i -> i <= ssc ? nominal_idxs[i][1] - 1 : 0, 4))

# ---- handle ----------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
end
function Handle(device::Integer=0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:rbq_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, out)
    rc == 0 || error("rbq_create failed with code $rc (no sm_100 device? the engine has no CPU fallback)")
    h = Handle(out[])
    finalizer(x -> ccall((:rbq_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end
const HANDLE = Ref{Union{Nothing,Handle}}(nothing)
handle() = (HANDLE[] === nothing && (HANDLE[] = Handle(0)); HANDLE[]::Handle)

function check(rc::Integer, what)
    rc == 0 && return
    msg = unsafe_string(ccall((:rbq_last_error, LIB), Cstring, (Ptr{Cvoid},), handle().ptr))
    error("$what failed with code $rc: $msg")
end

state_dim(p::ModelParams) = Int(ccall((:rbq_state_dim, LIB), Cint, (Ref{ModelParams},), p))
control_dim(p::ModelParams) = Int(ccall((:rbq_control_dim, LIB), Cint, (Ref{ModelParams},), p))

# ---- per-knot methods ---------------------------------------------------------------------------------
function discrete_dynamics(p::ModelParams, x::AbstractVector{Float64}, u::AbstractVector{Float64}, t::Real, dt::Real)
    n = state_dim(p)
    xv, uv, out = Vector{Float64}(x), Vector{Float64}(u), Vector{Float64}(undef, n)   # SVectors are immutable: copy
    GC.@preserve xv uv out check(ccall((:rbq_discrete_dynamics, LIB), Cint,
        (Ptr{Cvoid}, Ref{ModelParams}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64}),
        handle().ptr, p, xv, uv, t, dt, out), "rbq_discrete_dynamics")
    return SVector{n}(out)
end

"∇f is column-major n×(n+m): exactly the block layout rbq_discrete_jacobian writes."
function discrete_jacobian!(∇f::AbstractMatrix{Float64}, p::ModelParams, x, u, t::Real, dt::Real)
    xv, uv = Vector{Float64}(x), Vector{Float64}(u)
    buf = ∇f isa Matrix{Float64} ? ∇f : Matrix{Float64}(undef, size(∇f)...)
    GC.@preserve xv uv buf check(ccall((:rbq_discrete_jacobian, LIB), Cint,
        (Ptr{Cvoid}, Ref{ModelParams}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64}),
        handle().ptr, p, xv, uv, t, dt, buf), "rbq_discrete_jacobian")
    buf === ∇f || (∇f .= buf)
    return ∇f
end

# ---- trajectory level (what makes the GPU worthwhile: one call per Altro pass, not one per knot) -----------
"X: n×N matrix (column k = knot k, i.e. knot-major in memory), U: m×(N-1), dts: N-1."
function rollout!(X::Matrix{Float64}, p::ModelParams, x0::AbstractVector, U::Matrix{Float64}, dts::Vector{Float64})
    N = size(X, 2)
    x0v = Vector{Float64}(x0)
    GC.@preserve X x0v U dts check(ccall((:rbq_rollout, LIB), Cint,
        (Ptr{Cvoid}, Ref{ModelParams}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        handle().ptr, p, 1, N, x0v, U, dts, X), "rbq_rollout")
    return X
end

"AB: n×(n+m)×(N-1) array; slice k is ∇f_k.  X: n×(N-1) (or n×N), U: m×(N-1), ts: knot times z.t (only spin25 reads them)."
function jacobians!(AB::Array{Float64,3}, p::ModelParams, X::Matrix{Float64}, U::Matrix{Float64}, dts::Vector{Float64},
                    ts::Union{Nothing,Vector{Float64}}=nothing)
    K = size(U, 2)
    tp = ts === nothing ? Ptr{Float64}(C_NULL) : pointer(ts)
    GC.@preserve AB X U dts ts check(ccall((:rbq_jacobians, LIB), Cint,
        (Ptr{Cvoid}, Ref{ModelParams}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        handle().ptr, p, K, X, U, tp, dts, AB), "rbq_jacobians")
    return AB
end

# ---- Monte-Carlo evaluation ------------------------------------------------------------------------------
"Batched run_sim_prop(…; dynamics_type=schroed): fidelities[(gate_count+1) × S] and statistics."
function mc_static(controls::Vector{Float64}, dt::Real, gate::Integer, gate_count::Integer, fq::Vector{Float64},
                   da::Union{Nothing,Vector{Float64}}, psi0::Matrix{Float64}; algo::Integer=0)
    S = length(fq)
    size(psi0) == (4, S) || error("psi0 must be 4×S")
    fid = Matrix{Float64}(undef, gate_count + 1, S)
    st = Ref{Stats}()
    dap = da === nothing ? Ptr{Float64}(C_NULL) : pointer(da)
    GC.@preserve controls fq da psi0 fid check(ccall((:rbq_mc_static, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Cint, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint,
         Ptr{Float64}, Ptr{Stats}),
        handle().ptr, controls, length(controls), dt, gate, gate_count, S, fq, dap, psi0, algo, fid, st), "rbq_mc_static")
    return fid, st[]
end

"Batched run_sim_prop(…; dynamics_type=schroedda): noise is noise_len × S (one column per sample, already scaled by namp)."
function mc_noise(controls::Vector{Float64}, dt::Real, gate::Integer, gate_count::Integer, fq::Real,
                  noise::Matrix{Float64}, noise_dt_inv::Real, psi0::Matrix{Float64}; algo::Integer=0)
    noise_len, S = size(noise)
    size(psi0) == (4, S) || error("psi0 must be 4×S")
    fid = Matrix{Float64}(undef, gate_count + 1, S)
    st = Ref{Stats}()
    GC.@preserve controls noise psi0 fid check(ccall((:rbq_mc_noise, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Cint, Int64, Int64, Float64, Ptr{Float64}, Int64, Float64, Ptr{Float64},
         Cint, Ptr{Float64}, Ptr{Stats}),
        handle().ptr, controls, length(controls), dt, gate, gate_count, S, fq, noise, noise_len, noise_dt_inv, psi0, algo,
        fid, st), "rbq_mc_noise")
    return fid, st[]
end

"Batched run_sim_prop(…; dynamics_type=lindbladt1): rho0 is 2×2×S ComplexF64 (column-major vec per sample, as get_vec_viso)."
function lindblad(controls::Vector{Float64}, dt::Real, gate::Integer, gate_count::Integer, fq::Real,
                  da::Union{Nothing,Vector{Float64}}, rho0::Array{ComplexF64,3}; channels::Integer=1,
                  t2_gamma_c::Real=0.0, t2_gamma_f::Real=0.0)
    S = size(rho0, 3)
    fid = Matrix{Float64}(undef, gate_count + 1, S)
    rho = similar(rho0)
    st = Ref{Stats}()
    dap = da === nothing ? Ptr{Float64}(C_NULL) : pointer(da)
    GC.@preserve controls da rho0 fid rho check(ccall((:rbq_lindblad, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Cint, Int64, Int64, Float64, Ptr{Float64}, Ptr{ComplexF64}, Cint,
         Float64, Float64, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Stats}),
        handle().ptr, controls, length(controls), dt, gate, gate_count, S, fq, dap, rho0, channels, t2_gamma_c, t2_gamma_f,
        fid, rho, st), "rbq_lindblad")
    return fid, rho, st[]
end

"Analytic gates (integrate_prop_zpiby2nodis!/ypiby2nodis/xpiby2nodis, spin.jl:562-570): amps/durs are the pulse's segments."
function mc_static_segments(amps::Vector{Float64}, durs::Vector{Float64}, gate::Integer, gate_count::Integer,
                            fq::Vector{Float64}, da::Union{Nothing,Vector{Float64}}, psi0::Matrix{Float64})
    S = length(fq)
    size(psi0) == (4, S) || error("psi0 must be 4×S")
    fid = Matrix{Float64}(undef, gate_count + 1, S)
    st = Ref{Stats}()
    dap = da === nothing ? Ptr{Float64}(C_NULL) : pointer(da)
    GC.@preserve amps durs fq da psi0 fid check(ccall((:rbq_mc_static_segments, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Stats}),
        handle().ptr, amps, durs, length(amps), gate, gate_count, S, fq, dap, psi0, fid, st), "rbq_mc_static_segments")
    return fid, st[]
end

"integrate_prop_xpiby2da! (spin.jl:819-905): segment j sees amps[j] + noise[noise_idx[j] + 1, s]; noise_idx is 0-based."
function mc_noise_segments(amps::Vector{Float64}, durs::Vector{Float64}, noise_idx::Vector{Int32}, gate::Integer,
                           gate_count::Integer, fq::Real, noise::Matrix{Float64}, psi0::Matrix{Float64})
    noise_len, S = size(noise)
    size(psi0) == (4, S) || error("psi0 must be 4×S")
    fid = Matrix{Float64}(undef, gate_count + 1, S)
    st = Ref{Stats}()
    GC.@preserve amps durs noise_idx noise psi0 fid check(ccall((:rbq_mc_noise_segments, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int64, Cint, Int64, Int64, Float64, Ptr{Float64}, Int64,
         Ptr{Float64}, Ptr{Float64}, Ptr{Stats}),
        handle().ptr, amps, durs, noise_idx, length(amps), gate, gate_count, S, fq, noise, noise_len, psi0, fid, st),
        "rbq_mc_noise_segments")
    return fid, st[]
end

"run_sim_deqjl's ODE solved exactly (spin.jl:424-467, 1277-1380): segment list (amps, durs[, noise_idx]), saves after save_after[g] segments."
function mc_timeline(amps::Vector{Float64}, durs::Vector{Float64}, save_after::Vector{Int64}, gate::Integer, fq::Vector{Float64},
                     psi0::Matrix{Float64}; da::Union{Nothing,Vector{Float64}}=nothing,
                     noise::Union{Nothing,Matrix{Float64}}=nothing, noise_idx::Union{Nothing,Vector{Int32}}=nothing)
    S = length(fq)
    size(psi0) == (4, S) || error("psi0 must be 4×S")
    nsave = length(save_after)
    fid = Matrix{Float64}(undef, nsave + 1, S)
    st = Ref{Stats}()
    dap = da === nothing ? Ptr{Float64}(C_NULL) : pointer(da)
    np = noise === nothing ? Ptr{Float64}(C_NULL) : pointer(noise)
    ip = noise_idx === nothing ? Ptr{Int32}(C_NULL) : pointer(noise_idx)
    nlen = noise === nothing ? 0 : size(noise, 1)
    GC.@preserve amps durs save_after fq psi0 da noise noise_idx fid check(ccall((:rbq_mc_timeline, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int64, Ptr{Int64}, Int64, Cint, Int64, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Stats}),
        handle().ptr, amps, durs, ip, length(amps), save_after, nsave, gate, S, fq, dap, np, nlen, psi0, fid, st), "rbq_mc_timeline")
    return fid, st[]
end

"integrate_prop_zpiby2t1!/ypiby2t1!/xpiby2t1! (spin.jl:582-594, 651-670, 765-785)."
function lindblad_segments(amps::Vector{Float64}, durs::Vector{Float64}, gate::Integer, gate_count::Integer, fq::Real,
                           da::Union{Nothing,Vector{Float64}}, rho0::Array{ComplexF64,3}; channels::Integer=1,
                           t2_gamma_c::Real=0.0)
    S = size(rho0, 3)
    fid = Matrix{Float64}(undef, gate_count + 1, S)
    rho = similar(rho0)
    st = Ref{Stats}()
    dap = da === nothing ? Ptr{Float64}(C_NULL) : pointer(da)
    GC.@preserve amps durs da rho0 fid rho check(ccall((:rbq_lindblad_segments, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Int64, Int64, Float64, Ptr{Float64}, Ptr{ComplexF64}, Cint,
         Float64, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Stats}),
        handle().ptr, amps, durs, length(amps), gate, gate_count, S, fq, dap, rho0, channels, t2_gamma_c, fid, rho, st),
        "rbq_lindblad_segments")
    return fid, rho, st[]
end

"""
iLQR backward pass (Altro `backwardpass!`) for B problems on the device.  AB is n×(n+m)×(N-1)×B exactly as `jacobians!`
fills it; Exx n×n×N×B (symmetric, both triangles stored), Ex n×N×B, Euu m×m×(N-1)×B, Eux m×n×(N-1)×B, Eu m×(N-1)×B.
Returns K (m×n×(N-1)×B), d (m×(N-1)×B), ΔV (2×B) and the failure flags (-1, or the 0-based knot where Quu_reg was not
positive definite: raise ρ and call again, as Altro's regularization_update! does).
"""
function backward_pass(AB::Array{Float64,4}, Exx::Array{Float64,4}, Ex::Array{Float64,3}, Euu::Array{Float64,4},
                       Eux::Array{Float64,4}, Eu::Array{Float64,3}; rho::Real=0.0, state_reg::Bool=false)
    n, N, B = size(Ex)
    m = size(Eu, 1)
    K = Array{Float64,4}(undef, m, n, N - 1, B)
    d = Array{Float64,3}(undef, m, N - 1, B)
    dV = Matrix{Float64}(undef, 2, B)
    failed = Vector{Int64}(undef, B)
    GC.@preserve AB Exx Ex Euu Eux Eu K d dV failed check(ccall((:rbq_backward_pass, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
        handle().ptr, n, m, N, B, AB, Exx, Ex, Euu, Eux, Eu, rho, state_reg ? 1 : 0, K, d, dV, failed), "rbq_backward_pass")
    return K, d, dV, failed
end

"Row of the constraint table of `al_cost_expansion_dev!` (rbq_constraint in include/rbq.h); knot ranges are 0-based, half-open."
struct RbqConstraint
    kind::Int32      # 0 BoundConstraint on [x; u] (±Inf = absent), 1 GoalConstraint, 2 NormConstraint
    p::Int32
    k0::Int64
    k1::Int64
    nidx::Int32
    idx_off::Int32
    val_off::Int64
    dual_off::Int64
end

"""
Augmented-Lagrangian cost expansion of every knot on the device (TrajectoryOptimization `cost_expansion!(E, ::ALObjective, Z)`
for a diagonal LQRObjective plus Bound/Goal/Norm constraints).  All array arguments are device pointers (e.g. `pointer(::CuArray)`
reinterpreted as `Ptr{Cvoid}`), laid out as in include/rbq.h; pass `C_NULL` for `Exx` to evaluate J and c only.
"""
function al_cost_expansion_dev!(n::Integer, m::Integer, N::Integer, B::Integer, X, U, Qdiag, Rdiag, Qfdiag, xf, w, ncon::Integer, cons,
                                idx_pool, val_pool, lambda, mu, dual_stride::Integer, Exx, Ex, Euu, Eux, Eu, c, J)
    check(ccall((:rbq_al_cost_expansion_dev, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint,
         Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
         Ptr{Cvoid}, Ptr{Cvoid}),
        handle().ptr, n, m, N, B, X, U, Qdiag, Rdiag, Qfdiag, xf, w, ncon, cons, idx_pool, val_pool, lambda, mu, dual_stride, Exx, Ex,
        Euu, Eux, Eu, c, J), "rbq_al_cost_expansion_dev")
end

"Multiplier / penalty update and largest violation on the device (Altro dual_update!, penalty_update!, max_violation); device pointers."
al_dual_update_dev!(N::Integer, B::Integer, ncon::Integer, cons, max_p::Integer, c, lambda, mu, dual_stride::Integer, phi::Real, mu_max::Real,
                    lambda_max::Real, update::Bool, max_violation) =
    check(ccall((:rbq_al_dual_update_dev, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Float64, Float64, Float64, Cint, Ptr{Cvoid}),
        handle().ptr, N, B, ncon, cons, max_p, c, lambda, mu, dual_stride, phi, mu_max, lambda_max, update ? 1 : 0, max_violation),
        "rbq_al_dual_update_dev")

"""
Host-array form of the same expansion for one trajectory: X n×N, U m×(N-1), `cons::Vector{RbqConstraint}`, `idx_pool::Vector{Int32}`
(0-based state indices), `val_pool`, `lambda`, `mu` flat in table order.  Returns (J, c, Exx, Ex, Euu, Eux, Eu).
"""
function al_cost_expansion(X::Matrix{Float64}, U::Matrix{Float64}, Qdiag::Vector{Float64}, Rdiag::Vector{Float64},
                           Qfdiag::Vector{Float64}, xf::Vector{Float64}, cons::Vector{RbqConstraint}, idx_pool::Vector{Int32},
                           val_pool::Vector{Float64}, lambda::Vector{Float64}, mu::Vector{Float64}; w=nothing)
    n, N = size(X)
    m = size(U, 1)
    Exx = Array{Float64,3}(undef, n, n, N); Ex = Matrix{Float64}(undef, n, N)
    Euu = Array{Float64,3}(undef, m, m, N - 1); Eux = Array{Float64,3}(undef, m, n, N - 1); Eu = Matrix{Float64}(undef, m, N - 1)
    c = Vector{Float64}(undef, length(lambda)); J = Vector{Float64}(undef, N)
    wp = w === nothing ? Ptr{Float64}(C_NULL) : pointer(w)
    GC.@preserve X U Qdiag Rdiag Qfdiag xf cons idx_pool val_pool lambda mu w Exx Ex Euu Eux Eu c J check(ccall(
        (:rbq_al_cost_expansion, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Cint, Ptr{RbqConstraint}, Ptr{Int32}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64,
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        handle().ptr, n, m, N, 1, X, U, Qdiag, Rdiag, Qfdiag, xf, wp, length(cons), cons, idx_pool, length(idx_pool), val_pool,
        length(val_pool), lambda, mu, 0, Exx, Ex, Euu, Eux, Eu, c, J), "rbq_al_cost_expansion")
    return J, c, Exx, Ex, Euu, Eux, Eu
end

"Page-lock a Julia array in place so that mc_static reads it over PCIe directly (zero-copy); unregister before it is freed."
host_register(A::Array) = check(ccall((:rbq_host_register, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), handle().ptr, A, sizeof(A)), "rbq_host_register")
host_unregister(A::Array) = check(ccall((:rbq_host_unregister, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle().ptr, A), "rbq_host_unregister")

"Performance hint for mc_static: the (qubit frequency, amplitude offset) the explicit samples scatter around (default FQ, 0)."
set_sweep_centre(fq0::Real, da0::Real=0.0) =
    check(ccall((:rbq_set_sweep_centre, LIB), Cint, (Ptr{Cvoid}, Float64, Float64), handle().ptr, fq0, da0), "rbq_set_sweep_centre")

"Batched run_sim_prop(...; seed, state_seed) with the samples drawn on the device (Philox keyed by the global sample index): seeds in, fidelities / statistics out."
function mc_static_philox(controls::Vector{Float64}, dt::Real, gate::Integer, gate_count::Integer, first::Integer, S::Integer,
                          seed::Integer; fq0::Real=FQ, fq_rel_sigma::Real=0.01, fq_rel_clip::Real=0.03, da_sigma::Real=2.574e-5,
                          algo::Integer=0, want_fid::Bool=true)
    fid = want_fid ? Matrix{Float64}(undef, gate_count + 1, S) : nothing
    st = Ref{Stats}()
    fp = want_fid ? pointer(fid) : Ptr{Float64}(C_NULL)
    GC.@preserve controls fid check(ccall((:rbq_mc_static_philox, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Cint, Int64, Int64, Int64, UInt64, Float64, Float64, Float64, Float64, Cint,
         Ptr{Float64}, Ptr{Stats}),
        handle().ptr, controls, length(controls), dt, gate, gate_count, first, S, seed, fq0, fq_rel_sigma, fq_rel_clip, da_sigma, algo,
        fp, st), "rbq_mc_static_philox")
    return fid, st[]
end

"run_sim_deqjl / run_sim_fine_deqjl for the density dynamics (lindbladnodis: channels 0, lindbladt1: 1, lindbladt2: 2) over a segment timeline."
function lindblad_timeline(amps::Vector{Float64}, durs::Vector{Float64}, save_after::Vector{Int64}, gate::Integer, fq::Real,
                           da::Union{Nothing,Vector{Float64}}, rho0::Array{ComplexF64,3}; channels::Integer=1,
                           t2_gamma_c::Real=0.0, t2_gamma_f::Real=0.0)
    S, nsave = size(rho0, 3), length(save_after)
    fid = Matrix{Float64}(undef, nsave + 1, S)
    rho = Array{ComplexF64,4}(undef, 2, 2, nsave + 1, S)
    st = Ref{Stats}()
    dap = da === nothing ? Ptr{Float64}(C_NULL) : pointer(da)
    GC.@preserve amps durs save_after da rho0 fid rho check(ccall((:rbq_lindblad_timeline, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Int64, Cint, Int64, Float64, Ptr{Float64}, Ptr{ComplexF64}, Cint,
         Float64, Float64, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Stats}),
        handle().ptr, amps, durs, length(amps), save_after, nsave, gate, S, fq, dap, rho0, channels, t2_gamma_c, t2_gamma_f, fid, rho, st),
        "rbq_lindblad_timeline")
    return fid, rho, st[]
end

"One tuning / A-B switch of the handle (same names and values as the RBQ_* environment variables latched at creation)."
set_option(name::AbstractString, value::Union{Nothing,AbstractString}=nothing) =
    check(ccall((:rbq_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Cstring), handle().ptr, name, value === nothing ? C_NULL : value),
          "rbq_set_option")

# ---- trajectory-level specialisations: what makes the drop-in FASTER, not just equal -------------------------------------
# A per-knot ccall costs 20-30 us (two synchronous PCIe round trips) -- more than the CPU step it replaces -- so the per-knot
# methods below exist for API completeness; the solver's two hot loops are intercepted whole:
#   TO.dynamics_expansion!(RK3, D, model, Z)   all N-1 Jacobians in ONE rbq_jacobians call           (Altro: every iLQR iteration, PN)
#   RD.rollout!(RK3, model, Z, x0)             the open-loop rollout in ONE rbq_rollout call        (Altro: initial rollout)
# The closed-loop forward pass (Altro `rollout!(solver, alpha)`) stays on the per-knot path unless the caller batches its
# line search through `rollout_closedloop` (include/rbq.h).
# [from memory -- RobotDynamics 0.2.1 / TrajectoryOptimization 0.3.2 @ SchusterLab/rbqoc-stable are not vendored]: a KnotPoint
# has fields z (the concatenated [x; u] StaticVector), _x, _u (index ranges), dt, t; a DynamicsExpansion has the n x (n+m) matrix
# field ∇f and views A_, B_ of it plus tmpA, tmpB.  julia/selfcheck.jl asserts these names before anything else runs.
const KNOT_FIELDS = (:z, :_x, :_u, :dt, :t)
const EXPANSION_FIELDS = (:∇f, :A_, :B_, :tmpA, :tmpB)

mutable struct TrajBuffers              # page-locked is not needed here: plain Julia arrays, reused across iterations
    n::Int; m::Int; N::Int
    X::Matrix{Float64}; U::Matrix{Float64}; dts::Vector{Float64}; ts::Vector{Float64}; AB::Array{Float64,3}
end
const BUFFERS = Dict{Tuple{Int,Int,Int},TrajBuffers}()
function buffers(n, m, N)
    get!(BUFFERS, (n, m, N)) do
        TrajBuffers(n, m, N, zeros(n, N), zeros(m, N - 1), zeros(N - 1), zeros(N - 1), zeros(n, n + m, N - 1))
    end
end

"Gather a TO.Traj (vector of KnotPoints) into the engine's knot-major arrays."
function gather!(b::TrajBuffers, Z)
    for k = 1:b.N
        zk = Z[k]
        b.X[:, k] .= RD.state(zk)
        if k < b.N
            b.U[:, k] .= RD.control(zk)
            b.dts[k] = zk.dt
            b.ts[k] = zk.t
        end
    end
    return b
end

"All knot-point Jacobians of a trajectory in one engine call, scattered into the solver's DynamicsExpansion objects."
function dynamics_expansion!(D::AbstractVector, p::ModelParams, Z)
    N = length(Z)
    n, m = state_dim(p), control_dim(p)
    b = gather!(buffers(n, m, N), Z)
    jacobians!(b.AB, p, b.X, b.U, b.dts, b.ts)
    for k in eachindex(D)
        D[k].∇f .= view(b.AB, :, :, k)
        D[k].tmpA .= D[k].A_      # what TO.dynamics_expansion! does after every discrete_jacobian! (error-state expansion)
        D[k].tmpB .= D[k].B_
    end
    return nothing
end

"Open-loop rollout of a trajectory in one engine call: Z[k+1].x = discrete_dynamics(RK3, model, Z[k])."
function rollout_traj!(p::ModelParams, Z, x0)
    N = length(Z)
    n, m = state_dim(p), control_dim(p)
    b = gather!(buffers(n, m, N), Z)
    rollout!(b.X, p, x0, b.U, b.dts)
    for k = 1:N
        RD.set_state!(Z[k], SVector{n}(view(b.X, :, k)))
    end
    return nothing
end

# ---- the macro a reference script uses ---------------------------------------------------------------------
"""
    RbqEngine.@accelerate Model params_expr
    RbqEngine.@accelerate Model{DA,TO} where {DA,TO} params_expr

Adds, for the script's own `Model` type, methods that forward to the engine:
`RD.discrete_dynamics(::Type{RD.RK3}, ::Model, x, u, t, dt)` for Float64 arguments (dual-number calls from ForwardDiff
still reach the script's generic method), `RD.discrete_jacobian!(::Type{RD.RK3}, ∇f, ::Model, z)`, and the two
trajectory-level loops `TO.dynamics_expansion!(RK3, D, ::Model, Z)` / `RD.rollout!(RK3, ::Model, Z, x0)`.
`params_expr` is evaluated inside each method with the method's own `model` argument in scope (the argument is
deliberately NOT hygienic), e.g. `RbqEngine.params_spin30(model.S, model.nominal_idxs, model.fq_cov, model.alpha,
model.sample_state_count)`; type parameters named in the `where` clause are in scope too
(`RbqEngine.params_spin15(DA, TO)`).
"""
macro accelerate(M, pexpr)
    wh = Any[]
    if M isa Expr && M.head == :where          # `Model{DA,TO} where {DA,TO}` arrives as one :where expression
        wh = M.args[2:end]
        M = M.args[1]
    end
    model = esc(:model)                         # the method argument is visible to params_expr on purpose
    MT = esc(M)
    whs = map(esc, wh)
    P = esc(pexpr)
    TOexp = esc(:(TO.dynamics_expansion!))      # the reference scripts bind `const TO = TrajectoryOptimization` (spin13.jl:15)
    quote
        function RobotDynamics.discrete_dynamics(::Type{RobotDynamics.RK3}, $model::$MT,
                                                 x::StaticVector{<:Any,Float64}, u::StaticVector{<:Any,Float64},
                                                 t::Real, dt::Real) where {$(whs...)}
            return RbqEngine.discrete_dynamics($P, x, u, t, dt)
        end
        function RobotDynamics.discrete_jacobian!(::Type{RobotDynamics.RK3}, ∇f, $model::$MT,
                                                  z::RobotDynamics.AbstractKnotPoint{Float64}) where {$(whs...)}
            return RbqEngine.discrete_jacobian!(∇f, $P, RobotDynamics.state(z), RobotDynamics.control(z), z.t, z.dt)
        end
        function RobotDynamics.rollout!(::Type{RobotDynamics.RK3}, $model::$MT,
                                        Z::AbstractVector{<:RobotDynamics.AbstractKnotPoint{Float64}}, x0) where {$(whs...)}
            return RbqEngine.rollout_traj!($P, Z, x0)
        end
        function $TOexp(::Type{RobotDynamics.RK3}, D::AbstractVector, $model::$MT,
                        Z::AbstractVector{<:RobotDynamics.AbstractKnotPoint{Float64}}) where {$(whs...)}
            return RbqEngine.dynamics_expansion!(D, $P, Z)
        end
    end
end

end # module
