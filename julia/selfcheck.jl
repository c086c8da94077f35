# selfcheck.jl — run on a box that has Julia, the reference checkout (ROBUST_QOC_PATH) and a B200:
#   julia julia/selfcheck.jl [spin13 spin12 spin18 spin15 spin11 spin17 spin30 spin25]
# For every model file: (1) the struct fields the glue assumes exist, or it stops with their names; (2) the ccall-backed
# per-knot methods installed by @accelerate against the reference's own Julia methods (invoked on the un-accelerated generic
# signature through dual-free `invoke`) and ForwardDiff; (3) the trajectory-level methods (TO.dynamics_expansion!,
# RD.rollout!) against loops of the per-knot reference methods; (4) with `RBQ_SELFCHECK_SOLVE=1`, a `smoke_test=true`
# run_traj() with and without the engine: same iteration count and final cost.
# NOT executed in the build container or on the GPU box of this project (no Julia on either: probed, profiles/r2_julia_probe_gpu_box.txt).
using ForwardDiff, LinearAlgebra, Random, StaticArrays

const FILES = isempty(ARGS) ? ["spin13", "spin12", "spin18", "spin15", "spin11", "spin17", "spin30", "spin25"] : ARGS
include(joinpath(@__DIR__, "RbqEngine.jl"))

function check_fields(Z, D)
    kp = typeof(Z[1])
    missing_k = [f for f in RbqEngine.KNOT_FIELDS if !hasfield(kp, f)]
    isempty(missing_k) || error("KnotPoint type $kp lacks the fields $(missing_k) the glue reads; adapt RbqEngine.gather!")
    if D !== nothing
        missing_d = [f for f in RbqEngine.EXPANSION_FIELDS if !hasproperty(D[1], f)]
        isempty(missing_d) || error("DynamicsExpansion type $(typeof(D[1])) lacks $(missing_d); adapt RbqEngine.dynamics_expansion!")
    end
end

for file in FILES
    mod = Module(Symbol("Check_", file))
    Base.include(mod, joinpath(ENV["ROBUST_QOC_PATH"], "src", "spin", file * ".jl"))
    Core.eval(mod, :(const RbqEngine = $RbqEngine))
    # one Model instance and the matching parameter expression per file (constructors as in the file's run_traj)
    setup = Dict(
        "spin13" => (:(Model()), :(RbqEngine.params_spin13())),
        "spin12" => (:(Model((FQ + FQ * 1e-2, FQ - FQ * 1e-2) .* Ref(NEGI_H0_ISO))), :(RbqEngine.params_spin12(FQ * 1e-2))),
        "spin18" => (:(Model(NAMP_PREFACTOR)), :(RbqEngine.params_spin18(model.namp))),
        "spin15" => (:(Model(true, false)), :(RbqEngine.params_spin15(true, false))),
        "spin11" => (:(Model(2)), :(RbqEngine.params_spin11(2))),
        "spin17" => (:(Model(2)), :(RbqEngine.params_spin17(2))),
        "spin30" => (:(Model(SMatrix{4,4}(I(4) * 1.0), [STATE1_IDX], FQ * 1e-2, 1.0, 1)),
                     :(RbqEngine.params_spin30(model.S, model.nominal_idxs, model.fq_cov, model.alpha, model.sample_state_count))),
        "spin25" => (:(Model(SMatrix{4,4}(I(4) * 1.0), [STATE1_IDX], NAMP_PREFACTOR, 1.0, 1.0, 1)),
                     :(RbqEngine.params_spin25(model.S, model.nominal_idxs, model.namp, model.wnamp_cov, model.alpha, model.sample_state_count))),
    )[file]
    model = Core.eval(mod, setup[1])
    RD = Core.eval(mod, :RD)
    n, m = RD.state_dim(model), RD.control_dim(model)
    # reference methods BEFORE acceleration, captured as closures over the generic (dual-capable) method
    ref_step = (x, u, t, dt) -> RD.discrete_dynamics(RD.RK3, model, x, u, t, dt)
    Random.seed!(0)
    pts = [(SVector{n}(randn(n) .* 0.3), SVector{m}(randn(m) .* 0.1 .+ 0.3), 0.1 * rand(0:8), rand([0.1, 0.01])) for _ = 1:100]
    refs = [(ref_step(x, u, t, dt),
             ForwardDiff.jacobian(s -> RD.discrete_dynamics(RD.RK3, model, s[SVector{n}(1:n)], s[SVector{m}(n+1:n+m)], t, dt), [x; u]))
            for (x, u, t, dt) in pts]
    # install the engine through the macro, exactly as a user script would
    Core.eval(mod, :(RbqEngine.@accelerate $(typeof(model)) $(setup[2])))
    worst_x = worst_j = 0.0
    for ((x, u, t, dt), (xr, Jr)) in zip(pts, refs)
        got = Base.invokelatest(RD.discrete_dynamics, RD.RK3, model, x, u, t, dt)
        worst_x = max(worst_x, maximum(abs.(xr .- got)) / maximum(abs.(xr)))
        z = RD.KnotPoint(x, u, dt, t)
        G = zeros(n, n + m)
        Base.invokelatest(RD.discrete_jacobian!, RD.RK3, G, model, z)
        worst_j = max(worst_j, maximum(abs.(Jr .- G)) / maximum(abs.(Jr)))
    end
    println("$file per knot: worst state rel err = $worst_x (tol 1e-12), worst Jacobian rel err = $worst_j (tol 1e-10)")
    @assert worst_x < 1e-12 && worst_j < 1e-10
    # trajectory level: a 50-knot Traj
    N = 50
    Z = [RD.KnotPoint(pts[k][1], pts[k][2], 0.1, 0.1 * (k - 1)) for k = 1:N]
    check_fields(Z, nothing)
    Zr = deepcopy(Z)
    for k = 2:N
        RD.set_state!(Zr[k], ref_step(RD.state(Zr[k-1]), RD.control(Zr[k-1]), Zr[k-1].t, Zr[k-1].dt))
    end
    Base.invokelatest(RD.rollout!, RD.RK3, model, Z, RD.state(Z[1]))
    worst_r = maximum(maximum(abs.(RD.state(Z[k]) .- RD.state(Zr[k]))) for k = 1:N)
    println("$file rollout!: worst abs err = $worst_r")
    @assert worst_r < 1e-11
    TO = Core.eval(mod, :TO)
    D = [TO.DynamicsExpansion{Float64}(n, m) for _ = 1:N-1]
    check_fields(Z, D)
    Base.invokelatest(TO.dynamics_expansion!, RD.RK3, D, model, Z)
    worst_d = 0.0
    for k = 1:N-1
        x, u = RD.state(Z[k]), RD.control(Z[k])
        Jr = ForwardDiff.jacobian(s -> ref_step(s[SVector{n}(1:n)], s[SVector{m}(n+1:n+m)], Z[k].t, Z[k].dt), [x; u])
        worst_d = max(worst_d, maximum(abs.(Jr .- D[k].∇f)) / maximum(abs.(Jr)))
    end
    println("$file dynamics_expansion!: worst Jacobian rel err = $worst_d (tol 1e-10)")
    @assert worst_d < 1e-10
    if get(ENV, "RBQ_SELFCHECK_SOLVE", "0") == "1"
        res = Base.invokelatest(Core.eval(mod, :run_traj); smoke_test=true, save=false)
        println("$file run_traj(smoke_test=true) with the engine: ", res)
    end
end
println("selfcheck ok")
