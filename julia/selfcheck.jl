# selfcheck.jl — run on a box that has Julia, the reference checkout (ROBUST_QOC_PATH) and a B200:
#   julia julia/selfcheck.jl
# Compares the ccall-backed methods with the reference's own (un-overridden) Julia methods on random knot points.
# Not executed in the build container (no Julia there).
include(joinpath(ENV["ROBUST_QOC_PATH"], "src", "spin", "spin13.jl"))
include(joinpath(@__DIR__, "RbqEngine.jl"))
using ForwardDiff, LinearAlgebra, Random

model = Model()
n, m = RD.state_dim(model), RD.control_dim(model)
p = RbqEngine.params_spin13()
Random.seed!(0)
worst_x, worst_j = 0.0, 0.0
for trial = 1:100
    x = SVector{n}(randn(n) .* 0.3)
    u = SVector{m}(randn(m) .* 0.1)
    dt = rand([0.1, 0.01])
    ref = RD.discrete_dynamics(RK3, model, x, u, 0.0, dt)
    got = RbqEngine.discrete_dynamics(p, x, u, 0.0, dt)
    global worst_x = max(worst_x, maximum(abs.(ref .- got)) / maximum(abs.(ref)))
    J = ForwardDiff.jacobian(s -> RD.discrete_dynamics(RK3, model, s[SVector{n}(1:n)], s[SVector{m}(n+1:n+m)], 0.0, dt), [x; u])
    G = RbqEngine.discrete_jacobian!(zeros(n, n + m), p, x, u, 0.0, dt)
    global worst_j = max(worst_j, maximum(abs.(J .- G)) / maximum(abs.(J)))
end
println("spin13: worst state rel err = $worst_x (tol 1e-12), worst Jacobian rel err = $worst_j (tol 1e-10)")
@assert worst_x < 1e-12 && worst_j < 1e-10
