"""TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header): numpy restatement of the augmented-Lagrangian cost expansion
that Altro / TrajectoryOptimization perform per knot in `cost_expansion!(E, ::ALObjective, Z)` for the reference's problem
class (spin13.jl:122-153: diagonal LQRObjective + BoundConstraint + GoalConstraint + NormConstraint).

The solver packages are external and un-vendored (SchusterLab forks pinned in Manifest.toml:26-32, 1523-1529), so this
follows their published algorithm -- PARITY UNPINNED against the fork.  What is pinned here (tests/test_oracle_cpu.py): the
gradient against finite differences of the merit function it states, and the Gauss-Newton Hessian against the exact one
for the constraints that are linear.

    J   = sum_k w_k l_k(x_k, u_k) + sum_con lambda'c + 1/2 c' I_mu c,      I_mu = diag(a .* mu)
    a_j = 1 (equality) | (c_j >= 0 or lambda_j > 0) (inequality)
    Exx += cx' I_mu cx,  Ex += cx'(lambda + I_mu c)

Constraints are dicts: {"kind": "bound", "k0", "k1", "zmax", "zmin"} | {"kind": "goal", "k0", "k1", "idx", "target"} |
{"kind": "norm", "k0", "k1", "idx", "val"}; duals[i] = (lambda, mu) arrays of shape (k1 - k0, p).
"""
import numpy as np


def con_len(con, n, m):
    return {"bound": 2 * (n + m), "goal": len(con.get("idx", ())), "norm": 1}[con["kind"]]


def evaluate(con, x, u, n, m):
    """c and its Jacobians (cx, cu) at one knot; u is None at the last knot."""
    kind = con["kind"]
    if kind == "bound":
        nz = n + m
        z = np.concatenate([x, u if u is not None else np.full(m, np.nan)])
        c = np.concatenate([z - con["zmax"], con["zmin"] - z])
        if u is None:
            c[n:nz] = -np.inf
            c[nz + n:] = -np.inf
        J = np.vstack([np.eye(nz), -np.eye(nz)])
        return c, J[:, :n], J[:, n:]
    if kind == "goal":
        idx = np.asarray(con["idx"])
        cx = np.zeros((len(idx), n))
        cx[np.arange(len(idx)), idx] = 1.0
        return x[idx] - np.asarray(con["target"]), cx, np.zeros((len(idx), m))
    if kind == "norm":
        idx = np.asarray(con["idx"])
        cx = np.zeros((1, n))
        cx[0, idx] = 2.0 * x[idx]
        return np.array([x[idx] @ x[idx] - con["val"]]), cx, np.zeros((1, m))
    raise ValueError(kind)


def expansion(X, U, Qd, Rd, Qfd, xf, cons, duals, w=None):
    """Returns J (N,), Exx (N, n, n), Ex (N, n), Euu (N-1, m, m), Eux (N-1, m, n), Eu (N-1, m), c (list of (k1-k0, p))."""
    N, n = X.shape
    m = U.shape[1]
    w = np.ones(N) if w is None else np.asarray(w, dtype=float)
    J = np.zeros(N)
    Exx = np.zeros((N, n, n)); Ex = np.zeros((N, n))
    Euu = np.zeros((N - 1, m, m)); Eux = np.zeros((N - 1, m, n)); Eu = np.zeros((N - 1, m))
    cvals = [np.zeros((c["k1"] - c["k0"], con_len(c, n, m))) for c in cons]
    for k in range(N):
        term = k == N - 1
        q = w[k] * (Qfd if term else Qd)
        dx = X[k] - xf
        J[k] = 0.5 * np.sum(q * dx * dx)
        Exx[k] = np.diag(q); Ex[k] = q * dx
        if not term:
            r = w[k] * Rd
            J[k] += 0.5 * np.sum(r * U[k] * U[k])
            Euu[k] = np.diag(r); Eu[k] = r * U[k]
        for ci, con in enumerate(cons):
            if not (con["k0"] <= k < con["k1"]):
                continue
            lam, mu = duals[ci][0][k - con["k0"]], duals[ci][1][k - con["k0"]]
            c, cx, cu = evaluate(con, X[k], None if term else U[k], n, m)
            cvals[ci][k - con["k0"]] = c
            present = np.isfinite(c)
            act = present & ((c >= 0) | (lam > 0)) if con["kind"] == "bound" else present
            cs = np.where(present, c, 0.0)
            imu = np.where(act, mu, 0.0)
            lam_p = np.where(present, lam, 0.0)
            J[k] += lam_p @ cs + 0.5 * np.sum(imu * cs * cs)
            g = lam_p + imu * cs
            Exx[k] += cx.T @ (imu[:, None] * cx); Ex[k] += cx.T @ g
            if not term:
                Euu[k] += cu.T @ (imu[:, None] * cu); Eu[k] += cu.T @ g
                Eux[k] += cu.T @ (imu[:, None] * cx)
    return J, Exx, Ex, Euu, Eux, Eu, cvals
