"""CPU oracle (test infrastructure only) for the iLQR backward pass.

Restates Altro.jl 0.2.0 `backwardpass!` (ilqr/backwardpass.jl of the SchusterLab/rbqoc-stable fork pinned at
Manifest.toml:26-32 -- an un-vendored dependency, absent from /root/reference; restated from its published algorithm, the
iLQR recursion of Tassa et al. 2012 as Altro implements it) with plain numpy, one problem at a time:

    S_N = E_N.Q, s_N = E_N.q
    for k = N-1 .. 1:
        Qx = E.q + A's      Qu = E.r + B's
        Qxx = E.Q + A'SA    Quu = E.R + B'SB    Qux = E.H + B'SA
        state regularisation:   Quu_reg = Quu + rho B'B,  Qux_reg = Qux + rho B'A
        control regularisation: Quu_reg = Quu + rho I,    Qux_reg = Qux
        not positive definite -> stop (the caller raises rho and restarts)
        K = -Quu_reg \\ Qux_reg,  d = -Quu_reg \\ Qu
        s = Qx + K'Quu d + K'Qu + Qux'd
        S = Qxx + K'Quu K + K'Qux + Qux'K;  S = (S + S')/2
        dV += [d'Qu, d'Quu d / 2]

"parity unpinned": no output of the Julia package is available here; the recursion is pinned by its defining property (for a
linear-quadratic problem one backward + one forward pass lands on the optimum, tests/test_oracle_cpu.py).
"""
import numpy as np

REG_CONTROL, REG_STATE = 0, 1


def backward_pass(AB, Exx, Ex, Euu, Eux, Eu, rho=0.0, reg_type=REG_CONTROL):
    """AB (N-1, n, n+m) = [A_k B_k], Exx (N, n, n), Ex (N, n), Euu (N-1, m, m), Eux (N-1, m, n), Eu (N-1, m).
    Returns K (N-1, m, n), d (N-1, m), dV (2,), fail (-1 or the 0-based knot)."""
    N, n = Ex.shape
    m = Eu.shape[1] if N > 1 else 1
    K = np.zeros((max(N - 1, 0), m, n))
    d = np.zeros((max(N - 1, 0), m))
    dV = np.zeros(2)
    S = Exx[N - 1].copy()
    S = np.tril(S) + np.tril(S, -1).T          # the engine reads the lower triangle
    s = Ex[N - 1].copy()
    for k in range(N - 2, -1, -1):
        M = AB[k]
        A, B = M[:, :n], M[:, n:]
        Q = Exx[k]
        Q = np.tril(Q) + np.tril(Q, -1).T
        R = Euu[k]
        R = np.tril(R) + np.tril(R, -1).T
        H = Eux[k]
        Qx = Ex[k] + A.T @ s
        Qu = Eu[k] + B.T @ s
        Qxx = Q + A.T @ S @ A
        Quu = R + B.T @ S @ B
        Qux = H + B.T @ S @ A
        if reg_type == REG_STATE:
            Quu_reg = Quu + rho * (B.T @ B)
            Qux_reg = Qux + rho * (B.T @ A)
        else:
            Quu_reg = Quu + rho * np.eye(m)
            Qux_reg = Qux
        try:
            if not np.all(np.isfinite(Quu_reg)):
                raise np.linalg.LinAlgError
            L = np.linalg.cholesky(Quu_reg)
        except np.linalg.LinAlgError:
            return K, d, dV, k
        Kk = -np.linalg.solve(L.T, np.linalg.solve(L, Qux_reg))
        dk = -np.linalg.solve(L.T, np.linalg.solve(L, Qu))
        s = Qx + Kk.T @ (Quu @ dk) + Kk.T @ Qu + Qux.T @ dk
        S = Qxx + Kk.T @ Quu @ Kk + Kk.T @ Qux + Qux.T @ Kk
        S = 0.5 * (S + S.T)
        dV += [dk @ Qu, 0.5 * dk @ Quu @ dk]
        K[k] = Kk
        d[k] = dk
    return K, d, dV, -1
