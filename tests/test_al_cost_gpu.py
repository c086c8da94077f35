"""GPU parity: rbq_al_cost_expansion_dev against the numpy restatement (oracle/al_cost.py) of the solver's
cost_expansion!(E, ::ALObjective, Z).  Tolerance: relative 1e-12 (the same few products and sums per entry, summed in a
different order only for the knot cost and the norm constraint)."""
import numpy as np
import pytest
import torch

import rbqoc_b200 as rbq
from rbqoc_b200 import spin
from test_oracle_cpu import _al_problem

pytestmark = pytest.mark.gpu


def _run(eng, dev, X, U, Qd, Rd, Qfd, xf, w, table, lam, mu, expand=True):
    N, n = X.shape
    m = U.shape[1]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    z = lambda *s: torch.full(s, float("nan"), dtype=torch.float64, device=dev)
    table.upload(dev)
    Exx, Ex, Euu, Eux, Eu = (z(N, n, n), z(N, n), z(N - 1, m, m), z(N - 1, n, m), z(N - 1, m)) if expand else (None,) * 5
    c, J = z(table.dual_size), z(N)
    eng.al_cost_expansion_dev(n, m, N, d(X), d(U), d(Qd), d(Rd), d(Qfd), d(xf), None if w is None else d(w), table, d(lam), d(mu),
                              Exx, Ex, Euu, Eux, Eu, c, J)
    torch.cuda.synchronize()
    host = lambda t: None if t is None else t.cpu().numpy()
    return tuple(host(t) for t in (J, Exx, Ex, Euu, Eux, Eu, c))


@pytest.mark.parametrize("seed,N,n,m", [(1, 9, 11, 1), (2, 40, 11, 1), (3, 6, 11, 1)])
def test_al_cost_expansion_matches_oracle(seed, N, n, m):
    from oracle import al_cost

    dev = torch.device("cuda", 0)
    eng = rbq.Engine(0); eng.use_torch_stream()
    rng = np.random.default_rng(seed)
    X, U, Qd, Rd, Qfd, xf, w, table, lam, mu, duals = _al_problem(rng, N, n, m)
    Jr, Exxr, Exr, Euur, Euxr, Eur, cr = al_cost.expansion(X, U, Qd, Rd, Qfd, xf, table.specs, duals, w)
    J, Exx, Ex, Euu, Eux, Eu, c = _run(eng, dev, X, U, Qd, Rd, Qfd, xf, w, table, lam, mu)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    assert rel(J, Jr) < 1e-12 and rel(Ex, Exr) < 1e-12 and rel(Eu, Eur) < 1e-12
    # blocks are column-major on the device; Exx and Euu are symmetric, Eux (m x n column-major) is zero
    assert rel(np.swapaxes(Exx, -1, -2), Exxr) < 1e-12 and rel(Euu, Euur) < 1e-12 and np.all(Eux == 0)
    for (o, kn, p), ref in zip(table.dual_slices(), cr):
        got = c[o:o + kn * p].reshape(kn, p)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and rel(got[fin], ref[fin]) < 1e-12
    # values only (merit function of a line-search candidate): same J and c, no blocks touched
    J2, *_, c2 = _run(eng, dev, X, U, Qd, Rd, Qfd, xf, w, table, lam, mu, expand=False)
    assert np.array_equal(J2, J) and np.array_equal(c2, c, equal_nan=True)
    # unit weights when w is NULL
    J3 = _run(eng, dev, X, U, Qd, Rd, Qfd, xf, None, table, lam, mu)[0]
    assert rel(J3, al_cost.expansion(X, U, Qd, Rd, Qfd, xf, table.specs, duals, None)[0]) < 1e-12


def test_al_cost_blocks_feed_the_backward_pass_on_device():
    """cost expansion -> Jacobians -> backward pass without leaving the device: the gains equal those of the host path fed
    with the oracle's blocks."""
    from oracle import al_cost, riccati

    dev = torch.device("cuda", 0)
    eng = rbq.Engine(0); eng.use_torch_stream()
    model = spin.Spin13Model()
    n, m, N = model.n, model.m, 61
    dtv = np.full(N - 1, 0.1)
    U = 0.05 * np.sin(np.arange(N - 1) * 0.3)[:, None]
    X = eng.rollout(model.params, spin.spin13_x0(), U, dtv)
    xf = np.zeros(n); xf[:8] = X[-1, :8]
    table = spin.spin13_constraints(N, xf)
    rng = np.random.default_rng(4)
    lam = 0.01 * rng.standard_normal(table.dual_size); mu = np.full(table.dual_size, 10.0)
    Qd = np.array([1.0] * 8 + [1.0, 1.0, 0.1]); Rd = np.array([0.1]); Qfd = Qd * N
    duals = [(lam[o:o + k * p].reshape(k, p), mu[o:o + k * p].reshape(k, p)) for o, k, p in table.dual_slices()]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    table.upload(dev)
    dX, dU, ddt = d(X), d(U), d(dtv)
    Exx = torch.empty(N, n * n, dtype=torch.float64, device=dev); Ex = torch.empty(N, n, dtype=torch.float64, device=dev)
    Euu = torch.empty(N - 1, m * m, dtype=torch.float64, device=dev); Eux = torch.empty(N - 1, m * n, dtype=torch.float64, device=dev)
    Eu = torch.empty(N - 1, m, dtype=torch.float64, device=dev)
    c = torch.empty(table.dual_size, dtype=torch.float64, device=dev); J = torch.empty(N, dtype=torch.float64, device=dev)
    AB = torch.empty(N - 1, n * (n + m), dtype=torch.float64, device=dev)
    K = torch.empty(N - 1, n * m, dtype=torch.float64, device=dev); dd = torch.empty(N - 1, m, dtype=torch.float64, device=dev)
    dV = torch.empty(2, dtype=torch.float64, device=dev); fl = torch.empty(1, dtype=torch.int64, device=dev)
    eng.al_cost_expansion_dev(n, m, N, dX, dU, d(Qd), d(Rd), d(Qfd), d(xf), None, table, d(lam), d(mu), Exx, Ex, Euu, Eux, Eu, c, J)
    eng.jacobians_dev(model.params, N - 1, dX, dU, ddt, AB)
    eng.backward_pass_dev(n, m, N, 1, AB, Exx, Ex, Euu, Eux, Eu, 1e-3, rbq.REG_CONTROL, K, dd, dV, fl)
    torch.cuda.synchronize()
    assert int(fl[0]) == -1
    Jr, Exxr, Exr, Euur, Euxr, Eur, cr = al_cost.expansion(X, U, Qd, Rd, Qfd, xf, table.specs, duals, None)
    ABh = eng.jacobians(model.params, X[:-1], U, dtv)
    Kr, dr, dVr, fr = riccati.backward_pass(ABh, Exxr, Exr, Euur, Euxr, Eur, rho=1e-3, reg_type=rbq.REG_CONTROL)
    Kh = K.cpu().numpy().reshape(N - 1, n, m).swapaxes(-1, -2)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    assert rel(Kh, Kr) < 1e-9 and rel(dd.cpu().numpy(), dr) < 1e-9 and np.allclose(dV.cpu().numpy(), dVr, rtol=1e-9)
    assert abs(float(J.sum()) - Jr.sum()) < 1e-12 * abs(Jr.sum())


def test_al_cost_expansion_batch_equals_single_calls():
    """B trajectories in one launch: shared multipliers (line-search candidates) and per-trajectory multipliers
    (independent problems) both reproduce the single-trajectory results bit for bit."""
    dev = torch.device("cuda", 0)
    eng = rbq.Engine(0); eng.use_torch_stream()
    rng = np.random.default_rng(21)
    X, U, Qd, Rd, Qfd, xf, w, table, lam, mu, duals = _al_problem(rng, 12, 11, 1)
    B = 3
    Xs = np.stack([X + 0.05 * b for b in range(B)]); Us = np.stack([U * (1 + b) for b in range(B)])
    lams = np.stack([lam * (1 + 0.5 * b) for b in range(B)]); mus = np.stack([mu * (1 + b) for b in range(B)])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    z = lambda *s: torch.full(s, float("nan"), dtype=torch.float64, device=dev)
    table.upload(dev)
    N, n, m = 12, 11, 1
    for shared in (True, False):
        out = [z(B, N, n, n), z(B, N, n), z(B, N - 1, m, m), z(B, N - 1, n, m), z(B, N - 1, m), z(B, table.dual_size), z(B, N)]
        eng.al_cost_expansion_dev(n, m, N, d(Xs), d(Us), d(Qd), d(Rd), d(Qfd), d(xf), d(w), table, d(lam if shared else lams),
                                  d(mu if shared else mus), *out, B=B, dual_stride=0 if shared else table.dual_size)
        torch.cuda.synchronize()
        for b in range(B):
            single = _run(eng, dev, Xs[b], Us[b], Qd, Rd, Qfd, xf, w, table, lam if shared else lams[b], mu if shared else mus[b])
            J1, Exx1, Ex1, Euu1, Eux1, Eu1, c1 = single
            got = [t[b].cpu().numpy() for t in out]
            assert np.array_equal(got[6], J1) and np.array_equal(got[0], Exx1) and np.array_equal(got[1], Ex1)
            assert np.array_equal(got[2], Euu1) and np.array_equal(got[4], Eu1) and np.array_equal(got[5], c1)


def test_al_cost_expansion_host_pointer_form(engine):
    """rbq_al_cost_expansion (host arrays in, host arrays out) returns what the oracle states; batch with per-trajectory
    multipliers; values-only mode; argument errors come back as codes, not crashes."""
    from oracle import al_cost

    rng = np.random.default_rng(31)
    X, U, Qd, Rd, Qfd, xf, w, table, lam, mu, duals = _al_problem(rng, 10, 11, 1)
    Jr, Exxr, Exr, Euur, Euxr, Eur, cr = al_cost.expansion(X, U, Qd, Rd, Qfd, xf, table.specs, duals, w)
    got = engine.al_cost_expansion(X, U, Qd, Rd, Qfd, xf, table, lam, mu, w=w)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    assert rel(got["J"], Jr) < 1e-12 and rel(got["Exx"], Exxr) < 1e-12 and rel(got["Ex"], Exr) < 1e-12
    assert rel(got["Euu"], Euur) < 1e-12 and rel(got["Eu"], Eur) < 1e-12 and np.all(got["Eux"] == 0)
    cflat = np.concatenate([c.reshape(-1) for c in cr])
    fin = np.isfinite(cflat)
    assert np.array_equal(np.isneginf(got["c"]), np.isneginf(cflat)) and rel(got["c"][fin], cflat[fin]) < 1e-12
    vals = engine.al_cost_expansion(X, U, Qd, Rd, Qfd, xf, table, lam, mu, w=w, expand=False)
    assert np.array_equal(vals["J"], got["J"]) and "Exx" not in vals
    # two independent problems with their own multipliers
    Xs, Us = np.stack([X, X * 0.9]), np.stack([U, -U])
    lams, mus = np.stack([lam, 2 * lam]), np.stack([mu, 3 * mu])
    both = engine.al_cost_expansion(Xs, Us, Qd, Rd, Qfd, xf, table, lams, mus, w=w, shared_duals=False)
    one = engine.al_cost_expansion(Xs[1], Us[1], Qd, Rd, Qfd, xf, table, lams[1], mus[1], w=w)
    assert np.array_equal(both["J"][0], got["J"]) and np.array_equal(both["J"][1], one["J"]) and np.array_equal(both["Exx"][1], one["Exx"])
    # a table whose knot range leaves the trajectory is refused
    bad = spin.ConstraintTable(11, 1, 10).norm(0, 10, range(4))
    bad.rows[0].k1 = 11
    with pytest.raises(rbq.RbqError):
        engine.al_cost_expansion(X, U, Qd, Rd, Qfd, xf, bad, np.zeros(11), np.ones(11))


def test_al_dual_update_matches_numpy():
    """lambda <- clamp(lambda + mu c) (projected at 0 for bounds), mu <- min(phi mu, mu_max), largest violation; absent bounds
    (-inf) untouched; violation-only mode leaves the duals alone; clamps at lambda_max."""
    dev = torch.device("cuda", 0)
    eng = rbq.Engine(0); eng.use_torch_stream()
    rng = np.random.default_rng(41)
    N = 57
    table = spin.spin13_constraints(N, rng.standard_normal(11)).upload(dev)
    D = table.dual_size
    c = rng.standard_normal(D); lam = rng.standard_normal(D); mu = 10.0 ** rng.uniform(0, 9, D)
    ineq = np.zeros(D, dtype=bool)
    for row, (o, kn, p) in zip(table.rows, table.dual_slices()):
        ineq[o:o + kn * p] = row.kind == rbq.CON_BOUND
    absent = ineq & (rng.uniform(size=D) < 0.5)
    c[absent] = -np.inf
    lmax, mmax, phi = 5.0e3, 1.0e8, 10.0
    live = ~absent
    lam_ref, mu_ref = lam.copy(), mu.copy()
    l = lam + mu * np.where(live, c, 0.0)
    lam_ref[live] = np.clip(l, np.where(ineq, 0.0, -lmax), lmax)[live]
    mu_ref[live] = np.minimum(phi * mu, mmax)[live]
    viol_ref = max(np.max(np.abs(c[live & ~ineq])), np.max(np.maximum(c[live & ineq], 0.0)))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dc, dl, dm, dv = d(c), d(lam), d(mu), torch.empty(1, dtype=torch.float64, device=dev)
    eng.al_dual_update_dev(N, table, dc, dl, dm, dv, phi=phi, mu_max=mmax, lambda_max=lmax, update=False)
    torch.cuda.synchronize()
    assert float(dv[0]) == viol_ref and np.array_equal(dl.cpu().numpy(), lam) and np.array_equal(dm.cpu().numpy(), mu)
    eng.al_dual_update_dev(N, table, dc, dl, dm, dv, phi=phi, mu_max=mmax, lambda_max=lmax)
    torch.cuda.synchronize()
    assert float(dv[0]) == viol_ref
    assert np.allclose(dl.cpu().numpy(), lam_ref, rtol=1e-12, atol=1e-9) and np.array_equal(dm.cpu().numpy(), mu_ref)
    assert np.any(lam_ref == lmax) and np.any(mu_ref == mmax) and np.any(lam_ref[ineq & live] == 0.0)
