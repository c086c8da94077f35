"""GPU parity: augmented-state step maps, rollouts and knot-point Jacobians vs the CPU oracle.

Tolerances are the north star's: relative 1e-12 on states, 1e-10 on Jacobians (scaled by the largest
entry of the block).  Every call goes through the C ABI (librbq.so) via rbqoc_b200.Engine.
"""
import numpy as np
import pytest

import rbqoc_b200 as rbq
from rbqoc_b200 import spin
from _cases import MODEL_KW, golden_model_cases, relerr

pytestmark = pytest.mark.gpu

STATE_RTOL = 1e-12
JAC_RTOL = 1e-10


def _models():
    return [
        ("spin13", spin.Spin13Model()),
        ("spin12", spin.Spin12Model()),
        ("spin18", spin.Spin18Model()),
        ("spin15", spin.Spin15Model(False, False)),
        ("spin15_DA", spin.Spin15Model(True, False)),
        ("spin15_TO", spin.Spin15Model(False, True)),
        ("spin15_DA_TO", spin.Spin15Model(True, True)),
        ("spin11_DO0", spin.Spin11Model(0)),
        ("spin11_DO1", spin.Spin11Model(1)),
        ("spin11_DO2", spin.Spin11Model(2)),
        ("spin17_DO1", spin.Spin17Model(1)),
        ("spin17_DO2", spin.Spin17Model(2)),
        ("spin30", spin.Spin30Model(S_diag=(1.0, 2.0, 3.0, 4.0))),
        ("spin30_x2", spin.Spin30Model(sample_state_count=2, nominal_idxs=[1, 5], alpha=1.3)),
        ("spin25", spin.Spin25Model(S_diag=(1.0, 2.0, 3.0, 4.0), namp=2.574e-2, nominal_idxs=[5])),
        ("spin25_x2", spin.Spin25Model(sample_state_count=2, nominal_idxs=[1, 5], alpha=1.1, namp=1e-2, wnamp_cov=0.7)),
    ]


MODELS = _models()
IDS = [m[0] for m in MODELS]


def _random_point(model, rng, a=None):
    """A generic augmented state: unit-ish psi blocks, a in the control range, spin30 sigma points spread."""
    n, m = model.n, model.m
    x = rng.standard_normal(n) * 0.3
    mid = model.params.model_id
    if mid in (30, 25):
        base = 15 if mid == 30 else 17
        if mid == 30:
            x[:12] = np.concatenate([spin.IS1_ISO, spin.IS2_ISO, spin.IS3_ISO]) + 1e-3 * rng.standard_normal(12)
            x[12:15] = rng.uniform(-0.4, 0.4, 3)
        else:
            x[:8] = np.concatenate([spin.IS1_ISO, spin.IS2_ISO]) + 1e-3 * rng.standard_normal(8)
            x[8:11] = rng.uniform(-0.4, 0.4, 3)
            x[11:17] = rng.uniform(-1.0, 1.0, 6)       # pink-filter taps
        for c in range(model.params.sample_state_count):
            off = base + 41 * c
            for j in range(10):
                s = spin.IS1_ISO + 0.1 * rng.standard_normal(4)
                x[off + 4 * j: off + 4 * j + 4] = s / np.linalg.norm(s)
            x[off + 40] = rng.uniform(0, 1e-2)
        if a is not None:
            x[13 if mid == 30 else 9] = a
    else:
        x[8:11] = rng.uniform(-0.4, 0.4, 3)
        if a is not None:
            x[9] = a
    u = rng.uniform(-0.2, 0.2, m)
    if mid == 15 and model.params.time_optimal:
        u[1] = rng.uniform(0.2, 0.4)     # dt = u2^2 in [0.04, 0.16]
    return x, u


def _seed(name, k=0):
    import zlib

    return (zlib.crc32(name.encode()) + k) % 2**32


def _relerr(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("name,model", MODELS, ids=IDS)
def test_state_and_control_dims(engine, oracle, name, model):
    assert model.n == oracle.state_dim(model.params)
    assert model.m == oracle.control_dim(model.params)


@pytest.mark.parametrize("name,model", MODELS, ids=IDS)
@pytest.mark.parametrize("dt", [0.1, 0.01])
def test_discrete_dynamics_matches_oracle(engine, oracle, name, model, dt):
    rng = np.random.default_rng(_seed(name))
    for a in (None, 0.0, 0.5, -0.37, 1e-9):
        x, u = _random_point(model, rng, a)
        got = spin.discrete_dynamics(spin.RK3, model, x, u, 0.0, dt, engine=engine)
        ref = oracle.step(model.params, x, u, dt)
        assert got.shape == ref.shape
        assert _relerr(got, ref) < STATE_RTOL, (name, a)


@pytest.mark.parametrize("name,model", MODELS, ids=IDS)
@pytest.mark.parametrize("dt", [0.1, 0.01])
def test_discrete_jacobian_matches_oracle_forward_mode(engine, oracle, name, model, dt):
    rng = np.random.default_rng(_seed(name, 7))
    for a in (None, 0.0, 0.5, -0.37, -0.0):
        x, u = _random_point(model, rng, a)
        z = spin.KnotPoint(x, u, dt)
        got = spin.discrete_jacobian(spin.RK3, model, z, engine=engine)
        ref = oracle.jacobian(model.params, x, u, dt)
        assert got.shape == (model.n, model.n + model.m)
        assert _relerr(got, ref) < JAC_RTOL, (name, a)


@pytest.mark.parametrize("name,model", MODELS, ids=IDS)
@pytest.mark.parametrize("dt,a", [(0.5, 0.3), (1.0, 0.3), (2.0, 0.5), (5.0, -0.45), (0.3, 0.49)],
                         ids=["x0.22_series8", "x0.89_sincos", "x9.9_sincos", "x50_sincos", "x0.21_series8"])
def test_large_rotation_angles(engine, oracle, name, model, dt, a):
    """x = (pi dt)^2 (fq^2 + a^2) far outside the optimisation grids: the long series and the sincos branch of the closed
    form (value and tangents) against the oracle's Pade-7/9/13 + scaling-and-squaring."""
    if model.params.model_id == 15 and model.params.time_optimal:
        pytest.skip("dt is a control for the time-optimal model (covered below)")
    rng = np.random.default_rng(_seed(name, 31))
    x, u = _random_point(model, rng, a)
    got = engine.discrete_dynamics(model.params, x, u, 0.0, dt)
    ref = oracle.step(model.params, x, u, dt)
    assert _relerr(got, ref) < STATE_RTOL, name
    J = engine.discrete_jacobian(model.params, x, u, 0.0, dt)
    Jr = oracle.jacobian(model.params, x, u, dt)
    assert _relerr(J, Jr) < JAC_RTOL, name


def test_time_optimal_large_dt_control(engine, oracle):
    """spin15 TO: dt = u2^2 up to 4 ns, so the tangent of the sincos branch w.r.t. dt is exercised too."""
    model = spin.Spin15Model(True, True)
    rng = np.random.default_rng(17)
    for u2 in (0.7, 1.0, 1.5, 2.0):
        x, u = _random_point(model, rng, 0.4)
        u[1] = u2
        got = engine.discrete_dynamics(model.params, x, u, 0.0, 0.1)
        ref = oracle.step(model.params, x, u, 0.1)
        assert _relerr(got, ref) < STATE_RTOL
        J = engine.discrete_jacobian(model.params, x, u, 0.0, 0.1)
        Jr = oracle.jacobian(model.params, x, u, 0.1)
        assert _relerr(J, Jr) < JAC_RTOL


@pytest.mark.parametrize("name,model", MODELS, ids=IDS)
def test_rollout_and_batched_jacobians(engine, oracle, name, model):
    rng = np.random.default_rng(_seed(name, 11))
    N = 301
    dt = np.full(N - 1, 0.1)
    B = 3
    x0 = np.stack([_random_point(model, rng)[0] for _ in range(B)])
    if model.params.model_id != 30:
        x0[:, 0:8] = np.concatenate([spin.IS1_ISO, spin.IS2_ISO])
        x0[:, 8:11] = 0.0
    else:
        x0[:, 12:15] = 0.0
    if model.params.model_id == 25:
        x0[:, 11:17] = 0.0          # the reference starts the filter from rest
    U = rng.uniform(-2e-3, 2e-3, (B, N - 1, model.m))
    if model.params.model_id == 15 and model.params.time_optimal:
        U[:, :, 1] = rng.uniform(0.25, 0.35, (B, N - 1))
    X = engine.rollout(model.params, x0, U, dt)
    Xr = oracle.rollout(model.params, x0, U, dt)
    assert X.shape == (B, N, model.n)
    assert _relerr(X, Xr) < 5e-12, name    # 300 chained steps
    # all knots of trajectory 0 in one Jacobian call
    tk = spin.knot_times(dt)
    AB = engine.jacobians(model.params, Xr[0, :-1], U[0], dt, t=tk)
    ABr = oracle.jacobians(model.params, Xr[0, :-1], U[0], dt, t=tk)
    assert AB.shape == (N - 1, model.n, model.n + model.m)
    scale = np.max(np.abs(ABr), axis=(1, 2), keepdims=True)
    assert np.max(np.abs(AB - ABr) / scale) < JAC_RTOL, name


@pytest.mark.parametrize("case", golden_model_cases(), ids=lambda c: c[0])
def test_step_and_jacobian_match_50_digit_golden(engine, case):
    """Committed fixtures from the rounding-free referee (tests/golden/make_golden.py)."""
    import rbqoc_b200 as rbq

    key, name, dt, x, u, xn, J, t = case
    p = spin._params(**MODEL_KW[name])
    got = engine.discrete_dynamics(p, x, u, t, dt)
    assert relerr(got, xn) < 2e-15, key
    Jg = engine.discrete_jacobian(p, x, u, t, dt)
    assert Jg.shape == J.shape
    assert relerr(Jg, J) < 1e-13, key


@pytest.mark.parametrize("name,model", MODELS, ids=IDS)
def test_closed_loop_rollout(engine, oracle, name, model):
    """The iLQR line search: u_k = uref_k + K_k (x_k - xref_k) + alpha d_k for several alpha at once."""
    rng = np.random.default_rng(_seed(name, 5))
    N = 181
    dt = np.full(N - 1, 0.1)
    x0 = _random_point(model, rng)[0]
    if model.params.model_id != 30:
        x0[0:8] = np.concatenate([spin.IS1_ISO, spin.IS2_ISO])
        x0[8:11] = 0.0
    else:
        x0[12:15] = 0.0
    if model.params.model_id == 25:
        x0[11:17] = 0.0
    Uref = rng.uniform(-2e-3, 2e-3, (N - 1, model.m))
    if model.params.model_id == 15 and model.params.time_optimal:
        Uref[:, 1] = rng.uniform(0.25, 0.35, N - 1)
    Xref = oracle.rollout(model.params, x0, Uref, dt)
    # gains scaled per state so that the closed loop stays well conditioned (the derivative states of spin11/17 grow
    # like (pi t)^2; an O(1e-2) gain on them makes the loop amplify rounding noise by many orders of magnitude)
    K = 1e-2 * rng.standard_normal((N - 1, model.m, model.n)) / (1.0 + np.max(np.abs(Xref), axis=0))
    d = 1e-3 * rng.standard_normal((N - 1, model.m))
    alphas = np.array([1.0, 0.5, 0.25, 0.125, 0.0])
    X, Uo = engine.rollout_closedloop(model.params, Xref, Uref, K, d, dt, alphas)
    Kc = np.ascontiguousarray(K.transpose(0, 2, 1))
    Xr, Ur = oracle.rollout_closedloop(model.params, Xref, Uref, Kc, d, dt, alphas)
    assert X.shape == (5, N, model.n) and Uo.shape == (5, N - 1, model.m)
    assert _relerr(X, Xr) < 5e-12, name
    assert _relerr(Uo, Ur) < 5e-12, name
    assert _relerr(X[-1], Xref) < 1e-13      # alpha = 0 reproduces the reference trajectory


def test_rollout_large_batch_uses_thread_per_trajectory(engine, oracle):
    """Above 65536 trajectories the engine switches from one warp to one thread per trajectory."""
    model = spin.Spin13Model()
    B, N = 70000, 5
    rng = np.random.default_rng(1)
    x0 = np.tile(spin.spin13_x0(), (B, 1))
    U = rng.uniform(-1e-2, 1e-2, (B, N - 1, 1))
    dt = np.full(N - 1, 0.1)
    X = engine.rollout(model.params, x0, U, dt)
    idx = rng.integers(0, B, 64)
    Xr = oracle.rollout(model.params, x0[idx], U[idx], dt)
    assert _relerr(X[idx], Xr) < 1e-13


@pytest.mark.parametrize("name,model", MODELS, ids=IDS)
def test_jacobians_staged_kernel_equals_direct(engine, oracle, name, model, monkeypatch):
    """Large batches go through the shared-memory staged kernel (coalesced stores); same arithmetic, same bits."""
    rng = np.random.default_rng(_seed(name, 21))
    K = 301     # not a multiple of the knots-per-block of any model: exercises the ragged last block
    pts = [_random_point(model, rng) for _ in range(K)]
    X = np.stack([p[0] for p in pts])
    U = np.stack([p[1] for p in pts])
    dt = rng.choice([0.1, 0.01], K)
    engine.set_option("RBQ_JAC_STRUCTURED", "0")     # the generic forward-mode kernels (the unscented models default to the structured ones)
    engine.set_option("RBQ_JAC_STAGED", "0")
    direct = engine.jacobians(model.params, X, U, dt).copy()
    engine.set_option("RBQ_JAC_STAGED", "1")
    staged = engine.jacobians(model.params, X, U, dt).copy()
    assert np.array_equal(direct, staged), name
    ref = oracle.jacobians(model.params, X[:40], U[:40], dt[:40])
    scale = np.max(np.abs(ref), axis=(1, 2), keepdims=True)
    assert np.max(np.abs(staged[:40] - ref) / scale) < JAC_RTOL, name


def test_pinned_host_buffers(engine):
    model = spin.Spin13Model()
    X = engine.pinned_empty((50, 11))
    X[:] = np.tile(spin.spin13_x0(), (50, 1))
    U = engine.pinned_empty((50, 1))
    U[:] = 0.01
    dt = engine.pinned_empty(50)
    dt[:] = 0.1
    AB = engine.jacobians(model.params, X, U, dt)
    assert AB.shape == (50, 11, 12) and np.all(AB[:, 10, 11] == 0.1)


def test_degenerate_sizes(engine):
    """Empty and one-knot inputs (the edge cases a solver hits with N = 1 or an empty batch)."""
    model = spin.Spin12Model()
    X = engine.rollout(model.params, spin.spin12_x0(), np.empty((0, 1)), np.empty(0))
    assert X.shape == (1, 43) and np.array_equal(X[0], spin.spin12_x0())
    AB = engine.jacobians(model.params, np.empty((0, 43)), np.empty((0, 1)), np.empty(0))
    assert AB.shape == (0, 43, 44)
    Xb = engine.rollout(model.params, np.empty((0, 43)), np.empty((0, 3, 1)), np.full(3, 0.1))
    assert Xb.shape == (0, 4, 43)


def test_spin25_reads_its_time_argument(engine, oracle):
    """spin25.jl:196-214: Int(div(time, dt)) picks the pink filter's start-up branch, for the step and for its Jacobian."""
    model = spin.Spin25Model(namp=2.574e-2)
    rng = np.random.default_rng(3)
    x, u = _random_point(model, rng)
    dt = 0.1
    outs = []
    for k, t in enumerate(spin.knot_times(np.full(7, dt))):
        got = engine.discrete_dynamics(model.params, x, u, t, dt)
        ref = oracle.step(model.params, x, u, dt, t=t)
        assert _relerr(got, ref) < STATE_RTOL, k
        J = engine.discrete_jacobian(model.params, x, u, t, dt)
        Jr = oracle.jacobian(model.params, x, u, dt, t=t)
        assert _relerr(J, Jr) < JAC_RTOL, k
        outs.append(got[14])
        # d y_k / d (xp1, xp2, xp3, yp1, yp2, yp3): taps the start-up branch ignores have zero columns
        active = {0: 6, 1: 0, 2: 2, 3: 4}.get(k, 6)
        nz = np.count_nonzero(J[14, 11:17])
        assert nz == active, (k, nz)
    assert len(set(np.round(outs[:5], 12))) >= 4          # the branches really differ


def test_gate_error_expansion_matches_oracle(engine, oracle):
    """Cost-side primitives (spin.jl:1040-1092): batched err / gradient / Hessian, per-knot targets or one shared target."""
    rng = np.random.default_rng(12)
    K = 1000
    s1, s2 = rng.standard_normal((K, 4)), rng.standard_normal((K, 4))
    err, grad, hess = engine.gate_error_expansion(s1, s2)
    for k in range(0, K, 37):
        e, g, h = oracle.gate_error_expansion(s1[k], s2[k])
        assert abs(err[k] - e) < 1e-13 * max(1.0, abs(e))
        assert np.max(np.abs(grad[k] - g)) < 1e-13 * max(1.0, np.max(np.abs(g)))
        assert np.max(np.abs(hess[k] - h)) < 1e-13 * max(1.0, np.max(np.abs(h)))
    tgt = spin.IS3_ISO
    err1, grad1, hess1 = engine.gate_error_expansion(s1, tgt)
    e, g, h = oracle.gate_error_expansion(s1[5], tgt)
    assert abs(err1[5] - e) < 1e-13 and np.allclose(grad1[5], g, atol=1e-13) and np.allclose(hess1[5], h, atol=1e-13)
    # mirrors of the reference names
    assert np.allclose(spin.gate_error_iso2(s1[:3], tgt, engine=engine), err1[:3])
    assert np.allclose(spin.jacobian_gate_error_iso2(s1[:3], tgt, engine=engine), grad1[:3])
    assert np.allclose(spin.hessian_gate_error_iso2(tgt, engine=engine)[0], hess1[0])
    # a state equal to its target has zero gate error and zero gradient component along the error
    e0, _, _ = engine.gate_error_expansion(tgt[None, :], tgt)
    assert abs(e0[0]) < 1e-15


def test_sampling_cost_expansion_matches_reference_formulas(engine, oracle):
    """TO.stage_cost / TO.gradient! of the spin12 `Cost` (spin12.jl:102-140), restated here term by term with the oracle's
    gate-error primitives."""
    rng = np.random.default_rng(8)
    K, n = 257, 43
    X = rng.standard_normal((K, n)) * 0.5
    U = rng.standard_normal((K, 1)) * 0.1
    Q = rng.uniform(0.1, 2.0, n)
    R = np.array([0.37])
    xf = rng.standard_normal(n)
    targets = rng.standard_normal((4, 4))
    targets /= np.linalg.norm(targets, axis=1, keepdims=True)
    q_ss = np.array([1.0, 2.0, 0.5, 3.0])
    J, gx, gu = engine.sampling_cost_expansion(X, U, Q, R, xf, targets, q_ss)
    q = -Q * xf                                             # spin12.jl:77
    c = 0.5 * xf @ (Q * xf)                                 # spin12.jl:78
    for k in range(0, K, 16):
        x = X[k]
        cost = 0.5 * x @ (Q * x) + q @ x + c
        grad = Q * x + q
        for s in range(8):
            e, g, _ = oracle.gate_error_expansion(x[11 + 4 * s: 15 + 4 * s], targets[s % 4])
            cost += q_ss[s % 4] * e
            grad[11 + 4 * s: 15 + 4 * s] += q_ss[s % 4] * g
        cost += 0.5 * U[k, 0] * R[0] * U[k, 0]             # spin12.jl:116-119
        assert abs(J[k] - cost) < 1e-12 * max(1.0, abs(cost))
        assert np.max(np.abs(gx[k] - grad)) < 1e-12 * max(1.0, np.max(np.abs(grad)))
        assert abs(gu[k, 0] - R[0] * U[k, 0]) < 1e-15
    # terminal cost: no control term
    Jt, gxt, gut = engine.sampling_cost_expansion(X[:3], None, Q, None, xf, targets, q_ss)
    assert gut is None and np.allclose(Jt, J[:3] - 0.5 * R[0] * U[:3, 0] ** 2, rtol=1e-13) and np.array_equal(gxt, gx[:3])


def test_jacobian_structure_spin13(engine):
    """A = blkdiag(E, E, chain) + the d/da column; B = dt e_11."""
    model = spin.Spin13Model()
    x = np.concatenate([spin.IS3_ISO, spin.IS4_ISO, [0.1, 0.3, -0.2]])
    AB = engine.discrete_jacobian(model.params, x, np.array([0.05]), 0.0, 0.1)
    A, B = AB[:, :11], AB[:, 11]
    assert np.allclose(A[0:4, 0:4], A[4:8, 4:8], atol=0, rtol=0)
    assert np.allclose(A[0:4, 0:4] @ A[0:4, 0:4].T, np.eye(4), atol=1e-15)
    assert np.all(A[0:4, 4:9] == 0) and np.all(A[8:11, 0:8] == 0)
    assert np.allclose(A[8:11, 8:11], [[1, 0.1, 0], [0, 1, 0.1], [0, 0, 1]], atol=0)
    assert np.allclose(B, np.r_[np.zeros(10), 0.1], atol=0)


def test_error_codes(engine):
    import rbqoc_b200

    bad = spin.Spin11Model(5)
    with pytest.raises(ValueError):
        bad.n
    p = spin.Spin13Model().params
    with pytest.raises(ValueError):
        engine.discrete_dynamics(p, np.zeros(5), np.zeros(1), 0.0, 0.1)
    # NaN inputs are returned, not hidden (iLQR relies on seeing them)
    x = np.full(11, np.nan)
    out = engine.discrete_dynamics(p, x, np.zeros(1), 0.0, 0.1)
    assert np.all(np.isnan(out[:10]))
    assert isinstance(rbqoc_b200.RbqError(-1, "x"), RuntimeError)


@pytest.mark.parametrize("name,model", [mm for mm in MODELS if mm[1].params.model_id in (13, 15, 12, 18)], ids=lambda v: v if isinstance(v, str) else "")
def test_packed_batch_rollout_equals_warp_rollout(engine, oracle, name, model, monkeypatch):
    """Batches of the independent-block models run one thread per (trajectory, block) instead of one warp per trajectory: the
    same expressions per block, so the same bits; and the oracle's rollout within the usual tolerance."""
    rng = np.random.default_rng(_seed(name, 31))
    B, N = 37, 120
    dt = rng.choice([0.1, 0.01], N - 1)
    x0 = np.stack([_random_point(model, rng)[0] for _ in range(B)])
    U = rng.uniform(-2e-3, 2e-3, (B, N - 1, model.m))
    if model.params.model_id == 15 and model.params.time_optimal:
        U[:, :, 1] = rng.uniform(0.25, 0.35, (B, N - 1))
    engine.set_option("RBQ_ROLLOUT_PACKED", "0")
    Xw = engine.rollout(model.params, x0, U, dt)
    engine.set_option("RBQ_ROLLOUT_PACKED", "1")
    Xp = engine.rollout(model.params, x0, U, dt)
    assert np.array_equal(Xw, Xp), name
    Xr = oracle.rollout(model.params, x0[:5], U[:5], dt)
    assert _relerr(Xp[:5], Xr) < 5e-12, name


@pytest.mark.parametrize("name,model", [mm for mm in MODELS if mm[1].params.model_id in (30, 25) and mm[1].params.sample_state_count == 1],
                         ids=lambda v: v if isinstance(v, str) else "")
@pytest.mark.parametrize("K", [1, 63, 3001])
def test_structured_unscented_jacobians_equal_generic_forward_mode(engine, oracle, name, model, K):
    """spin30 / spin25 with one sample state assemble [A_k B_k] from the block structure (value path once per knot, sparse
    seeds through mean -> covariance -> Cholesky -> resample -> normalise; one bulk store per block).  Same derivatives as the
    generic tangent kernel and as the oracle's dual numbers; K = 3001 makes every persistent block loop over several knots
    (both stage buffers in use), K = 1 and 63 leave blocks idle."""
    rng = np.random.default_rng(_seed(name, 77) + K)
    pts = [_random_point(model, rng) for _ in range(min(K, 200))]
    X = np.stack([pts[i % len(pts)][0] for i in range(K)]) * (1 + 1e-3 * rng.standard_normal((K, 1)))
    U = np.stack([pts[i % len(pts)][1] for i in range(K)])
    dt = rng.choice([0.1, 0.01], K)
    t = dt * rng.integers(0, 9, K)                      # spin25: knot index 0..8 selects the filter start-up branch
    structured = engine.jacobians(model.params, X, U, dt, t=t).copy()
    engine.set_option("RBQ_JAC_STRUCTURED", "0")
    generic = engine.jacobians(model.params, X, U, dt, t=t).copy()
    scale = np.max(np.abs(generic), axis=(1, 2), keepdims=True)
    assert np.max(np.abs(structured - generic) / scale) < 1e-12, name
    pick = rng.integers(0, K, min(K, 24))
    ref = oracle.jacobians(model.params, X[pick], U[pick], dt[pick], t=t[pick])
    rs = np.max(np.abs(ref), axis=(1, 2), keepdims=True)
    assert np.max(np.abs(structured[pick] - ref) / rs) < JAC_RTOL, name


def test_structured_unscented_jacobians_unaligned_output_falls_back(engine, oracle):
    """A device output buffer that starts on an odd double cannot take bulk stores: the launcher must use the generic kernel."""
    import torch

    model = spin.Spin30Model()
    n, m, K = model.n, model.m, 40
    rng = np.random.default_rng(3)
    pts = [_random_point(model, rng) for _ in range(K)]
    X = np.stack([p[0] for p in pts]); U = np.stack([p[1] for p in pts]); dt = np.full(K, 0.1)
    dev = torch.device("cuda", 0)
    engine.use_torch_stream()
    dX, dU, ddt = (torch.from_numpy(a).to(dev) for a in (X, U, dt))
    buf = torch.zeros(K * n * (n + m) + 1, dtype=torch.float64, device=dev)
    engine.jacobians_dev(model.params, K, dX, dU, ddt, buf[1:])
    engine.synchronize()
    got = buf[1:].cpu().numpy().reshape(K, n + m, n).transpose(0, 2, 1)
    ref = engine.jacobians(model.params, X, U, dt)
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))
    engine.set_stream(0)


@pytest.mark.parametrize("name,model", [mm for mm in MODELS if mm[1].params.model_id in (13, 15, 12, 18)], ids=lambda v: v if isinstance(v, str) else "")
@pytest.mark.parametrize("N", [2, 66, 1001, 5681])
def test_scan_rollout_equals_stepwise_rollout_and_oracle(engine, oracle, name, model, N):
    """Open-loop rollouts of one (or a few) trajectories run as a prefix product of quaternions over the knots
    (rbq_scan_rollout.cuh), the control chain as three prefix sums; the step-by-step kernels and the oracle walk the knots.
    Equal to summation / product-order rounding."""
    rng = np.random.default_rng(_seed(name, 5) + N)
    n, m = model.n, model.m
    B = 3
    x0 = np.stack([_random_point(model, rng)[0] for _ in range(B)])
    x0[:, 8:11] = rng.uniform(-0.05, 0.05, (B, 3))
    U = rng.uniform(-2e-3, 2e-3, (B, N - 1, m))
    if model.params.model_id == 15 and model.params.time_optimal:
        U[:, :, 1] = rng.uniform(0.2, 0.4, (B, N - 1))
    dt = rng.choice([0.1, 0.01], N - 1)
    engine.set_option("RBQ_ROLLOUT_SCAN", "1")
    Xs = engine.rollout(model.params, x0, U, dt)
    engine.set_option("RBQ_ROLLOUT_SCAN", "0")
    engine.set_option("RBQ_ROLLOUT_PACKED", "0")
    Xw = engine.rollout(model.params, x0, U, dt)
    cb = slice(8, 11)
    assert _relerr(Xs[:, :, cb], Xw[:, :, cb]) < 1e-14, name          # the control chain: prefix sums against the sequential walk
    assert _relerr(Xs, Xw) < 1e-13 + 2e-16 * N, name
    Xr = oracle.rollout(model.params, x0, U, dt)
    assert _relerr(Xs, Xr) < 1e-12, name


# ---- Runge-Kutta steps of the continuous dynamics (RBQ_INTEGRATOR_RK2/3/4; spin15_standalone.jl:263-312, rbqoc.jl:41-54) -------
RK_MODELS = [mm for mm in MODELS if mm[1].params.model_id in (13, 15, 11, 17)]


@pytest.mark.parametrize("name,model", RK_MODELS, ids=lambda v: v if isinstance(v, str) else "")
@pytest.mark.parametrize("integ", [spin.IntegratorType.rk2, spin.IntegratorType.rk3, spin.IntegratorType.rk4], ids=["rk2", "rk3", "rk4"])
def test_runge_kutta_steps_jacobians_and_rollouts_match_oracle(engine, oracle, name, model, integ):
    mrk = spin.with_integrator(model, integ)
    rng = np.random.default_rng(_seed(name, 300 + int(integ)))
    n, m = mrk.n, mrk.m
    for dt in (0.1, 0.01):
        x, u = _random_point(mrk, rng)
        xn = engine.discrete_dynamics(mrk.params, x, u, 0.0, dt)
        assert _relerr(xn, oracle.step(mrk.params, x, u, dt)) < 1e-13, (name, dt)
        J = engine.discrete_jacobian(mrk.params, x, u, 0.0, dt)
        Jr = oracle.jacobian(mrk.params, x, u, dt)
        assert np.max(np.abs(J - Jr)) / np.max(np.abs(Jr)) < JAC_RTOL, (name, dt)
    B, N = 3, 40
    x0 = np.stack([_random_point(mrk, rng)[0] for _ in range(B)])
    U = rng.uniform(-2e-2, 2e-2, (B, N - 1, m))
    if mrk.params.model_id == 15 and mrk.params.time_optimal:
        U[:, :, 1] = rng.uniform(0.8, 1.2, (B, N - 1))
    dtv = np.full(N - 1, 0.05)
    X = engine.rollout(mrk.params, x0, U, dtv)
    assert _relerr(X, oracle.rollout(mrk.params, x0, U, dtv)) < 1e-12, name
    K = 70
    Xk = np.stack([_random_point(mrk, rng)[0] for _ in range(K)]); Uk = np.stack([_random_point(mrk, rng)[1] for _ in range(K)])
    AB = engine.jacobians(mrk.params, Xk, Uk, np.full(K, 0.1))
    ABr = oracle.jacobians(mrk.params, Xk, Uk, np.full(K, 0.1))
    assert np.max(np.abs(AB - ABr)) / np.max(np.abs(ABr)) < JAC_RTOL, name


def test_runge_kutta_is_refused_where_it_does_not_exist(engine):
    with pytest.raises(ValueError):
        spin.with_integrator(spin.Spin30Model(), spin.IntegratorType.rk4)
    mrk = spin.with_integrator(spin.Spin13Model(), spin.IntegratorType.rk3)
    N = 5
    X = np.tile(spin.spin13_x0(), (N, 1)); U = np.zeros((N - 1, 1))
    with pytest.raises(rbq.RbqError):       # the closed-loop kernels implement the exponential step only
        engine.rollout_closedloop(mrk.params, X, U, np.zeros((N - 1, 1, 11)), np.zeros((N - 1, 1)), np.full(N - 1, 0.1), np.array([1.0]))
