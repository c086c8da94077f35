"""GPU parity: the iLQR backward pass (rbq_backward_pass) against the numpy restatement of Altro's backwardpass!
(oracle/riccati.py).  Tolerance: relative 1e-10 on gains and value decrease (the recursion multiplies ~N products of
n x n matrices; both sides are plain FP64 sums in different orders)."""
import numpy as np
import pytest

import rbqoc_b200 as rbq
from rbqoc_b200 import spin
from test_oracle_cpu import _lq_problem

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("n,m,N", [(11, 1, 301), (11, 2, 90), (12, 2, 64), (12, 1, 50), (15, 1, 70), (19, 1, 45), (43, 1, 181), (56, 1, 120), (58, 1, 40),
                                   (3, 4, 17), (64, 2, 9), (5, 1, 1)])
@pytest.mark.parametrize("reg", [rbq.REG_CONTROL, rbq.REG_STATE], ids=["control_reg", "state_reg"])
def test_backward_pass_matches_oracle(engine, n, m, N, reg):
    from oracle import riccati

    rng = np.random.default_rng(n * 1000 + m * 10 + N)
    A, Bm, AB, Q, q, R, Eux, H, r = _lq_problem(rng, N, n, m)
    rho = 0.37
    K, d, dV, fail = engine.backward_pass(AB, Q, q, R, Eux, r, rho=rho, reg_type=reg)
    Kr, dr, dVr, failr = riccati.backward_pass(AB, Q, q, R, Eux, r, rho=rho, reg_type=reg)
    assert fail == failr == -1
    if N > 1:
        assert _rel(K, Kr) < 1e-10 and _rel(d, dr) < 1e-10
    assert np.allclose(dV, dVr, rtol=1e-10, atol=1e-300)


def test_backward_pass_batch_failure_flags_and_jacobian_layout(engine):
    """A batch where one problem has an indefinite Quu: its flag names the knot, the others are unaffected.  Then the real
    thing: Jacobians of a spin13 rollout straight from rbq_jacobians into the pass (same memory layout, no transpose)."""
    from oracle import riccati

    rng = np.random.default_rng(3)
    N, n, m, B = 33, 11, 1, 5
    probs = [_lq_problem(rng, N, n, m) for _ in range(B)]
    st = lambda i: np.stack([p[i] for p in probs])
    AB, Q, q, R, Eux, r = st(2), st(3), st(4), st(5), st(6), st(8)
    R[3, 20] = -1e3
    K, d, dV, fail = engine.backward_pass(AB, Q, q, R, Eux, r)
    assert list(fail) == [-1, -1, -1, 20, -1]
    for b in (0, 1, 2, 4):
        Kr, dr, dVr, _ = riccati.backward_pass(AB[b], Q[b], q[b], R[b], Eux[b], r[b])
        assert _rel(K[b], Kr) < 1e-10 and _rel(d[b], dr) < 1e-10 and np.allclose(dV[b], dVr, rtol=1e-10)
    Kr, dr, _, failr = riccati.backward_pass(AB[3], Q[3], q[3], R[3], Eux[3], r[3])
    assert failr == 20 and _rel(K[3, 21:], Kr[21:]) < 1e-10          # gains after the failing knot were already final
    # spin13: LQR objective of spin13.jl:128-135 (Q = diag(qs), R = qs[5], Qf = N Q) expanded about a rollout
    model = spin.Spin13Model()
    Nk = 101
    U = 1e-2 * rng.standard_normal((Nk - 1, 1))
    X = spin.rollout(model, spin.spin13_x0(), U, 0.1, engine=engine)
    ABj = spin.dynamics_expansion(model, X, U, 0.1, engine=engine)
    xf = np.zeros(11); xf[:8] = [1 / np.sqrt(2), 0, -1 / np.sqrt(2), 0, 0, 1 / np.sqrt(2), 0, 1 / np.sqrt(2)]
    Qd = np.array([1.0] * 8 + [1.0, 1.0, 0.1])
    Exx = np.tile(np.diag(Qd), (Nk, 1, 1)); Exx[-1] *= Nk
    Ex = (X - xf) * Qd; Ex[-1] *= Nk
    Euu = np.full((Nk - 1, 1, 1), 0.1); Eux0 = np.zeros((Nk - 1, 1, 11)); Eu = 0.1 * U
    K, d, dV, fail = engine.backward_pass(ABj, Exx, Ex, Euu, Eux0, Eu, rho=1e-3)
    Kr, dr, dVr, failr = riccati.backward_pass(ABj, Exx, Ex, Euu, Eux0, Eu, rho=1e-3)
    assert fail == failr == -1 and _rel(K, Kr) < 1e-9 and _rel(d, dr) < 1e-9 and np.allclose(dV, dVr, rtol=1e-9)
    assert dV[0] < 0                                                  # a descent direction
    with pytest.raises(Exception):
        engine.backward_pass(np.zeros((3, 65, 70)), np.zeros((4, 65, 65)), np.zeros((4, 65)), np.zeros((3, 5, 5)), np.zeros((3, 5, 65)),
                             np.zeros((3, 5)))


def test_backward_pass_one_warp_form_for_small_models(engine, monkeypatch):
    """n <= 16 in large batches runs one warp per problem (warp barriers only); same recursion, same tolerance, and both forms
    agree with each other to rounding."""
    from oracle import riccati

    rng = np.random.default_rng(77)
    for n, m, N in ((11, 1, 140), (12, 2, 60), (12, 1, 33), (15, 1, 40), (16, 3, 25), (3, 1, 9)):
        A, Bm, AB, Q, q, R, Eux, H, r = _lq_problem(rng, N, n, m)
        st = lambda x: np.stack([x] * 3)
        engine.set_option("RBQ_BP_WARPS", "1")
        K1, d1, dV1, f1 = engine.backward_pass(st(AB), st(Q), st(q), st(R), st(Eux), st(r), rho=0.2, reg_type=rbq.REG_STATE)
        engine.set_option("RBQ_BP_WARPS", "8")
        K8, d8, dV8, f8 = engine.backward_pass(st(AB), st(Q), st(q), st(R), st(Eux), st(r), rho=0.2, reg_type=rbq.REG_STATE)
        engine.set_option("RBQ_BP_WARPS", None)
        Kr, dr, dVr, fr = riccati.backward_pass(AB, Q, q, R, Eux, r, rho=0.2, reg_type=rbq.REG_STATE)
        assert list(f1) == list(f8) == [-1] * 3 and fr == -1
        for K, d, dV in ((K1, d1, dV1), (K8, d8, dV8)):
            assert _rel(K[1], Kr) < 1e-10 and _rel(d[2], dr) < 1e-10 and np.allclose(dV[0], dVr, rtol=1e-10)
    # a failing problem in the warp form
    A, Bm, AB, Q, q, R, Eux, H, r = _lq_problem(rng, 30, 11, 1)
    R[12] = -1e3
    engine.set_option("RBQ_BP_WARPS", "1")
    assert engine.backward_pass(AB, Q, q, R, Eux, r)[3] == 12


@pytest.mark.parametrize("n,N", [(11, 60), (15, 30), (36, 33), (43, 21), (44, 20), (56, 24), (58, 17)])
def test_backward_pass_staging_modes_and_specialised_shapes_agree(engine, monkeypatch, n, N):
    """The launcher picks a staging mode (cp.async / bulk TMA copies, dense or padded Jacobian block) and, for the reference
    models' shapes, a kernel with n known at compile time.  Every choice is the same recursion: all agree with the oracle, on a
    batch whose second and third problems start on odd doubles when n is odd (the shifted bulk copy of the cost block)."""
    from oracle import riccati

    rng = np.random.default_rng(900 + n)
    probs = [_lq_problem(rng, N, n, 1) for _ in range(3)]
    stack = lambda i: np.stack([p[i] for p in probs])
    AB, Q, q, R, Eux, r = (stack(i) for i in (2, 3, 4, 5, 6, 8))
    ref = [riccati.backward_pass(p[2], p[3], p[4], p[5], p[6], p[8], rho=0.11, reg_type=rbq.REG_CONTROL) for p in probs]
    outs = []
    for env in ({}, {"RBQ_BP_GENERIC": "1"}, {"RBQ_BP_TMA": "0"}, {"RBQ_BP_TMA": "2"}, {"RBQ_BP_GENERIC": "1", "RBQ_BP_TMA": "0"}):
        for k, v in env.items():
            engine.set_option(k, v)
        outs.append(engine.backward_pass(AB, Q, q, R, Eux, r, rho=0.11, reg_type=rbq.REG_CONTROL))
        for k in env:
            engine.set_option(k, None)
    for K, d, dV, fail in outs:
        assert list(fail) == [-1] * 3
        for b, (Kr, dr, dVr, fr) in enumerate(ref):
            assert _rel(K[b], Kr) < 1e-10 and _rel(d[b], dr) < 1e-10 and np.allclose(dV[b], dVr, rtol=1e-10)


def test_backward_pass_dev_unaligned_buffers_fall_back(engine):
    """Device buffers that start on an odd double cannot be bulk-copied (16-byte alignment): the launcher must fall back to
    the cp.async staging, with the same result."""
    import torch
    from oracle import riccati

    n, m, N = 56, 1, 12
    rng = np.random.default_rng(5)
    A, Bm, AB, Q, q, R, Eux, H, r = _lq_problem(rng, N, n, m)
    Kr, dr, dVr, fr = riccati.backward_pass(AB, Q, q, R, Eux, r, rho=0.3, reg_type=rbq.REG_CONTROL)
    dev = torch.device("cuda", 0)
    eng = rbq.Engine(0)
    eng.use_torch_stream()

    def odd(a):     # column-major blocks, placed one double into a fresh allocation
        flat = np.ascontiguousarray(np.swapaxes(a, -1, -2)).reshape(-1) if a.ndim >= 3 else np.ascontiguousarray(a).reshape(-1)
        buf = torch.empty(flat.size + 1, dtype=torch.float64, device=dev)
        buf[1:] = torch.from_numpy(flat).to(dev)
        assert buf[1:].data_ptr() % 16 == 8
        return buf[1:]

    K = torch.empty((N - 1) * n * m, dtype=torch.float64, device=dev)
    d = torch.empty((N - 1) * m, dtype=torch.float64, device=dev)
    dV = torch.empty(2, dtype=torch.float64, device=dev)
    fail = torch.empty(1, dtype=torch.int64, device=dev)
    eng.backward_pass_dev(n, m, N, 1, odd(AB), odd(Q), odd(q), odd(R), odd(Eux), odd(r), 0.3, rbq.REG_CONTROL, K, d, dV, fail)
    torch.cuda.synchronize()
    assert int(fail[0]) == -1
    Kh = K.cpu().numpy().reshape(N - 1, n, m).swapaxes(-1, -2)
    assert _rel(Kh, Kr) < 1e-10 and _rel(d.cpu().numpy().reshape(N - 1, m), dr) < 1e-10
    assert np.allclose(dV.cpu().numpy(), dVr, rtol=1e-10)


@pytest.mark.parametrize("n,N", [(43, 61), (56, 47), (58, 33)])
@pytest.mark.parametrize("reg", [rbq.REG_CONTROL, rbq.REG_STATE], ids=["control_reg", "state_reg"])
def test_backward_pass_cluster_form_matches_oracle_and_single_cta(engine, n, N, reg):
    """Few large problems run as a cluster of 2 or 4 CTAs per problem (column strips of T and G split over the CTAs, S kept
    in every CTA and updated through distributed shared memory).  Same recursion: every cluster size agrees with the oracle at
    the usual 1e-10 and with the one-CTA kernel to rounding; a failing knot is reported by every form."""
    from oracle import riccati

    rng = np.random.default_rng(4000 + n)
    probs = [_lq_problem(rng, N, n, 1) for _ in range(3)]
    stack = lambda i: np.stack([p[i] for p in probs])
    AB, Q, q, R, Eux, r = (stack(i) for i in (2, 3, 4, 5, 6, 8))
    ref = [riccati.backward_pass(p[2], p[3], p[4], p[5], p[6], p[8], rho=0.23, reg_type=reg) for p in probs]
    outs = {}
    for cl in ("1", "2", "4"):
        engine.set_option("RBQ_BP_CLUSTER", cl)
        outs[cl] = engine.backward_pass(AB, Q, q, R, Eux, r, rho=0.23, reg_type=reg)
    engine.set_option("RBQ_BP_CLUSTER", None)
    auto = engine.backward_pass(AB[0], Q[0], q[0], R[0], Eux[0], r[0], rho=0.23, reg_type=reg)      # one problem: the default picks a cluster
    assert auto[3] == -1 and _rel(auto[0], ref[0][0]) < 1e-10 and _rel(auto[1], ref[0][1]) < 1e-10
    for cl, (K, d, dV, fail) in outs.items():
        assert list(fail) == [-1] * 3, cl
        for b, (Kr, dr, dVr, fr) in enumerate(ref):
            assert _rel(K[b], Kr) < 1e-10 and _rel(d[b], dr) < 1e-10 and np.allclose(dV[b], dVr, rtol=1e-10), (cl, b)
        assert _rel(K, outs["1"][0]) < 1e-11
    R2 = R.copy(); R2[1, N // 2] = -1e12
    failr = riccati.backward_pass(AB[1], Q[1], q[1], R2[1], Eux[1], r[1], rho=0.23, reg_type=reg)[3]
    assert failr == N // 2
    for cl in ("1", "2", "4"):
        engine.set_option("RBQ_BP_CLUSTER", cl)
        K, d, dV, fail = engine.backward_pass(AB, Q, q, R2, Eux, r, rho=0.23, reg_type=reg)
        assert list(fail) == [-1, failr, -1], cl
        assert _rel(K[1, failr + 1:], outs["1"][0][1, failr + 1:]) < 1e-11
    engine.set_option("RBQ_BP_CLUSTER", None)
