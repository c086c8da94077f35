"""End-to-end consistency of the hot-path pieces: a plain iLQR loop on the reference's spin13 problem data (LQR objective of
spin13.jl:118-135, Z/2 target of spin13.jl:80-93, N = 301 knots, dt = 0.1 ns, U0 = 1e-4) driven entirely through the engine:

    rbq_rollout -> (LQR cost blocks on the host) -> rbq_jacobians -> rbq_backward_pass -> rbq_rollout_closedloop (line search)

Parity tests compare each piece with its oracle; this test checks that their CONVENTIONS agree with each other (layouts and
signs of A, B, K, d, the closed-loop law u = u_ref + K (x - x_ref) + alpha d): if any were off, the predicted decrease
alpha dV1 + alpha^2 dV2 would not match the realised one and the cost would not go down."""
import os

import numpy as np
import pytest

import rbqoc_b200 as rbq
from rbqoc_b200 import spin

pytestmark = pytest.mark.gpu


def test_ilqr_iterations_on_spin13_decrease_the_cost_as_predicted(engine):
    model = spin.Spin13Model()
    n, m, N, dt = 11, 1, 301, 0.1
    qs = [1.0, 1.0, 1.0, 1e-1, 1e-1]                                   # spin13.jl:66
    Qd = np.array([qs[0]] * 8 + [qs[1], qs[2], qs[3]])
    R = qs[4]
    g = spin.GATES[spin.GateType.zpiby2]
    xf = np.zeros(n)
    xf[0:4] = spin.get_vec_iso(g @ np.array([1.0, 0.0]))
    xf[4:8] = spin.get_vec_iso(g @ np.array([0.0, 1.0]))
    x0 = spin.spin13_x0()
    dts = np.full(N - 1, dt)

    def cost(X, U):
        dx = X - xf
        w = np.ones(N); w[-1] = N                                      # Qf = N Q (spin13.jl:134)
        return 0.5 * np.sum(w[:, None] * Qd * dx * dx) + 0.5 * R * np.sum(U * U)

    U = np.full((N - 1, m), 1e-4)
    X = spin.rollout(model, x0, U, dt, engine=engine)
    J = cost(X, U)
    J0 = J
    alphas = 0.5 ** np.arange(10)
    checked = 0
    for it in range(12):
        AB = spin.dynamics_expansion(model, X, U, dt, engine=engine)
        Exx = np.tile(np.diag(Qd), (N, 1, 1)); Exx[-1] *= N
        Ex = (X - xf) * Qd; Ex[-1] *= N
        Euu = np.full((N - 1, 1, 1), R); Eux = np.zeros((N - 1, 1, n)); Eu = R * U
        rho = 0.0
        while True:
            K, d, dV, fail = engine.backward_pass(AB, Exx, Ex, Euu, Eux, Eu, rho=rho)
            if fail < 0:
                break
            rho = max(10.0 * rho, 1e-6)                                # Altro's regularization_update!(:increase)
            assert rho < 1e6
        assert dV[0] < 0
        Xc, Uc = engine.rollout_closedloop(model.params, X, U, K, d, dts, alphas)
        Jc = np.array([cost(Xc[i], Uc[i]) for i in range(alphas.shape[0])])
        # quadratic model of the decrease along the search direction: dJ(alpha) ~ alpha dV1 + alpha^2 dV2; for small alpha the
        # linearisation error of the dynamics is O(alpha^2), so the ratio tends to 1
        small = alphas.shape[0] - 1
        pred = alphas[small] * dV[0] + alphas[small] ** 2 * dV[1]
        if abs(pred) > 1e-9 * J:
            assert (Jc[small] - J) / pred == pytest.approx(1.0, abs=0.15), (it, Jc[small] - J, pred)
            checked += 1
        ok = np.nonzero(Jc < J)[0]
        assert ok.size > 0, (it, J, Jc)
        i = int(ok[0])
        assert np.allclose(Xc[i, 0], x0)
        X, U, J = Xc[i], Uc[i], float(Jc[i])
    assert checked >= 6
    assert J < 0.2 * J0                                                # the loop makes real progress towards the Z/2 target
    # the closed-loop states ARE a rollout of the closed-loop controls
    assert np.max(np.abs(spin.rollout(model, x0, U, dt, engine=engine) - X)) < 1e-12


def test_device_resident_iteration_matches_host_pointer_iteration(engine):
    """One iLQR iteration with every array on the device (rbq_jacobians_dev -> rbq_backward_pass_dev ->
    rbq_rollout_closedloop_dev on a torch stream): the Jacobians and gains never cross PCIe, and the candidates are the bits
    the host-pointer calls return."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    model = spin.Spin12Model()
    n, m, N, dt = model.n, model.m, 181, 0.1                         # spin12.jl defaults: T = 18 ns
    rng = np.random.default_rng(5)
    U = 1e-3 * rng.standard_normal((N - 1, m))
    dts = np.full(N - 1, dt)
    X = spin.rollout(model, spin.spin12_x0(), U, dt, engine=engine)
    Qd = np.ones(n)
    xf = X[-1] * 0.9
    Exx = np.tile(np.diag(Qd), (N, 1, 1)); Ex = (X - xf) * Qd
    Euu = np.full((N - 1, 1, 1), 0.1); Eux = np.zeros((N - 1, m, n)); Eu = 0.1 * U
    alphas = 0.5 ** np.arange(4)
    # host-pointer reference of the same iteration
    AB = spin.dynamics_expansion(model, X, U, dt, engine=engine)
    K, d, dV, fail = engine.backward_pass(AB, Exx, Ex, Euu, Eux, Eu, rho=1e-3)
    assert fail == -1
    Xc, Uc = engine.rollout_closedloop(model.params, X, U, K, d, dts, alphas)
    # device-resident
    eng = rbq.Engine(0)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    col = lambda x: np.ascontiguousarray(np.swapaxes(x, -1, -2))
    with torch.cuda.stream(stream):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dX, dU, ddt, dal = t(X), t(U), t(dts), t(alphas)
        dExx, dEx, dEuu, dEux, dEu = t(col(Exx)), t(Ex), t(col(Euu)), t(col(Eux)), t(Eu)
        dAB = torch.empty(N - 1, n * (n + m), dtype=torch.float64, device=dev)
        dK = torch.empty(N - 1, n * m, dtype=torch.float64, device=dev); dd = torch.empty(N - 1, m, dtype=torch.float64, device=dev)
        ddV = torch.empty(2, dtype=torch.float64, device=dev); dfl = torch.empty(1, dtype=torch.int64, device=dev)
        dXc = torch.empty(4, N, n, dtype=torch.float64, device=dev); dUc = torch.empty(4, N - 1, m, dtype=torch.float64, device=dev)
    stream.synchronize()
    eng.jacobians_dev(model.params, N - 1, dX, dU, ddt, dAB)          # reads the first N-1 states of X in place
    eng.backward_pass_dev(n, m, N, 1, dAB, dExx, dEx, dEuu, dEux, dEu, 1e-3, rbq.REG_CONTROL, dK, dd, ddV, dfl)
    eng.rollout_closedloop_dev(model.params, N, dX, dU, dK, dd, ddt, 4, dal, dXc, dUc)
    eng.synchronize()
    assert int(dfl.cpu()[0]) == -1
    assert np.array_equal(dK.cpu().numpy().reshape(N - 1, n, m).transpose(0, 2, 1), K)
    assert np.array_equal(dXc.cpu().numpy(), Xc) and np.array_equal(dUc.cpu().numpy(), Uc)
    eng.close()


def test_al_ilqr_on_device_reduces_merit_as_predicted_and_constraint_violation(engine):
    """Augmented-Lagrangian iLQR on the spin13 problem with the reference's constraint list (spin13.jl:136-152), every inner
    iteration on the device: rbq_al_cost_expansion_dev -> rbq_jacobians_dev -> rbq_backward_pass_dev ->
    rbq_rollout_closedloop_dev -> rbq_al_cost_expansion_dev (values only, per candidate).  The merit function the cost
    kernel reports must fall as the backward pass predicts (its expansion IS the quadratic model the pass minimises), and
    the outer multiplier / penalty updates must drive the constraint violation down."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    eng = rbq.Engine(0); eng.use_torch_stream()
    model = spin.Spin13Model()
    n, m, N, dt = 11, 1, 301, 0.1
    g = spin.GATES[spin.GateType.zpiby2]
    xf = np.zeros(n)
    xf[0:4] = spin.get_vec_iso(g @ np.array([1.0, 0.0])); xf[4:8] = spin.get_vec_iso(g @ np.array([0.0, 1.0]))
    table = spin.spin13_constraints(N, xf).upload(dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    e = lambda *s: torch.empty(*s, dtype=torch.float64, device=dev)
    dQ, dR, dQf, dxf = t([1.0] * 8 + [1.0, 1.0, 0.1]), t([0.1]), t(np.array([1.0] * 8 + [1.0, 1.0, 0.1]) * N), t(xf)
    ddt = t(np.full(N - 1, dt))
    nal = 8
    dal = t(0.5 ** np.arange(nal))
    dU = t(np.full((N - 1, m), 1e-4))
    dX = e(1, N, n)
    eng.rollout_dev(model.params, 1, N, t(spin.spin13_x0()), dU, ddt, dX)
    dX = dX[0]
    lam = torch.zeros(table.dual_size, dtype=torch.float64, device=dev); mu = torch.ones(table.dual_size, dtype=torch.float64, device=dev)
    ineq = torch.zeros(table.dual_size, dtype=torch.bool, device=dev)
    for row, (o, kn, p) in zip(table.rows, table.dual_slices()):
        ineq[o:o + kn * p] = row.kind == rbq.CON_BOUND
    Exx, Ex, Euu, Eux, Eu = e(N, n * n), e(N, n), e(N - 1, m * m), e(N - 1, m * n), e(N - 1, m)
    c, J, cc, Jc = e(table.dual_size), e(N), e(nal, table.dual_size), e(nal, N)
    AB, K, d, dV, fl = e(N - 1, n * (n + m)), e(N - 1, n * m), e(N - 1, m), e(2), torch.empty(1, dtype=torch.int64, device=dev)
    Xc, Uc = e(nal, N, n), e(nal, N - 1, m)

    def violation(cv):
        v = torch.where(ineq, torch.clamp(cv, min=0.0), cv.abs())
        return float(torch.where(torch.isfinite(v), v, torch.zeros_like(v)).max())

    checked, viol = 0, []
    for outer in range(4):
        for inner in range(10):
            eng.al_cost_expansion_dev(n, m, N, dX, dU, dQ, dR, dQf, dxf, None, table, lam, mu, Exx, Ex, Euu, Eux, Eu, c, J)
            eng.jacobians_dev(model.params, N - 1, dX, dU, ddt, AB)
            rho = 0.0
            while True:
                eng.backward_pass_dev(n, m, N, 1, AB, Exx, Ex, Euu, Eux, Eu, rho, rbq.REG_CONTROL, K, d, dV, fl)
                if int(fl[0]) < 0:
                    break
                rho = max(10.0 * rho, 1e-6)
                assert rho < 1e8
            eng.rollout_closedloop_dev(model.params, N, dX, dU, K, d, ddt, nal, dal, Xc, Uc)
            merit = float(J.sum())
            dVh = dV.cpu().numpy()
            assert dVh[0] <= 0
            # the merit function of all candidates in one batched, values-only call (shared multipliers)
            eng.al_cost_expansion_dev(n, m, N, Xc, Uc, dQ, dR, dQf, dxf, None, table, lam, mu, None, None, None, None, None, cc, Jc, B=nal)
            Jcand = [float(v) for v in Jc.sum(dim=1)]
            a = float(dal[nal - 1])
            pred = a * dVh[0] + a * a * dVh[1]
            if abs(pred) > 1e-10 * abs(merit):
                # small step: the realised change of the merit function is the predicted one (the active set rarely flips)
                assert (Jcand[-1] - merit) / pred == pytest.approx(1.0, abs=0.05), (outer, inner, Jcand[-1] - merit, pred)
                checked += 1
            better = [i for i in range(nal) if Jcand[i] < merit]
            if not better:
                break
            i = better[0]
            dX, dU = Xc[i].clone(), Uc[i].clone()
            if os.environ.get("RBQ_TEST_VERBOSE"):
                print(outer, inner, "merit", merit, "->", Jcand[i], "alpha", float(dal[i]), "pred small", pred, "real small", Jcand[-1] - merit)
        eng.al_cost_expansion_dev(n, m, N, dX, dU, dQ, dR, dQf, dxf, None, table, lam, mu, None, None, None, None, None, c, J)
        torch.cuda.synchronize()
        viol.append(violation(c))
        if os.environ.get("RBQ_TEST_VERBOSE"):
            print("outer", outer, "violation", viol[-1], "merit", float(J.sum()))
        # dual update, then penalty x10, on the device; checked against the same update in torch
        cfin = torch.where(torch.isfinite(c), c, torch.zeros_like(c))
        lam_ref = torch.where(ineq, torch.clamp(lam + mu * cfin, min=0.0), lam + mu * cfin)
        mu_ref = torch.where(torch.isfinite(c), mu * 10.0, mu)
        dviol = torch.empty(1, dtype=torch.float64, device=dev)
        eng.al_dual_update_dev(N, table, c, lam, mu, dviol, phi=10.0)
        torch.cuda.synchronize()
        assert float(dviol[0]) == viol[-1]
        assert torch.allclose(lam, lam_ref, rtol=1e-12, atol=1e-14) and torch.equal(mu, mu_ref)
    assert checked >= 10
    assert viol[-1] < 0.5 * viol[0] and all(b <= a for a, b in zip(viol, viol[1:])), viol
    eng.close()
