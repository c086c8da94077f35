"""CPU suite, part 1: the oracle itself.

The reference has no tests, fixtures or stored pulses and cannot run here (Julia), so the oracle is
"parity unpinned" against reference OUTPUTS.  It is pinned instead against
  (a) the known answers embedded in the reference (analytic gates, spin.jl:557,608-612; D1 table;
      1 % detuning gate error),
  (b) golden vectors from a 50-digit restatement of the same mathematics (tests/golden/make_golden.py),
  (c) scipy's expm,
  (d) pulses produced by the reference's own spin14.py generator functions.
"""
import numpy as np
import pytest
import scipy.linalg
import scipy.signal

from _cases import GOLDEN, MODEL_KW, golden_model_cases, relerr

FQ = 1.4e-2


@pytest.fixture(scope="module")
def orc(oracle):
    return oracle


def test_expm_matches_scipy_and_picks_reference_branches(orc):
    rng = np.random.default_rng(0)
    H0, H1 = np.pi * np.array(orc_N0()), np.pi * np.array(orc_N1())
    seen = set()
    for dt, a in [(0.01, 0.0), (0.01, 0.49), (0.1, 0.02), (0.1, 0.5), (0.5, 0.3), (1.0, 0.3), (2.0, 0.5), (5.0, 0.5), (40.0, 0.5)]:
        G = dt * (FQ * H0 + a * H1)
        E, deg = orc.expm4(G)
        seen.add(deg)
        assert np.max(np.abs(E - scipy.linalg.expm(G))) < 5e-14 * max(1.0, np.abs(G).sum(axis=0).max())
        assert np.max(np.abs(E @ E.T - np.eye(4))) < 1e-13
    assert seen == {3, 5, 7, 9, 13}          # every StaticArrays branch exercised
    A = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    Ec, _ = orc.expm4c(0.3 * A)
    assert np.max(np.abs(Ec - scipy.linalg.expm(0.3 * A))) < 1e-13


REF_SPIN26 = "/root/reference/src/spin/spin26.jl"


@pytest.mark.skipif(not __import__("os").path.exists(REF_SPIN26), reason="reference tree not present (GPU box)")
def test_pade13_table_equals_the_reference_in_tree_copy(orc):
    """The reference carries Higham's degree-13 table at spin26.jl:25-40; the oracle's expm uses the same numbers."""
    import re

    src = open(REF_SPIN26).read()
    body = src[src.index("const B = ("):]
    body = body[:body.index(")")]
    ref = [float(t) for t in re.findall(r"\b\d+\b", body.split("(", 1)[1])]
    assert len(ref) == 14
    assert list(orc.pade13_coeffs()) == ref
    # the lower-degree tables are the classical ones: p_m(x) = sum_j (2m-j)! m! / ((2m)! (m-j)! j!) x^j, scaled to integers
    from math import factorial as f

    for m, table in ((3, [120, 60, 12, 1]), (5, [30240, 15120, 3360, 420, 30, 1])):
        c = [f(2 * m - j) * f(m) / (f(2 * m) * f(m - j) * f(j)) for j in range(m + 1)]
        assert np.allclose(np.array(c) / c[-1] * table[-1], table, rtol=1e-15)


def orc_N0():
    return [[0, 0, 1, 0], [0, 0, 0, -1], [-1, 0, 0, 0], [0, 1, 0, 0]]


def orc_N1():
    return [[0, 0, 0, 1], [0, 0, 1, 0], [0, -1, 0, 0], [-1, 0, 0, 0]]


def test_analytic_gates_embedded_in_reference(orc):
    """spin.jl:557: exp(TTOT_ZPIBY2 * FQ * H0) = Z/2;  spin.jl:608-612: three-segment Y/2."""
    Z = orc.propagator(FQ, 0.0, 17.857142857142858)
    assert np.max(np.abs(Z - orc.gate_iso(1))) < 5e-16
    tx, tz, A = 2.1656249366575766, 15.1423305995572655, 0.125
    Y = orc.propagator(FQ, -A, tx) @ orc.propagator(FQ, 0.0, tz) @ orc.propagator(FQ, A, tx)
    assert np.max(np.abs(Y - orc.gate_iso(2))) < 1e-15


def test_t1_table_and_d1_pins(orc):
    """D1 = int dt/T1(a): trials.xlsx spin14 (5.745 / 5.071 / 15.887 e-5), paper main.tex:1077."""
    assert orc.t1_ns(0.0) == pytest.approx(310.88e3, rel=1e-15)
    assert orc.t1_reduced_us(0.0) == pytest.approx(310.88, rel=1e-15)
    assert orc.t1_ns(5.0) == pytest.approx(1597.923e3, rel=1e-15)      # flat extrapolation below the table
    assert 17.857142857142858 / orc.t1_ns(0.0) == pytest.approx(5.7441e-5, rel=1e-4)
    tx, tz = 2.1656249366575766, 15.1423305995572655
    d1_y = 2 * tx / orc.t1_ns(0.125) + tz / orc.t1_ns(0.0)
    assert d1_y == pytest.approx(5.0722e-5, rel=1e-4)
    d1_x = 2 * d1_y + 17.857142857142858 / orc.t1_ns(0.0)
    assert d1_x == pytest.approx(15.888e-5, rel=1e-4)


def test_detuning_gate_error_pin(orc):
    """paper main.tex:857-859: analytic Z/2 at +-1 % detuning, mean over random states ~ 4.7e-5."""
    rng = np.random.default_rng(1)
    S = 1000
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    psi0 = np.concatenate([psi.real, psi.imag], axis=1)
    n = 100
    ge = []
    for f in (1.01 * FQ, 0.99 * FQ):
        fid = orc.mc_static(np.zeros(n), 17.857142857142858 / n, 1, 1, np.full(S, f), None, psi0)
        ge.append(1 - fid[:, 1])
    assert np.mean(ge) == pytest.approx(4.66e-5, rel=0.03)


@pytest.mark.parametrize("case", golden_model_cases(), ids=lambda c: c[0])
def test_oracle_step_and_jacobian_match_50_digit_golden(orc, case):
    key, name, dt, x, u, xn, J, t = case
    p = orc.make_params(**MODEL_KW[name])
    got = orc.step(p, x, u, dt, t=t)
    assert relerr(got, xn) < 2e-15, key
    Jg = orc.jacobian(p, x, u, dt, t=t)
    assert Jg.shape == J.shape
    assert relerr(Jg, J) < 1e-13, key


def test_forwarddiff_conventions_at_kinks(orc):
    """abs'(+0.0) = +1, abs'(-0.0) = -1 (ForwardDiff); the T1 table uses the interval searchsortedlast picks."""
    p = orc.make_params(model_id=15, decay_aware=1)
    x = np.zeros(12)
    x[:8] = [1, 0, 0, 0, 0, 1, 0, 0]
    slope_last = (310.88 - 560.251) / (0.5 - 0.492)            # last table interval, us per fbfq
    for a, sgn in ((0.0, 1.0), (-0.0, -1.0)):
        x[9] = a
        J = orc.jacobian(p, x, np.zeros(1), 0.1)
        # d/da [dt / T1(fbfq(a))] = -dt T1' fbfq'(a) / T1^2, fbfq' = -sgn * 0.202407
        want = -0.1 * slope_last * (-sgn * 0.202407) / 310.88**2
        assert J[11, 9] == pytest.approx(want, rel=1e-12)


def test_mc_static_and_lindblad_match_50_digit_golden(orc):
    z = np.load(GOLDEN + "/mc_static.npz")
    fid, states = orc.mc_static(z["controls"], float(z["dt"]), int(z["gate"]), int(z["gate_count"]), z["fq"], z["da"], z["psi0"],
                                want_states=True)
    # ~1950 generic Pade exponentials and 4x4 products per sample: the reference algorithm's own rounding
    assert np.max(np.abs(fid - z["fid"])) < 2e-12
    assert np.max(np.abs(states - z["states"])) < 2e-12
    z = np.load(GOLDEN + "/lindblad.npz")
    G = int(z["gate_count"])
    fid, fidc, rho = orc.lindblad_t1(z["controls"], float(z["dt"]), int(z["gate"]), G, float(z["fq"]), z["da"], z["rho0"])
    assert np.max(np.abs(rho - z["rho"][:, G])) < 3e-12
    # closed-form 2x2 Uhlmann fidelity == the reference's sqrtm route where the latter is well conditioned (mixed states)
    assert np.max(np.abs(fid[2:] - fidc[2:])) < 1e-9


def test_fidelity_and_cyclic_targets(orc):
    s = np.array([0.6, 0.0, 0.0, 0.8])
    assert orc.fidelity_vec_iso2(s, s) == pytest.approx(1.0, abs=1e-15)
    assert orc.fidelity_vec_iso2(s, np.array([0.0, 0.6, -0.8, 0.0])) < 1.0
    # an exact gate applied k times against the 4-cyclic targets stays at fidelity 1 (spin.jl:1198-1240)
    fid = orc.mc_static(np.zeros(50), 17.857142857142858 / 50, 1, 9, np.array([FQ]), None, np.array([[0.6, 0.0, 0.0, 0.8]]))
    assert np.max(np.abs(fid - 1)) < 1e-13


def test_gate_error_expansion_is_consistent(orc):
    """gate_error_iso2 = 1 - fidelity_vec_iso2; the analytic gradient (spin.jl:1060-1071) matches central differences; the
    reference's Hessian (spin.jl:1074-1092) is constant in s1 and symmetric."""
    rng = np.random.default_rng(4)
    for _ in range(10):
        s1, s2 = rng.standard_normal(4), rng.standard_normal(4)
        err, grad, hess = orc.gate_error_expansion(s1, s2)
        assert err == pytest.approx(1 - orc.fidelity_vec_iso2(s1, s2), abs=1e-14)
        h = 1e-6
        fd = np.array([(orc.gate_error_expansion(s1 + h * e, s2)[0] - orc.gate_error_expansion(s1 - h * e, s2)[0]) / (2 * h)
                       for e in np.eye(4)])
        assert np.allclose(grad, fd, rtol=1e-6, atol=1e-7)
        assert np.array_equal(hess, hess.T)
        # second differences: the exact Hessian of 1 - a^2 - b^2 in s1 (what the closed form of spin.jl:1074-1092 encodes)
        fd2 = np.array([(orc.gate_error_expansion(s1 + h * e, s2)[1] - orc.gate_error_expansion(s1 - h * e, s2)[1]) / (2 * h)
                        for e in np.eye(4)])
        assert np.allclose(hess, fd2, rtol=1e-6, atol=1e-7)


def test_philox_known_answer(orc):
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    assert list(orc.philox4x32_10([0, 0, 0, 0], [0, 0])) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert list(orc.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert list(orc.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_pink_filter_is_the_reference_iir(orc):
    x = np.zeros(16)
    x[0] = 1.0
    y = orc.pink_filter(x, 1.0)
    b = [0.049922035, -0.095993537, 0.050612699, -0.004408786]
    a = [1.0, -2.494956002, 2.017265875, -0.522189400]
    assert np.allclose(y, scipy.signal.lfilter(b, a, x), rtol=1e-13, atol=1e-18)
    assert np.allclose(orc.pink_filter(x, 10.0), 10.0 * y, rtol=1e-15)


def test_stats_definition(orc):
    e = np.array([1e-17, 1e-16, 3.2e-5, 0.5, 1.0, np.nan, 0.0])
    st = orc.stats(e)
    assert st.count == 6 and st.nonfinite == 1
    assert st.hist[0] == 3            # 1e-17 (clamped), 1e-16 and 0.0
    assert st.hist[int(np.floor((np.log10(3.2e-5) + 16) * 16))] == 1
    assert st.hist[255] == 1          # 1.0 clamps into the last bin
    assert st.min == 0.0 and st.max == 1.0


def _spin():
    from rbqoc_b200 import spin     # plain data (segment lists, constants): importing it needs no GPU

    return spin


def test_analytic_gate_segments_are_exact_and_match_50_digit_golden(orc):
    """integrate_prop_zpiby2nodis!/…t1! family (spin.jl:562-594, 651-670, 765-785) as segment lists: at the nominal
    qubit frequency the analytic gates are exact (the known answer the reference embeds); with detuning/amplitude
    offsets and T1 dissipation the oracle must sit on the 50-digit referee."""
    spin = _spin()
    z = np.load(GOLDEN + "/segments.npz")
    G = int(z["gate_count"])
    for gate in (spin.GateType.zpiby2, spin.GateType.ypiby2, spin.GateType.xpiby2):
        amps, durs = spin.ANALYTIC_SEGMENTS[gate]
        fid = orc.mc_static_segments(amps, durs, gate, G, z["fq"], z["da"], z["psi0"])
        assert np.max(np.abs(fid[0] - 1.0)) < 5e-15, gate.name            # sample 0: fq = FQ, da = 0
        assert np.max(np.abs(fid - z[f"fid_{gate.name}"])) < 1e-14, gate.name
        fidc, rho = orc.lindblad_segments(amps, durs, gate, G, FQ, z["da"], z["rho0"])
        assert np.max(np.abs(rho - z[f"rho_{gate.name}"][:, G])) < 5e-14, gate.name
        assert np.all(fidc[:, 0] > 1 - 1e-12) and np.all(np.diff(fidc[0]) < 0)   # T1 decay lowers a pure state's fidelity


def test_xpiby2da_segment_list_and_oracle(orc):
    """integrate_prop_xpiby2da! (spin.jl:819-905): 569 noise bins + 6 amplitude switches; the same noise indices every gate."""
    spin = _spin()
    amps, durs, nidx = spin.xpiby2da_segments()
    assert amps.shape == (575,) and nidx[0] == 0 and nidx[-1] == 568
    assert abs(durs.sum() - spin.TTOT_XPIBY2_JL) < 1e-13
    assert durs[21] == spin.TX_YPIBY2 - 2.1 and durs[22] == 2.2 - spin.TX_YPIBY2 and nidx[21] == nidx[22] == 21
    assert durs[-1] == spin.TTOT_XPIBY2_JL - 56.8
    z = np.load(GOLDEN + "/segments.npz")
    G = int(z["da_gate_count"])
    fid = orc.mc_noise_segments(amps, durs, nidx, spin.GateType.xpiby2, G, FQ, z["da_noise"], z["psi0"][:3])
    # 575 Pade exponentials per gate: the reference algorithm's own rounding floor (~2e-16 per knot)
    assert np.max(np.abs(fid - z["da_fid"])) < 1e-12
    # without noise it is the exact X/2 gate
    fid0 = orc.mc_noise_segments(amps, durs, nidx, spin.GateType.xpiby2, 4, FQ, np.zeros((3, 569)), z["psi0"][:3])
    assert np.max(np.abs(fid0 - 1.0)) < 1e-12
    with pytest.raises(ValueError):
        orc.mc_noise_segments(amps, durs, nidx, spin.GateType.xpiby2, 1, FQ, np.zeros((1, 100)), z["psi0"][:1])


def _lq_problem(rng, N, n, m):
    """Random time-varying LQ problem: x' = A_k x + B_k u, cost sum (x'Qx/2 + q'x + u'Ru/2 + r'u + u'Hx)."""
    A = 0.9 * np.eye(n) + 0.1 * rng.standard_normal((N - 1, n, n))
    Bm = rng.standard_normal((N - 1, n, m))
    AB = np.concatenate([A, Bm], axis=2)                                  # [A_k B_k]
    Q = rng.standard_normal((N, n, n)); Q = Q @ Q.transpose(0, 2, 1) / n + 0.1 * np.eye(n)
    R = rng.standard_normal((N - 1, m, m)); R = R @ R.transpose(0, 2, 1) + np.eye(m)
    H = 0.05 * rng.standard_normal((N - 1, m, n))
    q = rng.standard_normal((N, n)); r = rng.standard_normal((N - 1, m))
    return A, Bm, AB, Q, q, R, H, H, r


def test_riccati_oracle_solves_lq_problems_in_one_pass():
    """Defining property of the recursion (Altro backwardpass!): on a linear-quadratic problem the gains of one backward
    pass, rolled forward from x0 with u = K x + d, give the global optimum: the gradient of the total cost with respect to
    every control vanishes, and the predicted decrease dV matches the realised one."""
    from oracle import riccati

    rng = np.random.default_rng(5)
    N, n, m = 12, 5, 2
    A, Bm, AB, Q, q, R, Eux, H, r = _lq_problem(rng, N, n, m)
    K, d, dV, fail = riccati.backward_pass(AB, Q, q, R, Eux, r)
    assert fail == -1

    def cost(U, x0):
        x, J = x0, 0.0
        for k in range(N - 1):
            J += 0.5 * x @ Q[k] @ x + q[k] @ x + 0.5 * U[k] @ R[k] @ U[k] + r[k] @ U[k] + U[k] @ H[k] @ x
            x = A[k] @ x + Bm[k] @ U[k]
        return J + 0.5 * x @ Q[N - 1] @ x + q[N - 1] @ x

    x0 = np.zeros(n)                       # expansion point: (x, u) = 0, so the E blocks ARE the cost derivatives there
    x, U = x0, np.zeros((N - 1, m))
    for k in range(N - 1):
        U[k] = K[k] @ x + d[k]
        x = A[k] @ x + Bm[k] @ U[k]
    h = 1e-6
    grad = np.array([[(cost(U + h * np.eye(1, (N - 1) * m, i).reshape(N - 1, m), x0) -
                       cost(U - h * np.eye(1, (N - 1) * m, i).reshape(N - 1, m), x0)) / (2 * h)] for i in range((N - 1) * m)])
    assert np.max(np.abs(grad)) < 1e-7
    assert cost(U, x0) - cost(np.zeros((N - 1, m)), x0) == pytest.approx(dV[0] + dV[1], rel=1e-10)
    # an indefinite Quu stops the pass at that knot; control regularisation repairs it
    R2 = R.copy(); R2[7] = -50.0 * np.eye(m)
    assert riccati.backward_pass(AB, Q, q, R2, Eux, r)[3] == 7
    assert riccati.backward_pass(AB, Q, q, R2, Eux, r, rho=100.0)[3] == -1


def test_deqjl_timeline_is_the_exact_solution_of_the_reference_ode(orc):
    """run_sim_deqjl (spin.jl:1277-1380) hands Vern9 the right-hand side of spin.jl:424-467:
        d psi/dt = (fq NEGI_H0 + (controls[floor(t dt_inv) mod count] + noise[floor(t noise_dt_inv)]) NEGI_H1) psi.
    The segment list of spin.deqjl_timeline + the oracle's per-segment exponentials must reproduce an adaptive high-order
    integration of exactly that right-hand side (scipy DOP853, restarted at the breakpoints so that it keeps its order)."""
    import scipy.integrate

    spin = _spin()
    rng = np.random.default_rng(12)
    H0, H1 = np.pi * np.array(orc_N0()), np.pi * np.array(orc_N1())
    cdi, ndi, gate_time, gates = 10.0, 4.0, 2.37, 3                     # 23 control knots per period, wrap-around of 0.07
    controls = 0.3 * rng.standard_normal(40)
    count = int(np.floor(gate_time * cdi))
    nlen = int(np.ceil(gate_time * gates * ndi)) + 1
    noise = 0.05 * rng.standard_normal((2, nlen))
    fq = np.array([FQ, 1.03 * FQ])
    psi = rng.random((2, 2)) + 1j * rng.random((2, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    psi0 = np.concatenate([psi.real, psi.imag], axis=1)
    amps, durs, nidx, save_after = spin.deqjl_timeline(controls, cdi, gate_time, gates, ndi)
    assert abs(durs.sum() - gates * gate_time) < 1e-12 and save_after[-1] == durs.shape[0]
    fid = orc.mc_timeline(amps, durs, nidx, save_after, spin.GateType.xpiby2, fq, None, noise, psi0)

    def rhs(t, y, s):                                                  # the reference's indexing, literally
        a = controls[int(np.floor(t * cdi)) % count] + noise[s, int(np.floor(t * ndi))]
        return (fq[s] * H0 + a * H1) @ y

    edges = np.concatenate([[0.0], np.cumsum(durs)])
    G = np.array(spin.GATES[spin.GateType.xpiby2])
    for s in range(2):
        y = psi0[s].copy()
        want = [1.0]
        k = 0
        for j in range(durs.shape[0]):
            tm = 0.5 * (edges[j] + edges[j + 1])                       # constant on the segment: evaluate the indices at its midpoint
            sol = scipy.integrate.solve_ivp(lambda t, y: rhs(tm, y, s), (edges[j], edges[j + 1]), y, method="DOP853", rtol=1e-13, atol=1e-14)
            y = sol.y[:, -1]
            if j + 1 == save_after[k]:
                k += 1
                c = y[:2] + 1j * y[2:]
                tgt = np.linalg.matrix_power(G, k % 4) @ (psi0[s, :2] + 1j * psi0[s, 2:])
                want.append(abs(np.vdot(tgt, c)) ** 2)
        assert np.max(np.abs(fid[s] - np.array(want))) < 1e-10
    # the timeline indices are the reference's: spot-check interior times of every segment
    for j in rng.integers(0, durs.shape[0], 50):
        t = edges[j] + rng.uniform(0.05, 0.95) * durs[j]
        assert amps[j] == controls[int(np.floor(t * cdi)) % count] and nidx[j] == int(np.floor(t * ndi))


# ---- augmented-Lagrangian cost expansion (oracle/al_cost.py) ----------------------------------------------------------
def _al_problem(rng, N=9, n=11, m=1):
    """spin13-shaped problem with every constraint type, some bounds active, non-zero multipliers."""
    from rbqoc_b200 import spin

    X = 0.6 * rng.standard_normal((N, n)); U = rng.standard_normal((N - 1, m))
    X[:, 9] = np.linspace(-0.8, 0.8, N)                       # the bounded control state crosses both bounds
    xf = rng.standard_normal(n)
    table = spin.spin13_constraints(N, xf)
    table.bound(0, N, u_max=[0.3], u_min=[-0.2])              # a control bound that also covers the last knot
    lam = rng.standard_normal(table.dual_size) * 0.3
    mu = 10.0 ** rng.uniform(-1, 2, table.dual_size)
    Qd, Rd, Qfd = rng.uniform(0.1, 2, n), rng.uniform(0.1, 2, m), rng.uniform(1, 20, n)
    w = rng.uniform(0.05, 0.2, N)
    duals = [(lam[o:o + k * p].reshape(k, p), mu[o:o + k * p].reshape(k, p)) for o, k, p in table.dual_slices()]
    return X, U, Qd, Rd, Qfd, xf, w, table, lam, mu, duals


def test_al_cost_gradient_matches_finite_differences_of_the_merit_function():
    from oracle import al_cost

    rng = np.random.default_rng(11)
    X, U, Qd, Rd, Qfd, xf, w, table, lam, mu, duals = _al_problem(rng)
    J, Exx, Ex, Euu, Eux, Eu, c = al_cost.expansion(X, U, Qd, Rd, Qfd, xf, table.specs, duals, w)
    assert np.all(Eux == 0)
    h = 1e-6
    for k, i in ((0, 3), (2, 9), (4, 9), (5, 0), (7, 9), (8, 2), (8, 8), (3, 6)):
        Xp, Xm = X.copy(), X.copy(); Xp[k, i] += h; Xm[k, i] -= h
        fd = (al_cost.expansion(Xp, U, Qd, Rd, Qfd, xf, table.specs, duals, w)[0].sum()
              - al_cost.expansion(Xm, U, Qd, Rd, Qfd, xf, table.specs, duals, w)[0].sum()) / (2 * h)
        assert abs(fd - Ex[k, i]) < 1e-6 * max(1.0, abs(fd)), (k, i, fd, Ex[k, i])
    for k in (0, 3, 7):
        Up, Um = U.copy(), U.copy(); Up[k, 0] += h; Um[k, 0] -= h
        fd = (al_cost.expansion(X, Up, Qd, Rd, Qfd, xf, table.specs, duals, w)[0].sum()
              - al_cost.expansion(X, Um, Qd, Rd, Qfd, xf, table.specs, duals, w)[0].sum()) / (2 * h)
        assert abs(fd - Eu[k, 0]) < 1e-6 * max(1.0, abs(fd))


def test_al_cost_hessian_exact_for_linear_constraints_and_active_set_rule():
    from oracle import al_cost

    rng = np.random.default_rng(12)
    X, U, Qd, Rd, Qfd, xf, w, table, lam, mu, duals = _al_problem(rng)
    lin = [i for i, s in enumerate(table.specs) if s["kind"] != "norm"]
    specs = [table.specs[i] for i in lin]; dl = [duals[i] for i in lin]
    J, Exx, Ex, Euu, Eux, Eu, c = al_cost.expansion(X, U, Qd, Rd, Qfd, xf, specs, dl, w)
    h = 1e-4
    k = 4                                                  # second difference of the merit function along e_9 at an interior knot
    f = lambda d: al_cost.expansion(X + d, U, Qd, Rd, Qfd, xf, specs, dl, w)[0].sum()
    e = np.zeros_like(X); e[k, 9] = h
    assert abs((f(e) - 2 * f(0 * e) + f(-e)) / h ** 2 - Exx[k, 9, 9]) < 1e-5 * Exx[k, 9, 9]
    # inactive inequality (c < 0, lambda <= 0): no penalty curvature; active through lambda > 0 alone: curvature mu
    o, kn, p = table.dual_slices()[0]
    kk = 3 - table.specs[0]["k0"]
    cu = X[3, 9] - 0.5
    assert cu < 0
    lam2 = dl[0][0].copy(); lam2[kk, 9] = -0.1
    E0 = al_cost.expansion(X, U, Qd, Rd, Qfd, xf, specs, [(lam2, dl[0][1])] + dl[1:], w)[1][3, 9, 9]
    lam2[kk, 9] = +0.1
    E1 = al_cost.expansion(X, U, Qd, Rd, Qfd, xf, specs, [(lam2, dl[0][1])] + dl[1:], w)[1][3, 9, 9]
    assert np.isclose(E1 - E0, dl[0][1][kk, 9])
    # an absent bound writes -inf and contributes nothing
    assert np.all(np.isneginf(c[0][:, 0])) and np.isfinite(J).all()


# ---- pulse I/O mirror (rbqoc.jl:98-139, spin.jl:1485-1530) --------------------------------------------------------------
def test_grab_and_sample_controls_follow_the_reference_save_formats():
    from rbqoc_b200 import spin

    N, n, dt = 31, 11, 0.1
    t = np.arange(N) * dt
    a = 0.02 * t ** 3 - 0.05 * t ** 2 + 0.1 * t                     # a cubic: an interpolating cubic spline reproduces it
    astates = np.zeros((N, n)); astates[:, 9] = a
    acontrols = np.zeros((N - 1, 1)); acontrols[:, 0] = (0.12 * t - 0.1)[:-1]
    save = {"save_type": 1, "astates": astates, "acontrols": acontrols, "controls_idx": [10], "d2controls_dt2_idx": [1],
            "dt": dt, "evolution_time": (N - 1) * dt}
    c, dt_inv, T = spin.grab_controls(save)
    assert c.shape == (N - 1, 1) and np.array_equal(c[:, 0], a[:-1]) and dt_inv == pytest.approx(10.0) and T == pytest.approx(3.0)
    cs, ds, Tf = spin.sample_controls(save)
    count = int(np.floor(Tf * 100.0))
    assert Tf == pytest.approx(3.0) and cs.shape == (count, 1) and ds.shape == (count, 1)
    ts = np.arange(count) * 1e-2
    # the reference fits the spline on the first N-1 knots and evaluates up to T - dt_pref: the last 0.1 ns is extrapolated
    assert np.max(np.abs(cs[:, 0] - (0.02 * ts ** 3 - 0.05 * ts ** 2 + 0.1 * ts))) < 1e-12
    assert np.max(np.abs(ds[:, 0] - (0.12 * ts - 0.1))) < 1e-12
    # per-knot dt stored as a control column (time-optimal runs)
    ac2 = np.hstack([acontrols, np.full((N - 1, 1), dt)])
    save2 = dict(save, acontrols=ac2, dt_idx=[2]); del save2["dt"]
    cs2, _, Tf2 = spin.sample_controls(save2)
    assert Tf2 == pytest.approx(Tf) and np.allclose(cs2, cs, atol=1e-13)
    # the resampled and the python formats
    c3, dti3, T3 = spin.grab_controls({"save_type": 2, "controls_sample": cs, "evolution_time_sample": Tf})
    assert c3.shape == (count - 1, 1) and dti3 == 100.0 and T3 == Tf
    c4, _, _ = spin.grab_controls({"save_type": 3, "controls": cs.T, "evolution_time": Tf})
    assert np.array_equal(c4, cs)
    with pytest.raises(FileNotFoundError):
        spin.read_save("/nonexistent/pulse.h5")


def test_reference_constraint_lists_have_the_reference_layout():
    """spin13.jl:136-152, spin12.jl:280-296, spin30.jl:301-320: knot ranges (1-based 2:N-2, N-1:N-1, N:N, 2:N-1) and the index
    sets of the goal and unit-norm constraints."""
    from rbqoc_b200 import _lib, spin

    N = 21
    for make, n, ctrl, goal, blocks in (
            (spin.spin13_constraints, 11, 9, list(range(9)), [0, 4]),
            (spin.spin12_constraints, 43, 9, list(range(9)), [0, 4] + list(range(11, 43, 4))),
            (spin.spin30_constraints, 56, 13, list(range(8)) + [12], list(range(15, 55, 4)) + [0, 4, 8])):
        t = make(N, np.arange(n, dtype=float))
        assert t.n == n and t.ncon == 3 + len(blocks)
        b0, b1, g = t.specs[0], t.specs[1], t.specs[2]
        assert (b0["k0"], b0["k1"]) == (1, N - 2) and (b1["k0"], b1["k1"]) == (N - 2, N - 1) and (g["k0"], g["k1"]) == (N - 1, N)
        fin = np.isfinite(b0["zmax"])
        assert list(np.nonzero(fin)[0]) == [ctrl] and b0["zmax"][ctrl] == 0.5 and b0["zmin"][ctrl] == -0.5
        assert b1["zmax"][ctrl] == 0.0 and b1["zmin"][ctrl] == 0.0
        assert g["idx"] == goal and np.array_equal(g["target"], np.asarray(goal, dtype=float))
        norms = t.specs[3:]
        assert sorted(s["idx"][0] for s in norms) == sorted(blocks)
        assert all(s["kind"] == "norm" and (s["k0"], s["k1"]) == (1, N - 1) and s["val"] == 1.0 and len(s["idx"]) == 4 for s in norms)
        # the packed rows agree with the description, duals laid out back to back
        off = 0
        for row, spec in zip(t.rows, t.specs):
            assert row.dual_off == off and (row.k0, row.k1) == (spec["k0"], spec["k1"])
            off += (row.k1 - row.k0) * row.p
        assert off == t.dual_size and t.rows[0].kind == _lib.CON_BOUND and t.rows[0].p == 2 * (n + 1)


# ---- density ODEs with dephasing (dynamics_lindbladt1_deqjl / dynamics_lindbladt2_deqjl, spin.jl:505-518, 542-554) ----------
def _mixed_rho(S, seed):
    rng = np.random.default_rng(seed)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho = np.einsum("si,sj->sij", psi, psi.conj())
    lam = rng.uniform(0.6, 1.0, S)[:, None, None]
    return lam * rho + (1 - lam) * np.eye(2) / 2


def test_density_ode_oracle_equals_the_vec_propagator_oracle_for_t1(orc):
    """Two different reference functions, one answer: the ODE right-hand side on the 4x4 iso density (spin.jl:505-518)
    solved exactly, and integrate_prop_lindbladt1!'s complex vec-Liouvillian exponentials (spin.jl:521-539)."""
    from rbqoc_b200 import spin

    controls = spin.pulse_sin(301)[:-1]
    rho0 = _mixed_rho(5, 3)
    da = 1e-3 * np.random.default_rng(1).standard_normal(5)
    _, fc1, r1 = orc.lindblad_t1(controls, 0.1, 3, 3, FQ, da, rho0)
    fc2, r2 = orc.lindblad_knots(controls, 0.1, 3, 3, FQ, da, rho0, channels=1)
    assert np.max(np.abs(r1 - r2)) < 5e-14 and np.max(np.abs(fc1 - fc2)) < 5e-14


def test_density_ode_oracle_matches_an_adaptive_integration_of_the_reference_rhs(orc):
    """scipy DOP853 (rtol = atol = 1e-13) on the right-hand side written as the reference writes it, T2 with the drive on and
    T1+T2: what Vern9 at reltol 1e-12 would return, to the solver's tolerance."""
    from scipy.integrate import solve_ivp

    from rbqoc_b200 import spin

    N0 = np.pi * np.array(orc_N0()); N1 = np.pi * np.array(orc_N1())
    GE = np.array([[0, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 1], [0, 0, 0, 0]], dtype=float); EG = GE.T
    NGG = -0.5 * np.diag([1.0, 0, 1, 0]); NEE = -0.5 * np.diag([0.0, 1, 0, 1])
    NEG_DOP = -np.array([[0, 1, 0, 1], [1, 0, 1, 0], [0, 1, 0, 1], [1, 0, 1, 0]], dtype=float)
    controls = spin.pulse_sin(61)[:-1]
    dt, nk = 0.1, 60
    gc, gf = spin.GAMMAC, 30 * spin.GAMMAF_PREFACTOR
    rho0 = _mixed_rho(2, 9)
    for channels in (2, 3):
        fid, rho = orc.lindblad_knots(controls, dt, 1, 2, FQ, None, rho0, channels=channels, gamma_c=gc, gamma_f=gf)

        def rhs(t, y):
            m = y.reshape(4, 4)
            k = int(np.floor(t / dt + 1e-9)) % nk
            a = controls[k]
            h = FQ * N0 + a * N1
            out = h @ m - m @ h + (gc + 2 * gf * gf * t) * NEG_DOP * m
            if channels & 1:
                g1 = 1.0 / orc.t1_ns(a)
                out = out + g1 * (GE @ m @ EG + NEE @ m + m @ NEE) + g1 * (EG @ m @ GE + NGG @ m + m @ NGG)
            return out.reshape(-1)

        for s in range(2):
            r = rho0[s]
            y = np.block([[r.real, -r.imag], [r.imag, r.real]]).reshape(-1)
            t = 0.0
            for _ in range(2 * nk):                          # restart at every knot: the right-hand side jumps there
                sol = solve_ivp(rhs, (t, t + dt), y, method="DOP853", rtol=1e-13, atol=1e-13, first_step=dt / 4)
                y, t = sol.y[:, -1], t + dt
            m = y.reshape(4, 4)
            got = m[:2, :2] + 1j * m[2:, :2]
            assert np.max(np.abs(got - rho[s])) < 2e-11, (channels, s)


def test_density_ode_oracle_matches_50_digit_golden(orc):
    z = np.load(GOLDEN + "/lindblad_t2.npz")
    G = int(z["gate_count"])
    for name, ch in (("t2", 2), ("t1t2", 3), ("t2c", 2), ("t2big", 2)):
        _, rho = orc.lindblad_knots(z["controls"], float(z["dt"]), int(z["gate"]), G, float(z["fq"]), z["da"], z["rho0"], channels=ch,
                                    gamma_c=float(z["gamma_c"]), gamma_f=float(z["gf_" + name]))
        assert np.max(np.abs(rho - z["rho_" + name][:, G])) < 1e-15, name


def test_runge_kutta_steps_converge_at_their_order(orc):
    """The RK restatement (RobotDynamics tableaux over the continuous dynamics of spin15_standalone.jl:263-312) against an
    independent fine-grid solution of the same ODE (constant u, so the amplitude is the exact quadratic a(t); the iso states by
    exponential midpoint steps 400x finer): the error falls at the tableau's order -- the tableaux are the right ones, the
    continuous dynamics the right vector field."""
    from rbqoc_b200 import spin

    base = spin.Spin13Model()
    x0 = spin.spin13_x0(); x0[8:11] = [0.0, 0.2, 0.05]
    T, u = 2.0, 0.01
    H0, H1 = np.pi * np.array(orc_N0()), np.pi * np.array(orc_N1())
    nf = 16000
    hf = T / nf
    psi = np.stack([x0[0:4], x0[4:8]], axis=1)
    for j in range(nf):
        tm = (j + 0.5) * hf
        a = x0[9] + x0[10] * tm + 0.5 * u * tm * tm
        psi = scipy.linalg.expm(hf * (FQ * H0 + a * H1)) @ psi
    ref = np.concatenate([psi[:, 0], psi[:, 1]])
    errs = {}
    for integ, order in ((spin.IntegratorType.rk2, 2), (spin.IntegratorType.rk3, 3), (spin.IntegratorType.rk4, 4)):
        e = []
        for N in (21, 41):
            dt = T / (N - 1)
            X = orc.rollout(spin.with_integrator(base, integ).params, x0, np.full((N - 1, 1), u), np.full(N - 1, dt))
            e.append(np.max(np.abs(X[-1, :8] - ref)))
            # the control chain is a polynomial of degree <= 2 in t: RK3 and RK4 integrate it exactly
            if order >= 3:
                assert abs(X[-1, 9] - (x0[9] + x0[10] * T + 0.5 * u * T * T)) < 1e-14
        errs[order] = e
        assert np.log2(e[0] / e[1]) > order - 0.4, (order, e)
    assert errs[4][1] < errs[3][1] < errs[2][1]
