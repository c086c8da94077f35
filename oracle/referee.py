"""50-digit mpmath referee for the oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

The reference ships no golden vectors and its Julia toolchain is absent ("parity unpinned"), so the
FP64 oracle (oracle.cpp) is pinned against this rounding-free restatement of the same mathematics:
exact matrix exponentials (mpmath.expm at 50 digits), the step maps written from the reference
files, central-difference Jacobians in 50-digit arithmetic (h = 1e-18, truncation 1e-36).
Used by tests/golden/make_golden.py to produce the committed fixtures and by the CPU test-suite.
"""
from __future__ import annotations

import mpmath as mp
import numpy as np

mp.mp.dps = 50

PI = mp.pi
FQ = mp.mpf("1.4e-2")
N0 = mp.matrix([[0, 0, 1, 0], [0, 0, 0, -1], [-1, 0, 0, 0], [0, 1, 0, 0]])   # spin.jl:234 without pi
N1 = mp.matrix([[0, 0, 0, 1], [0, 0, 1, 0], [0, -1, 0, 0], [-1, 0, 0, 0]])   # spin.jl:236 without pi

T1_US = [mp.mpf(s) for s in ("1597.923 1627.93 301.86 269.03 476.33 1783.19 2131.76 2634.50 4364.68 2587.82 1661.915 "
                             "1794.468 2173.88 1188.83 1576.493 965.183 560.251 310.88").split()]
FBFQ = [mp.mpf(s) for s in "0.26 0.28 0.32 0.34 0.36 0.38 0.4 0.42 0.44 0.46 0.465 0.47 0.475 0.48 0.484 0.488 0.492 0.5".split()]


def f(x):
    """binary64 -> mpf exactly."""
    return mp.mpf(float(x))


def propagator(fq, a, dt):
    """exp(dt (fq NEGI_H0_ISO + a NEGI_H1_ISO))  (spin13.jl:45-46)."""
    return mp.expm(dt * PI * (fq * N0 + a * N1))


def t1_us(a):
    """amp_t1_reduced_spline (spin.jl:366,396): gridded-linear in fbfq = -|a| 0.202407 + 0.5, flat outside."""
    x = -abs(a) * mp.mpf("0.202407") + mp.mpf("0.5")
    x = min(max(x, FBFQ[0]), FBFQ[-1])
    i = 0
    while i < 16 and FBFQ[i + 1] <= x:
        i += 1
    w = (x - FBFQ[i]) / (FBFQ[i + 1] - FBFQ[i])
    return (1 - w) * T1_US[i] + w * T1_US[i + 1]


def _mv(E, v):
    return [sum(E[i, k] * v[k] for k in range(4)) for i in range(4)]


def step(model, x, u, dt, **kw):
    """Augmented-state step maps in 50 digits.  model in {"spin13","spin12","spin18","spin15","spin11","spin17","spin30"}.
    x, u: sequences of mpf.  kw mirrors rbq_model_params."""
    fq = kw.get("fq", FQ)
    x = list(x)
    if model == "spin30":
        return _step_spin30(x, u, dt, fq, kw)
    if model == "spin25":
        return _step_spin25(x, u, dt, fq, kw)
    to = kw.get("time_optimal", False)
    if model == "spin15" and to:
        dt = u[1] * u[1]                                           # spin15.jl:52-54
    a = x[9]
    E = propagator(fq, a, dt)
    out = _mv(E, x[0:4]) + _mv(E, x[4:8]) + [x[8] + a * dt, x[9] + x[10] * dt, x[10] + u[0] * dt]
    if model == "spin13":
        return out
    if model in ("spin12", "spin18"):
        if model == "spin12":
            fs = kw["fq_samples"]
            Ep, En = propagator(fs[0], a, dt), propagator(fs[1], a, dt)
        else:
            Ep, En = propagator(fq, a + kw["namp"], dt), propagator(fq, a - kw["namp"], dt)
        for s in range(8):
            out += _mv(Ep if s < 4 else En, x[11 + 4 * s: 15 + 4 * s])
        return out
    if model == "spin15":
        if kw.get("decay_aware", False):
            out.append(x[11] + dt / t1_us(a))                      # spin15.jl:66-69
        return out
    if model in ("spin11", "spin17"):
        NH = PI * (N0 if model == "spin11" else N1)
        do = kw.get("derivative_order", 0)
        if do >= 1:
            w = [x[11 + i] + dt * sum(NH[i, k] * x[k] for k in range(4)) for i in range(4)]
            out += _mv(E, w)
        if do >= 2:
            w = [x[15 + i] + 2 * dt * sum(NH[i, k] * x[11 + k] for k in range(4)) for i in range(4)]
            out += _mv(E, w)
        return out
    raise ValueError(model)


PFIR_B = [mp.mpf(t) for t in ("0.049922035", "-0.095993537", "0.050612699", "-0.004408786")]   # spin.jl:166-169
PFIR_A = [mp.mpf(1)] + [mp.mpf(t) for t in ("-2.494956002", "2.017265875", "-0.522189400")]    # spin.jl:170-173


def _step_spin25(x, u, dt, fq, kw):
    """spin25.jl:181-235; kw["time"] is the knot time, kw["fq_cov"] carries wnamp_cov."""
    camp = x[9]
    E = propagator(fq, camp, dt)
    out = _mv(E, x[0:4]) + _mv(E, x[4:8]) + [x[8] + camp * dt, x[9] + x[10] * dt, x[10] + u[0] * dt]
    xk = mp.sqrt(kw["fq_cov"])
    time = kw.get("time", mp.mpf(0))
    kp = int(mp.nint((time - mp.fmod(time, dt)) / dt))
    xp, yp = x[11:14], x[14:17]
    yk = PFIR_B[0] * xk
    if kp != 1:
        yk += PFIR_B[1] * xp[0] - PFIR_A[1] * yp[0]
        if kp != 2:
            yk += PFIR_B[2] * xp[1] - PFIR_A[2] * yp[1]
            if kp != 3:
                yk += PFIR_B[3] * xp[2] - PFIR_A[3] * yp[2]
    out += [xk, xp[0], xp[1], yk, yp[0], yp[1]]
    namp = kw["namp"] * yk
    Ep, En = propagator(fq, camp + namp, dt), propagator(fq, camp - namp, dt)
    return out + _unscented(x, 17, E, Ep, En, kw)


def _step_spin30(x, u, dt, fq, kw):
    a = x[13]
    E = propagator(fq, a, dt)
    out = _mv(E, x[0:4]) + _mv(E, x[4:8]) + _mv(E, x[8:12]) + [x[12] + a * dt, x[13] + x[14] * dt, x[14] + u[0] * dt]
    sig = mp.sqrt(kw["fq_cov"])
    Ep, En = propagator(fq + sig, a, dt), propagator(fq - sig, a, dt)
    return out + _unscented(x, 15, E, Ep, En, kw)


def _unscented(x, base, E, Ep, En, kw):
    out = []
    fq_cov, alpha = kw["fq_cov"], kw.get("alpha", mp.mpf(1))
    S = kw.get("S_diag", [mp.mpf(1)] * 4)
    for c in range(kw.get("sample_state_count", 1)):
        off = base + 41 * c
        s = [_mv(Ep if j == 4 else En if j == 9 else E, x[off + 4 * j: off + 4 * j + 4]) for j in range(10)]
        sm = [sum(s[j][i] for j in range(10)) / 10 for i in range(4)]
        d = [[s[j][i] - sm[i] for i in range(4)] for j in range(10)]
        P = mp.matrix(5, 5)
        for i in range(4):
            for k in range(4):
                P[i, k] = sum(d[j][i] * d[j][k] for j in range(10)) / (2 * alpha * alpha)
        P[4, 4] = fq_cov
        L = alpha * mp.cholesky(P)
        new = []
        for j in range(4):
            new.append([sm[i] + L[i, j] for i in range(4)])
        new.append(list(sm))
        for j in range(4):
            new.append([sm[i] - L[i, j] for i in range(4)])
        new.append(list(sm))
        for v in new:
            nrm = mp.sqrt(sum(t * t for t in v))
            out += [t / nrm for t in v]
        nom = kw.get("nominal_offset", [0] * 4)[c]
        dn = [sm[i] - x[nom + i] for i in range(4)]
        out.append(sum(P[i, i] * S[i] for i in range(4)) + sum(dn[i] * S[i] * dn[i] for i in range(4)))
    return out


def jacobian(model, x, u, dt, h=mp.mpf("1e-18"), **kw):
    """Central differences in 50-digit arithmetic over [x;u]: n x (n+m).  Not valid AT a kink of
    abs()/the T1 table (use points away from them)."""
    x, u = list(x), list(u)
    n, m = len(x), len(u)
    cols = []
    for c in range(n + m):
        xp, xm, up, um = list(x), list(x), list(u), list(u)
        if c < n:
            xp[c] += h; xm[c] -= h
        else:
            up[c - n] += h; um[c - n] -= h
        fp, fm = step(model, xp, up, dt, **kw), step(model, xm, um, dt, **kw)
        cols.append([(p - q) / (2 * h) for p, q in zip(fp, fm)])
    return [[cols[c][r] for c in range(n + m)] for r in range(len(cols[0]))]


def fidelity_vec_iso2(s1, s2):
    a = sum(p * q for p, q in zip(s1, s2))
    b = s1[0] * s2[2] + s1[1] * s2[3] - s1[2] * s2[0] - s1[3] * s2[1]
    return a * a + b * b


GATES = {
    1: lambda: mp.matrix([[(1 - 1j), 0], [0, (1 + 1j)]]) / mp.sqrt(2),
    2: lambda: mp.matrix([[1, -1], [1, 1]]) / mp.sqrt(2),
    3: lambda: mp.matrix([[1, -1j], [-1j, 1]]) / mp.sqrt(2),
    4: lambda: mp.matrix([[0, -1j], [-1j, 0]]),
}


def mat_iso(g):
    G = mp.matrix(4, 4)
    for i in range(2):
        for j in range(2):
            G[i, j] = mp.re(g[i, j]); G[i, j + 2] = -mp.im(g[i, j])
            G[i + 2, j] = mp.im(g[i, j]); G[i + 2, j + 2] = mp.re(g[i, j])
    return G


def sim_schroed(controls, dt, gate, gate_count, fq, da, psi0):
    """integrate_prop_schroed! + compute_fidelities (spin.jl:438-452, 1198-1240) in 50 digits."""
    prop = mp.eye(4)
    for a in controls:
        prop = propagator(fq, a + da, dt) * prop
    G = mat_iso(GATES[gate]())
    tg = [mp.matrix(psi0)]
    for _ in range(3):
        tg.append(G * tg[-1])
    st = mp.matrix(psi0)
    fid = [fidelity_vec_iso2(list(st), list(tg[0]))]
    for g in range(1, gate_count + 1):
        st = prop * st
        fid.append(fidelity_vec_iso2(list(st), list(tg[g % 4])))
    return fid, list(st)


def lindblad_generator(fq, a):
    """FQ_NEGI_H0_VISO + a NEGI_H1_VISO + gamma_1 (LOPP_VISO + LOPM_VISO)  (spin.jl:240-276, 529-531), column-major vec."""
    I2 = mp.eye(2)
    sz = mp.matrix([[1, 0], [0, -1]])
    sx = mp.matrix([[0, 1], [1, 0]])
    H = -1j * PI * (fq * sz + a * sx)

    def kron(A, B):
        K = mp.matrix(4, 4)
        for i in range(2):
            for j in range(2):
                for k in range(2):
                    for l in range(2):
                        K[2 * i + k, 2 * j + l] = A[i, j] * B[k, l]
        return K

    L = kron(I2, H) - kron(H.T, I2)
    gg, ge = mp.matrix([[1, 0], [0, 0]]), mp.matrix([[0, 1], [0, 0]])
    eg, ee = mp.matrix([[0, 0], [1, 0]]), mp.matrix([[0, 0], [0, 1]])
    gamma = 1 / (t1_us(a) * 1000)
    L += gamma * (kron(ge, ge) - kron(I2, ee) / 2 - kron(ee, I2) / 2)
    L += gamma * (kron(eg, eg) - kron(I2, gg) / 2 - kron(gg, I2) / 2)
    return L


def sim_lindblad(controls, dt, gate_count, fq, da, rho0):
    """integrate_prop_lindbladt1! (spin.jl:521-539): returns the vec'd densities after each gate."""
    prop = mp.eye(4)
    for a in controls:
        prop = mp.expm(lindblad_generator(fq, a + da) * dt) * prop
    v = mp.matrix([rho0[0][0], rho0[1][0], rho0[0][1], rho0[1][1]])
    out = [list(v)]
    for _ in range(gate_count):
        v = prop * v
        out.append(list(v))
    return out


def sim_density_ode(controls, dt, gate_count, fq, da, rho0, t1=False, t2=False, gamma_c=0, gamma_f=0):
    """Exact solution, in 50 digits, of the iso-density ODEs the reference gives to Vern9: dynamics_lindbladt1_deqjl
    (spin.jl:505-518) and dynamics_lindbladt2_deqjl (spin.jl:542-554), pulse repeated `gate_count` times on its knot grid.
    The right-hand side is written exactly as the reference writes it (4x4 real iso density, commutator with negi_h, the
    GE/EG dissipators, the elementwise NEG_DOP_ISO mask with rate GAMMAC + 2 gammaf^2 t); inside a knot it is
    L_A rho + (t - t0) L_1 rho, whose Taylor coefficients obey (n+1) c_{n+1} = L_A c_n + L_1 c_{n-1}.
    Returns the complex 2x2 densities after 0..gate_count gates."""
    GE = mp.matrix([[0, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 1], [0, 0, 0, 0]])
    EG = GE.T
    NGG = mp.diag([1, 0, 1, 0]) * mp.mpf(-0.5)
    NEE = mp.diag([0, 1, 0, 1]) * mp.mpf(-0.5)
    DOP = [[0, 1, 0, 1], [1, 0, 1, 0], [0, 1, 0, 1], [1, 0, 1, 0]]
    rho = mp.matrix(4, 4)
    for i in range(2):
        for j in range(2):
            rho[i, j] = mp.re(rho0[i][j]); rho[i, j + 2] = -mp.im(rho0[i][j])
            rho[i + 2, j] = mp.im(rho0[i][j]); rho[i + 2, j + 2] = mp.re(rho0[i][j])
    g1coef = 2 * gamma_f * gamma_f if t2 else mp.mpf(0)

    def mask(m, c):
        out = mp.matrix(4, 4)
        for i in range(4):
            for j in range(4):
                out[i, j] = -c * DOP[i][j] * m[i, j]
        return out

    def unpack(m):
        return [[mp.mpc(m[0, 0], m[2, 0]), mp.mpc(m[0, 1], m[2, 1])], [mp.mpc(m[1, 0], m[3, 0]), mp.mpc(m[1, 1], m[3, 1])]]

    out = [unpack(rho)]
    step = 0
    for _ in range(gate_count):
        for a in controls:
            c1 = a + da
            H = PI * (fq * N0 + c1 * N1)
            gamma1 = 1 / (t1_us(c1) * 1000) if t1 else mp.mpf(0)
            gam0 = (gamma_c + g1coef * (step * dt)) if t2 else mp.mpf(0)

            def rhs(m):
                r = H * m - m * H
                if t1:
                    r += gamma1 * (GE * m * EG + NEE * m + m * NEE) + gamma1 * (EG * m * GE + NGG * m + m * NGG)
                if t2:
                    r += mask(m, gam0)
                return r

            cprev, c, acc, hp = mp.matrix(4, 4), rho.copy(), rho.copy(), mp.mpf(1)
            for n in range(80):
                nxt = (rhs(c) + mask(cprev, g1coef)) / (n + 1)
                cprev, c = c, nxt
                hp *= dt
                acc += c * hp
                if max(abs(x) for x in c) * hp < mp.mpf("1e-45"):
                    break
            rho = acc
            step += 1
        out.append(unpack(rho))
    return out


def to_f64(v):
    return np.array([float(t) for t in v], dtype=np.float64)
