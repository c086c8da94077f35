"""Regenerates the committed golden fixtures (run in the build container; needs mpmath, and
/root/reference for the pulse fixture):

  models.npz        step maps + Jacobians of every model from the 50-digit referee (oracle/referee.py)
  mc_static.npz     integrate_prop_schroed! + compute_fidelities on the analytic Y/2 pulse, 50 digits
  lindblad.npz      integrate_prop_lindbladt1! on the same pulse, 50 digits
  segments.npz      the analytic gates (spin.jl:562-594, 651-670, 765-785) and integrate_prop_xpiby2da! (spin.jl:819-905)
                    as segment lists, 50 digits
  spin14_pulses.npz controls produced by the REFERENCE's own generator functions
                    (/root/reference/src/spin/spin14.py:62-220, imported with h5py/matplotlib stubbed)

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import referee as ref  # noqa: E402

F = ref.f


def model_cases():
    rng = np.random.default_rng(20261019)
    cases = {}

    def psi_blocks(k):
        v = rng.standard_normal(4 * k) * 0.5
        return v

    specs = [
        ("spin13", "spin13", 11, 1, {}),
        ("spin12", "spin12", 43, 1, {"fq_samples": [1.4e-2 * 1.01, 1.4e-2 * 0.99]}),
        ("spin18", "spin18", 43, 1, {"namp": 5.21e-6 / 0.202407}),
        ("spin15_DA", "spin15", 12, 1, {"decay_aware": True}),
        ("spin15_DA_TO", "spin15", 12, 2, {"decay_aware": True, "time_optimal": True}),
        ("spin11_DO2", "spin11", 19, 1, {"derivative_order": 2}),
        ("spin17_DO2", "spin17", 19, 1, {"derivative_order": 2}),
        ("spin30", "spin30", 56, 1, {"fq_cov": 1.4e-4, "alpha": 1.0, "S_diag": [1.0, 2.0, 3.0, 4.0], "sample_state_count": 1,
                                     "nominal_offset": [0, 0, 0, 0]}),
        ("spin30_x2", "spin30", 97, 1, {"fq_cov": 1.4e-4, "alpha": 1.3, "S_diag": [1.0, 1.0, 1.0, 1.0], "sample_state_count": 2,
                                        "nominal_offset": [0, 4, 0, 0]}),
        # spin25: time-dependent filter start-up branch (knot 2) and the steady branch (knot 7); namp exaggerated so that the
        # amplitude-offset sigma points matter
        ("spin25_k2", "spin25", 58, 1, {"fq_cov": 1.0, "namp": 2.574e-2, "alpha": 1.0, "S_diag": [1.0, 2.0, 3.0, 4.0],
                                        "sample_state_count": 1, "nominal_offset": [4, 0, 0, 0], "knot": 2}),
        ("spin25_k7", "spin25", 58, 1, {"fq_cov": 1.0, "namp": 2.574e-2, "alpha": 1.2, "S_diag": [1.0, 1.0, 1.0, 1.0],
                                        "sample_state_count": 1, "nominal_offset": [0, 0, 0, 0], "knot": 7}),
    ]
    for name, model, n, m, kw in specs:
        for dt in (0.1, 0.01):
            x = rng.standard_normal(n) * 0.3
            if model in ("spin30", "spin25"):
                base = 15 if model == "spin30" else 17
                if model == "spin30":
                    x[12:15] = rng.uniform(-0.4, 0.4, 3)
                else:
                    x[8:11] = rng.uniform(-0.4, 0.4, 3)
                for c in range(kw["sample_state_count"]):
                    off = base + 41 * c
                    for j in range(10):
                        s = np.array([1.0, 0, 0, 0]) + 0.1 * rng.standard_normal(4)
                        x[off + 4 * j: off + 4 * j + 4] = s / np.linalg.norm(s)
                    x[off + 40] = 0.003
            else:
                x[8:11] = rng.uniform(-0.4, 0.4, 3)
            u = rng.uniform(-0.2, 0.2, m)
            if kw.get("time_optimal"):
                u[1] = 0.31
            mkw = {}
            time = 0.0
            if "knot" in kw:
                for _ in range(kw["knot"]):
                    time = time + dt          # accumulated like the knot times of a trajectory
                mkw["time"] = F(time)
            for k, v in kw.items():
                if k == "knot":
                    continue
                if isinstance(v, list) and k != "nominal_offset":
                    mkw[k] = [F(t) for t in v]
                elif isinstance(v, float):
                    mkw[k] = F(v)
                else:
                    mkw[k] = v
            xm, um, dtm = [F(t) for t in x], [F(t) for t in u], F(dt)
            xn = ref.step(model, xm, um, dtm, **mkw)
            J = ref.jacobian(model, xm, um, dtm, **mkw)
            key = f"{name}_dt{dt}"
            cases[key + "_x"] = x
            cases[key + "_u"] = u
            cases[key + "_t"] = np.float64(time)
            cases[key + "_xn"] = ref.to_f64(xn)
            cases[key + "_J"] = np.array([[float(t) for t in row] for row in J])
            print("model case", key, flush=True)
    return cases


def reference_pulses():
    """Import the reference's spin14.py with its missing plotting/IO dependencies stubbed and call its generators."""
    path = "/root/reference/src/spin/spin14.py"
    for name in ("h5py", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    src = open(path).read()
    mod = types.ModuleType("spin14_ref")
    mod.__file__ = path
    mod.__name__ = "spin14_ref"          # so that the `if __name__ == "__main__"` guard stays off
    exec(compile(src, path, "exec"), mod.__dict__)
    out = {}
    for nm, gen, ttot in (("ypiby2", mod.gen_controls_ypiby2, mod.TTOT_YPIBY2), ("xpiby2", mod.gen_controls_xpiby2, mod.TTOT_XPIBY2)):
        count = int(np.floor(ttot * mod.DT_INV))                        # spin14.py:116-118
        times = np.arange(0, count, 1) * mod.DT
        controls = np.vstack([gen(t, shape=mod.Shape.SQUARE) for t in times]).astype(np.float64)
        controls[0, 0] = 0
        controls[-1, 0] = 0
        out[nm] = controls[:, 0]
        out[nm + "_evolution_time"] = np.float64(ttot)
    out["dt_inv"] = np.float64(mod.DT_INV)
    return out


def segment_cases():
    """Analytic-gate integrators in 50 digits.  Segment lists come from rbqoc_b200.spin (plain data: amplitudes/durations)."""
    from rbqoc_b200 import spin

    rng = np.random.default_rng(99)
    out = {}
    S, G = 5, 6
    fq = 1.4e-2 * (1 + 0.01 * rng.standard_normal(S)); fq[0] = 1.4e-2
    da = 2.574e-5 * rng.standard_normal(S); da[0] = 0.0
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    psi0 = np.concatenate([psi.real, psi.imag], axis=1)
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    lam = np.array([1.0, 1.0, 0.9, 0.8, 0.6])[:, None, None]
    rho0 = lam * rho0 + (1 - lam) * np.eye(2) / 2
    out.update(fq=fq, da=da, psi0=psi0, rho0=rho0, gate_count=G)
    for gate in (spin.GateType.zpiby2, spin.GateType.ypiby2, spin.GateType.xpiby2):
        amps, durs = spin.ANALYTIC_SEGMENTS[gate]
        Gm = ref.mat_iso(ref.GATES[int(gate)]())
        fid = np.empty((S, G + 1)); rho = np.empty((S, G + 1, 2, 2), dtype=np.complex128)
        for s in range(S):
            prop = mp.eye(4)
            for a, d in zip(amps, durs):
                prop = ref.propagator(F(fq[s]), F(a) + F(da[s]), F(d)) * prop
            tg = [mp.matrix([F(t) for t in psi0[s]])]
            for _ in range(3):
                tg.append(Gm * tg[-1])
            st = tg[0].copy()
            fid[s, 0] = float(ref.fidelity_vec_iso2(list(st), list(tg[0])))
            for g in range(1, G + 1):
                st = prop * st
                fid[s, g] = float(ref.fidelity_vec_iso2(list(st), list(tg[g % 4])))
            lp = mp.eye(4)
            for a, d in zip(amps, durs):
                lp = mp.expm(ref.lindblad_generator(ref.FQ, F(a) + F(da[s])) * F(d)) * lp
            v = mp.matrix([mp.mpc(complex(rho0[s, 0, 0])), mp.mpc(complex(rho0[s, 1, 0])), mp.mpc(complex(rho0[s, 0, 1])),
                           mp.mpc(complex(rho0[s, 1, 1]))])
            for g in range(G + 1):
                rho[s, g] = [[complex(v[0]), complex(v[2])], [complex(v[1]), complex(v[3])]]
                v = lp * v
        out[f"fid_{gate.name}"] = fid
        out[f"rho_{gate.name}"] = rho
        print("segments", gate.name, flush=True)
    # integrate_prop_xpiby2da!: noise exaggerated (1e-3 GHz) so that every segment's offset matters
    amps, durs, nidx = spin.xpiby2da_segments()
    S2, G2 = 3, 3
    noise = 1e-3 * rng.standard_normal((S2, 569))
    fid = np.empty((S2, G2 + 1))
    Gm = ref.mat_iso(ref.GATES[int(spin.GateType.xpiby2)]())
    for s in range(S2):
        tg = [mp.matrix([F(t) for t in psi0[s]])]
        for _ in range(3):
            tg.append(Gm * tg[-1])
        st = tg[0].copy()
        fid[s, 0] = float(ref.fidelity_vec_iso2(list(st), list(tg[0])))
        for g in range(1, G2 + 1):
            for a, d, k in zip(amps, durs, nidx):
                st = ref.propagator(ref.FQ, F(a) + F(noise[s, k]), F(d)) * st
            fid[s, g] = float(ref.fidelity_vec_iso2(list(st), list(tg[g % 4])))
        print("xpiby2da sample", s, flush=True)
    out.update(da_noise=noise, da_fid=fid, da_gate_count=G2)
    return out


def main():
    if "--segments-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "segments.npz"), **segment_cases())
        return
    np.savez_compressed(os.path.join(HERE, "models.npz"), **model_cases())
    if "--models-only" in sys.argv:
        return
    pulses = reference_pulses() if os.path.exists("/root/reference/src/spin/spin14.py") else None
    if pulses is not None:
        np.savez_compressed(os.path.join(HERE, "spin14_pulses.npz"), **pulses)
        controls = pulses["ypiby2"]
    else:
        controls = np.load(os.path.join(HERE, "spin14_pulses.npz"))["ypiby2"]
    dt = 0.01
    rng = np.random.default_rng(77)
    S, G = 6, 5
    fq = 1.4e-2 * (1 + 0.01 * rng.standard_normal(S))
    da = 2.574e-5 * rng.standard_normal(S)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    psi0 = np.concatenate([psi.real, psi.imag], axis=1)
    cm = [F(a) for a in controls]
    fid = np.empty((S, G + 1)); states = np.empty((S, 4))
    for s in range(S):
        f_, st = ref.sim_schroed(cm, F(dt), 2, G, F(fq[s]), F(da[s]), [F(t) for t in psi0[s]])
        fid[s] = ref.to_f64(f_); states[s] = ref.to_f64(st)
        print("mc sample", s, flush=True)
    np.savez_compressed(os.path.join(HERE, "mc_static.npz"), controls=controls, dt=dt, gate=2, gate_count=G, fq=fq, da=da, psi0=psi0,
                        fid=fid, states=states)
    S, G = 4, 6
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    lam = np.array([1.0, 1.0, 0.8, 0.7])[:, None, None]
    rho0 = lam * rho0 + (1 - lam) * np.eye(2) / 2
    da = 1e-3 * rng.standard_normal(S)
    rho = np.empty((S, G + 1, 2, 2), dtype=np.complex128)
    for s in range(S):
        out = ref.sim_lindblad(cm, F(dt), G, ref.FQ, F(da[s]), [[mp.mpc(complex(rho0[s, i, j])) for j in range(2)] for i in range(2)])
        for g, v in enumerate(out):
            rho[s, g] = [[complex(v[0]), complex(v[2])], [complex(v[1]), complex(v[3])]]
        print("lindblad sample", s, flush=True)
    np.savez_compressed(os.path.join(HERE, "segments.npz"), **segment_cases())
    np.savez_compressed(os.path.join(HERE, "lindblad.npz"), controls=controls, dt=dt, gate=2, gate_count=G, fq=1.4e-2, da=da, rho0=rho0,
                        rho=rho)


if __name__ == "__main__":
    main()
