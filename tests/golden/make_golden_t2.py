"""Golden vectors of the dephasing (T2) and T1+T2 density ODEs with the drive ON, from the 50-digit referee
(oracle/referee.py: sim_density_ode restates dynamics_lindbladt1_deqjl / dynamics_lindbladt2_deqjl, spin.jl:505-518, 542-554).

    python tests/golden/make_golden_t2.py        -> tests/golden/lindblad_t2.npz
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import referee as ref  # noqa: E402

F = ref.f
GAMMAC = 1 / 3e5                                  # spin.jl:162
GAMMAF = 5.21e-6 * 4 * 2 * np.pi                   # GAMMAF_PREFACTOR = FBFQ_NAMP * SQRTLNIR * 2 pi, spin.jl:161-165


def main():
    rng = np.random.default_rng(4242)
    N = 161
    k = np.arange(N)
    controls = (0.30 * np.sin(2 * np.pi * k / (N - 1)) + 0.10 * np.sin(6 * np.pi * k / (N - 1)))[:-1]
    dt = 0.1
    S, G = 3, 3
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    lam = np.array([1.0, 0.9, 0.6])[:, None, None]
    rho0 = lam * rho0 + (1 - lam) * np.eye(2) / 2
    da = 1e-3 * rng.standard_normal(S)
    out = dict(controls=controls, dt=dt, gate=3, gate_count=G, fq=1.4e-2, da=da, rho0=rho0, gamma_c=GAMMAC, gamma_f=GAMMAF)
    # gamma_f exaggerated 30x in the last case so that the time-dependent rate is well above rounding
    for name, t1, t2, gc, gf in (("t2", False, True, GAMMAC, GAMMAF), ("t1t2", True, True, GAMMAC, GAMMAF),
                                 ("t2c", False, True, GAMMAC, 0.0), ("t2big", False, True, GAMMAC, 30 * GAMMAF)):
        rho = np.empty((S, G + 1, 2, 2), dtype=np.complex128)
        for s in range(S):
            res = ref.sim_density_ode([F(a) for a in controls], F(dt), G, ref.FQ, F(da[s]),
                                      [[mp.mpc(complex(rho0[s, i, j])) for j in range(2)] for i in range(2)],
                                      t1=t1, t2=t2, gamma_c=F(gc), gamma_f=F(gf))
            for g, m in enumerate(res):
                rho[s, g] = [[complex(m[0][0]), complex(m[0][1])], [complex(m[1][0]), complex(m[1][1])]]
            print(name, "sample", s, flush=True)
        out["rho_" + name] = rho
        out["gf_" + name] = gf
    # realistic offsets (2.574e-5 GHz, spin.jl:164): the regime the engine's per-knot Frechet table serves; T1 and T1 + constant T2
    da_small = 2.574e-5 * rng.standard_normal(S)
    out["da_small"] = da_small
    for name, t2, gc in (("t1_small", False, 0.0), ("t1t2c_small", True, GAMMAC)):
        rho = np.empty((S, G + 1, 2, 2), dtype=np.complex128)
        for s in range(S):
            res = ref.sim_density_ode([F(a) for a in controls], F(dt), G, ref.FQ, F(da_small[s]),
                                      [[mp.mpc(complex(rho0[s, i, j])) for j in range(2)] for i in range(2)],
                                      t1=True, t2=t2, gamma_c=F(gc), gamma_f=F(0.0))
            for g, m in enumerate(res):
                rho[s, g] = [[complex(m[0][0]), complex(m[0][1])], [complex(m[1][0]), complex(m[1][1])]]
            print(name, "sample", s, flush=True)
        out["rho_" + name] = rho
    np.savez_compressed(os.path.join(HERE, "lindblad_t2.npz"), **out)


if __name__ == "__main__":
    main()
