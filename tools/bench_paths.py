"""Secondary measurements (device-resident, CUDA events): every kernel family of the hot path besides the headline
sweep.  Prints one JSON line per path; results are summarised in profiles/README.md.

    python tools/bench_paths.py            # on a B200
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin

dev = torch.device("cuda", 0)
eng = rbq.Engine(0)
eng.use_torch_stream()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def out(name, ms, units, unit_name, **kw):
    print(json.dumps({"path": name, "ms": round(ms, 4), "per_s": units / ms * 1e3, "unit": unit_name, **kw}), flush=True)


peak = eng.fp64_peak(20000)
print(json.dumps({"fp64_peak_tflops": peak}), flush=True)
stats = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)

# ---- static sweeps: both grids, both algorithms --------------------------------------------------------------
for name, controls, dt, S in (("P_sin(1001) dt=0.1", spin.pulse_sin(1001)[:-1], 0.1, 1 << 20),
                              ("P_X2(5680) dt=0.01", spin.controls_xpiby2(100.0), 0.01, 1 << 20)):
    fq, da, psi0 = eng.philox_samples(0, S, 1, spin.FQ, 0.01, 0.03, 2.574e-5)
    dc, dfq, dda, dpsi = d(controls), d(fq), d(da), d(psi0)
    dfid = torch.empty(S, 2, dtype=torch.float64, device=dev)
    nk = controls.shape[0]
    for algo, an, reps in ((rbq.ALGO_SU2, "su2", 10), (rbq.ALGO_PADE, "pade (reference algorithm on GPU)", 2)):
        Sa = S if algo == rbq.ALGO_SU2 else S // 8
        ms = timed(lambda: eng.mc_static_dev(dc, nk, dt, rbq.GATE_XPIBY2, 1, Sa, dfq, dda, dpsi, algo, dfid, stats), reps=reps)
        out(f"mc_static {an} {name}", ms, Sa * nk, "trajectory-steps/s", S=Sa, knots=nk)
    ms = timed(lambda: eng.mc_static_philox_dev(dc, nk, dt, rbq.GATE_XPIBY2, 1, 0, S, 1, spin.FQ, 0.01, 0.03, 2.574e-5,
                                                rbq.ALGO_SU2, None, stats))
    out(f"mc_static_philox su2 {name} (stats only)", ms, S * nk, "trajectory-steps/s", S=S, knots=nk)

# ---- time-dependent noise (gen_3b shape: X/2 pulse, 0.1 ns noise grid) ------------------------------------------
controls, dt, gates, ndi = spin.controls_xpiby2(100.0), 0.01, 20, 10.0
nk = controls.shape[0]
S = 1 << 16
nlen = int(np.ceil(nk * dt * gates * ndi)) + 1
dc = d(controls)
dnoise = torch.empty(S, nlen, dtype=torch.float64, device=dev)
noise_h = eng.pink_noise(0, 4096, 3, spin.NAMP_PREFACTOR, ndi, nlen)
dnoise.view(-1, 4096, nlen)[:] = d(noise_h)
_, _, psi0 = eng.philox_samples(0, S, 3, spin.FQ, 0, 0, 0)
dpsi = d(psi0)
dfid = torch.empty(S, gates + 1, dtype=torch.float64, device=dev)
ms = timed(lambda: eng.mc_noise_dev(dc, nk, dt, rbq.GATE_XPIBY2, gates, S, spin.FQ, dnoise, nlen, ndi, dpsi, rbq.ALGO_SU2, dfid, stats),
           reps=3, warm=1)
out("mc_noise su2 (explicit noise rows) P_X2 x 20 gates", ms, S * nk * gates, "trajectory-steps/s", S=S, knots=nk, gates=gates)
ms = timed(lambda: eng.mc_pink_philox_dev(dc, nk, dt, rbq.GATE_XPIBY2, gates, 0, S, 3, spin.FQ, spin.NAMP_PREFACTOR, ndi,
                                          rbq.ALGO_SU2, dfid, stats), reps=3, warm=1)
out("mc_pink_philox su2 (seeds in: noise rows materialised on the device, gate-parallel form) P_X2 x 20 gates", ms, S * nk * gates, "trajectory-steps/s", S=S, knots=nk,
    gates=gates)

# gen_3b shape proper: 1000 seeds x 200 gates x 5680 knots (figures.jl:504-505, 551-561): sequential vs gate-parallel form
S3, gates3 = 1000, 200
nlen3 = int(np.ceil(nk * dt * gates3 * ndi)) + 1
dnoise3 = torch.empty(S3, nlen3, dtype=torch.float64, device=dev)
dnoise3[:] = d(eng.pink_noise(0, S3, 7, spin.NAMP_PREFACTOR, ndi, nlen3))
dfid3 = torch.empty(S3, gates3 + 1, dtype=torch.float64, device=dev)
for mode in ("0", "1"):
    eng.set_option("RBQ_NOISE_GATE_PARALLEL", mode)
    ms = timed(lambda: eng.mc_noise_dev(dc, nk, dt, rbq.GATE_XPIBY2, gates3, S3, spin.FQ, dnoise3, nlen3, ndi, dpsi, rbq.ALGO_SU2, dfid3, stats),
               reps=2, warm=1)
    out(f"mc_noise gen_3b shape (1000 samples x 200 gates x 5680 knots), {'gate-parallel' if mode == '1' else 'sequential'} form", ms,
        S3 * nk * gates3, "trajectory-steps/s", S=S3, knots=nk, gates=gates3)
eng.set_option("RBQ_NOISE_GATE_PARALLEL", None)

# ---- Lindblad T1: 1e5 realisations (BASELINE configs[3]) ----------------------------------------------------------
S = 100_000
rng = np.random.default_rng(0)
psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
psi /= np.linalg.norm(psi, axis=1, keepdims=True)
rho0 = np.einsum("si,sj->sij", psi, psi.conj())
vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4)).view(np.float64)
da = 2.574e-5 * rng.standard_normal(S)
for name, controls, dt, gates in (("P_sin(1001) dt=0.1", spin.pulse_sin(1001)[:-1], 0.1, 1), ("P_Y2(1947) dt=0.01", spin.controls_ypiby2(100.0), 0.01, 200)):
    dc, dda, drho = d(controls), d(da), d(vec)
    dfid = torch.empty(S, gates + 1, dtype=torch.float64, device=dev)
    dro = torch.empty(S, 8, dtype=torch.float64, device=dev)
    nk = controls.shape[0]
    ms = timed(lambda: eng.lindblad_dev(dc, nk, dt, rbq.GATE_YPIBY2, gates, S, spin.FQ, dda, drho, rbq.LINDBLAD_T1, 0.0, 0.0, dfid, dro, stats),
               reps=5)
    out(f"lindblad T1 {name} gates={gates}", ms, S * nk, "density-steps/s", S=S, knots=nk)

# ---- trajectory level: rollout + Jacobians (BASELINE configs[0..2]) ---------------------------------------------------
for mname, model, x0, N in (("spin13", spin.Spin13Model(), spin.spin13_x0(), 1001), ("spin12", spin.Spin12Model(), spin.spin12_x0(), 1001),
                            ("spin30", spin.Spin30Model(), spin.spin30_x0(), 1001)):
    n, m = model.n, model.m
    dtv = np.full(N - 1, 0.1)
    U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None] * np.ones((1, m))
    X = eng.rollout(model.params, x0, U, dtv)
    for B in (1, 64):
        dX0 = d(np.tile(x0, (B, 1)))
        dU = d(np.tile(U, (B, 1, 1)))
        ddt = d(dtv)
        dX = torch.empty(B, N, n, dtype=torch.float64, device=dev)
        ms = timed(lambda: eng.rollout_dev(model.params, B, N, dX0, dU, ddt, dX), reps=5)
        out(f"rollout {mname} N={N} B={B}", ms, B * (N - 1), "knot-steps/s", n=n)
        K = B * (N - 1)
        dXk = d(np.tile(X[:-1], (B, 1)))
        dUk = d(np.tile(U, (B, 1)))
        ddk = d(np.tile(dtv, B))
        dAB = torch.empty(K, n * (n + m), dtype=torch.float64, device=dev)
        ms = timed(lambda: eng.jacobians_dev(model.params, K, dXk, dUk, ddk, dAB), reps=5)
        out(f"jacobians {mname} N={N} B={B}", ms, K, "knot-jacobians/s", n=n, bytes_out=K * n * (n + m) * 8,
            gbs=K * n * (n + m) * 8 / ms * 1e-6)
    # host-pointer calls (what a per-iteration ccall costs, PCIe included)
    ms_h = timed(lambda: eng.jacobians(model.params, X[:-1], U, dtv), reps=5)
    out(f"jacobians {mname} N={N} host pointers (H2D+D2H inside)", ms_h, N - 1, "knot-jacobians/s", n=n)
    ms_h = timed(lambda: eng.rollout(model.params, x0, U, dtv), reps=5)
    out(f"rollout {mname} N={N} host pointers", ms_h, N - 1, "knot-steps/s", n=n)

# ---- iLQR backward pass fed by the Jacobians above (device resident) --------------------------------
for mname, model, x0, N in (("spin13", spin.Spin13Model(), spin.spin13_x0(), 1001), ("spin12", spin.Spin12Model(), spin.spin12_x0(), 1001),
                            ("spin30", spin.Spin30Model(), spin.spin30_x0(), 1001)):
    n, m = model.n, model.m
    dtv = np.full(N - 1, 0.1)
    U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None] * np.ones((1, m))
    X = eng.rollout(model.params, x0, U, dtv)
    AB = eng.jacobians(model.params, X[:-1], U, dtv)                    # (N-1, n, n+m) view of the column-major blocks
    Exx = np.tile(np.eye(n), (N, 1, 1)); Exx[-1] *= N
    Ex = 0.1 * X; Euu = np.full((N - 1, m, m), 0.0) + 0.1 * np.eye(m); Eux = np.zeros((N - 1, m, n)); Eu = 0.1 * U
    col = lambda x: np.ascontiguousarray(np.swapaxes(x, -1, -2))
    for B in (1, 148, 592):
        rep = lambda x: d(np.tile(x[None], (B,) + (1,) * x.ndim))
        dAB, dExx, dEx, dEuu, dEux, dEu = rep(col(AB)), rep(col(Exx)), rep(Ex), rep(col(Euu)), rep(col(Eux)), rep(Eu)
        dK = torch.empty(B, N - 1, n * m, dtype=torch.float64, device=dev); dd = torch.empty(B, N - 1, m, dtype=torch.float64, device=dev)
        ddV = torch.empty(B, 2, dtype=torch.float64, device=dev); dfl = torch.empty(B, dtype=torch.int64, device=dev)
        ms = timed(lambda: eng.backward_pass_dev(n, m, N, B, dAB, dExx, dEx, dEuu, dEux, dEu, 1e-3, rbq.REG_CONTROL, dK, dd, ddV, dfl), reps=3, warm=1)
        flops = 2.0 * (n * n * (n + m) + 0.5 * (n + m) * (n + m) * n) * (N - 1) * B
        out(f"backward_pass {mname} N={N} B={B}", ms, B * (N - 1), "knot-updates/s", n=n, us_per_knot=ms * 1e3 / (N - 1),
            tflops=flops / ms * 1e-9)
        assert int(dfl.max().item()) == -1
