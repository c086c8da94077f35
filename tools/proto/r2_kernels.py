"""Round-2 kernel probe: event timings of the Lindblad table sweep, the structured unscented Jacobians and the scan rollout
(run plain, then under ncu for the launch list / full capture)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
what = sys.argv[1:] or ["lindblad", "jac30", "rollout"]
if what == ["all"]:
    what = ["mc_static", "lindblad", "jac30", "rollout"]
reps = int(os.environ.get("REPS", "5"))
def timed(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
N = 1001
if "mc_static" in what:
    S = 1 << 20
    controls = spin.pulse_sin(N)[:-1]
    fq, da, psi0 = eng.philox_samples(0, S, 2026, spin.FQ, 0.01, 0.03, 2.574e-5)
    dcn, dfq, dda, dpsi = d(controls), d(fq), d(da), d(psi0)
    dfid = torch.empty(S, 2, dtype=torch.float64, device=dev)
    dst = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
    ms = timed(lambda: eng.mc_static_dev(dcn, N - 1, 0.1, rbq.GATE_ZPIBY2, 1, S, dfq, dda, dpsi, rbq.ALGO_SU2, dfid, dst))
    print(f"mc_static S={S}: {ms:.4f} ms  {S * (N - 1) / ms * 1e-6:.2f} G trajectory-steps/s", flush=True)
if "lindblad" in what:
    S = 100_000
    rng = np.random.default_rng(0)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2)); psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4)).view(np.float64)
    controls = spin.pulse_sin(N)[:-1]
    dcn, dda, drho = d(controls), d(2.574e-5 * rng.standard_normal(S)), d(vec)
    dfid = torch.empty(S, 2, dtype=torch.float64, device=dev); dro = torch.empty(S, 8, dtype=torch.float64, device=dev)
    dst = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
    for opt in (None, "0"):
        eng.set_option("RBQ_LINDBLAD_TABLE", opt)
        ms = timed(lambda: eng.lindblad_dev(dcn, N - 1, 0.1, rbq.GATE_XPIBY2, 1, S, spin.FQ, dda, drho, rbq.LINDBLAD_T1, 0.0, 0.0, dfid, dro, dst))
        print(f"lindblad S={S} table={opt}: {ms:.4f} ms  {S * (N - 1) / ms * 1e-6:.2f} G density-steps/s", flush=True)
    eng.set_option("RBQ_LINDBLAD_TABLE", None)
    for spt in ("1", "2"):
        eng.set_option("RBQ_LINDBLAD_SPT", spt)
        for Sx in (S, 20000, 50000):
            ms = timed(lambda: eng.lindblad_dev(dcn, N - 1, 0.1, rbq.GATE_XPIBY2, 1, Sx, spin.FQ, dda, drho, rbq.LINDBLAD_T1, 0.0, 0.0, dfid, dro, dst))
            print(f"lindblad S={Sx} spt={spt}: {ms:.4f} ms  {Sx * (N - 1) / ms * 1e-6:.2f} G density-steps/s", flush=True)
    eng.set_option("RBQ_LINDBLAD_SPT", None)
if "jac30" in what:
    for mname, model, x0, B in (("spin30", spin.Spin30Model(), spin.spin30_x0(), 128), ("spin25", spin.Spin25Model(), spin.spin25_x0(), 128)):
        n, m = model.n, model.m
        dtv = np.full(N - 1, 0.1)
        U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None] * np.ones((1, m))
        X = eng.rollout(model.params, x0, U, dtv)
        for Bk in (1, B):
            K = Bk * (N - 1)
            dXk, dUk, ddk = d(np.tile(X[:-1], (Bk, 1))), d(np.tile(U, (Bk, 1))), d(np.tile(dtv, Bk))
            dtk = d(np.tile(spin.knot_times(dtv), Bk))
            dAB = torch.empty(K, n * (n + m), dtype=torch.float64, device=dev)
            for opt in (None, "0"):
                eng.set_option("RBQ_JAC_STRUCTURED", opt)
                ms = timed(lambda: eng.jacobians_dev(model.params, K, dXk, dUk, ddk, dAB, t=dtk))
                print(f"{mname} K={K} structured={opt}: {ms:.4f} ms  {K * n * (n + m) * 8 / ms * 1e-9:.2f} TB/s written", flush=True)
            eng.set_option("RBQ_JAC_STRUCTURED", None)
if "rollout" in what:
    for mname, model, x0 in (("spin13", spin.Spin13Model(), spin.spin13_x0()), ("spin12", spin.Spin12Model(), spin.spin12_x0()),
                             ("spin15DA", spin.Spin15Model(True, False), np.concatenate([spin.spin13_x0(), [0.0]]))):
        n, m = model.n, model.m
        dtv = np.full(N - 1, 0.1)
        U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None] * np.ones((1, m))
        dx0, dU, ddt = d(x0[None]), d(U), d(dtv)
        dX = torch.empty(N, n, dtype=torch.float64, device=dev)
        for opt in ("1", "0"):
            eng.set_option("RBQ_ROLLOUT_SCAN", opt)
            ms = timed(lambda: eng.rollout_dev(model.params, 1, N, dx0, dU, ddt, dX))
            print(f"{mname} rollout N={N} scan={opt}: {ms * 1e3:.1f} us", flush=True)
        eng.set_option("RBQ_ROLLOUT_SCAN", None)
