"""Times the Lindblad sweep of BASELINE configs[3] alone (1e5 realisations x P_sin(1001)); device resident, CUDA events."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
S = 100000
rng = np.random.default_rng(0)
psi = rng.random((S, 2)) + 1j * rng.random((S, 2)); psi /= np.linalg.norm(psi, axis=1, keepdims=True)
rho0 = np.einsum("si,sj->sij", psi, psi.conj())
vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4)).view(np.float64)
da = 2.574e-5 * rng.standard_normal(S)
stats = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
for name, controls, dt, gates in (("P_sin(1001) dt=0.1", spin.pulse_sin(1001)[:-1], 0.1, 1), ("P_Y2(1947) dt=0.01", spin.controls_ypiby2(100.0), 0.01, 200)):
    dc, dda, drho = d(controls), d(da), d(vec)
    dfid = torch.empty(S, gates + 1, dtype=torch.float64, device=dev); dro = torch.empty(S, 8, dtype=torch.float64, device=dev)
    nk = controls.shape[0]
    f = lambda: eng.lindblad_dev(dc, nk, dt, rbq.GATE_YPIBY2, gates, S, spin.FQ, dda, drho, rbq.LINDBLAD_T1, 0.0, 0.0, dfid, dro, stats)
    best = 1e9
    for rep in range(3):
        for _ in range(3): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): f()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 20)
    print(json.dumps({"path": f"lindblad T1 {name} gates={gates}", "ms": round(best, 4), "per_s": S * nk / best * 1e3}), flush=True)
