"""Seed-in noise sweep (rbq_mc_pink_philox_dev) with the pink-noise rows materialised first against generated inside the sweep."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
controls, dt, ndi = spin.controls_xpiby2(100.0), 0.01, 10.0
nk = controls.shape[0]
dc = d(controls)
stats = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
for S, gates in ((65536, 20), (65536, 1), (1 << 20, 1), (1000, 200)):
    dfid = torch.empty(S, gates + 1, dtype=torch.float64, device=dev)
    res = {}
    for mat in ("1", "0"):
        eng.set_option("RBQ_NOISE_MATERIALISE", mat)
        f = lambda: eng.mc_pink_philox_dev(dc, nk, dt, rbq.GATE_XPIBY2, gates, 0, S, 7, spin.FQ, spin.NAMP_PREFACTOR, ndi, rbq.ALGO_SU2, dfid, stats)
        for _ in range(2): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): f()
        b.record(); torch.cuda.synchronize()
        res[mat] = (a.elapsed_time(b) / 5, dfid.clone())
    print(json.dumps({"S": S, "gates": gates, "knots": nk, "materialised_ms": round(res["1"][0], 4), "on_the_fly_ms": round(res["0"][0], 4),
                      "steps_per_s_materialised": S * gates * nk / res["1"][0] * 1e3, "max_abs_diff": float((res["1"][1] - res["0"][1]).abs().max())}), flush=True)
