"""GPU tuning aid: end-to-end time of rbq_mc_static (pinned host buffers) for several chunk schedules."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin

S, NK, DT = 1 << 20, 1000, 0.1
eng = rbq.Engine(0)
fq, da, psi0 = eng.philox_samples(0, S, 1, spin.FQ, 0.01, 0.03, 2.574e-5)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
fq, da, psi0 = pin(fq), pin(da), pin(psi0)
fid = torch.empty(S, 2, dtype=torch.float64).pin_memory().numpy()
controls = spin.pulse_sin(NK + 1)[:-1]
for sched in sys.argv[1:] or ["1,2", "1,3", "1,6", "2,2", "1,1", "7,7"]:
    eng.set_option("RBQ_MC_CHUNK_WAVES", sched)
    for _ in range(3):
        eng.mc_static(controls, DT, rbq.GATE_ZPIBY2, 1, fq, da, psi0, fid_out=fid)
    t0 = time.perf_counter()
    for _ in range(20):
        eng.mc_static(controls, DT, rbq.GATE_ZPIBY2, 1, fq, da, psi0, fid_out=fid)
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"chunks (first,rest waves) {sched}: {ms:.4f} ms  {S * NK / ms * 1e3:.4e} steps/s", flush=True)
for s_small in (1024, 65536):
    t0 = time.perf_counter()
    for _ in range(50):
        eng.mc_static(controls, DT, rbq.GATE_ZPIBY2, 1, fq[:s_small], da[:s_small], psi0[:s_small], fid_out=fid[:s_small])
    print(f"S={s_small}: {(time.perf_counter() - t0) / 50 * 1e3:.4f} ms per call")
