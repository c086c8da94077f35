"""GPU tuning aid: times the closed-form MC sweep for every compiled (samples/thread, min blocks/SM) variant."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin

S, NK, DT = 1 << 20, 1000, 0.1
if len(sys.argv) > 1:
    NK = int(sys.argv[1]); DT = float(sys.argv[2])
dev = torch.device("cuda", 0)
eng = rbq.Engine(0)
eng.use_torch_stream()
fq, da, psi0 = eng.philox_samples(0, S, 1, spin.FQ, 0.01, 0.03, 2.574e-5)
d = lambda a: torch.from_numpy(a).to(dev)
d_c, d_fq, d_da, d_psi0 = d(spin.pulse_sin(NK + 1)[:-1]), d(fq), d(da), d(psi0)
d_fid = torch.empty(S, 2, dtype=torch.float64, device=dev)
d_st = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
for v in sys.argv[3:] or ["112", "122", "132", "142", "212", "222", "232", "242", "141", "161", "181", "241", "261", "281"]:
    eng.set_option("RBQ_MC_VARIANT", v)
    for _ in range(3):
        eng.mc_static_dev(d_c, NK, DT, rbq.GATE_ZPIBY2, 1, S, d_fq, d_da, d_psi0, rbq.ALGO_SU2, d_fid, d_st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        eng.mc_static_dev(d_c, NK, DT, rbq.GATE_ZPIBY2, 1, S, d_fq, d_da, d_psi0, rbq.ALGO_SU2, d_fid, d_st)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"variant {v}: {ms:.4f} ms  {S * NK / ms * 1e3:.4e} steps/s", flush=True)
print("fp64 peak TFLOP/s:", eng.fp64_peak(20000))
