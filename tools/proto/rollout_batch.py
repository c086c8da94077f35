import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
N = 1001
for mname, model, x0 in (("spin13", spin.Spin13Model(), spin.spin13_x0()), ("spin12", spin.Spin12Model(), spin.spin12_x0())):
    n, m = model.n, model.m
    dtv = np.full(N - 1, 0.1); U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None] * np.ones((1, m))
    for B in (16, 64, 1024, 4096, 32768):
        dX0, dU, ddt = d(np.tile(x0, (B, 1))), d(np.tile(U, (B, 1, 1))), d(dtv)
        dX = torch.empty(B, N, n, dtype=torch.float64, device=dev)
        for mode in ("0", "1"):
            os.environ["RBQ_ROLLOUT_PACKED"] = mode
            for _ in range(2): eng.rollout_dev(model.params, B, N, dX0, dU, ddt, dX)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3): eng.rollout_dev(model.params, B, N, dX0, dU, ddt, dX)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3
            print(f"{mname} B={B:6d} packed={mode}: {ms:8.4f} ms  {B * (N - 1) / ms * 1e3:.3e} knot-steps/s", flush=True)
