"""One open-loop spin30 rollout of 1001 knots on the device (for ncu source-level captures of warp_rollout_kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
N = 1001
model, x0 = spin.Spin30Model(), spin.spin30_x0()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
dU = d(1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None]); ddt = d(np.full(N - 1, 0.1)); dx0 = d(x0)
dX = torch.empty(1, N, model.n, dtype=torch.float64, device=dev)
for _ in range(4):
    eng.rollout_dev(model.params, 1, N, dx0, dU, ddt, dX)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(10):
    a.record(); eng.rollout_dev(model.params, 1, N, dx0, dU, ddt, dX); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(f"spin30 open-loop rollout, {N} knots: {best * 1e3:.1f} us ({best * 1e3 / (N - 1):.3f} us per knot)")
