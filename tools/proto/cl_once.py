"""Closed-loop rollout of 8 candidates and the open-loop rollout, device resident, 1001 knots: ms per call."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
N = 1001
for name, model, x0 in (("spin13", spin.Spin13Model(), spin.spin13_x0()), ("spin12", spin.Spin12Model(), spin.spin12_x0()),
                        ("spin30", spin.Spin30Model(), spin.spin30_x0())):
    n, m = model.n, model.m
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dtv = np.full(N - 1, 0.1); U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None]
    X = eng.rollout(model.params, x0, U, dtv)
    rng = np.random.default_rng(0)
    dX, dU, ddt = d(X), d(U), d(dtv)
    dK = d(1e-3 * rng.standard_normal((N - 1, n, m))); dd = d(1e-4 * rng.standard_normal((N - 1, m))); dal = d(0.5 ** np.arange(8))
    dXc = torch.empty(8, N, n, dtype=torch.float64, device=dev); dUc = torch.empty(8, N - 1, m, dtype=torch.float64, device=dev)
    dXo = torch.empty(1, N, n, dtype=torch.float64, device=dev); dx0 = d(x0)
    def timed(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(10):
            a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
        return best
    print(name, "closed-loop x8 %.3f ms" % timed(lambda: eng.rollout_closedloop_dev(model.params, N, dX, dU, dK, dd, ddt, 8, dal, dXc, dUc)),
          "open-loop %.3f ms" % timed(lambda: eng.rollout_dev(model.params, 1, N, dx0, dU, ddt, dXo)))
