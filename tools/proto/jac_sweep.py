import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
N = 1001
for mname, model, x0, B in (("spin13", spin.Spin13Model(), spin.spin13_x0(), 1024), ("spin12", spin.Spin12Model(), spin.spin12_x0(), 256),
                            ("spin30", spin.Spin30Model(), spin.spin30_x0(), 128)):
    n, m = model.n, model.m
    dtv = np.full(N - 1, 0.1)
    U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None] * np.ones((1, m))
    X = eng.rollout(model.params, x0, U, dtv)
    K = B * (N - 1)
    dXk, dUk, ddk = d(np.tile(X[:-1], (B, 1))), d(np.tile(U, (B, 1))), d(np.tile(dtv, B))
    dAB = torch.empty(K, n * (n + m), dtype=torch.float64, device=dev)
    for kb in sys.argv[1:]:
        os.environ["RBQ_JAC_STAGE_KB"] = kb
        for _ in range(2): eng.jacobians_dev(model.params, K, dXk, dUk, ddk, dAB)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): eng.jacobians_dev(model.params, K, dXk, dUk, ddk, dAB)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{mname} K={K} stage={kb} KB: {ms:.4f} ms  {K * n * (n + m) * 8 / ms * 1e-9:.2f} TB/s written", flush=True)
