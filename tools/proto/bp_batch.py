"""One-warp against eight-warp form of the backward pass for n = 11 over batch sizes (threshold of launch_backward_pass)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
n, m, N = 11, 1, 1001
rng = np.random.default_rng(0)
for B in (1, 74, 148, 222, 296, 370, 444, 592, 1184):
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    AB = d(0.1 * rng.standard_normal((B, N - 1, n + m, n))); Exx = d(np.tile(np.eye(n), (B, N, 1, 1))); Ex = d(rng.standard_normal((B, N, n)))
    Euu = d(np.ones((B, N - 1, m, m))); Eux = d(np.zeros((B, N - 1, n, m))); Eu = d(rng.standard_normal((B, N - 1, m)))
    K = torch.empty(B, N - 1, n * m, dtype=torch.float64, device=dev); dd = torch.empty(B, N - 1, m, dtype=torch.float64, device=dev)
    dV = torch.empty(B, 2, dtype=torch.float64, device=dev); fl = torch.empty(B, dtype=torch.int64, device=dev)
    res = {}
    for w in ("1", "8"):
        os.environ["RBQ_BP_WARPS"] = w
        for _ in range(2):
            eng.backward_pass_dev(n, m, N, B, AB, Exx, Ex, Euu, Eux, Eu, 1e-3, 0, K, dd, dV, fl)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            a.record(); eng.backward_pass_dev(n, m, N, B, AB, Exx, Ex, Euu, Eux, Eu, 1e-3, 0, K, dd, dV, fl); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res[w] = best
    print(f"B={B:5d}  one-warp {res['1']:7.3f} ms   eight-warp {res['8']:7.3f} ms")
