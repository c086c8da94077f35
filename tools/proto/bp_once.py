import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rbqoc_b200 as rbq
dev = torch.device("cuda", 0)
eng = rbq.Engine(0); eng.use_torch_stream()
n, m, N, B = int(sys.argv[1]), int(os.environ.get("BP_M", "1")), (int(sys.argv[2]) if len(sys.argv) > 2 else 201), (int(sys.argv[3]) if len(sys.argv) > 3 else 1)
rng = np.random.default_rng(0)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
AB = d(0.1 * rng.standard_normal((B, N - 1, n + m, n))); Exx = d(np.tile(np.eye(n), (B, N, 1, 1))); Ex = d(rng.standard_normal((B, N, n)))
Euu = d(np.ones((B, N - 1, m, m))); Eux = d(np.zeros((B, N - 1, n, m))); Eu = d(rng.standard_normal((B, N - 1, m)))
K = torch.empty(B, N - 1, n * m, dtype=torch.float64, device=dev); dd = torch.empty(B, N - 1, m, dtype=torch.float64, device=dev)
dV = torch.empty(B, 2, dtype=torch.float64, device=dev); fl = torch.empty(B, dtype=torch.int64, device=dev)
for _ in range(3):
    eng.backward_pass_dev(n, m, N, B, AB, Exx, Ex, Euu, Eux, Eu, 1e-3, 0, K, dd, dV, fl)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(10):
    a.record(); eng.backward_pass_dev(n, m, N, B, AB, Exx, Ex, Euu, Eux, Eu, 1e-3, 0, K, dd, dV, fl); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print("n", n, "us per knot (best of 10)", best * 1e3 / (N - 1), "fail", fl.tolist())


if hasattr(eng._lib, "rbq_debug_bp_prof"):
    import ctypes as C
    buf = (C.c_longlong * 32)()
    eng._lib.rbq_debug_bp_prof(buf, 1)
    eng.backward_pass_dev(n, m, N, B, AB, Exx, Ex, Euu, Eux, Eu, 1e-3, 0, K, dd, dV, fl); torch.cuda.synchronize()
    eng._lib.rbq_debug_bp_prof(buf, 0)
    names = ["wait AB + sync/mirror", "T = S AB", "G tiles (thread 0)", "g rows (thread 0)", "barrier", "gains", "barrier_or", "symmetrise", "stage next jacobian", "V pass", "sv/dV (other threads)", "cluster wait (top)"]
    for off, who in ((0, "first CTA"), (16, "gains CTA (cluster form)")):
        tot = sum(buf[off:off + 12])
        if tot == 0: continue
        print(" ", who)
        for i in range(12): print(f"  {names[i]:20s} {buf[off + i] / (N - 1):9.0f} cycles/knot  {100 * buf[off + i] / tot:5.1f}%")
        print("  total cycles/knot", tot / (N - 1))
