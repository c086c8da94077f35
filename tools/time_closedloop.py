import sys, time, numpy as np
sys.path.insert(0, ".")
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
eng = rbq.Engine(0)
for name, model, x0 in (("spin13", spin.Spin13Model(), spin.spin13_x0()), ("spin30", spin.Spin30Model(), spin.spin30_x0())):
    n, m = model.n, model.m
    for N in (101, 1001, 4001):
        dtv = np.full(N - 1, 0.1)
        U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None] * np.ones((1, m))
        X = eng.rollout(model.params, x0, U, dtv)
        K = 1e-3 * np.ones((N - 1, m, n)); d = 1e-4 * np.ones((N - 1, m))
        for na in (1, 8):
            alphas = 0.5 ** np.arange(na)
            eng.rollout_closedloop(model.params, X, U, K, d, dtv, alphas)
            t0 = time.perf_counter()
            for _ in range(10): eng.rollout_closedloop(model.params, X, U, K, d, dtv, alphas)
            tc = (time.perf_counter() - t0) / 10 * 1e3
            t0 = time.perf_counter()
            for _ in range(10): eng.rollout(model.params, x0, U, dtv)
            to = (time.perf_counter() - t0) / 10 * 1e3
            print(f"{name} N={N} nalpha={na}: closed-loop {tc:.3f} ms, open-loop {to:.3f} ms")
