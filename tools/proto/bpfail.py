import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import rbqoc_b200 as rbq
from test_oracle_cpu import _lq_problem
eng = rbq.Engine(0)
for n, N in ((43, 61), (56, 47), (58, 33)):
    rng = np.random.default_rng(4000 + n)
    probs = [_lq_problem(rng, N, n, 1) for _ in range(3)]
    stack = lambda i: np.stack([p[i] for p in probs])
    AB, Q, q, R, Eux, r = (stack(i) for i in (2, 3, 4, 5, 6, 8))
    R2 = R.copy(); R2[1, N // 2] = -1e3
    for cl in ("1", "2", "4"):
        eng.set_option("RBQ_BP_CLUSTER", cl)
        K, d, dV, fail = eng.backward_pass(AB, Q, q, R2, Eux, r, rho=0.23, reg_type=rbq.REG_CONTROL)
        print(n, cl, list(fail), flush=True)
