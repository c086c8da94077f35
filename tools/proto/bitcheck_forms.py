import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import rbqoc_b200 as rbq
from rbqoc_b200 import spin
eng = rbq.Engine(0)
for nk, dt in ((1000, 0.1), (1001 - 4, 0.1), (5680, 0.01)):
    controls = spin.pulse_sin(nk + 1)[:-1] if dt == 0.1 else spin.controls_xpiby2(100.0)
    S = 200000 + 77
    fq, da, psi0 = eng.philox_samples(0, S, 5, spin.FQ, 0.01, 0.03, 2.574e-5)
    da[np.arange(S) % 6400 < 32] += 0.2
    out = {}
    for v in ("161", "3161", "2161"):
        eng.set_option("RBQ_MC_VARIANT", v)
        f, st = eng.mc_static(controls, dt, rbq.GATE_XPIBY2, 2, fq, da, psi0)
        out[v] = (f.copy(), st.sum, list(st.hist[:]))
    print(nk, dt, "3161 == 161 bitwise:", np.array_equal(out["3161"][0], out["161"][0]), out["3161"][1] == out["161"][1], out["3161"][2] == out["161"][2],
          " 2161 vs 161 max:", float(np.max(np.abs(out["2161"][0] - out["161"][0]))))
