"""Shared helpers: the golden model cases as (name, ModelParams-builder, arrays)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FQ = 1.4e-2
MODEL_KW = {
    "spin13": dict(model_id=13),
    "spin12": dict(model_id=12, fq_samples=(FQ * 1.01, FQ * 0.99)),
    "spin18": dict(model_id=18, namp=5.21e-6 / 0.202407),
    "spin15_DA": dict(model_id=15, decay_aware=1),
    "spin15_DA_TO": dict(model_id=15, decay_aware=1, time_optimal=1),
    "spin11_DO2": dict(model_id=11, derivative_order=2),
    "spin17_DO2": dict(model_id=17, derivative_order=2),
    "spin30": dict(model_id=30, fq_cov=1.4e-4, alpha=1.0, S_diag=(1.0, 2.0, 3.0, 4.0), sample_state_count=1,
                   nominal_offset=(0, 0, 0, 0)),
    "spin30_x2": dict(model_id=30, fq_cov=1.4e-4, alpha=1.3, S_diag=(1.0, 1.0, 1.0, 1.0), sample_state_count=2,
                      nominal_offset=(0, 4, 0, 0)),
    "spin25_k2": dict(model_id=25, fq_cov=1.0, namp=2.574e-2, alpha=1.0, S_diag=(1.0, 2.0, 3.0, 4.0), sample_state_count=1,
                      nominal_offset=(4, 0, 0, 0)),
    "spin25_k7": dict(model_id=25, fq_cov=1.0, namp=2.574e-2, alpha=1.2, S_diag=(1.0, 1.0, 1.0, 1.0), sample_state_count=1,
                      nominal_offset=(0, 0, 0, 0)),
}


def golden_model_cases():
    z = np.load(os.path.join(GOLDEN, "models.npz"))
    out = []
    for name in MODEL_KW:
        for dt in (0.1, 0.01):
            k = f"{name}_dt{dt}"
            t = float(z[k + "_t"]) if (k + "_t") in z else 0.0
            out.append((k, name, dt, z[k + "_x"], z[k + "_u"], z[k + "_xn"], z[k + "_J"], t))
    return out


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
