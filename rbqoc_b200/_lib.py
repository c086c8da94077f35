"""ctypes binding of librbq.so (the C ABI declared in include/rbq.h).

This is the same binding a Julia `ccall` stub makes; nothing here computes.
The library is built in-tree by `__graft_entry__.build()` / `make -C rbqoc_b200/csrc`.  There is
no CPU fallback: if the library is missing, or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbq.so")

HIST_BINS = 256
MAX_SAMPLE_STATES = 4

# enums of include/rbq.h
GATE_ZPIBY2, GATE_YPIBY2, GATE_XPIBY2, GATE_XPI = 1, 2, 3, 4
MODEL_SPIN11, MODEL_SPIN12, MODEL_SPIN13, MODEL_SPIN15 = 11, 12, 13, 15
MODEL_SPIN17, MODEL_SPIN18, MODEL_SPIN25, MODEL_SPIN30 = 17, 18, 25, 30
ALGO_SU2, ALGO_PADE = 0, 1
LINDBLAD_T1, LINDBLAD_T2 = 1, 2
INTEGRATOR_EXP, INTEGRATOR_RK2, INTEGRATOR_RK3, INTEGRATOR_RK4 = 0, 2, 3, 4
REG_CONTROL, REG_STATE = 0, 1
CON_BOUND, CON_GOAL, CON_NORM = 0, 1, 2
E_NULL, E_MODEL, E_SIZE, E_GATE, E_ALGO, E_NODEVICE, E_ALIGN = -1, -2, -3, -4, -5, -6, -7


class ModelParams(C.Structure):
    """`rbq_model_params`: POD mirror of the reference's immutable `Model` structs."""

    _fields_ = [
        ("model_id", C.c_int32),
        ("derivative_order", C.c_int32),
        ("decay_aware", C.c_int32),
        ("time_optimal", C.c_int32),
        ("sample_state_count", C.c_int32),
        ("algo", C.c_int32),
        ("nominal_offset", C.c_int32 * MAX_SAMPLE_STATES),
        ("integrator", C.c_int32),
        ("reserved", C.c_int32),
        ("fq", C.c_double),
        ("fq_samples", C.c_double * 2),
        ("namp", C.c_double),
        ("fq_cov", C.c_double),
        ("alpha", C.c_double),
        ("S_diag", C.c_double * 4),
    ]


class Constraint(C.Structure):
    """rbq_constraint (include/rbq.h): one row of the constraint table of rbq_al_cost_expansion_dev."""
    _fields_ = [("kind", C.c_int32), ("p", C.c_int32), ("k0", C.c_int64), ("k1", C.c_int64), ("nidx", C.c_int32),
                ("idx_off", C.c_int32), ("val_off", C.c_int64), ("dual_off", C.c_int64)]


class Stats(C.Structure):
    """`rbq_stats`: additive infidelity statistics of a sweep."""

    _fields_ = [
        ("count", C.c_int64),
        ("nonfinite", C.c_int64),
        ("sum", C.c_double),
        ("sumsq", C.c_double),
        ("min", C.c_double),
        ("max", C.c_double),
        ("hist", C.c_int64 * HIST_BINS),
    ]

    @property
    def mean(self) -> float:
        return self.sum / self.count if self.count else float("nan")

    @property
    def std(self) -> float:
        if self.count < 2:
            return float("nan")
        m = self.mean
        return max(self.sumsq / self.count - m * m, 0.0) ** 0.5


_P = C.c_void_p
_I64 = C.c_int64
_D = C.c_double
_MP = C.POINTER(ModelParams)
_SP = C.POINTER(Stats)

# name -> (restype, argtypes); the order follows include/rbq.h
SIGNATURES = {
    "rbq_version": (C.c_int, []),
    "rbq_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "rbq_destroy": (C.c_int, [_P]),
    "rbq_last_error": (C.c_char_p, [_P]),
    "rbq_set_stream": (C.c_int, [_P, _P]),
    "rbq_set_sweep_centre": (C.c_int, [_P, _D, _D]),
    "rbq_synchronize": (C.c_int, [_P]),
    "rbq_set_option": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "rbq_host_alloc": (C.c_int, [_P, _I64, C.POINTER(_P)]),
    "rbq_host_register": (C.c_int, [_P, _P, _I64]),
    "rbq_host_unregister": (C.c_int, [_P, _P]),
    "rbq_host_free": (C.c_int, [_P, _P]),
    "rbq_launch_count": (_I64, [_P]),
    "rbq_state_dim": (C.c_int, [_MP]),
    "rbq_control_dim": (C.c_int, [_MP]),
    "rbq_discrete_dynamics": (C.c_int, [_P, _MP, _P, _P, _D, _D, _P]),
    "rbq_discrete_jacobian": (C.c_int, [_P, _MP, _P, _P, _D, _D, _P]),
    "rbq_rollout": (C.c_int, [_P, _MP, _I64, _I64, _P, _P, _P, _P]),
    "rbq_rollout_closedloop": (C.c_int, [_P, _MP, _I64, _P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "rbq_rollout_closedloop_dev": (C.c_int, [_P, _MP, _I64, _P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "rbq_jacobians": (C.c_int, [_P, _MP, _I64, _P, _P, _P, _P, _P]),
    "rbq_rollout_dev": (C.c_int, [_P, _MP, _I64, _I64, _P, _P, _P, _P]),
    "rbq_jacobians_dev": (C.c_int, [_P, _MP, _I64, _P, _P, _P, _P, _P]),
    "rbq_gate_error_expansion": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _P, _P]),
    "rbq_gate_error_expansion_dev": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _P, _P]),
    "rbq_sampling_cost_expansion": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rbq_mc_static": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _P, _P, _P, C.c_int, _P, _SP]),
    "rbq_mc_static_dev": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _P, _P, _P, C.c_int, _P, _P]),
    "rbq_mc_static_philox_dev": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _I64, C.c_uint64, _D, _D, _D, _D,
                                           C.c_int, _P, _P]),
    "rbq_mc_static_philox": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _I64, C.c_uint64, _D, _D, _D, _D,
                                       C.c_int, _P, _SP]),
    "rbq_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "rbq_group_destroy": (C.c_int, [_P]),
    "rbq_group_size": (C.c_int, [_P]),
    "rbq_group_handle": (_P, [_P, C.c_int]),
    "rbq_group_last_error": (C.c_char_p, [_P]),
    "rbq_group_mc_static_philox": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _I64, C.c_uint64, _D, _D, _D, _D, C.c_int, _P, _SP]),
    "rbq_philox_samples": (C.c_int, [_P, _I64, _I64, C.c_uint64, _D, _D, _D, _D, _P, _P, _P]),
    "rbq_stats_reset_dev": (C.c_int, [_P, _P]),
    "rbq_stats_merge": (C.c_int, [_SP, _SP]),
    "rbq_mc_noise": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _D, _P, _I64, _D, _P, C.c_int, _P, _SP]),
    "rbq_mc_noise_dev": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _D, _P, _I64, _D, _P, C.c_int, _P, _P]),
    "rbq_mc_pink_philox_dev": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _I64, C.c_uint64, _D, _D, _D, C.c_int,
                                         _P, _P]),
    "rbq_mc_pink_philox": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _I64, C.c_uint64, _D, _D, _D, C.c_int,
                                     _P, _SP]),
    "rbq_pink_noise": (C.c_int, [_P, _I64, _I64, C.c_uint64, _D, _D, _I64, _P]),
    "rbq_lindblad": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _D, _P, _P, C.c_int, _D, _D, _P, _P, _SP]),
    "rbq_lindblad_dev": (C.c_int, [_P, _P, _I64, _D, C.c_int, _I64, _I64, _D, _P, _P, C.c_int, _D, _D, _P, _P, _P]),
    "rbq_mc_static_segments": (C.c_int, [_P, _P, _P, _I64, C.c_int, _I64, _I64, _P, _P, _P, _P, _SP]),
    "rbq_mc_noise_segments": (C.c_int, [_P, _P, _P, _P, _I64, C.c_int, _I64, _I64, _D, _P, _I64, _P, _P, _SP]),
    "rbq_mc_timeline": (C.c_int, [_P, _P, _P, _P, _I64, _P, _I64, C.c_int, _I64, _P, _P, _P, _I64, _P, _P, _SP]),
    "rbq_mc_timeline_states": (C.c_int, [_P, _P, _P, _P, _I64, _P, _I64, C.c_int, _I64, _P, _P, _P, _I64, _P, _P, _P, _SP]),
    "rbq_lindblad_timeline": (C.c_int, [_P, _P, _P, _I64, _P, _I64, C.c_int, _I64, _D, _P, _P, C.c_int, _D, _D, _P, _P, _SP]),
    "rbq_lindblad_segments": (C.c_int, [_P, _P, _P, _I64, C.c_int, _I64, _I64, _D, _P, _P, C.c_int, _D, _P, _P, _SP]),
    "rbq_backward_pass": (C.c_int, [_P, C.c_int, C.c_int, _I64, _I64, _P, _P, _P, _P, _P, _P, _D, C.c_int, _P, _P, _P, _P]),
    "rbq_backward_pass_dev": (C.c_int, [_P, C.c_int, C.c_int, _I64, _I64, _P, _P, _P, _P, _P, _P, _D, C.c_int, _P, _P, _P, _P]),
    "rbq_al_cost_expansion_dev": (C.c_int, [_P, C.c_int, C.c_int, _I64, _I64] + [_P] * 7 + [C.c_int] + [_P] * 5 + [_I64] + [_P] * 7),
    "rbq_al_dual_update_dev": (C.c_int, [_P, _I64, _I64, C.c_int, _P, _I64, _P, _P, _P, _I64, _D, _D, _D, C.c_int, _P]),
    "rbq_al_cost_expansion": (C.c_int, [_P, C.c_int, C.c_int, _I64, _I64] + [_P] * 7 + [C.c_int, _P, _P, _I64, _P, _I64, _P, _P, _I64] + [_P] * 7),
    "rbq_fp64_peak": (C.c_int, [_P, C.c_int, C.POINTER(_D)]),
}


class RbqError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        kind = "argument error" if code < 0 else "CUDA error"
        super().__init__(f"{where}: {kind} {code}" + (f" ({detail})" if detail else ""))


_lib = None


def load():
    """dlopen librbq.so and attach signatures.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RbqError(E_NODEVICE, "rbqoc_b200", f"{LIB_PATH} is missing: run __graft_entry__.build() "
                           "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
