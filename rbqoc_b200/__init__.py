"""rbqoc_b200 — B200-native batched propagation engine for rbqoc's hot path.

csrc/     hand-written sm_100a kernels + the C ABI (include/rbq.h) -> librbq.so
_lib.py   ctypes binding of the C ABI
engine.py numpy / torch-tensor marshalling over one rbq_handle
spin.py   host-side mirror of the reference's spin-model / simulation interface
dist.py   sample sharding across ranks + the one small all-gather of statistics
"""
from . import _lib
from ._lib import (ALGO_PADE, ALGO_SU2, GATE_XPI, GATE_XPIBY2, GATE_YPIBY2, GATE_ZPIBY2, LINDBLAD_T1, LINDBLAD_T2,
                   REG_CONTROL, REG_STATE, CON_BOUND, CON_GOAL, CON_NORM, Constraint, ModelParams, RbqError, Stats)
from .engine import Engine, EngineGroup, stats_from_words, stats_merge, STATS_INT64_WORDS

__all__ = ["Engine", "EngineGroup", "ModelParams", "Stats", "RbqError", "stats_from_words", "stats_merge", "STATS_INT64_WORDS",
           "ALGO_SU2", "ALGO_PADE", "GATE_ZPIBY2", "GATE_YPIBY2", "GATE_XPIBY2", "GATE_XPI", "LINDBLAD_T1",
           "LINDBLAD_T2", "REG_CONTROL", "REG_STATE", "CON_BOUND", "CON_GOAL", "CON_NORM", "Constraint"]
