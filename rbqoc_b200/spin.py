"""Host-side mirror of the reference's spin-qubit interface for the propagation hot path.

Same names, argument meaning and error behaviour as the Julia side, so that parity tests read like
the reference's own scripts; everything numeric is forwarded to the CUDA engine (no CPU arithmetic
beyond building inputs).  Citations are file:line in SchusterLab/rbqoc.

  constants / operators        src/spin/spin.jl:122-345
  Model, state_dim, ...        src/spin/spin13.jl:37-40, spin12.jl:52-56, spin30.jl:60-70, spin15.jl:40-46,
                               spin11.jl:51-57, spin17.jl, spin18.jl:50-55, spin25.jl:67-75
  discrete_dynamics            RD.discrete_dynamics(::Type{RK3}, model, x, u, t, dt)  (spin13.jl:44-58 ...)
  discrete_jacobian            RD.discrete_jacobian!(::Type{RK3}, ∇f, model, z) (RobotDynamics default: ForwardDiff)
  run_sim_prop / evaluate_*    src/spin/spin.jl:1553-1672
  analytic pulses              src/spin/spin14.py:60-241
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import ModelParams
from .engine import Engine

# ---- constants (spin.jl:122-173) --------------------------------------------------------------------
HDIM = 2
HDIM_ISO = 4
FQ = 1.4e-2                       # GHz, spin.jl:148
MAX_CONTROL_NORM_0 = 0.5          # spin.jl:156
FBFQ_A = 0.202407
FBFQ_B = 0.5
AYPIBY2 = 1.25e-1
FBFQ_NAMP = 5.21e-6               # spin.jl:163
NAMP_PREFACTOR = FBFQ_NAMP / FBFQ_A   # spin.jl:164
GAMMAC = 1 / 3e5                  # spin.jl:162
SQRTLNIR = 4                      # spin.jl:163
GAMMAF_PREFACTOR = FBFQ_NAMP * SQRTLNIR * 2 * math.pi   # spin.jl:165: Gamma(t) = GAMMAC + 2 GAMMAF_PREFACTOR^2 t (spin.jl:549-552)
DT_NOISE_INV = 1e1
TTOT_ZPIBY2 = 17.857142857142858  # spin.jl:557


class GateType(enum.IntEnum):     # spin.jl:29-34
    zpiby2 = 1
    ypiby2 = 2
    xpiby2 = 3
    xpi = 4


class DynamicsType(enum.IntEnum):  # spin.jl:36-58 (same integer values): the members run_sim_prop integrates (DT_INT, spin.jl:986-996)
    schroed = 1
    lindbladnodis = 2
    lindbladt1 = 3
    ypiby2nodis = 4
    ypiby2t1 = 5
    xpiby2nodis = 6
    xpiby2t1 = 7
    zpiby2nodis = 9
    zpiby2t1 = 10
    xpiby2da = 13
    lindbladt2 = 15
    schroedda = 18


class RK3:                        # integrator tag the reference always dispatches on (rbqoc.jl:49-54)
    pass


class RK2:                        # RobotDynamics' other tags (IT_RDI, rbqoc.jl:49-54)
    pass


class RK4:
    pass


class IntegratorType(enum.IntEnum):   # rbqoc.jl:41-46
    rk2 = 1
    rk3 = 2
    rk4 = 3


_INTEGRATOR_ID = {None: _lib.INTEGRATOR_EXP, IntegratorType.rk2: _lib.INTEGRATOR_RK2, IntegratorType.rk3: _lib.INTEGRATOR_RK3,
                  IntegratorType.rk4: _lib.INTEGRATOR_RK4}


def with_integrator(model: "Model", integrator) -> "Model":
    """The model with its CONTINUOUS dynamics (RD.dynamics of the reference's RK variants, spin15_standalone.jl:263-312,
    spin19.jl:66-129, spin21.jl:63-109) stepped by RobotDynamics' explicit RK2 / RK3 / RK4 (`integrator_type` of those files'
    run_traj) instead of the file's own exponential discrete_dynamics.  integrator None restores the exponential step.
    spin13 / spin15 / spin11 / spin17 only."""
    import copy

    m = copy.deepcopy(model)
    m.params.integrator = _INTEGRATOR_ID[None if integrator is None else IntegratorType(integrator)]
    if control_dim(m) < 0:
        raise ValueError("this model has no Runge-Kutta variant")
    return m


# real isomorphism helpers (rbqoc.jl:189-210)
def get_vec_iso(v):
    v = np.asarray(v, dtype=np.complex128)
    return np.concatenate([v.real, v.imag])


def get_vec_uniso(v):
    v = np.asarray(v, dtype=np.float64)
    h = v.shape[0] // 2
    return v[:h] + 1j * v[h:]


def get_mat_iso(m):
    m = np.asarray(m, dtype=np.complex128)
    return np.block([[m.real, -m.imag], [m.imag, m.real]])


SIGMAX = np.array([[0, 1], [1, 0]], dtype=np.complex128)
SIGMAZ = np.array([[1, 0], [0, -1]], dtype=np.complex128)
NEGI_H0_ISO = math.pi * get_mat_iso(-1j * SIGMAZ)     # spin.jl:234
NEGI_H1_ISO = math.pi * get_mat_iso(-1j * SIGMAX)     # spin.jl:236
_R2 = 1 / math.sqrt(2)
GATES = {                                              # spin.jl:301-312
    GateType.zpiby2: np.array([[1 - 1j, 0], [0, 1 + 1j]]) * _R2,
    GateType.ypiby2: np.array([[1, -1], [1, 1]], dtype=np.complex128) * _R2,
    GateType.xpiby2: np.array([[1, -1j], [-1j, 1]]) * _R2,
    GateType.xpi: np.array([[0, -1j], [-1j, 0]]),
}
IS1_ISO = np.array([1.0, 0, 0, 0])                     # spin.jl:325-332
IS2_ISO = np.array([0.0, 1, 0, 0])
IS3_ISO = np.array([1.0, 0, 0, 1]) * _R2
IS4_ISO = np.array([1.0, -1, 0, 0]) * _R2


# ---- models ---------------------------------------------------------------------------------------------
@dataclass
class Model:
    """One of the reference's `Model <: RD.AbstractModel` structs, flattened to `rbq_model_params`."""

    params: ModelParams
    name: str = ""

    @property
    def n(self) -> int:
        return state_dim(self)

    @property
    def m(self) -> int:
        return control_dim(self)


def _params(**kw) -> ModelParams:
    p = ModelParams()
    p.fq = FQ
    p.alpha = 1.0
    for k, v in kw.items():
        if k in ("nominal_offset", "fq_samples", "S_diag"):
            arr = getattr(p, k)
            for i, x in enumerate(v):
                arr[i] = x
        else:
            setattr(p, k, v)
    return p


def Spin13Model(fq=FQ) -> Model:
    return Model(_params(model_id=_lib.MODEL_SPIN13, fq=fq), "spin13")


def Spin12Model(fq_cov=FQ * 1e-2, fq=FQ) -> Model:
    """spin12.jl:204-205: h0_samples = (FQ + fq_cov, FQ - fq_cov) (fq_cov used as a deviation)."""
    return Model(_params(model_id=_lib.MODEL_SPIN12, fq=fq, fq_samples=(fq + fq_cov, fq - fq_cov)), "spin12")


def Spin18Model(namp=NAMP_PREFACTOR, fq=FQ) -> Model:
    return Model(_params(model_id=_lib.MODEL_SPIN18, fq=fq, namp=namp), "spin18")


def Spin15Model(DA=False, TO=False, fq=FQ) -> Model:
    return Model(_params(model_id=_lib.MODEL_SPIN15, fq=fq, decay_aware=int(DA), time_optimal=int(TO)), "spin15")


def Spin11Model(DO=0, fq=FQ) -> Model:
    return Model(_params(model_id=_lib.MODEL_SPIN11, fq=fq, derivative_order=int(DO)), "spin11")


def Spin17Model(DO=0, fq=FQ) -> Model:
    return Model(_params(model_id=_lib.MODEL_SPIN17, fq=fq, derivative_order=int(DO)), "spin17")


def Spin30Model(S_diag=(1.0, 1.0, 1.0, 1.0), nominal_idxs=None, fq_cov=FQ * 1e-2, alpha=1.0, sample_state_count=1,
                fq=FQ) -> Model:
    """spin30.jl:60-66.  nominal_idxs: 1-based first index of each sample state's nominal block
    (default STATE1_IDX for every sample state, as run_traj builds it at spin30.jl:216-217)."""
    if nominal_idxs is None:
        nominal_idxs = [1] * sample_state_count
    if len(nominal_idxs) != sample_state_count:
        raise ValueError("one nominal index block per sample state")
    return Model(_params(model_id=_lib.MODEL_SPIN30, fq=fq, fq_cov=fq_cov, alpha=alpha,
                         sample_state_count=int(sample_state_count), S_diag=tuple(S_diag),
                         nominal_offset=tuple(int(i) - 1 for i in nominal_idxs)), "spin30")


def Spin25Model(S_diag=(1.0, 1.0, 1.0, 1.0), nominal_idxs=None, namp=NAMP_PREFACTOR, wnamp_cov=1.0, alpha=1.0,
                sample_state_count=1, fq=FQ) -> Model:
    """spin25.jl:67-74: unscented sampling of the amplitude noise; the noise amplitude is namp times the output of the
    pink filter driven by sqrt(wnamp_cov), whose taps are part of the state.  The only model that reads its time argument."""
    if nominal_idxs is None:
        nominal_idxs = [1] * sample_state_count
    if len(nominal_idxs) != sample_state_count:
        raise ValueError("one nominal index block per sample state")
    return Model(_params(model_id=_lib.MODEL_SPIN25, fq=fq, namp=namp, fq_cov=wnamp_cov, alpha=alpha,
                         sample_state_count=int(sample_state_count), S_diag=tuple(S_diag),
                         nominal_offset=tuple(int(i) - 1 for i in nominal_idxs)), "spin25")


def state_dim(model: Model) -> int:
    n = Engine.state_dim(model.params)
    if n < 0:
        raise ValueError("inconsistent model parameters")
    return n


def control_dim(model: Model) -> int:
    m = Engine.control_dim(model.params)
    if m < 0:
        raise ValueError("inconsistent model parameters")
    return m


@dataclass
class KnotPoint:
    """RobotDynamics.AbstractKnotPoint: z = [x; u], t, dt."""

    x: np.ndarray
    u: np.ndarray
    dt: float
    t: float = 0.0


_default_engine = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def discrete_dynamics(_rk, model: Model, x, u=None, t=0.0, dt=None, engine: Engine | None = None):
    """RD.discrete_dynamics(::Type{RK3}, model, x, u, t, dt) or (…, model, z::KnotPoint)."""
    eng = engine or default_engine()
    if isinstance(x, KnotPoint):
        z = x
        return eng.discrete_dynamics(model.params, z.x, z.u, z.t, z.dt)
    return eng.discrete_dynamics(model.params, x, u, t, dt)


def discrete_jacobian(_rk, model: Model, z: KnotPoint, engine: Engine | None = None):
    """RD.discrete_jacobian!(::Type{RK3}, ∇f, model, z): returns ∇f = [A B], n x (n+m)."""
    eng = engine or default_engine()
    return eng.discrete_jacobian(model.params, z.x, z.u, z.t, z.dt)


def rollout(model: Model, x0, U, dt, engine: Engine | None = None):
    """TO.rollout!: X[0]=x0, X[k+1] = discrete_dynamics(RK3, model, X[k], U[k], t, dt[k])."""
    U = np.asarray(U, dtype=np.float64)
    dts = np.full(U.shape[-2], float(dt)) if np.isscalar(dt) else dt
    return (engine or default_engine()).rollout(model.params, x0, U, dts)


def knot_times(dts):
    """z.t of a TO.Traj: t_1 = 0, t_{k+1} = t_k + dt_k (accumulated, as the engine's rollouts do)."""
    dts = np.asarray(dts, dtype=np.float64)
    t = np.zeros(dts.shape[0])
    acc = 0.0
    for k in range(dts.shape[0]):
        t[k] = acc
        acc = acc + dts[k]
    return t


def dynamics_expansion(model: Model, X, U, dt, engine: Engine | None = None):
    """TO.dynamics_expansion!: ∇f_k for k = 1..N-1 in one call.  X (N,n) or (N-1,n); U (N-1,m)."""
    X, U = np.asarray(X, dtype=np.float64), np.asarray(U, dtype=np.float64)
    K = U.shape[0]
    dts = np.full(K, float(dt)) if np.isscalar(dt) else np.asarray(dt, dtype=np.float64)
    return (engine or default_engine()).jacobians(model.params, X[:K], U, dts, t=knot_times(dts))


# initial augmented states the reference's run_traj builds
def spin13_x0():
    return np.concatenate([IS1_ISO, IS2_ISO, np.zeros(3)])                       # spin13.jl:78-84


def spin12_x0():
    return np.concatenate([IS1_ISO, IS2_ISO, np.zeros(3)] + [IS1_ISO, IS2_ISO, IS3_ISO, IS4_ISO] * 2)   # spin12.jl:213-218


def spin30_x0(state_cov=1e-2, rng=None):
    """spin30.jl:222-239: nominal IS1..IS3 + zeros(3) + 10 sigma points around IS1 + penalty slot."""
    rng = rng or np.random.default_rng(0)
    pts = []
    for _ in range(10):
        s = IS1_ISO + math.sqrt(state_cov) * rng.standard_normal(4)
        pts.append(s / np.linalg.norm(s))
    return np.concatenate([IS1_ISO, IS2_ISO, IS3_ISO, np.zeros(3)] + pts + [np.zeros(1)])


def spin25_x0(state_cov=1e-2, rng=None):
    """spin25.jl run_traj: nominal IS1, IS2 + zeros(3) + zero filter taps + 10 sigma points around IS1 + penalty slot."""
    rng = rng or np.random.default_rng(0)
    pts = []
    for _ in range(10):
        s = IS1_ISO + math.sqrt(state_cov) * rng.standard_normal(4)
        pts.append(s / np.linalg.norm(s))
    return np.concatenate([IS1_ISO, IS2_ISO, np.zeros(3), np.zeros(6)] + pts + [np.zeros(1)])


# ---- analytic pulses (spin14.py:60-241), sampled at DT = 1/dt_inv ------------------------------------------
T_XZ_YPIBY2 = 2.1656249366575766
T_Z_YPIBY2 = 15.142330599557274
TTOT_YPIBY2 = 2 * T_XZ_YPIBY2 + T_Z_YPIBY2
TTOT_XPIBY2 = 2 * TTOT_YPIBY2 + TTOT_ZPIBY2


def gen_controls_ypiby2(t):
    if t <= T_XZ_YPIBY2:
        return AYPIBY2
    if t <= T_XZ_YPIBY2 + T_Z_YPIBY2:
        return 0.0
    return -AYPIBY2


def gen_controls_xpiby2(t):
    if t <= TTOT_YPIBY2:
        return -gen_controls_ypiby2(t)
    if t <= TTOT_YPIBY2 + TTOT_ZPIBY2:
        return 0.0
    return gen_controls_ypiby2(t - (TTOT_YPIBY2 + TTOT_ZPIBY2))


def _sampled(gen, evolution_time, dt_inv):
    count = int(math.floor(evolution_time * dt_inv))
    ts = np.arange(count) / dt_inv
    c = np.array([gen(t) for t in ts])
    c[0] = 0.0
    c[-1] = 0.0                                                                   # spin14.py:121-122
    return c


def controls_ypiby2(dt_inv=100.0):
    return _sampled(gen_controls_ypiby2, TTOT_YPIBY2, dt_inv)


def controls_xpiby2(dt_inv=100.0):
    return _sampled(gen_controls_xpiby2, TTOT_XPIBY2, dt_inv)


def controls_zpiby2(dt_inv=100.0):
    return np.zeros(int(math.floor(TTOT_ZPIBY2 * dt_inv)))


# the analytic gates as (amplitude, duration) segments with the reference's exact durations (spin.jl:557, 603-612, 691-697)
TX_YPIBY2 = 2.1656249366575766
TZ_YPIBY2 = 15.1423305995572655
TTOT_YPIBY2_JL = 19.4735804728724204                 # spin.jl:605 (spin14.py's sum above differs in the last ulp)
TTOT_XPIBY2_JL = 4 * TX_YPIBY2 + 2 * TZ_YPIBY2 + TTOT_ZPIBY2   # spin.jl:690
ANALYTIC_SEGMENTS = {
    GateType.zpiby2: ([0.0], [TTOT_ZPIBY2]),
    GateType.ypiby2: ([AYPIBY2, 0.0, -AYPIBY2], [TX_YPIBY2, TZ_YPIBY2, TX_YPIBY2]),
    # X/2 = (-Y/2) Z/2 (Y/2) in time order: -Y/2 first (spin14.py:150-158, spin.jl:691-697)
    GateType.xpiby2: ([-AYPIBY2, 0.0, AYPIBY2, 0.0, AYPIBY2, 0.0, -AYPIBY2],
                      [TX_YPIBY2, TZ_YPIBY2, TX_YPIBY2, TTOT_ZPIBY2, TX_YPIBY2, TZ_YPIBY2, TX_YPIBY2]),
}
ANALYTIC_DYNAMICS = {   # dynamics type -> (gate whose segments it integrates, density?)   (DT_GTM / DT_ST, spin.jl:76-91, 1000-1013)
    DynamicsType.zpiby2nodis: (GateType.zpiby2, False), DynamicsType.zpiby2t1: (GateType.zpiby2, True),
    DynamicsType.ypiby2nodis: (GateType.ypiby2, False), DynamicsType.ypiby2t1: (GateType.ypiby2, True),
    DynamicsType.xpiby2nodis: (GateType.xpiby2, False), DynamicsType.xpiby2t1: (GateType.xpiby2, True),
}


def xpiby2da_segments():
    """The segment list integrate_prop_xpiby2da! walks (spin.jl:819-905): the analytic X/2 pulse cut at every 0.1 ns noise
    bin (noise_dt_inv = 10 assumed there) and at the six amplitude switches.  Returns (amps, durs, noise_idx) with 0-based
    noise indices; durations are formed with the reference's own expressions (`T1_XPIBY2 - 2.1`, `2.2 - T1_XPIBY2`, ...)."""
    t = [TX_YPIBY2]                                   # T1..T6 by the reference's chain of additions (spin.jl:691-697)
    for d in (TZ_YPIBY2, TX_YPIBY2, TTOT_ZPIBY2, TX_YPIBY2, TZ_YPIBY2):
        t.append(t[-1] + d)
    level = [-AYPIBY2, 0.0, AYPIBY2, 0.0, AYPIBY2, 0.0, -AYPIBY2]
    amps, durs, nidx = [], [], []
    j = 1                                             # 1-based noise knot, as in the reference
    for seg in range(7):
        if seg < 6:
            jb = int(math.floor(t[seg] * 10.0)) + 1   # the bin that contains the switch: 22, 174, 195, 374, 395, 547
            while j < jb:
                amps.append(level[seg]); durs.append(1e-1); nidx.append(j - 1); j += 1
            lo, hi = (jb - 1) / 10.0, jb / 10.0         # the literals 2.1 / 2.2, 17.3 / 17.4, ... of the reference
            amps.append(level[seg]); durs.append(t[seg] - lo); nidx.append(jb - 1)
            amps.append(level[seg + 1]); durs.append(hi - t[seg]); nidx.append(jb - 1)
            j = jb + 1
        else:
            while j <= 568:
                amps.append(level[seg]); durs.append(1e-1); nidx.append(j - 1); j += 1
            amps.append(level[seg]); durs.append(TTOT_XPIBY2_JL - 56.8); nidx.append(568)
    return np.array(amps), np.array(durs), np.array(nidx, dtype=np.int32)


def pulse_sin(N):
    """Synthetic benchmark pulse P_sin(N) of SURVEY §8(d): a_k on N knots, a_0 = a_{N-1} = 0."""
    k = np.arange(N)
    return 0.30 * np.sin(2 * np.pi * k / (N - 1)) + 0.10 * np.sin(6 * np.pi * k / (N - 1))


# ---- sample generation (spin.jl:1119-1166) ---------------------------------------------------------------------
_FIXED_STATES = {-1: [1, 0], -2: [0, 1], -3: [_R2, 1j * _R2], -4: [_R2, -_R2], -5: [_R2, _R2]}


def gen_rand_state_iso(seed=0, engine: Engine | None = None):
    """Seeds -1..-5 are the reference's fixed states.  For seed >= 0 the reference seeds Julia's
    MersenneTwister, which no other runtime reproduces; here seed k is sample index k of the engine's
    Philox stream (key 0), drawn with the same recipe `rand(2) + im*rand(2)`, normalised."""
    if seed in _FIXED_STATES:
        return get_vec_iso(_FIXED_STATES[seed])
    if seed < 0:
        raise ValueError("unknown fixed state seed")
    _, _, psi0 = (engine or default_engine()).philox_samples(seed, 1, 0, FQ, 0.0, 0.0, 0.0)
    return psi0[0]


def gen_rand_density(seed=0, engine: Engine | None = None):
    psi = get_vec_uniso(gen_rand_state_iso(seed, engine))
    return np.outer(psi, psi.conj())


# ---- cost-side primitives (spin.jl:1040-1092), batched ---------------------------------------------------------------
def gate_error_iso2(s1, s2, engine: Engine | None = None):
    return (engine or default_engine()).gate_error_expansion(s1, s2, want_grad=False, want_hess=False)[0]


def jacobian_gate_error_iso2(s1, s2, engine: Engine | None = None):
    return (engine or default_engine()).gate_error_expansion(s1, s2, want_hess=False)[1]


def hessian_gate_error_iso2(s2, engine: Engine | None = None):
    s2 = np.asarray(s2, dtype=np.float64)
    s1 = np.zeros((1, 4)) if s2.ndim == 1 else np.zeros_like(s2)
    return (engine or default_engine()).gate_error_expansion(s1, s2, want_grad=False)[2]


# ---- simulation entry points (spin.jl:1553-1672), batched over samples -------------------------------------------
@dataclass
class SimResult:
    fidelities: np.ndarray
    stats: object = None
    extra: dict = field(default_factory=dict)


def run_sim_prop(gate_count, gate_type, controls, dt_inv, fq=FQ, da=None, namp=NAMP_PREFACTOR, noise_dt_inv=DT_NOISE_INV,
                 seed=0, state_seed=None, dynamics_type=DynamicsType.schroed, noise=None, engine: Engine | None = None,
                 algo=_lib.ALGO_SU2, gate_time=None, t2_gamma_c=GAMMAC, t2_gamma_f=GAMMAF_PREFACTOR):
    """run_sim_prop for one (seed, state_seed) like the reference, or for arrays of them (batched).

    `controls`/`dt_inv`/`gate_time` replace `save_file_path` (what grab_controls returns, rbqoc.jl:98-123; gate_time
    defaults to len(controls)/dt_inv and only sets the length of the noise row, spin.jl:1567-1570);
    `fq` replaces `negi_h0 = fq * NEGI_H0_ISO`.  lindbladt2 uses the reference's dephasing rate
    Gamma(t) = GAMMAC + 2 GAMMAF_PREFACTOR^2 t (spin.jl:549-552) unless t2_gamma_c / t2_gamma_f say otherwise.
    Returns a dict with "fidelities" of shape (gate_count+1,) for scalar seeds or (S, gate_count+1) for array seeds."""
    eng = engine or default_engine()
    gate_type = GateType(gate_type)
    dynamics_type = DynamicsType(dynamics_type)
    scalar = np.isscalar(seed)
    seeds = np.atleast_1d(np.asarray(seed, dtype=np.int64))
    sseeds = seeds if state_seed is None else np.broadcast_to(np.atleast_1d(np.asarray(state_seed, dtype=np.int64)), seeds.shape)
    S = seeds.shape[0]
    controls = None if controls is None else np.ascontiguousarray(controls, dtype=np.float64).reshape(-1)
    dt = 1.0 / dt_inv if dt_inv else 0.0
    fqs = np.broadcast_to(np.asarray(fq, dtype=np.float64), (S,)).copy()
    psi0 = np.stack([gen_rand_state_iso(int(s), eng) for s in sseeds])
    if dynamics_type == DynamicsType.schroed:
        das = None if da is None else np.broadcast_to(np.asarray(da, dtype=np.float64), (S,)).copy()
        fid, st = eng.mc_static(controls, dt, gate_type, gate_count, fqs, das, psi0, algo=algo)
    elif dynamics_type == DynamicsType.schroedda:
        gt = controls.shape[0] * dt if gate_time is None else float(gate_time)
        nlen = int(math.ceil(gt * gate_count * noise_dt_inv)) + 1                         # spin.jl:1567-1570
        if noise is None:
            noise = np.concatenate([eng.pink_noise(int(s), 1, 0, namp, noise_dt_inv, nlen) for s in seeds])
        if not np.all(fqs == fqs[0]):
            raise ValueError("schroedda sweeps share one fq")
        fid, st = eng.mc_noise(controls, dt, gate_type, gate_count, float(fqs[0]), noise, noise_dt_inv, psi0, algo=algo)
    elif dynamics_type in (DynamicsType.lindbladt1, DynamicsType.lindbladt2):
        rho0 = np.stack([np.outer(get_vec_uniso(p), get_vec_uniso(p).conj()) for p in psi0])
        das = None if da is None else np.broadcast_to(np.asarray(da, dtype=np.float64), (S,)).copy()
        ch = _lib.LINDBLAD_T1 if dynamics_type == DynamicsType.lindbladt1 else _lib.LINDBLAD_T2
        t2 = ch == _lib.LINDBLAD_T2
        fid, rho, st = eng.lindblad(controls, dt, gate_type, gate_count, float(fqs[0]), das, rho0, channels=ch,
                                    t2_gamma_c=t2_gamma_c if t2 else 0.0, t2_gamma_f=t2_gamma_f if t2 else 0.0)
    elif dynamics_type in ANALYTIC_DYNAMICS:
        # analytic gates (DT_AN, spin.jl:94-114): `controls`/`dt_inv` are ignored like the reference ignores save_file_path.
        # xpiby2nodis is an empty loop in the reference (spin.jl:721-727); here it integrates the X/2 segments for real.
        g, density = ANALYTIC_DYNAMICS[dynamics_type]
        amps, durs = ANALYTIC_SEGMENTS[g]
        das = None if da is None else np.broadcast_to(np.asarray(da, dtype=np.float64), (S,)).copy()
        if density:
            rho0 = np.stack([np.outer(get_vec_uniso(p), get_vec_uniso(p).conj()) for p in psi0])
            fid, rho, st = eng.lindblad_segments(amps, durs, gate_type, gate_count, float(fqs[0]), das, rho0)
        else:
            fid, st = eng.mc_static_segments(amps, durs, gate_type, gate_count, fqs, das, psi0)
    elif dynamics_type == DynamicsType.xpiby2da:
        amps, durs, nidx = xpiby2da_segments()
        nlen = int(math.ceil(TTOT_XPIBY2_JL * gate_count * noise_dt_inv)) + 1             # spin.jl:1569
        if noise is None:
            noise = np.concatenate([eng.pink_noise(int(s), 1, 0, namp, noise_dt_inv, max(nlen, 569)) for s in seeds])
        fid, st = eng.mc_noise_segments(amps, durs, nidx, gate_type, gate_count, float(fqs[0]), noise, psi0)
    else:
        raise ValueError("unsupported dynamics type")
    return {
        "gate_count": gate_count, "gate_type": int(gate_type), "seed": seed, "state_seed": state_seed,
        "fidelities": fid[0] if scalar else fid, "stats": st, "namp": namp, "noise_dt_inv": noise_dt_inv,
    }


def deqjl_timeline(controls, controls_dt_inv, gate_time, gate_count, noise_dt_inv=None, save_times=None):
    """Segment list of the ODE run_sim_deqjl integrates (spin.jl:424-467, 1277-1380).  Its right-hand side reads
    controls[floor(t controls_dt_inv) mod control_knot_count] (the pulse repeats with period knot_count / dt_inv, which is NOT
    the gate time when gate_time * dt_inv is not an integer -- the reference's own wrap-around, kept) and, for schroedda,
    noise_offsets[floor(t noise_dt_inv)]; states are saved at multiples of gate_time (or at `save_times`, as
    run_sim_fine_deqjl's save_step grid).  Between the breakpoints of those three grids the right-hand side is constant.
    Returns (amps, durs, noise_idx or None, save_after)."""
    controls = np.asarray(controls, dtype=np.float64).reshape(-1)
    count = int(math.floor(gate_time * controls_dt_inv))                  # control_knot_count, spin.jl:1299
    if count < 1 or count > controls.shape[0]:
        raise ValueError("gate_time * controls_dt_inv must select between 1 and len(controls) knots")
    t_end = gate_time * gate_count
    saves = gate_time * np.arange(1, gate_count + 1) if save_times is None else np.asarray(save_times, dtype=np.float64)
    if saves.size and (np.any(np.diff(saves) <= 0) or saves[0] <= 0 or saves[-1] > t_end * (1 + 1e-15)):
        raise ValueError("save times must be increasing inside (0, evolution time]")
    pts = [np.arange(0, int(math.floor(t_end * controls_dt_inv)) + 1) / controls_dt_inv, saves]
    if noise_dt_inv:
        pts.append(np.arange(0, int(math.floor(t_end * noise_dt_inv)) + 1) / noise_dt_inv)
    t = np.unique(np.concatenate(pts))
    t = t[(t >= 0.0) & (t <= t_end)]
    if t[-1] < t_end:
        t = np.append(t, t_end)
    durs = np.diff(t)
    keep = durs > 0.0
    mid = 0.5 * (t[:-1] + t[1:])[keep]                                    # the index of a segment is that of its midpoint
    durs = durs[keep]
    amps = controls[np.floor(mid * controls_dt_inv).astype(np.int64) % count]
    nidx = np.floor(mid * noise_dt_inv).astype(np.int32) if noise_dt_inv else None
    ends = np.cumsum(durs)
    save_after = np.searchsorted(ends, saves - 1e-12 * max(t_end, 1.0), side="left") + 1
    return amps, durs, nidx, np.minimum(save_after, durs.shape[0]).astype(np.int64)


def gen_rand_density_iso(seed=0, engine: Engine | None = None):
    """spin.jl:1141-1158 as the complex 2x2 density (the reference stores its 4x4 real iso image): seed 0 is the fixed state
    (|0> + |1>)/sqrt(2), any other seed a random pure state (Philox here, see gen_rand_state_iso)."""
    if seed == 0:
        psi = np.array([_R2, _R2], dtype=np.complex128)
    else:
        psi = get_vec_uniso(gen_rand_state_iso(seed, engine))
    return np.outer(psi, psi.conj())


_DENSITY_CHANNELS = {DynamicsType.lindbladnodis: 0, DynamicsType.lindbladt1: _lib.LINDBLAD_T1, DynamicsType.lindbladt2: _lib.LINDBLAD_T2}


def _deqjl_run(gate_count, gate_type, controls, controls_dt_inv, gate_time, fq, da, namp, noise_dt_inv, seed, state_seed,
               dynamics_type, noise, engine, save_times, want_states, t2_gamma_c, t2_gamma_f):
    eng = engine or default_engine()
    gate_type = GateType(gate_type)
    dynamics_type = DynamicsType(dynamics_type)
    seeds = np.atleast_1d(np.asarray(seed, dtype=np.int64))
    sseeds = seeds if state_seed is None else np.broadcast_to(np.atleast_1d(np.asarray(state_seed, dtype=np.int64)), seeds.shape)
    S = seeds.shape[0]
    fqs = np.broadcast_to(np.asarray(fq, dtype=np.float64), (S,)).copy()
    das = None if da is None else np.broadcast_to(np.asarray(da, dtype=np.float64), (S,)).copy()
    noisy = dynamics_type == DynamicsType.schroedda
    amps, durs, nidx, save_after = deqjl_timeline(controls, controls_dt_inv, gate_time, gate_count, noise_dt_inv if noisy else None,
                                                  save_times)
    if dynamics_type in _DENSITY_CHANNELS:
        if not np.all(fqs == fqs[0]):
            raise ValueError("the density dynamics use one fq (FQ_NEGI_H0_ISO, spin.jl:495-512)")
        rho0 = np.stack([gen_rand_density_iso(int(s), eng) for s in sseeds])
        ch = _DENSITY_CHANNELS[dynamics_type]
        t2 = ch == _lib.LINDBLAD_T2
        fid, rho, st = eng.lindblad_timeline(amps, durs, save_after, gate_type, float(fqs[0]), rho0, da=das, channels=ch,
                                             t2_gamma_c=t2_gamma_c if t2 else 0.0, t2_gamma_f=t2_gamma_f if t2 else 0.0)
        return fid, st, rho
    if dynamics_type not in (DynamicsType.schroed, DynamicsType.schroedda):
        raise ValueError("run_sim_deqjl mirrors schroed, schroedda, lindbladnodis, lindbladt1 and lindbladt2")
    psi0 = np.stack([gen_rand_state_iso(int(s), eng) for s in sseeds])
    if noisy and noise is None:
        nlen = int(math.ceil(gate_time * gate_count * noise_dt_inv)) + 1                  # spin.jl:1305
        noise = np.concatenate([eng.pink_noise(int(s), 1, 0, namp, noise_dt_inv, nlen) for s in seeds])
    out = eng.mc_timeline(amps, durs, save_after, gate_type, fqs, psi0, da=das, noise=noise if noisy else None,
                          noise_idx=nidx if noisy else None, want_states=want_states)
    return (out[0], out[1], out[2]) if want_states else (out[0], out[1], None)


def run_sim_deqjl(gate_count, gate_type, controls, controls_dt_inv, gate_time, fq=FQ, da=None, namp=NAMP_PREFACTOR,
                  noise_dt_inv=DT_NOISE_INV, seed=0, state_seed=None, dynamics_type=DynamicsType.schroed, noise=None,
                  engine: Engine | None = None, t2_gamma_c=GAMMAC, t2_gamma_f=GAMMAF_PREFACTOR):
    """run_sim_deqjl (spin.jl:1277-1380), batched over seeds, for dynamics_type schroed / schroedda (iso states) and
    lindbladnodis / lindbladt1 / lindbladt2 (densities, gen_rand_density_iso: seed 0 is the fixed |+> state).  The reference
    lets Vern9 integrate a right-hand side that is piecewise constant in the controls to reltol = abstol = 1e-12; this evaluates
    the exact solution of the same ODE (the ordered product of segment exponentials, `deqjl_timeline`; the running dephasing
    rate of lindbladt2 by a fourth-order Magnus step per segment), so the two agree to the solver's tolerance.  `gate_time`
    replaces the evolution time grab_controls reads from the pulse file.  Density runs also return "densities"."""
    scalar = np.isscalar(seed)
    fid, st, extra = _deqjl_run(gate_count, gate_type, controls, controls_dt_inv, gate_time, fq, da, namp, noise_dt_inv, seed,
                                state_seed, dynamics_type, noise, engine, None, False, t2_gamma_c, t2_gamma_f)
    res = {"gate_count": gate_count, "gate_type": int(GateType(gate_type)), "gate_time": gate_time, "seed": seed,
           "fidelities": fid[0] if scalar else fid, "stats": st, "dt": 1.0 / controls_dt_inv, "namp": namp,
           "noise_dt_inv": noise_dt_inv, "dynamics_type": int(DynamicsType(dynamics_type))}
    if extra is not None:
        res["densities"] = extra[0] if scalar else extra
    return res


def run_sim_fine_deqjl(controls, controls_dt_inv, gate_time, gate_count=1, gate_type=GateType.xpiby2, save_step=1e-1, fq=FQ,
                       da=None, namp=NAMP_PREFACTOR, noise_dt_inv=DT_NOISE_INV, seed=0, dynamics_type=DynamicsType.schroed,
                       noise=None, engine: Engine | None = None, t2_gamma_c=GAMMAC, t2_gamma_f=GAMMAF_PREFACTOR):
    """run_sim_fine_deqjl (spin.jl:1389-1478; the Fig. 4b / 3d driver, figures.jl:638-646, 760-764): the states themselves on
    the save_step grid, t = 0, save_step, 2 save_step, ... <= gate_time * gate_count.  Returns {"states": (n_save, 4) iso states
    or (n_save, 2, 2) complex densities (the reference stores their 4x4 iso image: get_mat_iso), "save_times", ...}; batched
    over seeds like run_sim_deqjl (leading axis S for array seeds)."""
    scalar = np.isscalar(seed)
    t_end = gate_time * gate_count
    count = int(math.floor(t_end / save_step * (1 + 1e-15)))
    save_times = save_step * np.arange(1, count + 1)
    save_times = save_times[save_times <= t_end * (1 + 1e-15)]
    fid, st, states = _deqjl_run(gate_count, gate_type, controls, controls_dt_inv, gate_time, fq, da, namp, noise_dt_inv, seed, None,
                                 dynamics_type, noise, engine, save_times, True, t2_gamma_c, t2_gamma_f)
    return {"gate_count": gate_count, "gate_time": gate_time, "seed": seed, "states": states[0] if scalar else states,
            "save_times": np.concatenate([[0.0], save_times]), "dt": 1.0 / controls_dt_inv, "namp": namp,
            "noise_dt_inv": noise_dt_inv, "dynamics_type": int(DynamicsType(dynamics_type))}


def evaluate_fqdev(controls, dt_inv, fq_cov=FQ * 1e-2, trial_count=1000, gate_type=GateType.zpiby2,
                   engine: Engine | None = None):
    """spin.jl:1611-1656 in two batched calls instead of 2*(trial_count+4) serial ones."""
    seeds = np.repeat(np.arange(trial_count), 2)
    fqs = np.tile([FQ + fq_cov, FQ - fq_cov], trial_count)
    res = run_sim_prop(1, gate_type, controls, dt_inv, fq=fqs, seed=seeds, engine=engine)
    bseeds = np.repeat(np.arange(-4, 0), 2)
    bres = run_sim_prop(1, gate_type, controls, dt_inv, fq=np.tile([FQ + fq_cov, FQ - fq_cov], 4), seed=bseeds, engine=engine)
    return {"gate_errors": 1 - res["fidelities"][:, -1], "gate_errors_basis": 1 - bres["fidelities"][:, -1]}


def evaluate_adev(controls, dt_inv, gate_count=100, gate_type=GateType.xpiby2, trial_count=100, engine: Engine | None = None):
    """spin.jl:1659-1672."""
    res = run_sim_prop(gate_count, gate_type, controls, dt_inv, seed=np.arange(1, trial_count + 1),
                       dynamics_type=DynamicsType.schroedda, engine=engine)
    return {"gate_errors": 1 - res["fidelities"][:, -1]}


# ---- constraint table of the AL cost expansion (spin13.jl:95-153 and siblings) ---------------------------------------
class ConstraintTable:
    """Host-side builder of the rbq_constraint table: the reference's ConstraintList (spin13.jl:145-152) restricted to
    the three constraint types it uses.  Knot ranges are 0-based and half-open here ([k0, k1)); the reference's `2:N-2`
    is `bound(1, N - 2, ...)`.  `specs` keeps the plain description (what tests hand to the oracle)."""

    def __init__(self, n: int, m: int, N: int):
        self.n, self.m, self.N = n, m, N
        self.rows, self.idx, self.val, self.specs = [], [], [], []
        self.dual_size = 0
        self.d_cons = self.d_idx = self.d_val = None

    @property
    def ncon(self) -> int:
        return len(self.rows)

    def _add(self, kind, p, k0, k1, idx, vals, spec):
        if not (0 <= k0 <= k1 <= self.N):
            raise ValueError("knot range outside the trajectory")
        self.rows.append(_lib.Constraint(kind, p, k0, k1, len(idx), len(self.idx), len(self.val), self.dual_size))
        self.idx += [int(i) for i in idx]
        self.val += [float(v) for v in vals]
        self.dual_size += (k1 - k0) * p
        self.specs.append(dict(spec, k0=k0, k1=k1))
        return self

    def bound(self, k0, k1, x_max=None, x_min=None, u_max=None, u_min=None):
        """BoundConstraint(n, m, x_max=, x_min=, u_max=, u_min=): missing sides are infinite."""
        full = lambda v, size, fill: np.full(size, fill) if v is None else np.asarray(v, dtype=float)
        zmax = np.concatenate([full(x_max, self.n, np.inf), full(u_max, self.m, np.inf)])
        zmin = np.concatenate([full(x_min, self.n, -np.inf), full(u_min, self.m, -np.inf)])
        return self._add(_lib.CON_BOUND, 2 * (self.n + self.m), k0, k1, [], np.concatenate([zmax, zmin]),
                         {"kind": "bound", "zmax": zmax, "zmin": zmin})

    def goal(self, k0, k1, xf, idx):
        """GoalConstraint(xf, idx): x[idx] = xf[idx]."""
        idx = [int(i) for i in idx]
        tgt = np.asarray(xf, dtype=float)[idx]
        return self._add(_lib.CON_GOAL, len(idx), k0, k1, idx, tgt, {"kind": "goal", "idx": idx, "target": tgt})

    def norm(self, k0, k1, idx, val=1.0):
        """NormConstraint(n, m, val, Equality(), idx): sum x[idx]^2 = val."""
        idx = [int(i) for i in idx]
        return self._add(_lib.CON_NORM, 1, k0, k1, idx, [val], {"kind": "norm", "idx": idx, "val": float(val)})

    def dual_slices(self):
        """(offset, knots, p) of every constraint inside the flat lambda / mu / c arrays."""
        return [(r.dual_off, r.k1 - r.k0, r.p) for r in self.rows]

    def upload(self, device):
        import ctypes as C

        import torch

        raw = bytes((_lib.Constraint * max(len(self.rows), 1))(*self.rows))
        self.d_cons = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        self.d_idx = torch.tensor(self.idx or [0], dtype=torch.int32, device=device)
        self.d_val = torch.tensor(self.val or [0.0], dtype=torch.float64, device=device)
        assert C.sizeof(_lib.Constraint) == 48
        return self


def _reference_constraints(n, N, xf, ctrl, goal_idx, norm_blocks, max_control):
    """The ConstraintList every run_traj of the reference builds (spin13.jl:136-152, spin12.jl:280-296, spin30.jl:301-320):
    amplitude bound on the control state over 2:N-2, zero control at N-1, goal at N, unit norm of the iso states over 2:N-1."""
    x_max = np.full(n, np.inf); x_min = np.full(n, -np.inf)
    x_max[ctrl], x_min[ctrl] = max_control, -max_control
    xb_max = np.full(n, np.inf); xb_min = np.full(n, -np.inf)
    xb_max[ctrl] = xb_min[ctrl] = 0.0
    t = ConstraintTable(n, 1, N)
    t.bound(1, N - 2, x_max=x_max, x_min=x_min)              # 2:N-2
    t.bound(N - 2, N - 1, x_max=xb_max, x_min=xb_min)        # N-1:N-1
    t.goal(N - 1, N, xf, goal_idx)                           # N:N
    for b in norm_blocks:
        t.norm(1, N - 1, range(b, b + 4))                    # 2:N-1
    return t


def spin13_constraints(N, xf, max_control=MAX_CONTROL_NORM_0):
    """spin13.jl:136-152 (states 0-3, 4-7; integral of the control 8; control 9; its derivative 10)."""
    return _reference_constraints(11, N, xf, 9, list(range(0, 9)), [0, 4], max_control)


def spin12_constraints(N, xf, max_control=MAX_CONTROL_NORM_0):
    """spin12.jl:280-296 (spin18 alike): spin13's list plus unit norm of the 8 sample states at 11 + 4i."""
    return _reference_constraints(43, N, xf, 9, list(range(0, 9)), [0, 4] + [11 + 4 * i for i in range(8)], max_control)


def spin30_constraints(N, xf, sample_state_count=1, max_control=MAX_CONTROL_NORM_0):
    """spin30.jl:301-320: three nominal states 0-11, integral 12, control 13, derivative 14, then chunks of 41 (10 sigma
    points and a penalty); goal on STATE1, STATE2 and the integral; unit norm of every sigma point and nominal state."""
    n = 15 + 41 * sample_state_count
    blocks = [15 + 41 * c + 4 * i for c in range(sample_state_count) for i in range(10)] + [0, 4, 8]
    return _reference_constraints(n, N, xf, 13, list(range(0, 8)) + [12], blocks, max_control)


# ---- pulse I/O on the caller's side of the path (rbqoc.jl:98-139, spin.jl:1485-1530) ---------------------------------
DT_PREF, DT_PREF_INV = 1e-2, 1e2                     # spin.jl:24-25


class SaveType(enum.IntEnum):                        # rbqoc.jl:27-31
    jl = 1
    samplejl = 2
    py = 3


def read_save(save):
    """rbqoc.jl:129-139.  `save` is either the mapping read_save would return (key -> array / scalar, arrays with the shapes
    Julia sees, index lists 1-based as stored) or a path to the .h5 file, read by `h5lite` (the HDF5 subset those files use;
    no libhdf5 / h5py in this image).  Arrays are transposed on the way in because HDF5.jl stores Julia's column-major
    arrays with reversed dimensions."""
    if isinstance(save, (str, bytes)) or hasattr(save, "__fspath__"):
        from . import h5lite
        f = h5lite.File(save)
        return {k: (v.T if isinstance(v, np.ndarray) and v.ndim > 1 else v) for k, v in f.items()}
    return save


def write_save(path, result):
    """The write side of the same schema (`spin13.jl:196-233`: `h5open(path, "cw") do f; for (k, v) in result; write(f, k, v)`):
    `result` holds arrays with the shapes Julia sees; they are stored with reversed dimensions as HDF5.jl does, so that
    `read_save(path)` returns `result` and h5py would show the transposed arrays.  Entries that are None (Julia `nothing`,
    e.g. `benchmark_result`) are skipped."""
    from . import h5lite
    out = {}
    for k, v in result.items():
        if v is None:
            continue
        if isinstance(v, (str, dict)):
            out[k] = v
        else:
            a = np.asarray(v)
            out[k] = np.ascontiguousarray(a.T) if a.ndim > 1 else (a if a.ndim else a[()])
    h5lite.write(path, out)


def _cols(idx):
    return np.atleast_1d(np.asarray(idx, dtype=np.int64)) - 1      # 1-based, as the reference stores them


def grab_controls(save):
    """rbqoc.jl:98-123: (controls, controls_dt_inv, evolution_time) of a saved pulse, for the three save formats."""
    s = read_save(save)
    st = SaveType(int(s["save_type"]))
    if st == SaveType.jl:
        controls = np.asarray(s["astates"], dtype=np.float64)[:-1][:, _cols(s["controls_idx"])]
        dt_inv = 1.0 / float(s["dt"]) if "dt" in s else DT_PREF_INV
        return controls, dt_inv, float(s["evolution_time"])
    if st == SaveType.samplejl:
        return np.asarray(s["controls_sample"], dtype=np.float64)[:-1], DT_PREF_INV, float(s["evolution_time_sample"])
    return np.asarray(s["controls"], dtype=np.float64).T, DT_PREF_INV, float(s["evolution_time"])


def sample_controls(save, dt=DT_PREF, dt_inv=DT_PREF_INV):
    """spin.jl:1485-1530: resample an optimised pulse (controls = the control column of the augmented states, and the second
    derivative that is the solver's actual control) onto the dt grid with interpolating cubic splines.  The reference uses
    Dierckx.Spline1D (k = 3, s = 0), i.e. FITPACK's curfit -- the routine scipy.interpolate.splrep wraps.
    Returns (controls_sample, d2controls_dt2_sample, final_time_sample)."""
    from scipy.interpolate import splev, splrep

    s = read_save(save)
    controls = np.asarray(s["astates"], dtype=np.float64)[:-1][:, _cols(s["controls_idx"])]
    d2 = np.asarray(s["acontrols"], dtype=np.float64)[:, _cols(s["d2controls_dt2_idx"])]
    knots = controls.shape[0]
    if "dt_idx" in s:
        dts = np.asarray(s["acontrols"], dtype=np.float64)[:, int(np.atleast_1d(s["dt_idx"])[0]) - 1]
    else:
        dts = float(s["dt"]) * np.ones(knots)
    time_axis = np.concatenate([[0.0], np.cumsum(dts)[:-1]])
    final_time = float(np.sum(dts))
    count = int(math.floor(final_time * dt_inv))
    t_sample = np.arange(count) * dt                      # the last control is dt before final_time
    out_c = np.empty((count, controls.shape[1])); out_d = np.empty((count, d2.shape[1]))
    for i in range(controls.shape[1]):
        out_c[:, i] = splev(t_sample, splrep(time_axis, controls[:, i], k=3, s=0))
        out_d[:, i] = splev(t_sample, splrep(time_axis, d2[:, i], k=3, s=0))
    return out_c, out_d, final_time
