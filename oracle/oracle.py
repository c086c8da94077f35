"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; rbqoc_b200 never does.

PARITY UNPINNED (see oracle.cpp header): the reference has no tests or golden vectors and its
Julia toolchain is absent, so the oracle is pinned by analytic known answers and an mpmath
referee (oracle/referee.py) rather than by reference outputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

HIST_BINS = 256
MAX_SAMPLE_STATES = 4

GATE_ZPIBY2, GATE_YPIBY2, GATE_XPIBY2, GATE_XPI = 1, 2, 3, 4


class ModelParams(C.Structure):
    """Mirror of `rbq_model_params` (include/rbq.h)."""

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


class Stats(C.Structure):
    _fields_ = [
        ("count", C.c_int64),
        ("nonfinite", C.c_int64),
        ("sum", C.c_double),
        ("sumsq", C.c_double),
        ("min", C.c_double),
        ("max", C.c_double),
        ("hist", C.c_int64 * HIST_BINS),
    ]


_NATIVE_DIR = os.path.join(_HERE, "_native")
_NATIVE_LIB = os.path.join(_NATIVE_DIR, "liboracle_native.so")
NATIVE_FLAGS = ["-O3", "-march=native", "-std=c++17", "-fPIC", "-fopenmp", "-shared"]


def _cpu_tag() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib

                    return hashlib.sha1(line.encode()).hexdigest()[:16]
    except OSError:
        pass
    return "unknown"


def use_native_build() -> str:
    """TIMING ARM ONLY (bench.py cpu_baseline / --impl reference): rebuild the restatement with the flags BASELINE.md §4
    names for the CPU baseline, `-O3 -march=native` (contraction left at the compiler's default), for the host this runs on,
    and route this module's calls to it.  The parity suite keeps liboracle.so (-O2 -ffp-contract=off, Julia's evaluation
    order).  The build lands in oracle/_native/ (git- and gpurun-ignored: it is specific to the CPU it was built on)."""
    global _lib
    tag_path = os.path.join(_NATIVE_DIR, "cpu.tag")
    src = os.path.join(_HERE, "oracle.cpp")
    fresh = (os.path.exists(_NATIVE_LIB) and os.path.exists(tag_path) and open(tag_path).read() == _cpu_tag()
             and os.path.getmtime(_NATIVE_LIB) >= os.path.getmtime(src))
    if not fresh:
        os.makedirs(_NATIVE_DIR, exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx] + NATIVE_FLAGS + ["-o", _NATIVE_LIB, src], check=True)
        with open(tag_path, "w") as f:
            f.write(_cpu_tag())
    _lib = None
    lib(_NATIVE_LIB)
    return _NATIVE_LIB


def set_num_threads(n: int = 0):
    lib().orc_set_num_threads(int(n))


def build(force: bool = False) -> str:
    """Compile liboracle.so with the Makefile next to this file (building the checker is not
    using it)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "oracle.cpp")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib(path: str | None = None):
    global _lib
    if _lib is None:
        if path is None and not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(path or _LIB_PATH)
        _lib.orc_fidelity_vec_iso2.restype = C.c_double
        _lib.orc_fidelity_mat.restype = C.c_double
        _lib.orc_t1_ns.restype = C.c_double
        _lib.orc_t1_reduced_us.restype = C.c_double
        _lib.orc_t1_ns.argtypes = [C.c_double]
        _lib.orc_t1_reduced_us.argtypes = [C.c_double]
        for name in ("orc_pade13_coeffs", "orc_gate_error_expansion", "orc_propagator", "orc_lindblad_generator", "orc_philox4x32_10", "orc_philox_samples",
                     "orc_pink_filter", "orc_pink_noise", "orc_stats", "orc_set_num_threads"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_params(**kw) -> ModelParams:
    p = ModelParams()
    p.fq = 1.4e-2
    p.alpha = 1.0
    for k, v in kw.items():
        if k in ("nominal_offset", "fq_samples", "S_diag"):
            arr = getattr(p, k)
            for i, x in enumerate(v):
                arr[i] = x
        else:
            setattr(p, k, v)
    return p


def state_dim(p) -> int:
    return lib().orc_state_dim(C.byref(p))


def control_dim(p) -> int:
    return lib().orc_control_dim(C.byref(p))


def num_threads() -> int:
    return lib().orc_num_threads()


def expm4(A):
    A = _f64(A)
    E = np.empty((4, 4))
    deg = lib().orc_expm4(_p(A), _p(E))
    return E, deg


def expm4c(A):
    A = np.ascontiguousarray(A, dtype=np.complex128)
    E = np.empty((4, 4), dtype=np.complex128)
    deg = lib().orc_expm4c(_p(A), _p(E))
    return E, deg


def pade13_coeffs():
    out = np.empty(14)
    lib().orc_pade13_coeffs(_p(out))
    return out


def propagator(fq, a, dt):
    E = np.empty((4, 4))
    lib().orc_propagator(C.c_double(fq), C.c_double(a), C.c_double(dt), _p(E))
    return E


def t1_ns(a):
    return lib().orc_t1_ns(float(a))


def t1_reduced_us(a):
    return lib().orc_t1_reduced_us(float(a))


def step(p, x, u, dt, t=0.0):
    x, u = _f64(x), _f64(u)
    xn = np.empty(state_dim(p))
    rc = lib().orc_step(C.byref(p), _p(x), _p(u), C.c_double(t), C.c_double(dt), _p(xn))
    if rc:
        raise ValueError(f"orc_step rc={rc}")
    return xn


def jacobian(p, x, u, dt, t=0.0):
    """Column-major n x (n+m) returned as a numpy (n, n+m) array."""
    x, u = _f64(x), _f64(u)
    n, m = state_dim(p), control_dim(p)
    AB = np.empty((n + m, n))
    rc = lib().orc_jacobian(C.byref(p), _p(x), _p(u), C.c_double(t), C.c_double(dt), _p(AB))
    if rc:
        raise ValueError(f"orc_jacobian rc={rc}")
    return AB.T.copy()


def jacobians(p, X, U, dt, t=None):
    X, U, dt = _f64(X), _f64(U), _f64(dt)
    t = None if t is None else _f64(t)
    n, m = state_dim(p), control_dim(p)
    K = X.shape[0]
    AB = np.empty((K, n + m, n))
    rc = lib().orc_jacobians(C.byref(p), C.c_int64(K), _p(X), _p(U), _p(t), _p(dt), _p(AB))
    if rc:
        raise ValueError(f"orc_jacobians rc={rc}")
    return np.ascontiguousarray(AB.transpose(0, 2, 1))


def rollout(p, x0, U, dt):
    """x0 (B,n) or (n,), U (B,N-1,m) or (N-1,m), dt (N-1,) -> X (B,N,n) or (N,n)."""
    x0, U, dt = _f64(x0), _f64(U), _f64(dt)
    single = x0.ndim == 1
    n, m = state_dim(p), control_dim(p)
    x0b = x0.reshape(-1, n)
    B = x0b.shape[0]
    Ub = U.reshape(B, -1, m)
    N = Ub.shape[1] + 1
    X = np.empty((B, N, n))
    rc = lib().orc_rollout(C.byref(p), C.c_int64(B), C.c_int64(N), _p(x0b), _p(Ub), _p(dt), _p(X))
    if rc:
        raise ValueError(f"orc_rollout rc={rc}")
    return X[0] if single else X


def rollout_closedloop(p, Xref, Uref, K, d, dt, alphas):
    Xref, Uref, K, d, dt, alphas = map(_f64, (Xref, Uref, K, d, dt, alphas))
    n, m = state_dim(p), control_dim(p)
    N = Xref.shape[0]
    na = alphas.shape[0]
    X = np.empty((na, N, n))
    Uo = np.empty((na, N - 1, m))
    rc = lib().orc_rollout_closedloop(C.byref(p), C.c_int64(N), _p(Xref), _p(Uref), _p(K), _p(d), _p(dt),
                                      C.c_int64(na), _p(alphas), _p(X), _p(Uo))
    if rc:
        raise ValueError(f"orc_rollout_closedloop rc={rc}")
    return X, Uo


def fidelity_vec_iso2(s1, s2):
    s1, s2 = _f64(s1), _f64(s2)
    return lib().orc_fidelity_vec_iso2(_p(s1), _p(s2))


def fidelity_mat(m1, m2, closed=False):
    m1 = np.ascontiguousarray(m1, dtype=np.complex128)
    m2 = np.ascontiguousarray(m2, dtype=np.complex128)
    return lib().orc_fidelity_mat(_p(m1), _p(m2), int(closed))


def gate_error_expansion(s1, s2):
    """One state/target pair: (err, grad[4], hess[4,4]) per spin.jl:1040-1092."""
    s1, s2 = _f64(s1), _f64(s2)
    err = C.c_double()
    grad, hess = np.empty(4), np.empty((4, 4))
    lib().orc_gate_error_expansion(_p(s1), _p(s2), C.byref(err), _p(grad), _p(hess))
    return err.value, grad, hess


def gate_iso(gate):
    G = np.empty((4, 4))
    rc = lib().orc_gate_iso(int(gate), _p(G))
    if rc:
        raise ValueError("unknown gate")
    return G


def mc_static(controls, dt, gate, gate_count, fq, da, psi0, nthreads=0, want_states=False):
    controls, fq, psi0 = _f64(controls), _f64(fq), _f64(psi0)
    da = None if da is None else _f64(da)
    S = fq.shape[0]
    fid = np.empty((S, gate_count + 1))
    states = np.empty((S, 4)) if want_states else None
    rc = lib().orc_mc_static(_p(controls), C.c_int64(controls.shape[0]), C.c_double(dt), int(gate),
                             C.c_int64(gate_count), C.c_int64(S), _p(fq), _p(da), _p(psi0), _p(fid), _p(states),
                             int(nthreads))
    if rc:
        raise ValueError(f"orc_mc_static rc={rc}")
    return (fid, states) if want_states else fid


def mc_noise(controls, dt, gate, gate_count, fq, noise, noise_dt_inv, psi0, nthreads=0):
    controls, noise, psi0 = _f64(controls), _f64(noise), _f64(psi0)
    S, nlen = noise.shape
    fid = np.empty((S, gate_count + 1))
    rc = lib().orc_mc_noise(_p(controls), C.c_int64(controls.shape[0]), C.c_double(dt), int(gate),
                            C.c_int64(gate_count), C.c_int64(S), C.c_double(fq), _p(noise), C.c_int64(nlen),
                            C.c_double(noise_dt_inv), _p(psi0), _p(fid), None, int(nthreads))
    if rc:
        raise ValueError(f"orc_mc_noise rc={rc}")
    return fid


def lindblad_t1(controls, dt, gate, gate_count, fq, da, rho0, nthreads=0):
    """rho0: (S,2,2) complex.  Returns fid (eigen route, as the reference), fid_closed, rho_out (S,2,2)."""
    controls = _f64(controls)
    rho0 = np.ascontiguousarray(rho0, dtype=np.complex128)
    S = rho0.shape[0]
    vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4))  # column-major vec
    da = None if da is None else _f64(da)
    fid = np.empty((S, gate_count + 1))
    fidc = np.empty((S, gate_count + 1))
    out = np.empty((S, 4), dtype=np.complex128)
    rc = lib().orc_lindblad_t1(_p(controls), C.c_int64(controls.shape[0]), C.c_double(dt), int(gate),
                               C.c_int64(gate_count), C.c_int64(S), C.c_double(fq), _p(da), _p(vec), _p(fid),
                               _p(fidc), _p(out), int(nthreads))
    if rc:
        raise ValueError(f"orc_lindblad_t1 rc={rc}")
    return fid, fidc, out.reshape(S, 2, 2).transpose(0, 2, 1).copy()


def mc_static_segments(amps, durs, gate, gate_count, fq, da, psi0):
    amps, durs, fq, psi0 = _f64(amps), _f64(durs), _f64(fq), _f64(psi0)
    da = None if da is None else _f64(da)
    S = fq.shape[0]
    fid = np.empty((S, gate_count + 1))
    rc = lib().orc_mc_static_segments(_p(amps), _p(durs), C.c_int64(amps.shape[0]), int(gate), C.c_int64(gate_count), C.c_int64(S),
                                      _p(fq), _p(da), _p(psi0), _p(fid))
    if rc:
        raise ValueError(f"orc_mc_static_segments rc={rc}")
    return fid


def mc_noise_segments(amps, durs, nidx, gate, gate_count, fq, noise, psi0):
    amps, durs, noise, psi0 = _f64(amps), _f64(durs), _f64(noise), _f64(psi0)
    nidx = np.ascontiguousarray(nidx, dtype=np.int32)
    S, nlen = noise.shape
    fid = np.empty((S, gate_count + 1))
    rc = lib().orc_mc_noise_segments(_p(amps), _p(durs), _p(nidx), C.c_int64(amps.shape[0]), int(gate), C.c_int64(gate_count),
                                     C.c_int64(S), C.c_double(fq), _p(noise), C.c_int64(nlen), _p(psi0), _p(fid))
    if rc:
        raise ValueError(f"orc_mc_noise_segments rc={rc}")
    return fid


def mc_timeline(amps, durs, nidx, save_after, gate, fq, da, noise, psi0, want_states=False):
    amps, durs, fq, psi0 = _f64(amps), _f64(durs), _f64(fq), _f64(psi0)
    save_after = np.ascontiguousarray(save_after, dtype=np.int64)
    da = None if da is None else _f64(da)
    nlen = 0
    if noise is not None:
        noise = _f64(noise)
        nidx = np.ascontiguousarray(nidx, dtype=np.int32)
        nlen = noise.shape[1]
    else:
        nidx = None
    S, nsave = fq.shape[0], save_after.shape[0]
    fid = np.empty((S, nsave + 1))
    states = np.empty((S, nsave + 1, 4)) if want_states else None
    rc = lib().orc_mc_timeline(_p(amps), _p(durs), _p(nidx), C.c_int64(amps.shape[0]), _p(save_after), C.c_int64(nsave), int(gate),
                               C.c_int64(S), _p(fq), _p(da), _p(noise), C.c_int64(nlen), _p(psi0), _p(fid), _p(states))
    if rc:
        raise ValueError(f"orc_mc_timeline rc={rc}")
    return (fid, states) if want_states else fid


def lindblad_segments(amps, durs, gate, gate_count, fq, da, rho0):
    amps, durs = _f64(amps), _f64(durs)
    rho0 = np.ascontiguousarray(rho0, dtype=np.complex128)
    S = rho0.shape[0]
    vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4))
    da = None if da is None else _f64(da)
    fidc = np.empty((S, gate_count + 1))
    out = np.empty((S, 4), dtype=np.complex128)
    rc = lib().orc_lindblad_segments(_p(amps), _p(durs), C.c_int64(amps.shape[0]), int(gate), C.c_int64(gate_count), C.c_int64(S),
                                     C.c_double(fq), _p(da), _p(vec), _p(fidc), _p(out))
    if rc:
        raise ValueError(f"orc_lindblad_segments rc={rc}")
    return fidc, out.reshape(S, 2, 2).transpose(0, 2, 1).copy()


def lindblad_ode(amps, durs, save_after, gate, fq, da, rho0, channels=1, gamma_c=0.0, gamma_f=0.0, nthreads=0):
    """Exact (long-double Taylor) solution of the reference's density ODEs dynamics_lindbladt1_deqjl / lindbladt2_deqjl
    (spin.jl:505-518, 542-554) over a piecewise-constant segment timeline.  rho0 (S,2,2) complex.
    Returns fid_closed (S, nsave+1), rho (S, nsave+1, 2, 2)."""
    amps, durs = _f64(amps), _f64(durs)
    save_after = np.ascontiguousarray(save_after, dtype=np.int64)
    nsave = save_after.shape[0]
    rho0 = np.ascontiguousarray(rho0, dtype=np.complex128)
    S = rho0.shape[0]
    vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4))
    da = None if da is None else _f64(da)
    fidc = np.empty((S, nsave + 1))
    out = np.empty((S, nsave + 1, 4), dtype=np.complex128)
    rc = lib().orc_lindblad_ode(_p(amps), _p(durs), C.c_int64(amps.shape[0]), _p(save_after), C.c_int64(nsave), int(gate),
                                C.c_int64(S), C.c_double(fq), _p(da), _p(vec), int(channels), C.c_double(gamma_c),
                                C.c_double(gamma_f), _p(fidc), _p(out), int(nthreads))
    if rc:
        raise ValueError(f"orc_lindblad_ode rc={rc}")
    return fidc, out.reshape(S, nsave + 1, 2, 2).transpose(0, 1, 3, 2).copy()


def lindblad_knots(controls, dt, gate, gate_count, fq, da, rho0, channels=1, gamma_c=0.0, gamma_f=0.0, nthreads=0):
    """lindblad_ode on the knot grid: `gate_count` repetitions of the pulse, saved after every gate (what run_sim_prop's
    lindblad types integrate, with the T2 rate running on through the gates).  Returns fid_closed (S, gate_count+1), rho_final."""
    controls = _f64(controls).reshape(-1)
    nk = controls.shape[0]
    amps = np.tile(controls, gate_count)
    durs = np.full(nk * gate_count, float(dt))
    save_after = nk * np.arange(1, gate_count + 1)
    fidc, rho = lindblad_ode(amps, durs, save_after, gate, fq, da, rho0, channels, gamma_c, gamma_f, nthreads)
    return fidc, rho[:, -1]


def lindblad_generator(fq, a):
    L = np.empty((4, 4), dtype=np.complex128)
    lib().orc_lindblad_generator(C.c_double(fq), C.c_double(a), _p(L))
    return L


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def philox_samples(first, S, seed, fq0, fq_rel_sigma, fq_rel_clip, da_sigma):
    fq, da, psi0 = np.empty(S), np.empty(S), np.empty((S, 4))
    lib().orc_philox_samples(C.c_int64(first), C.c_int64(S), C.c_uint64(seed), C.c_double(fq0),
                             C.c_double(fq_rel_sigma), C.c_double(fq_rel_clip), C.c_double(da_sigma), _p(fq),
                             _p(da), _p(psi0))
    return fq, da, psi0


def pink_filter(xs, dt_inv):
    xs = _f64(xs)
    ys = np.empty_like(xs)
    lib().orc_pink_filter(_p(xs), C.c_int64(xs.shape[0]), C.c_double(dt_inv), _p(ys))
    return ys


def pink_noise(first, S, seed, namp, noise_dt_inv, noise_len):
    out = np.empty((S, noise_len))
    lib().orc_pink_noise(C.c_int64(first), C.c_int64(S), C.c_uint64(seed), C.c_double(namp),
                         C.c_double(noise_dt_inv), C.c_int64(noise_len), _p(out))
    return out


def stats(e) -> Stats:
    e = _f64(e)
    st = Stats()
    lib().orc_stats(_p(e), C.c_int64(e.shape[0]), C.byref(st))
    return st
