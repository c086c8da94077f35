"""Engine: one `rbq_handle` (one CUDA device) behind numpy / torch-tensor friendly calls.

Every method is a thin argument-marshalling layer over one C-ABI entry point of include/rbq.h;
all arithmetic happens in librbq.so's sm_100a kernels.  Host arrays are numpy float64;
`*_dev` methods take torch CUDA tensors (device memory and streams are torch's job here,
nothing else is).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ModelParams, RbqError, Stats


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    # torch tensor on the device
    if a.dtype.itemsize != 8 or not a.is_contiguous():
        raise ValueError("device tensors must be contiguous 8-byte element tensors")
    return C.c_void_p(a.data_ptr())


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.rbq_create(int(device), C.byref(h))
        if rc:
            raise RbqError(rc, "rbq_create", "no sm_100 CUDA device; the engine has no CPU fallback")
        self._h = h
        self.device = int(device)

    def close(self):
        """Destroys the handle.  Page-locked arrays from pinned_empty own their memory (freed when the last view of the
        array is collected, through cudaFreeHost directly), so they stay valid after close()."""
        if getattr(self, "_h", None):
            self._lib.rbq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc, where):
        if rc:
            raise RbqError(rc, where, self._lib.rbq_last_error(self._h).decode(errors="replace"))

    # ---- plumbing ------------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        self._ck(self._lib.rbq_set_stream(self._h, C.c_void_p(cuda_stream)), "rbq_set_stream")

    def set_sweep_centre(self, fq0: float, da0: float = 0.0):
        """Performance hint: where the explicit samples of mc_static scatter (default FQ, 0); see include/rbq.h."""
        self._ck(self._lib.rbq_set_sweep_centre(self._h, float(fq0), float(da0)), "rbq_set_sweep_centre")

    def set_option(self, name: str, value=None):
        """One tuning / A-B switch of this handle (rbq_set_option): same names and values as the RBQ_* environment
        variables the library latches in rbq_create; value None restores the default."""
        v = None if value is None else str(value).encode()
        self._ck(self._lib.rbq_set_option(self._h, name.encode(), v), "rbq_set_option")

    def use_torch_stream(self):
        import torch

        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        self._ck(self._lib.rbq_synchronize(self._h), "rbq_synchronize")

    def pinned_empty(self, shape, dtype=np.float64):
        """numpy array backed by page-locked memory from rbq_host_alloc.  The memory belongs to the array: a finalizer on
        the ctypes buffer every view keeps alive returns it with rbq_host_free(NULL handle -> cudaFreeHost) when the last
        view is collected, whether or not this Engine still exists."""
        import weakref

        shape = tuple(int(x) for x in np.atleast_1d(shape))
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        self._ck(self._lib.rbq_host_alloc(self._h, max(nbytes, 1), C.byref(ptr)), "rbq_host_alloc")
        buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
        weakref.finalize(buf, self._lib.rbq_host_free, None, C.c_void_p(ptr.value))
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def host_register(self, arr: np.ndarray):
        """Page-lock a numpy array in place (rbq_host_register); pair with host_unregister before the array is freed."""
        self._ck(self._lib.rbq_host_register(self._h, arr.ctypes.data_as(C.c_void_p), arr.nbytes), "rbq_host_register")

    def host_unregister(self, arr: np.ndarray):
        self._ck(self._lib.rbq_host_unregister(self._h, arr.ctypes.data_as(C.c_void_p)), "rbq_host_unregister")

    @property
    def launch_count(self) -> int:
        return int(self._lib.rbq_launch_count(self._h))

    def fp64_peak(self, iters: int = 20000) -> float:
        out = C.c_double()
        self._ck(self._lib.rbq_fp64_peak(self._h, int(iters), C.byref(out)), "rbq_fp64_peak")
        return out.value

    # ---- model API (RobotDynamics) -------------------------------------------------------------------
    @staticmethod
    def state_dim(p: ModelParams) -> int:
        return _lib.load().rbq_state_dim(C.byref(p))

    @staticmethod
    def control_dim(p: ModelParams) -> int:
        return _lib.load().rbq_control_dim(C.byref(p))

    def _dims(self, p):
        n, m = self.state_dim(p), self.control_dim(p)
        if n < 0 or m < 0:
            raise RbqError(_lib.E_MODEL, "model parameters")
        return n, m

    def discrete_dynamics(self, p, x, u, t, dt):
        n, m = self._dims(p)
        x, u = _f64(x), _f64(u)
        if x.shape != (n,) or u.shape != (m,):
            raise ValueError(f"expected x of length {n} and u of length {m}")
        out = np.empty(n)
        self._ck(self._lib.rbq_discrete_dynamics(self._h, C.byref(p), _ptr(x), _ptr(u), float(t), float(dt), _ptr(out)),
                 "rbq_discrete_dynamics")
        return out

    def discrete_jacobian(self, p, x, u, t, dt):
        """n x (n+m) matrix [A B] (the C side writes it column-major, like RD's ∇f)."""
        n, m = self._dims(p)
        x, u = _f64(x), _f64(u)
        if x.shape != (n,) or u.shape != (m,):
            raise ValueError(f"expected x of length {n} and u of length {m}")
        ab = np.empty((n + m, n))
        self._ck(self._lib.rbq_discrete_jacobian(self._h, C.byref(p), _ptr(x), _ptr(u), float(t), float(dt), _ptr(ab)),
                 "rbq_discrete_jacobian")
        return ab.T

    def rollout(self, p, x0, U, dt):
        """x0 (B,n)|(n,), U (B,N-1,m)|(N-1,m), dt (N-1,) -> X (B,N,n)|(N,n)."""
        n, m = self._dims(p)
        x0, U, dt = _f64(x0), _f64(U), _f64(dt)
        single = x0.ndim == 1
        x0b = x0.reshape(-1, n)
        B = x0b.shape[0]
        if U.ndim < 2:
            raise ValueError("U must be (N-1, m) or (B, N-1, m)")
        N = U.shape[-2] + 1
        Ub = U.reshape(B, N - 1, m)
        if dt.shape != (N - 1,):
            raise ValueError("dt must have N-1 entries")
        X = np.empty((B, N, n))
        self._ck(self._lib.rbq_rollout(self._h, C.byref(p), B, N, _ptr(x0b), _ptr(Ub), _ptr(dt), _ptr(X)), "rbq_rollout")
        return X[0] if single else X

    def rollout_closedloop(self, p, Xref, Uref, K, d, dt, alphas):
        """K (N-1,m,n) gains (numpy row-major m x n per knot is converted to the column-major ABI layout)."""
        n, m = self._dims(p)
        Xref, Uref, d, dt, alphas = map(_f64, (Xref, Uref, d, dt, alphas))
        N = Xref.shape[0]
        Kc = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(N - 1, m, n).transpose(0, 2, 1))
        na = alphas.shape[0]
        X = np.empty((na, N, n))
        Uo = np.empty((na, N - 1, m))
        self._ck(self._lib.rbq_rollout_closedloop(self._h, C.byref(p), N, _ptr(Xref), _ptr(Uref), _ptr(Kc), _ptr(d), _ptr(dt),
                                                  na, _ptr(alphas), _ptr(X), _ptr(Uo)), "rbq_rollout_closedloop")
        return X, Uo

    def jacobians(self, p, X, U, dt, t=None, out=None):
        """X (K,n), U (K,m), dt (K,), t (K,) knot times or None -> (K, n, n+m).  Only spin25 reads t.
        out: optional (K, n+m, n) float64 buffer (e.g. from pinned_empty) that receives the column-major blocks."""
        n, m = self._dims(p)
        X, U, dt = _f64(X).reshape(-1, n), _f64(U).reshape(-1, m), _f64(dt)
        K = X.shape[0]
        if U.shape[0] != K or dt.shape != (K,):
            raise ValueError("X, U, dt disagree on the knot count")
        if t is not None:
            t = _f64(t)
            if t.shape != (K,):
                raise ValueError("t must have one entry per knot")
        if out is not None:
            if out.shape != (K, n + m, n) or out.dtype != np.float64 or not out.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous float64 array of shape (K, n+m, n)")
            ab = out
        else:
            ab = np.empty((K, n + m, n))
        self._ck(self._lib.rbq_jacobians(self._h, C.byref(p), K, _ptr(X), _ptr(U), _ptr(t), _ptr(dt), _ptr(ab)), "rbq_jacobians")
        return ab.transpose(0, 2, 1)

    def rollout_dev(self, p, B, N, x0, U, dt, X):
        self._ck(self._lib.rbq_rollout_dev(self._h, C.byref(p), B, N, _ptr(x0), _ptr(U), _ptr(dt), _ptr(X)), "rbq_rollout_dev")

    def rollout_closedloop_dev(self, p, N, Xref, Uref, K, d, dt, nalpha, alphas, X, Uout):
        """Device tensors throughout; K column-major m x n per knot, as backward_pass_dev writes it."""
        self._ck(self._lib.rbq_rollout_closedloop_dev(self._h, C.byref(p), int(N), _ptr(Xref), _ptr(Uref), _ptr(K), _ptr(d), _ptr(dt),
                                                      int(nalpha), _ptr(alphas), _ptr(X), _ptr(Uout)), "rbq_rollout_closedloop_dev")

    def jacobians_dev(self, p, K, X, U, dt, AB, t=None):
        self._ck(self._lib.rbq_jacobians_dev(self._h, C.byref(p), K, _ptr(X), _ptr(U), _ptr(t), _ptr(dt), _ptr(AB)), "rbq_jacobians_dev")

    # ---- cost-side primitives (spin.jl:1040-1092) ---------------------------------------------------------------
    def gate_error_expansion(self, s1, s2, want_grad=True, want_hess=True):
        """s1 (K,4) states, s2 (K,4) or (4,) targets -> err (K,), grad (K,4), hess (K,4,4)."""
        s1 = _f64(s1).reshape(-1, 4)
        s2 = _f64(s2)
        K = s1.shape[0]
        st2 = 0 if s2.ndim == 1 else 4
        if s2.ndim == 2 and s2.shape != (K, 4):
            raise ValueError("s2 must be (4,) or (K,4)")
        err = np.empty(K)
        grad = np.empty((K, 4)) if want_grad else None
        hess = np.empty((K, 4, 4)) if want_hess else None
        self._ck(self._lib.rbq_gate_error_expansion(self._h, K, _ptr(s1), 4, _ptr(s2), st2, _ptr(err), _ptr(grad), _ptr(hess)),
                 "rbq_gate_error_expansion")
        return err, grad, hess

    def sampling_cost_expansion(self, X, U, Qdiag, Rdiag, xf, targets, q_ss):
        """spin12/spin18 `Cost` (spin12.jl:59-161): X (K,43), U (K,1) or None -> J (K,), gx (K,43), gu (K,1) or None."""
        X = _f64(X).reshape(-1, 43)
        K = X.shape[0]
        U = None if U is None else _f64(U).reshape(K, 1)
        Qdiag, xf, targets, q_ss = _f64(Qdiag), _f64(xf), _f64(targets).reshape(4, 4), _f64(q_ss)
        Rdiag = None if Rdiag is None else _f64(Rdiag)
        J, gx = np.empty(K), np.empty((K, 43))
        gu = None if U is None else np.empty((K, 1))
        self._ck(self._lib.rbq_sampling_cost_expansion(self._h, K, _ptr(X), _ptr(U), _ptr(Qdiag), _ptr(Rdiag), _ptr(xf),
                                                       _ptr(targets), _ptr(q_ss), _ptr(J), _ptr(gx), _ptr(gu)),
                 "rbq_sampling_cost_expansion")
        return J, gx, gu

    # ---- Monte-Carlo sweeps ------------------------------------------------------------------------------
    def mc_static(self, controls, dt, gate, gate_count, fq, da, psi0, algo=_lib.ALGO_SU2, want_fid=True, want_stats=True,
                  fid_out=None):
        controls, fq, psi0 = _f64(controls), _f64(fq), _f64(psi0)
        da = None if da is None else _f64(da)
        S = fq.shape[0]
        if psi0.shape != (S, 4) or (da is not None and da.shape != (S,)):
            raise ValueError("fq (S,), da (S,), psi0 (S,4) expected")
        fid = fid_out if fid_out is not None else (np.empty((S, gate_count + 1)) if want_fid else None)
        st = Stats() if want_stats else None
        self._ck(self._lib.rbq_mc_static(self._h, _ptr(controls), controls.shape[0], float(dt), int(gate), int(gate_count), S,
                                         _ptr(fq), _ptr(da), _ptr(psi0), int(algo), _ptr(fid),
                                         C.byref(st) if st is not None else None), "rbq_mc_static")
        return fid, st

    def mc_static_philox(self, controls, dt, gate, gate_count, first, S, seed, fq0, fq_rel_sigma, fq_rel_clip, da_sigma,
                         algo=_lib.ALGO_SU2, want_fid=False, fid_out=None):
        """Seeds in, fidelities / statistics out (the shape of the reference's run_sim_prop(...; seed, state_seed)).
        fid_out: optional (S, gate_count+1) float64 buffer; when it is page-locked (pinned_empty, torch pin_memory) the kernel
        writes it directly."""
        controls = _f64(controls)
        fid = fid_out if fid_out is not None else (np.empty((S, gate_count + 1)) if want_fid else None)
        st = Stats()
        self._ck(self._lib.rbq_mc_static_philox(self._h, _ptr(controls), controls.shape[0], float(dt), int(gate),
                                                int(gate_count), int(first), int(S), int(seed), float(fq0),
                                                float(fq_rel_sigma), float(fq_rel_clip), float(da_sigma), int(algo),
                                                _ptr(fid), C.byref(st)), "rbq_mc_static_philox")
        return fid, st

    def philox_samples(self, first, S, seed, fq0, fq_rel_sigma, fq_rel_clip, da_sigma):
        fq, da, psi0 = np.empty(S), np.empty(S), np.empty((S, 4))
        self._ck(self._lib.rbq_philox_samples(self._h, int(first), int(S), int(seed), float(fq0), float(fq_rel_sigma),
                                              float(fq_rel_clip), float(da_sigma), _ptr(fq), _ptr(da), _ptr(psi0)),
                 "rbq_philox_samples")
        return fq, da, psi0

    def mc_noise(self, controls, dt, gate, gate_count, fq, noise, noise_dt_inv, psi0, algo=_lib.ALGO_SU2):
        controls, noise, psi0 = _f64(controls), _f64(noise), _f64(psi0)
        S, nlen = noise.shape
        fid = np.empty((S, gate_count + 1))
        st = Stats()
        self._ck(self._lib.rbq_mc_noise(self._h, _ptr(controls), controls.shape[0], float(dt), int(gate), int(gate_count), S,
                                        float(fq), _ptr(noise), nlen, float(noise_dt_inv), _ptr(psi0), int(algo), _ptr(fid),
                                        C.byref(st)), "rbq_mc_noise")
        return fid, st

    def mc_pink_philox(self, controls, dt, gate, gate_count, first, S, seed, fq, namp, noise_dt_inv, algo=_lib.ALGO_SU2,
                       want_fid=True):
        controls = _f64(controls)
        fid = np.empty((S, gate_count + 1)) if want_fid else None
        st = Stats()
        self._ck(self._lib.rbq_mc_pink_philox(self._h, _ptr(controls), controls.shape[0], float(dt), int(gate),
                                              int(gate_count), int(first), int(S), int(seed), float(fq), float(namp),
                                              float(noise_dt_inv), int(algo), _ptr(fid), C.byref(st)), "rbq_mc_pink_philox")
        return fid, st

    def pink_noise(self, first, S, seed, namp, noise_dt_inv, noise_len):
        out = np.empty((S, noise_len))
        self._ck(self._lib.rbq_pink_noise(self._h, int(first), int(S), int(seed), float(namp), float(noise_dt_inv),
                                          int(noise_len), _ptr(out)), "rbq_pink_noise")
        return out

    def lindblad(self, controls, dt, gate, gate_count, fq, da, rho0, channels=_lib.LINDBLAD_T1, t2_gamma_c=0.0,
                 t2_gamma_f=0.0):
        """rho0 (S,2,2) complex -> fid (S,gate_count+1), rho_out (S,2,2), stats."""
        controls = _f64(controls)
        rho0 = np.ascontiguousarray(rho0, dtype=np.complex128)
        S = rho0.shape[0]
        vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4))   # column-major vec(rho)
        da = None if da is None else _f64(da)
        fid = np.empty((S, gate_count + 1))
        out = np.empty((S, 4), dtype=np.complex128)
        st = Stats()
        self._ck(self._lib.rbq_lindblad(self._h, _ptr(controls), controls.shape[0], float(dt), int(gate), int(gate_count), S,
                                        float(fq), _ptr(da), _ptr(vec), int(channels), float(t2_gamma_c), float(t2_gamma_f),
                                        _ptr(fid), _ptr(out), C.byref(st)), "rbq_lindblad")
        return fid, out.reshape(S, 2, 2).transpose(0, 2, 1).copy(), st

    def lindblad_into(self, controls, dt, gate, gate_count, fq, da, vec, fid, rho_out, channels=_lib.LINDBLAD_T1, t2_gamma_c=0.0,
                      t2_gamma_f=0.0):
        """rbq_lindblad on buffers already in the ABI's layout and owned by the caller (page-locked ones from `pinned_empty`
        travel at the full PCIe rate): vec (S, 8) doubles = column-major vec(rho0) as (re, im) pairs, fid (S, gate_count+1),
        rho_out (S, 8).  Nothing is allocated, transposed or copied on the host.  Returns the statistics."""
        S = vec.shape[0]
        if vec.dtype != np.float64 or vec.shape != (S, 8) or fid.shape != (S, gate_count + 1) or rho_out.shape != (S, 8) or \
                not (vec.flags.c_contiguous and fid.flags.c_contiguous and rho_out.flags.c_contiguous):
            raise ValueError("lindblad_into: vec (S,8), fid (S,gate_count+1), rho_out (S,8) contiguous float64")
        st = Stats()
        self._ck(self._lib.rbq_lindblad(self._h, _ptr(controls), controls.shape[0], float(dt), int(gate), int(gate_count), S,
                                        float(fq), _ptr(da), _ptr(vec), int(channels), float(t2_gamma_c), float(t2_gamma_f),
                                        _ptr(fid), _ptr(rho_out), C.byref(st)), "rbq_lindblad")
        return st

    # ---- analytic gates: segments with their own durations (spin.jl:562-594, 651-670, 765-785) -------------------------
    def mc_static_segments(self, amps, durs, gate, gate_count, fq, da, psi0):
        amps, durs, fq, psi0 = _f64(amps), _f64(durs), _f64(fq), _f64(psi0)
        da = None if da is None else _f64(da)
        S = fq.shape[0]
        if amps.shape != durs.shape or psi0.shape != (S, 4):
            raise ValueError("amps/durs must have one entry per segment, psi0 (S,4)")
        fid = np.empty((S, gate_count + 1))
        st = Stats()
        self._ck(self._lib.rbq_mc_static_segments(self._h, _ptr(amps), _ptr(durs), amps.shape[0], int(gate), int(gate_count), S,
                                                  _ptr(fq), _ptr(da), _ptr(psi0), _ptr(fid), C.byref(st)), "rbq_mc_static_segments")
        return fid, st

    def mc_noise_segments(self, amps, durs, noise_idx, gate, gate_count, fq, noise, psi0):
        """integrate_prop_xpiby2da! (spin.jl:819-905): segment j sees amps[j] + noise[s, noise_idx[j]] in every gate."""
        amps, durs, noise, psi0 = _f64(amps), _f64(durs), _f64(noise), _f64(psi0)
        nidx = np.ascontiguousarray(noise_idx, dtype=np.int32)
        if noise.ndim != 2:
            raise ValueError("noise must be (S, noise_len)")
        S, nlen = noise.shape
        if amps.shape != durs.shape or nidx.shape != amps.shape or psi0.shape != (S, 4):
            raise ValueError("amps/durs/noise_idx must have one entry per segment, psi0 (S,4)")
        fid = np.empty((S, gate_count + 1))
        st = Stats()
        self._ck(self._lib.rbq_mc_noise_segments(self._h, _ptr(amps), _ptr(durs), _ptr(nidx), amps.shape[0], int(gate),
                                                 int(gate_count), S, float(fq), _ptr(noise), nlen, _ptr(psi0), _ptr(fid),
                                                 C.byref(st)), "rbq_mc_noise_segments")
        return fid, st

    def mc_timeline(self, amps, durs, save_after, gate, fq, psi0, da=None, noise=None, noise_idx=None, want_states=False):
        """run_sim_deqjl's piecewise-constant ODE solved exactly (spin.jl:424-467, 1277-1380): segments (amps, durs[, noise_idx]),
        fidelities at t = 0 and after save_after[g] segments -> (S, nsave + 1).  want_states: also the iso states at those
        times, (S, nsave + 1, 4) (run_sim_fine_deqjl's "states"), returned as a third value."""
        amps, durs, fq, psi0 = _f64(amps), _f64(durs), _f64(fq), _f64(psi0)
        save_after = np.ascontiguousarray(save_after, dtype=np.int64)
        da = None if da is None else _f64(da)
        S, nsave = fq.shape[0], save_after.shape[0]
        if amps.shape != durs.shape or psi0.shape != (S, 4) or (da is not None and da.shape != (S,)):
            raise ValueError("amps/durs one entry per segment, fq (S,), da (S,), psi0 (S,4)")
        nlen, nidx = 0, None
        if noise is not None:
            noise = _f64(noise)
            nidx = np.ascontiguousarray(noise_idx, dtype=np.int32)
            if noise.ndim != 2 or noise.shape[0] != S or nidx.shape != amps.shape:
                raise ValueError("noise (S, noise_len) with one noise_idx per segment")
            nlen = noise.shape[1]
        fid = np.empty((S, nsave + 1))
        states = np.empty((S, nsave + 1, 4)) if want_states else None
        st = Stats()
        self._ck(self._lib.rbq_mc_timeline_states(self._h, _ptr(amps), _ptr(durs), _ptr(nidx), amps.shape[0], _ptr(save_after), nsave,
                                                  int(gate), S, _ptr(fq), _ptr(da), _ptr(noise), nlen, _ptr(psi0), _ptr(fid),
                                                  _ptr(states), C.byref(st)), "rbq_mc_timeline_states")
        return (fid, st, states) if want_states else (fid, st)

    def lindblad_timeline(self, amps, durs, save_after, gate, fq, rho0, da=None, channels=_lib.LINDBLAD_T1, t2_gamma_c=0.0,
                          t2_gamma_f=0.0):
        """The density ODEs of run_sim_deqjl / run_sim_fine_deqjl (spin.jl:488-518, 542-554) over a segment timeline.
        rho0 (S,2,2) complex -> fid (S, nsave+1), rho (S, nsave+1, 2, 2), stats."""
        amps, durs = _f64(amps), _f64(durs)
        save_after = np.ascontiguousarray(save_after, dtype=np.int64)
        nsave = save_after.shape[0]
        rho0 = np.ascontiguousarray(rho0, dtype=np.complex128)
        S = rho0.shape[0]
        vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4))
        da = None if da is None else _f64(da)
        if amps.shape != durs.shape or (da is not None and da.shape != (S,)):
            raise ValueError("amps/durs one entry per segment, da (S,)")
        fid = np.empty((S, nsave + 1))
        out = np.empty((S, nsave + 1, 4), dtype=np.complex128)
        st = Stats()
        self._ck(self._lib.rbq_lindblad_timeline(self._h, _ptr(amps), _ptr(durs), amps.shape[0], _ptr(save_after), nsave, int(gate), S,
                                                 float(fq), _ptr(da), _ptr(vec), int(channels), float(t2_gamma_c), float(t2_gamma_f),
                                                 _ptr(fid), _ptr(out), C.byref(st)), "rbq_lindblad_timeline")
        return fid, out.reshape(S, nsave + 1, 2, 2).transpose(0, 1, 3, 2).copy(), st

    def lindblad_segments(self, amps, durs, gate, gate_count, fq, da, rho0, channels=_lib.LINDBLAD_T1, t2_gamma_c=0.0):
        amps, durs = _f64(amps), _f64(durs)
        rho0 = np.ascontiguousarray(rho0, dtype=np.complex128)
        S = rho0.shape[0]
        vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4))
        da = None if da is None else _f64(da)
        fid = np.empty((S, gate_count + 1))
        out = np.empty((S, 4), dtype=np.complex128)
        st = Stats()
        self._ck(self._lib.rbq_lindblad_segments(self._h, _ptr(amps), _ptr(durs), amps.shape[0], int(gate), int(gate_count), S,
                                                 float(fq), _ptr(da), _ptr(vec), int(channels), float(t2_gamma_c), _ptr(fid),
                                                 _ptr(out), C.byref(st)), "rbq_lindblad_segments")
        return fid, out.reshape(S, 2, 2).transpose(0, 2, 1).copy(), st

    # ---- iLQR backward pass (Altro backwardpass!, external) --------------------------------------------------------------
    def backward_pass(self, AB, Exx, Ex, Euu, Eux, Eu, rho=0.0, reg_type=_lib.REG_CONTROL):
        """AB (B, N-1, n, n+m) or (N-1, n, n+m) = [A_k B_k] as `jacobians` returns it (a view of the column-major blocks: no
        copy is made); Exx (.., N, n, n) symmetric, Ex (.., N, n), Euu (.., N-1, m, m), Eux (.., N-1, m, n), Eu (.., N-1, m).
        Returns K (.., N-1, m, n), d (.., N-1, m), dV (.., 2), fail (-1 or the knot where Quu_reg was not positive definite)."""
        AB, Exx, Ex, Euu, Eux, Eu = (np.asarray(x, dtype=np.float64) for x in (AB, Exx, Ex, Euu, Eux, Eu))
        single = Ex.ndim == 2
        if single:
            AB, Exx, Ex, Euu, Eux, Eu = (x[None] for x in (AB, Exx, Ex, Euu, Eux, Eu))
        B, N, n = Ex.shape
        m = Eu.shape[2]
        if AB.shape != (B, N - 1, n, n + m) or Exx.shape != (B, N, n, n) or Euu.shape != (B, N - 1, m, m) or \
                Eux.shape != (B, N - 1, m, n) or Eu.shape != (B, N - 1, m):
            raise ValueError("inconsistent shapes")
        col = lambda x: np.ascontiguousarray(np.swapaxes(x, -1, -2))          # column-major blocks for the ABI
        ABc, Exxc, Euuc, Euxc = col(AB), col(Exx), col(Euu), col(Eux)
        Ex, Eu = np.ascontiguousarray(Ex), np.ascontiguousarray(Eu)
        K = np.empty((B, N - 1, n, m)); d = np.empty((B, N - 1, m)); dV = np.empty((B, 2)); fail = np.empty(B, dtype=np.int64)
        self._ck(self._lib.rbq_backward_pass(self._h, n, m, N, B, _ptr(ABc), _ptr(Exxc), _ptr(Ex), _ptr(Euuc), _ptr(Euxc), _ptr(Eu),
                                             float(rho), int(reg_type), _ptr(K), _ptr(d), _ptr(dV), _ptr(fail)), "rbq_backward_pass")
        K = np.swapaxes(K, -1, -2)
        return (K[0], d[0], dV[0], int(fail[0])) if single else (K, d, dV, fail)

    def backward_pass_dev(self, n, m, N, B, d_AB, d_Exx, d_Ex, d_Euu, d_Eux, d_Eu, rho, reg_type, d_K, d_d, d_dV, d_fail):
        self._ck(self._lib.rbq_backward_pass_dev(self._h, int(n), int(m), int(N), int(B), _ptr(d_AB), _ptr(d_Exx), _ptr(d_Ex),
                                                 _ptr(d_Euu), _ptr(d_Eux), _ptr(d_Eu), float(rho), int(reg_type), _ptr(d_K),
                                                 _ptr(d_d), _ptr(d_dV), _ptr(d_fail)), "rbq_backward_pass_dev")

    def al_dual_update_dev(self, N, table, d_c, d_lambda, d_mu, d_max_violation, phi=10.0, mu_max=1e8, lambda_max=1e8, update=True,
                           B=1, dual_stride=0):
        """Multiplier / penalty update and largest constraint violation on the device (rbq_al_dual_update_dev)."""
        max_p = max((r.p for r in table.rows), default=1)
        self._ck(self._lib.rbq_al_dual_update_dev(self._h, int(N), int(B), int(table.ncon), C.c_void_p(table.d_cons.data_ptr()),
                                                  int(max_p), _ptr(d_c), _ptr(d_lambda), _ptr(d_mu), int(dual_stride), float(phi),
                                                  float(mu_max), float(lambda_max), 1 if update else 0, _ptr(d_max_violation)),
                 "rbq_al_dual_update_dev")

    def al_cost_expansion(self, X, U, Qdiag, Rdiag, Qfdiag, xf, table, lam, mu, w=None, expand=True, shared_duals=True):
        """Host-array form of the AL cost expansion.  X (N, n) or (B, N, n), U (N-1, m) or (B, N-1, m); `table` a
        spin.ConstraintTable; lam / mu flat (dual_size,) or (B, dual_size) when not shared.  Returns a dict with J (.., N),
        c (.., dual_size) and, if `expand`, Exx (.., N, n, n), Ex, Euu, Eux (.., N-1, m, n), Eu in math layout."""
        X = _f64(X); U = _f64(U)
        single = X.ndim == 2
        if single:
            X, U = X[None], U[None]
        B, N, n = X.shape
        m = U.shape[2]
        lam, mu = _f64(lam), _f64(mu)
        rows = (_lib.Constraint * max(table.ncon, 1))(*table.rows)
        idx = np.asarray(table.idx or [0], dtype=np.int32); val = np.asarray(table.val or [0.0], dtype=np.float64)
        J = np.empty((B, N)); c = np.empty((B, table.dual_size))
        if expand:
            Exx = np.empty((B, N, n, n)); Ex = np.empty((B, N, n)); Euu = np.empty((B, N - 1, m, m))
            Eux = np.empty((B, N - 1, n, m)); Eu = np.empty((B, N - 1, m))
        else:
            Exx = Ex = Euu = Eux = Eu = None
        self._ck(self._lib.rbq_al_cost_expansion(self._h, n, m, N, B, _ptr(X), _ptr(U), _ptr(_f64(Qdiag)), _ptr(_f64(Rdiag)),
                                                 _ptr(_f64(Qfdiag)), _ptr(_f64(xf)), _ptr(None if w is None else _f64(w)), table.ncon,
                                                 C.cast(rows, C.c_void_p), _ptr(idx), len(table.idx), _ptr(val), len(table.val),
                                                 _ptr(lam), _ptr(mu), 0 if shared_duals else table.dual_size, _ptr(Exx), _ptr(Ex),
                                                 _ptr(Euu), _ptr(Eux), _ptr(Eu), _ptr(c), _ptr(J)), "rbq_al_cost_expansion")
        out = {"J": J, "c": c}
        if expand:
            out.update(Exx=np.swapaxes(Exx, -1, -2), Ex=Ex, Euu=np.swapaxes(Euu, -1, -2), Eux=np.swapaxes(Eux, -1, -2), Eu=Eu)
        return {k: v[0] for k, v in out.items()} if single else out

    def al_cost_expansion_dev(self, n, m, N, d_X, d_U, d_Qdiag, d_Rdiag, d_Qfdiag, d_xf, d_w, table, d_lambda, d_mu,
                              d_Exx, d_Ex, d_Euu, d_Eux, d_Eu, d_c, d_J, B=1, dual_stride=0):
        """Augmented-Lagrangian cost blocks of all knots (rbq_al_cost_expansion_dev).  `table` is a spin.ConstraintTable
        already uploaded (`table.upload(device)`); d_Exx = None evaluates J and c only; B > 1: a batch of trajectories
        (line-search candidates with shared multipliers, or independent problems dual_stride apart)."""
        raw = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        self._ck(self._lib.rbq_al_cost_expansion_dev(self._h, int(n), int(m), int(N), int(B), _ptr(d_X), _ptr(d_U), _ptr(d_Qdiag),
                                                     _ptr(d_Rdiag), _ptr(d_Qfdiag), _ptr(d_xf), _ptr(d_w), int(table.ncon),
                                                     raw(table.d_cons), raw(table.d_idx), raw(table.d_val), _ptr(d_lambda),
                                                     _ptr(d_mu), int(dual_stride), _ptr(d_Exx), _ptr(d_Ex), _ptr(d_Euu), _ptr(d_Eux), _ptr(d_Eu),
                                                     _ptr(d_c), _ptr(d_J)), "rbq_al_cost_expansion_dev")

    # ---- device-pointer forms (torch CUDA tensors; asynchronous on the handle's stream) ------------------------
    def stats_reset_dev(self, d_stats):
        self._ck(self._lib.rbq_stats_reset_dev(self._h, _ptr(d_stats)), "rbq_stats_reset_dev")

    def mc_static_dev(self, d_controls, nknots, dt, gate, gate_count, S, d_fq, d_da, d_psi0, algo, d_fid, d_stats):
        self._ck(self._lib.rbq_mc_static_dev(self._h, _ptr(d_controls), int(nknots), float(dt), int(gate), int(gate_count),
                                             int(S), _ptr(d_fq), _ptr(d_da), _ptr(d_psi0), int(algo), _ptr(d_fid),
                                             _ptr(d_stats)), "rbq_mc_static_dev")

    def mc_static_philox_dev(self, d_controls, nknots, dt, gate, gate_count, first, S, seed, fq0, fq_rel_sigma, fq_rel_clip,
                             da_sigma, algo, d_fid, d_stats):
        self._ck(self._lib.rbq_mc_static_philox_dev(self._h, _ptr(d_controls), int(nknots), float(dt), int(gate),
                                                    int(gate_count), int(first), int(S), int(seed), float(fq0),
                                                    float(fq_rel_sigma), float(fq_rel_clip), float(da_sigma), int(algo),
                                                    _ptr(d_fid), _ptr(d_stats)), "rbq_mc_static_philox_dev")

    def mc_noise_dev(self, d_controls, nknots, dt, gate, gate_count, S, fq, d_noise, noise_len, noise_dt_inv, d_psi0, algo,
                     d_fid, d_stats):
        self._ck(self._lib.rbq_mc_noise_dev(self._h, _ptr(d_controls), int(nknots), float(dt), int(gate), int(gate_count),
                                            int(S), float(fq), _ptr(d_noise), int(noise_len), float(noise_dt_inv),
                                            _ptr(d_psi0), int(algo), _ptr(d_fid), _ptr(d_stats)), "rbq_mc_noise_dev")

    def mc_pink_philox_dev(self, d_controls, nknots, dt, gate, gate_count, first, S, seed, fq, namp, noise_dt_inv, algo,
                           d_fid, d_stats):
        self._ck(self._lib.rbq_mc_pink_philox_dev(self._h, _ptr(d_controls), int(nknots), float(dt), int(gate),
                                                  int(gate_count), int(first), int(S), int(seed), float(fq), float(namp),
                                                  float(noise_dt_inv), int(algo), _ptr(d_fid), _ptr(d_stats)),
                 "rbq_mc_pink_philox_dev")

    def lindblad_dev(self, d_controls, nknots, dt, gate, gate_count, S, fq, d_da, d_rho0, channels, t2_gamma_c, t2_gamma_f,
                     d_fid, d_rho_out, d_stats):
        self._ck(self._lib.rbq_lindblad_dev(self._h, _ptr(d_controls), int(nknots), float(dt), int(gate), int(gate_count),
                                            int(S), float(fq), _ptr(d_da), _ptr(d_rho0), int(channels), float(t2_gamma_c),
                                            float(t2_gamma_f), _ptr(d_fid), _ptr(d_rho_out), _ptr(d_stats)),
                 "rbq_lindblad_dev")


class EngineGroup:
    """One process, several GPUs: an `rbq_group` (one handle per listed device) for sweeps that shard by global sample index.
    The other supported shape is one process per GPU with torch.distributed (rbqoc_b200/dist.py)."""

    def __init__(self, devices):
        self._lib = _lib.load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        g = C.c_void_p()
        rc = self._lib.rbq_group_create(devs, len(devices), C.byref(g))
        if rc:
            raise RbqError(rc, "rbq_group_create", "every member needs an sm_100 CUDA device; there is no CPU fallback")
        self._g = g
        self.devices = [int(d) for d in devices]

    def close(self):
        if getattr(self, "_g", None):
            self._lib.rbq_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mc_static_philox(self, controls, dt, gate, gate_count, first, S, seed, fq0, fq_rel_sigma, fq_rel_clip, da_sigma,
                         algo=_lib.ALGO_SU2, want_fid=False):
        controls = _f64(controls)
        fid = np.empty((S, gate_count + 1)) if want_fid else None
        st = Stats()
        rc = self._lib.rbq_group_mc_static_philox(self._g, _ptr(controls), controls.shape[0], float(dt), int(gate), int(gate_count),
                                                  int(first), int(S), int(seed), float(fq0), float(fq_rel_sigma), float(fq_rel_clip),
                                                  float(da_sigma), int(algo), _ptr(fid), C.byref(st))
        if rc:
            raise RbqError(rc, "rbq_group_mc_static_philox", self._lib.rbq_group_last_error(self._g).decode(errors="replace"))
        return fid, st


STATS_INT64_WORDS = C.sizeof(Stats) // 8   # a device rbq_stats is a torch.int64 tensor of this many words


def stats_from_words(words) -> Stats:
    """Reinterpret STATS_INT64_WORDS int64 words (e.g. a torch tensor copied to the host) as Stats."""
    buf = np.ascontiguousarray(np.asarray(words, dtype=np.int64))
    st = Stats()
    C.memmove(C.byref(st), buf.ctypes.data, C.sizeof(Stats))
    return st


def stats_merge(a: Stats, b: Stats) -> Stats:
    _lib.load().rbq_stats_merge(C.byref(a), C.byref(b))
    return a
