// rbq_models.cuh — the augmented-state step maps of the reference's `Model` types, written once
// over a scalar type T (double, or Tan<NT> for the Jacobian kernels) and an IO policy that says
// where x, u come from and where x' goes (global trajectory arrays, or the [A_k B_k] block).
//
// Reference: RD.discrete_dynamics(::Type{RK3}, model, x, u, t, dt) of
//   spin13.jl:44-58, spin12.jl:165-192, spin18.jl:162-190, spin15.jl:50-73, spin11.jl:64-117,
//   spin17.jl:61-114, spin30.jl:86-202.
// The propagator is the closed-form SU(2) expression of rbq_math.cuh, not a generic exp.
#pragma once
#include "rbq_math.cuh"

namespace rbq {

// dimensions (RD.state_dim / RD.control_dim): spin13.jl:39-40, spin12.jl:26-33, spin15.jl:43-46,
// spin11.jl:53-55, spin18.jl:23-28, spin30.jl:29-32,67-70
__host__ __device__ inline int model_state_dim(const rbq_model_params& p) {
    switch (p.model_id) {
    case RBQ_MODEL_SPIN13: return 11;
    case RBQ_MODEL_SPIN12:
    case RBQ_MODEL_SPIN18: return 43;
    case RBQ_MODEL_SPIN15: return 11 + (p.decay_aware ? 1 : 0);
    case RBQ_MODEL_SPIN11:
    case RBQ_MODEL_SPIN17:
        if (p.derivative_order < 0 || p.derivative_order > 2) return RBQ_E_MODEL;
        return 11 + 4 * p.derivative_order;
    case RBQ_MODEL_SPIN30:
        if (p.sample_state_count < 1 || p.sample_state_count > RBQ_MAX_SAMPLE_STATES) return RBQ_E_MODEL;
        return 15 + 41 * p.sample_state_count;
    case RBQ_MODEL_SPIN25:   // spin25.jl:28-33,73-75: 2 states + 3 controls + 6 filter taps + 41 per sample state
        if (p.sample_state_count < 1 || p.sample_state_count > RBQ_MAX_SAMPLE_STATES) return RBQ_E_MODEL;
        return 17 + 41 * p.sample_state_count;
    default: return RBQ_E_MODEL;
    }
}
__host__ __device__ inline int model_control_dim(const rbq_model_params& p) {
    if (model_state_dim(p) < 0) return RBQ_E_MODEL;
    if (p.integrator != RBQ_INTEGRATOR_EXP) {   // RK steps exist for the models with a continuous RD.dynamics in the reference
        const bool known = p.integrator == RBQ_INTEGRATOR_RK2 || p.integrator == RBQ_INTEGRATOR_RK3 || p.integrator == RBQ_INTEGRATOR_RK4;
        const bool fam = p.model_id == RBQ_MODEL_SPIN13 || p.model_id == RBQ_MODEL_SPIN15 || p.model_id == RBQ_MODEL_SPIN11 || p.model_id == RBQ_MODEL_SPIN17;
        if (!known || !fam) return RBQ_E_MODEL;
    }
    if (p.model_id == RBQ_MODEL_SPIN15) return 1 + (p.time_optimal ? 1 : 0);
    return 1;
}

// psi' = E psi for the 4-block starting at `at`
template <class T, class IO> RBQ_DEV void advance_block(const Prop<T>& E, IO& io, int at) {
    const T a0 = io.x(at), a1 = io.x(at + 1), a2 = io.x(at + 2), a3 = io.x(at + 3);
    T o0, o1, o2, o3;
    prop_apply<T>(E, a0, a1, a2, a3, o0, o1, o2, o3);
    io.out(at, o0); io.out(at + 1, o1); io.out(at + 2, o2); io.out(at + 3, o3);
}
// [int a, a, da] Euler chain driven by u = d2a (spin13.jl:49-51)
template <class T, class DT, class IO> RBQ_DEV void control_chain(IO& io, int base, const DT& dt) {
    const T ia = io.x(base), a = io.x(base + 1), da = io.x(base + 2), u = io.u(0);
    io.out(base, ia + a * dt);
    io.out(base + 1, a + da * dt);
    io.out(base + 2, da + u * dt);
}
template <class T, class DT> RBQ_DEV Prop<T> prop_of(double f, const T& a, const DT& dt) {
    const DT w = RBQ_PI * dt;
    const T p = T(w * f), q = a * w;
    const T x = p * p + q * q;
    T c, s;
    su2_general(x, c, s);
    Prop<T> E; E.c = c; E.sp = s * p; E.sq = s * q;
    return E;
}

// ---- spin13 -----------------------------------------------------------------------------------
template <class T, class IO> RBQ_DEV void step_spin13(const rbq_model_params& p, IO& io, double dt) {
    const Prop<T> E = prop_of<T, double>(p.fq, io.x(9), dt);
    advance_block<T>(E, io, 0);
    advance_block<T>(E, io, 4);
    control_chain<T, double>(io, 8, dt);
}
// ---- spin12 (fq +- sigma) and spin18 (a +- namp) ------------------------------------------------
template <class T, class IO> RBQ_DEV void step_spin12_18(const rbq_model_params& p, IO& io, double dt) {
    const T a = io.x(9);
    const Prop<T> E = prop_of<T, double>(p.fq, a, dt);
    advance_block<T>(E, io, 0);
    advance_block<T>(E, io, 4);
    control_chain<T, double>(io, 8, dt);
    Prop<T> Ep, En;
    if (p.model_id == RBQ_MODEL_SPIN12) {
        Ep = prop_of<T, double>(p.fq_samples[0], a, dt);
        En = prop_of<T, double>(p.fq_samples[1], a, dt);
    } else {
        Ep = prop_of<T, double>(p.fq, a + p.namp, dt);
        En = prop_of<T, double>(p.fq, a - p.namp, dt);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) advance_block<T>(Ep, io, 11 + 4 * s);
#pragma unroll
    for (int s = 4; s < 8; ++s) advance_block<T>(En, io, 11 + 4 * s);
}
// ---- spin15: dt = u2^2 when time optimal, int gamma_1 state when decay aware ---------------------
template <class T, class DT, class IO> RBQ_DEV void step_spin15_dt(const rbq_model_params& p, IO& io, const DT& dt) {
    const T a = io.x(9);
    const Prop<T> E = prop_of<T, DT>(p.fq, a, dt);
    advance_block<T>(E, io, 0);
    advance_block<T>(E, io, 4);
    control_chain<T, DT>(io, 8, dt);
    if (p.decay_aware) {
        const T t1 = t1_interp<T>(a, 1.0);   // microseconds, spin15.jl:66-69 / spin.jl:396
        io.out(11, io.x(11) + (1.0 / t1) * dt);
    }
}
template <class T, class IO> RBQ_DEV void step_spin15(const rbq_model_params& p, IO& io, double dt) {
    if (p.time_optimal) {
        const T u2 = io.u(1);
        step_spin15_dt<T, T>(p, io, u2 * u2);
    } else {
        step_spin15_dt<T, double>(p, io, dt);
    }
}
// ---- spin11 (d/dfq, H0 source) and spin17 (d/da, H1 source): Lawson-Euler --------------------------
template <class T, class IO> RBQ_DEV void step_spin11_17(const rbq_model_params& p, IO& io, double dt) {
    const Prop<T> E = prop_of<T, double>(p.fq, io.x(9), dt);
    advance_block<T>(E, io, 0);
    advance_block<T>(E, io, 4);
    control_chain<T, double>(io, 8, dt);
    const bool h1 = p.model_id == RBQ_MODEL_SPIN17;
    const double w = RBQ_PI * dt;
#pragma unroll
    for (int o = 1; o <= 2; ++o) {
        if (p.derivative_order < o) break;
        // d^o psi' = E (d^o psi + o dt H d^(o-1) psi), source state: psi1 (o=1), d psi (o=2)
        const int src = (o == 1) ? 0 : 11, dst = 11 + 4 * (o - 1);
        const T s0 = io.x(src), s1 = io.x(src + 1), s2 = io.x(src + 2), s3 = io.x(src + 3);
        T g0, g1, g2, g3;
        if (h1) apply_N1<T>(s0, s1, s2, s3, g0, g1, g2, g3);
        else apply_N0<T>(s0, s1, s2, s3, g0, g1, g2, g3);
        const double f = w * (double)o;
        const T w0 = io.x(dst) + g0 * f, w1 = io.x(dst + 1) + g1 * f, w2 = io.x(dst + 2) + g2 * f,
                w3 = io.x(dst + 3) + g3 * f;
        T o0, o1, o2, o3;
        prop_apply<T>(E, w0, w1, w2, w3, o0, o1, o2, o3);
        io.out(dst, o0); io.out(dst + 1, o1); io.out(dst + 2, o2); io.out(dst + 3, o3);
    }
}
// ---- unscented transform of one sample-state chunk (spin30.jl:86-174, spin25.jl:91-178) --------------------------
// 10 sigma points at `off`, propagated with E except points 5 and 10 (Ep / En), then mean, covariance, Cholesky,
// resampling, normalisation and the penalty against the nominal block `nom` of the INPUT state.
template <class T, class IO>
RBQ_DEV void unscented_chunk(const rbq_model_params& p, IO& io, int off, int nom, const Prop<T>& E, const Prop<T>& Ep,
                             const Prop<T>& En) {
    const double cf = 1.0 / (2.0 * p.alpha * p.alpha);
    T s[10][4];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const Prop<T>& M = (j == 4) ? Ep : (j == 9) ? En : E;
        prop_apply<T>(M, io.x(off + 4 * j), io.x(off + 4 * j + 1), io.x(off + 4 * j + 2), io.x(off + 4 * j + 3),
                      s[j][0], s[j][1], s[j][2], s[j][3]);
    }
    T sm[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T acc = s[0][i];
#pragma unroll
        for (int j = 1; j < 10; ++j) acc = acc + s[j][i];
        sm[i] = acc * 0.1;
    }
#pragma unroll
    for (int j = 0; j < 10; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[j][i] = s[j][i] - sm[i];
    // lower triangle of the 4x4 sample covariance (row-major packed: 00 10 11 20 21 22 30 31 32 33)
    T P[10];
    {
        int q = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k <= i; ++k) {
                T acc = s[0][i] * s[0][k];
#pragma unroll
                for (int j = 1; j < 10; ++j) acc = acc + s[j][i] * s[j][k];
                P[q++] = acc * cf;
            }
    }
    // Cholesky of blkdiag(P, parameter variance): the 5th row/column decouples exactly, so only the 4x4 factor is
    // needed (spin30.jl:136-140, spin25.jl:141-146).  L packed like P.
    T L[10];
#define RBQ_TRI(i, k) ((i) * ((i) + 1) / 2 + (k))
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T d = P[RBQ_TRI(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) d = d - L[RBQ_TRI(j, k)] * L[RBQ_TRI(j, k)];
        const T ljj = tsqrt(d);
        L[RBQ_TRI(j, j)] = ljj;
        const T inv = 1.0 / ljj;
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            T t = P[RBQ_TRI(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) t = t - L[RBQ_TRI(i, k)] * L[RBQ_TRI(j, k)];
            L[RBQ_TRI(i, j)] = t * inv;
        }
    }
    // resample sm +- alpha L[:,j], s5 = s10 = sm, normalise each
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        T vp[4], vm[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j < 4 && i >= j) {
                const T col = L[RBQ_TRI(i, j)] * p.alpha;
                vp[i] = sm[i] + col; vm[i] = sm[i] - col;
            } else { vp[i] = sm[i]; vm[i] = sm[i]; }
        }
        const T np = 1.0 / tsqrt(vp[0] * vp[0] + vp[1] * vp[1] + vp[2] * vp[2] + vp[3] * vp[3]);
        const T nm = (j < 4) ? 1.0 / tsqrt(vm[0] * vm[0] + vm[1] * vm[1] + vm[2] * vm[2] + vm[3] * vm[3]) : np;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            io.out(off + 4 * j + i, vp[i] * np);
            io.out(off + 4 * (j + 5) + i, vm[i] * nm);
        }
    }
#undef RBQ_TRI
    // penalty = tr(P S) + (sm - psi_nom)' S (sm - psi_nom), psi_nom from the INPUT state
    T pen = P[0] * p.S_diag[0] + P[2] * p.S_diag[1] + P[5] * p.S_diag[2] + P[9] * p.S_diag[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T dn = sm[i] - io.x(nom + i);
        pen = pen + dn * dn * p.S_diag[i];
    }
    io.out(off + 40, pen);
}
// ---- spin30: unscented sampling in fq ----------------------------------------------------------------------------
template <class T, class IO> RBQ_DEV void step_spin30(const rbq_model_params& p, IO& io, double dt) {
    const T a = io.x(13);
    const Prop<T> E = prop_of<T, double>(p.fq, a, dt);
    advance_block<T>(E, io, 0);
    advance_block<T>(E, io, 4);
    advance_block<T>(E, io, 8);
    control_chain<T, double>(io, 12, dt);
    const double sig = sqrt(p.fq_cov);   // used as a deviation WITHOUT alpha, spin30.jl:102,108,113
    const Prop<T> Ep = prop_of<T, double>(p.fq + sig, a, dt);
    const Prop<T> En = prop_of<T, double>(p.fq - sig, a, dt);
    for (int c = 0; c < p.sample_state_count; ++c) unscented_chunk<T>(p, io, 15 + 41 * c, p.nominal_offset[c], E, Ep, En);
}
// knot index the reference derives from the time argument: Int(div(time, dt)) (spin25.jl:196); Julia's float `div` is
// round((x - rem(x, y)) / y)
__host__ __device__ inline long long knot_point_of(double time, double dt) {
    return (long long)rint((time - fmod(time, dt)) / dt);
}
// ---- spin25: unscented sampling in the amplitude, pink-filtered noise amplitude carried in the state ------------------
// state: psi1 0..3, psi2 4..7, [int a, a, da] 8..10, filter inputs xp1..3 11..13, outputs yp1..3 14..16, chunks at 17+41c
template <class T, class IO> RBQ_DEV void step_spin25(const rbq_model_params& p, IO& io, double time, double dt) {
    const T camp = io.x(9);
    const Prop<T> E = prop_of<T, double>(p.fq, camp, dt);
    advance_block<T>(E, io, 0);
    advance_block<T>(E, io, 4);
    control_chain<T, double>(io, 8, dt);
    // filter white noise (spin25.jl:194-219): x_k = sqrt(wnamp_cov) is a constant input, the taps are states
    const double xk = sqrt(p.fq_cov);    // `fq_cov` carries wnamp_cov for this model (include/rbq.h)
    const long long kp = knot_point_of(time, dt);
    const T xp1 = io.x(11), xp2 = io.x(12), xp3 = io.x(13), yp1 = io.x(14), yp2 = io.x(15), yp3 = io.x(16);
    T yk = T(0.049922035 * xk);
    if (kp == 1) {
    } else if (kp == 2) {
        yk = yk + xp1 * (-0.095993537) - yp1 * (-2.494956002);
    } else if (kp == 3) {
        yk = yk + xp1 * (-0.095993537) + xp2 * 0.050612699 - yp1 * (-2.494956002) - yp2 * 2.017265875;
    } else {
        yk = yk + xp1 * (-0.095993537) + xp2 * 0.050612699 + xp3 * (-0.004408786) - yp1 * (-2.494956002) - yp2 * 2.017265875 -
             yp3 * (-0.522189400);
    }
    io.out(11, T(xk)); io.out(12, xp1); io.out(13, xp2);
    io.out(14, yk); io.out(15, yp1); io.out(16, yp2);
    const T namp = yk * p.namp;
    const Prop<T> Ep = prop_of<T, double>(p.fq, camp + namp, dt);
    const Prop<T> En = prop_of<T, double>(p.fq, camp - namp, dt);
    for (int c = 0; c < p.sample_state_count; ++c) unscented_chunk<T>(p, io, 17 + 41 * c, p.nominal_offset[c], E, Ep, En);
}

// ---- explicit Runge-Kutta steps of the continuous dynamics (RBQ_INTEGRATOR_RK2/3/4, see include/rbq.h) ---------------------------
// RD.dynamics of the reference's RK variants (spin15_standalone.jl:263-312, spin19.jl:66-129, spin21.jl:63-109) on the layout of
// spin13 / spin15 / spin11 / spin17: psi1 0..3, psi2 4..7, [int a, a, da] 8..10, then int gamma (spin15 decay aware) or the
// derivative states of psi1 (4 per order).
constexpr int RK_MAX_N = 19;
template <class T> RBQ_DEV void apply_H(double wf, const T& wa, const T* x, T* y) {
    // (fq H0 + a H1) x with the pi folded into wf = pi fq, wa = pi a: rows of prop_apply without the cosine term
    y[0] = tfma(wa, x[3], x[2] * wf);
    y[1] = tfma(wa, x[2], x[3] * (-wf));
    y[2] = tfma(-wa, x[1], x[0] * (-wf));
    y[3] = tfma(-wa, x[0], x[1] * wf);
}
template <class T> RBQ_DEV void continuous_dynamics(const rbq_model_params& p, int n, const T* x, const T* u, T* dx) {
    const T a = x[9];
    const double wf = RBQ_PI * p.fq;
    const T wa = a * RBQ_PI;
    apply_H<T>(wf, wa, x, dx);
    apply_H<T>(wf, wa, x + 4, dx + 4);
    dx[8] = a; dx[9] = x[10]; dx[10] = u[0];
    if (p.model_id == RBQ_MODEL_SPIN15) {
        if (p.decay_aware) dx[11] = 1.0 / t1_interp<T>(a, 1e3);      // amp_t1_spline: ns (spin15_standalone.jl:272, 300)
        if (p.time_optimal) {
            const T s = u[1] * u[1];
            for (int i = 0; i < n; ++i) dx[i] = dx[i] * s;
        }
    } else if (p.model_id == RBQ_MODEL_SPIN11 || p.model_id == RBQ_MODEL_SPIN17) {
        const bool h1 = p.model_id == RBQ_MODEL_SPIN17;
        for (int o = 1; o <= p.derivative_order && o <= 2; ++o) {
            const T* src = o == 1 ? x : x + 11;
            T* dst = dx + 11 + 4 * (o - 1);
            T g[4];
            if (h1) apply_N1<T>(src[0], src[1], src[2], src[3], g[0], g[1], g[2], g[3]);
            else apply_N0<T>(src[0], src[1], src[2], src[3], g[0], g[1], g[2], g[3]);
            apply_H<T>(wf, wa, x + 11 + 4 * (o - 1), dst);
            const double f = RBQ_PI * (double)o;
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = dst[i] + g[i] * f;
        }
    }
}
template <class T, class IO> RBQ_DEV void step_rk(const rbq_model_params& p, IO& io, int n, int m, double dt) {
    T x[RK_MAX_N], u[2], k1[RK_MAX_N], k2[RK_MAX_N], k3[RK_MAX_N], xt[RK_MAX_N];
    for (int i = 0; i < n; ++i) x[i] = io.x(i);
    for (int j = 0; j < m; ++j) u[j] = io.u(j);
    continuous_dynamics<T>(p, n, x, u, k1);
    if (p.integrator == RBQ_INTEGRATOR_RK2) {            // midpoint
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * (0.5 * dt);
        continuous_dynamics<T>(p, n, xt, u, k2);
        for (int i = 0; i < n; ++i) io.out(i, x[i] + k2[i] * dt);
    } else if (p.integrator == RBQ_INTEGRATOR_RK3) {     // k3 at x - k1 + 2 k2, weights (1, 4, 1) / 6
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * (0.5 * dt);
        continuous_dynamics<T>(p, n, xt, u, k2);
        for (int i = 0; i < n; ++i) xt[i] = x[i] - k1[i] * dt + k2[i] * (2.0 * dt);
        continuous_dynamics<T>(p, n, xt, u, k3);
        for (int i = 0; i < n; ++i) io.out(i, x[i] + (k1[i] + k2[i] * 4.0 + k3[i]) * (dt / 6.0));
    } else {                                             // classic RK4
        T k4[RK_MAX_N];
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * (0.5 * dt);
        continuous_dynamics<T>(p, n, xt, u, k2);
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k2[i] * (0.5 * dt);
        continuous_dynamics<T>(p, n, xt, u, k3);
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k3[i] * dt;
        continuous_dynamics<T>(p, n, xt, u, k4);
        for (int i = 0; i < n; ++i) io.out(i, x[i] + (k1[i] + k2[i] * 2.0 + k3[i] * 2.0 + k4[i]) * (dt / 6.0));
    }
}
__host__ __device__ inline bool model_supports_rk(const rbq_model_params& p) {
    return p.model_id == RBQ_MODEL_SPIN13 || p.model_id == RBQ_MODEL_SPIN15 || p.model_id == RBQ_MODEL_SPIN11 || p.model_id == RBQ_MODEL_SPIN17;
}

// MODEL is the reference file number (RBQ_MODEL_*); one kernel instantiation per model family.
// `time` is used by spin25 only (every other reference model ignores its time argument).
// RK: the Runge-Kutta steps are compiled into their own kernel instantiations only (inside the exponential-step kernels their
// mere presence cost 150 registers and a 4.6 KB stack frame)
template <int MODEL, class T, class IO, bool RK = false> RBQ_DEV void model_step(const rbq_model_params& p, IO& io, double time, double dt) {
    if constexpr (RK) {
        step_rk<T>(p, io, model_state_dim(p), model_control_dim(p), dt);
        return;
    }
    if (MODEL == RBQ_MODEL_SPIN13) step_spin13<T>(p, io, dt);
    else if (MODEL == RBQ_MODEL_SPIN12) step_spin12_18<T>(p, io, dt);   // also spin18
    else if (MODEL == RBQ_MODEL_SPIN15) step_spin15<T>(p, io, dt);
    else if (MODEL == RBQ_MODEL_SPIN11) step_spin11_17<T>(p, io, dt);   // also spin17
    else if (MODEL == RBQ_MODEL_SPIN25) step_spin25<T>(p, io, time, dt);
    else step_spin30<T>(p, io, dt);
}

}  // namespace rbq
