// rbq_math.cuh — device arithmetic shared by every kernel of the engine (sm_100a only).
//
//  * closed-form iso propagator: for G = dt*(f*NEGI_H0_ISO + a*NEGI_H1_ISO) (spin.jl:234-238)
//    G^2 = -theta^2 I with theta^2 = (pi dt)^2 (f^2 + a^2), hence
//        exp(G) psi = cos(theta) psi + sinc(theta) G psi
//    cos and sinc are even entire functions of theta, evaluated as Taylor polynomials in
//    x = theta^2 (no sqrt, no division); the degree is picked from a bound on x.
//  * hand-written forward-mode tangents (Tan<NT>) replacing ForwardDiff duals
//  * the reference algorithm itself (generic 4x4 Pade + LU, StaticArrays `exp`) for the
//    RBQ_ALGO_PADE kernels
//  * Philox4x32-10, T1(flux) interpolation, fidelity helpers
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/rbq.h"

#define RBQ_PI 3.14159265358979323846264338327950288
#define RBQ_DEV __device__ __forceinline__

namespace rbq {

// ---------------------------------------------------------------------------------------------
// Tangent numbers: value + NT directional derivatives.  One lane of a warp carries NT columns of
// the Jacobian, so the register cost is (1+NT) doubles per live scalar.
// Conventions copied from ForwardDiff 0.10.23 where they are observable: abs'(+0.0)=+1,
// abs'(-0.0)=-1; comparisons look at the value only.
// ---------------------------------------------------------------------------------------------
template <int NT> struct Tan {
    double v;
    double d[NT];
    RBQ_DEV Tan() {}
    RBQ_DEV Tan(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < NT; ++i) d[i] = 0.0;
    }
};
template <int NT> RBQ_DEV Tan<NT> operator+(const Tan<NT>& a, const Tan<NT>& b) {
    Tan<NT> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int NT> RBQ_DEV Tan<NT> operator-(const Tan<NT>& a, const Tan<NT>& b) {
    Tan<NT> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int NT> RBQ_DEV Tan<NT> operator-(const Tan<NT>& a) {
    Tan<NT> r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int NT> RBQ_DEV Tan<NT> operator*(const Tan<NT>& a, const Tan<NT>& b) {
    Tan<NT> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = fma(a.d[i], b.v, a.v * b.d[i]);
    return r;
}
template <int NT> RBQ_DEV Tan<NT> operator/(const Tan<NT>& a, const Tan<NT>& b) {
    Tan<NT> r; double ib = 1.0 / b.v; r.v = a.v * ib;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
template <int NT> RBQ_DEV Tan<NT> operator+(const Tan<NT>& a, double b) { Tan<NT> r = a; r.v += b; return r; }
template <int NT> RBQ_DEV Tan<NT> operator+(double a, const Tan<NT>& b) { return b + a; }
template <int NT> RBQ_DEV Tan<NT> operator-(const Tan<NT>& a, double b) { Tan<NT> r = a; r.v -= b; return r; }
template <int NT> RBQ_DEV Tan<NT> operator-(double a, const Tan<NT>& b) { return (-b) + a; }
template <int NT> RBQ_DEV Tan<NT> operator*(const Tan<NT>& a, double b) {
    Tan<NT> r; r.v = a.v * b;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] * b;
    return r;
}
template <int NT> RBQ_DEV Tan<NT> operator*(double a, const Tan<NT>& b) { return b * a; }
template <int NT> RBQ_DEV Tan<NT> operator/(double a, const Tan<NT>& b) { return Tan<NT>(a) / b; }
template <int NT> RBQ_DEV Tan<NT> tsqrt(const Tan<NT>& a) {
    Tan<NT> r; r.v = sqrt(a.v); double s = 0.5 / r.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] * s;
    return r;
}
RBQ_DEV double tsqrt(double a) { return sqrt(a); }
template <int NT> RBQ_DEV Tan<NT> tabs(const Tan<NT>& a) { return signbit(a.v) ? -a : a; }
RBQ_DEV double tabs(double a) { return fabs(a); }
template <int NT> RBQ_DEV double val(const Tan<NT>& a) { return a.v; }
RBQ_DEV double val(double a) { return a; }
// fused multiply-add spelled once so both scalar types share the model code
RBQ_DEV double tfma(double a, double b, double c) { return fma(a, b, c); }
template <int NT> RBQ_DEV Tan<NT> tfma(const Tan<NT>& a, const Tan<NT>& b, const Tan<NT>& c) { return a * b + c; }
template <int NT> RBQ_DEV Tan<NT> tfma(double a, const Tan<NT>& b, const Tan<NT>& c) { return b * a + c; }

// ---------------------------------------------------------------------------------------------
// cos(sqrt(x)) and sin(sqrt(x))/sqrt(x) as degree-K polynomials in x (Taylor, Horner).
// Truncation below 1e-17 relative for  x <= XMAX[K]:
//   K=3: 8.0e-4   K=4: 8.0e-3   K=5: 4.0e-2   K=6: 1.3e-1   K=7: 3.4e-1   K=8: 7.3e-1
// (dt=0.01 grid: x <= 2.5e-4 -> K=3;  dt=0.1 grid: x <= 2.5e-2 -> K=5.)
// ---------------------------------------------------------------------------------------------
// 1/n! correctly rounded: n! is exactly representable in binary64 for every n used here (<= 21)
RBQ_DEV constexpr double inv_fact(int n) { double f = 1.0; for (int i = 2; i <= n; ++i) f *= (double)i; return 1.0 / f; }

template <int K> RBQ_DEV void su2_series(double x, double& c, double& s) {
    // alternating signs folded into the evaluation: cos = sum (-x)^k/(2k)!
    const double mx = -x;
    double pc = inv_fact(2 * K);
    double ps = inv_fact(2 * K + 1);
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        pc = fma(pc, mx, inv_fact(2 * k));
        ps = fma(ps, mx, inv_fact(2 * k + 1));
    }
    c = pc;
    s = ps;
}
template <int K, int NT> RBQ_DEV void su2_series(const Tan<NT>& x, Tan<NT>& c, Tan<NT>& s) {
    // value by Horner, derivative d/dx by Horner on the differentiated polynomial
    const double mx = -x.v;
    double pc = inv_fact(2 * K), ps = inv_fact(2 * K + 1);
    double dc = 0.0, ds = 0.0;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        dc = fma(dc, mx, pc);
        ds = fma(ds, mx, ps);
        pc = fma(pc, mx, inv_fact(2 * k));
        ps = fma(ps, mx, inv_fact(2 * k + 1));
    }
    c.v = pc; s.v = ps;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c.d[i] = -dc * x.d[i]; s.d[i] = -ds * x.d[i]; }
}
// any x >= 0: the shortest series whose truncation is below 1e-17 for this x, closed form above 0.5 (used by the
// trajectory kernels, whose dt may be a decision variable; the sweep kernels pick a class per warp instead)
RBQ_DEV void su2_general(double x, double& c, double& s) {
    if (x <= 8.0e-4) { su2_series<3>(x, c, s); return; }
    if (x <= 4.0e-2) { su2_series<5>(x, c, s); return; }
    if (x < 0.5) { su2_series<8>(x, c, s); return; }
    double th = sqrt(x), sn, cs;
    sincos(th, &sn, &cs);
    c = cs; s = sn / th;
}
template <int NT> RBQ_DEV void su2_general(const Tan<NT>& x, Tan<NT>& c, Tan<NT>& s) {
    if (x.v <= 8.0e-4) { su2_series<4, NT>(x, c, s); return; }   // one more term than the value needs: d/dx drops a power
    if (x.v <= 4.0e-2) { su2_series<6, NT>(x, c, s); return; }
    if (x.v < 0.5) { su2_series<9, NT>(x, c, s); return; }
    double th = sqrt(x.v), sn, cs;
    sincos(th, &sn, &cs);
    c.v = cs; s.v = sn / th;
    const double dc = -0.5 * s.v;                 // d cos(sqrt x)/dx
    const double ds = 0.5 * (cs - s.v) / x.v;     // d sinc(sqrt x)/dx
#pragma unroll
    for (int i = 0; i < NT; ++i) { c.d[i] = dc * x.d[i]; s.d[i] = ds * x.d[i]; }
}

// E = c*I + sp*N0 + sq*N1 with N0, N1 the signed permutations of spin.jl:234,236 (pi folded in)
template <class T> struct Prop { T c, sp, sq; };

template <class T> RBQ_DEV void prop_apply(const Prop<T>& E, const T& a0, const T& a1, const T& a2, const T& a3,
                                           T& o0, T& o1, T& o2, T& o3) {
    // rows of G psi: ( p psi3 + q psi4, -p psi4 + q psi3, -p psi1 - q psi2, p psi2 - q psi1 )
    o0 = tfma(E.sq, a3, tfma(E.sp, a2, E.c * a0));
    o1 = tfma(E.sq, a2, tfma(-E.sp, a3, E.c * a1));
    o2 = tfma(-E.sq, a1, tfma(-E.sp, a0, E.c * a2));
    o3 = tfma(-E.sq, a0, tfma(E.sp, a1, E.c * a3));
}
RBQ_DEV void prop_apply(const Prop<double>& E, double* psi) {
    double o0, o1, o2, o3;
    prop_apply<double>(E, psi[0], psi[1], psi[2], psi[3], o0, o1, o2, o3);
    psi[0] = o0; psi[1] = o1; psi[2] = o2; psi[3] = o3;
}
// G psi / dt for H0 or H1 alone (Lawson-Euler source terms of spin11.jl:85, spin17.jl dstate1)
template <class T> RBQ_DEV void apply_N0(const T& a0, const T& a1, const T& a2, const T& a3, T& o0, T& o1, T& o2, T& o3) {
    o0 = a2; o1 = -a3; o2 = -a0; o3 = a1;
}
template <class T> RBQ_DEV void apply_N1(const T& a0, const T& a1, const T& a2, const T& a3, T& o0, T& o1, T& o2, T& o3) {
    o0 = a3; o1 = a2; o2 = -a1; o3 = -a0;
}

// ---------------------------------------------------------------------------------------------
// Reference algorithm: exp(::SMatrix{4,4,Float64}) of StaticArrays 0.12.5 (Higham 2005 Pade,
// thresholds 0.015/0.25/0.95/2.1, degree 13 + scaling/squaring above), `\` by LU with partial
// pivoting.  Row-major 16 doubles, everything unrolled into registers.
// ---------------------------------------------------------------------------------------------
RBQ_DEV void mm4(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = A[4 * i] * B[j];
#pragma unroll
            for (int k = 1; k < 4; ++k) s = fma(A[4 * i + k], B[4 * k + j], s);
            C[4 * i + j] = s;
        }
}
RBQ_DEV void solve4(double* A, double* B, double* X) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {   // bring the largest |A[r][c]| to row c by pairwise swaps
            if (fabs(A[4 * r + c]) > fabs(A[4 * c + c])) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double t = A[4 * c + j]; A[4 * c + j] = A[4 * r + j]; A[4 * r + j] = t;
                    t = B[4 * c + j]; B[4 * c + j] = B[4 * r + j]; B[4 * r + j] = t;
                }
            }
        }
        const double ip = 1.0 / A[4 * c + c];
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
            const double f = A[4 * r + c] * ip;
#pragma unroll
            for (int j = c + 1; j < 4; ++j) A[4 * r + j] = fma(-f, A[4 * c + j], A[4 * r + j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) B[4 * r + j] = fma(-f, B[4 * c + j], B[4 * r + j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int r = 3; r >= 0; --r) {
            double s = B[4 * r + j];
#pragma unroll
            for (int k = r + 1; k < 4; ++k) s = fma(-A[4 * r + k], X[4 * k + j], s);
            X[4 * r + j] = s / A[4 * r + r];
        }
}
// U = A*(sum cu_k A2^k), V = sum cv_k A2^k, Horner in A2 like Julia's @evalpoly
template <int NC> RBQ_DEV void pade_uv(const double* A, const double* A2, const double (&cu)[NC], const double (&cv)[NC],
                                        double* U, double* V) {
    double Pu[16], Pv[16], T[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { Pu[i] = (i % 5 == 0) ? cu[NC - 1] : 0.0; Pv[i] = (i % 5 == 0) ? cv[NC - 1] : 0.0; }
#pragma unroll
    for (int k = NC - 2; k >= 0; --k) {
        mm4(A2, Pu, T);
#pragma unroll
        for (int i = 0; i < 16; ++i) Pu[i] = T[i] + ((i % 5 == 0) ? cu[k] : 0.0);
        mm4(A2, Pv, T);
#pragma unroll
        for (int i = 0; i < 16; ++i) Pv[i] = T[i] + ((i % 5 == 0) ? cv[k] : 0.0);
    }
    mm4(A, Pu, U);
#pragma unroll
    for (int i = 0; i < 16; ++i) V[i] = Pv[i];
}
static __device__ __noinline__ void expm4_pade13(const double* Ain, double* E) {
    double A[16];
    double nA = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) nA = fmax(nA, fabs(Ain[j]) + fabs(Ain[4 + j]) + fabs(Ain[8 + j]) + fabs(Ain[12 + j]));
    int si = 0;
    double s = log2(nA / 5.4);
    if (s > 0) si = (int)ceil(s);
    const double sc = ldexp(1.0, -si);
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] = Ain[i] * sc;
    const double C[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                          129060195264000.,   10559470521600.,    670442572800.,     33522128640.,
                          1323241920.,        40840800.,          960960.,           16380.,
                          182.,               1.};
    double A2[16], A4[16], A6[16], t1[16], t2[16], u[16], v[16], w[16];
    mm4(A, A, A2); mm4(A2, A2, A4); mm4(A2, A4, A6);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const double id = (i % 5 == 0) ? 1.0 : 0.0;
        t1[i] = C[13] * A6[i] + C[11] * A4[i] + C[9] * A2[i];
        t2[i] = C[12] * A6[i] + C[10] * A4[i] + C[8] * A2[i];
        u[i] = C[7] * A6[i] + C[5] * A4[i] + C[3] * A2[i] + C[1] * id;
        v[i] = C[6] * A6[i] + C[4] * A4[i] + C[2] * A2[i] + C[0] * id;
    }
    mm4(A6, t1, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) u[i] += w[i];
    mm4(A6, t2, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += w[i];
    mm4(A, u, w);  // U
    double P[16], Q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { P[i] = v[i] + w[i]; Q[i] = v[i] - w[i]; }
    solve4(Q, P, E);
    for (int t = 0; t < si; ++t) { mm4(E, E, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) E[i] = w[i]; }
}
RBQ_DEV void expm4_pade(const double* A, double* E) {
    double nA = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) nA = fmax(nA, fabs(A[j]) + fabs(A[4 + j]) + fabs(A[8 + j]) + fabs(A[12 + j]));
    if (nA > 2.1) { expm4_pade13(A, E); return; }
    double A2[16], U[16], V[16];
    mm4(A, A, A2);
    if (nA > 0.95) {
        const double cu[5] = {8821612800., 302702400., 2162160., 3960., 1.};
        const double cv[5] = {17643225600., 2075673600., 30270240., 110880., 90.};
        pade_uv<5>(A, A2, cu, cv, U, V);
    } else if (nA > 0.25) {
        const double cu[4] = {8648640., 277200., 1512., 1.};
        const double cv[4] = {17297280., 1995840., 25200., 56.};
        pade_uv<4>(A, A2, cu, cv, U, V);
    } else if (nA > 0.015) {
        const double cu[3] = {15120., 420., 1.};
        const double cv[3] = {30240., 3360., 30.};
        pade_uv<3>(A, A2, cu, cv, U, V);
    } else {
        const double cu[2] = {60., 1.};
        const double cv[2] = {120., 12.};
        pade_uv<2>(A, A2, cu, cv, U, V);
    }
    double P[16], Q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { P[i] = V[i] + U[i]; Q[i] = V[i] - U[i]; }
    solve4(Q, P, E);
}
// dense generator dt*(f*pi*N0 + a*pi*N1), row-major
RBQ_DEV void generator4(double p, double q, double* G) {
#pragma unroll
    for (int i = 0; i < 16; ++i) G[i] = 0.0;
    G[2] = p;  G[3] = q;
    G[6] = q;  G[7] = -p;
    G[8] = -p; G[9] = -q;
    G[12] = -q; G[13] = p;
}
RBQ_DEV void mv4(const double* E, double* psi) {
    double o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = fma(E[4 * i + 3], psi[3], fma(E[4 * i + 2], psi[2], fma(E[4 * i + 1], psi[1], E[4 * i] * psi[0])));
#pragma unroll
    for (int i = 0; i < 4; ++i) psi[i] = o[i];
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10; stream layout (counter = index_lo, index_hi, j, stream; key = seed) as documented in include/rbq.h
// ---------------------------------------------------------------------------------------------
RBQ_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
RBQ_DEV double u53(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
RBQ_DEV void philox_uniform2(uint64_t index, uint64_t seed, uint32_t j, uint32_t stream, double& u0, double& u1) {
    uint32_t o[4];
    philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), j, stream, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    u0 = u53(o[0], o[1]);
    u1 = u53(o[2], o[3]);
}
// 1/sqrt(x) for a normal, positive x without the library routine's call, special-case tests and second Newton step: the
// hardware seed (MUFU.RSQ64H, relative error below 2^-22) and one third-order correction
// y <- y + y e (1/2 + 3/8 e), e = 1 - x y^2, which leaves ~2^-66 before the final rounding.  Five dependent FP64 instructions
// after the seed; used where a reciprocal square root sits on a knot-to-knot dependency chain (unscented rollouts).
RBQ_DEV double rsqrt_lean(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}
RBQ_DEV void philox_normal2(uint64_t index, uint64_t seed, uint32_t j, uint32_t stream, double& z0, double& z1) {
    double u0, u1;
    philox_uniform2(index, seed, j, stream, u0, u1);
    const double r = sqrt(-2.0 * log(1.0 - u0));
    double sn, cs;
    sincos(2.0 * RBQ_PI * u1, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}
// gen_rand_state_iso recipe (spin.jl:1132-1136)
RBQ_DEV void sample_psi0(uint64_t index, uint64_t seed, double* psi) {
    philox_uniform2(index, seed, 0, 0, psi[0], psi[1]);
    philox_uniform2(index, seed, 1, 0, psi[2], psi[3]);
    const double n = sqrt(psi[0] * psi[0] + psi[1] * psi[1] + psi[2] * psi[2] + psi[3] * psi[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) psi[i] /= n;
}
RBQ_DEV void sample_static(uint64_t index, uint64_t seed, double fq0, double fq_rel_sigma, double fq_rel_clip,
                           double da_sigma, double& fq, double& da) {
    double z0, z1;
    philox_normal2(index, seed, 0, 1, z0, z1);
    double dev = fq_rel_sigma * z0;
    dev = fmin(fmax(dev, -fq_rel_clip), fq_rel_clip);
    fq = fq0 * (1.0 + dev);
    da = da_sigma * z1;
}

// ---------------------------------------------------------------------------------------------
// T1(flux): gridded-linear interpolation with flat extrapolation (spin.jl:192-206, 366, 393-396)
// ---------------------------------------------------------------------------------------------
static __device__ __constant__ double kT1us[18] = {1597.923, 1627.93, 301.86,  269.03,   476.33,   1783.19,
                                            2131.76,  2634.50, 4364.68, 2587.82,  1661.915, 1794.468,
                                            2173.88,  1188.83, 1576.493, 965.183, 560.251,  310.88};
static __device__ __constant__ double kFbfq[18] = {0.26, 0.28, 0.32,  0.34, 0.36,  0.38,  0.4,   0.42,  0.44,
                                            0.46, 0.465, 0.47, 0.475, 0.48, 0.484, 0.488, 0.492, 0.5};
template <class T> RBQ_DEV T t1_interp(const T& amplitude, double scale) {
    T x = -tabs(amplitude) * 0.202407 + 0.5;      // amp_fbfq_lo, spin.jl:366
    if (val(x) > kFbfq[17]) x = T(kFbfq[17]);     // Flat(): a clamped argument loses its tangent
    else if (val(x) < kFbfq[0]) x = T(kFbfq[0]);
    int i = 0;
    while (i < 16 && kFbfq[i + 1] <= val(x)) ++i;
    const double x0 = kFbfq[i], h = kFbfq[i + 1] - kFbfq[i];
    T w = (x - x0) * (1.0 / h);
    // (1-w) y_i + w y_{i+1}; written so both scalar types share it
    return (1.0 - w) * (kT1us[i] * scale) + w * (kT1us[i + 1] * scale);
}

// fidelity_vec_iso2, spin.jl:1035-1037
RBQ_DEV double fidelity_vec_iso2(const double* s1, const double* s2) {
    const double a = fma(s1[3], s2[3], fma(s1[2], s2[2], fma(s1[1], s2[1], s1[0] * s2[0])));
    const double b = fma(-s1[3], s2[1], fma(-s1[2], s2[0], fma(s1[1], s2[3], s1[0] * s2[2])));
    return fma(a, a, b * b);
}

}  // namespace rbq
