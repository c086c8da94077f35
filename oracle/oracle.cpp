// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference algorithm (SchusterLab/rbqoc) for the propagation hot
// path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (rbqoc_b200 + librbq.so) never does.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or stored pulses and its
// toolchain (Julia 1.5 + StaticArrays 0.12.5 + ForwardDiff 0.10.23 + RobotDynamics 0.2.1
// @SchusterLab/rbqoc-stable) is absent from this image, so this restatement is pinned only
// by (a) the analytic gates embedded in the reference (spin.jl:557,608-612), (b) scipy /
// mpmath cross-checks of the matrix exponential and (c) a 50-digit mpmath referee
// (oracle/referee.py).  See DESIGN.md "Oracle".
//
// Everything here follows the reference's *algorithm*: generic Higham-2005 Pade `exp` with
// StaticArrays' branch thresholds + LU solve, per-knot 4x4 propagator products, forward-mode
// dual numbers with ForwardDiff's conventions.  The engine under test uses a different
// algorithm (closed-form SU(2) series, lane-split tangents), which is what makes the
// comparison meaningful.  Citations are file:line in the reference tree.
#include "../include/rbq.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ------------------------------------------------------------------------------------------
// Dual numbers with ForwardDiff.Dual conventions (ForwardDiff 0.10.23, Manifest.toml:488-492).
// The partial count is a run-time value so one type serves every model (n+m <= 180).
// ------------------------------------------------------------------------------------------
constexpr int MAXP = 184;
static thread_local int g_np = 0;

struct Dual {
    double v;
    double d[MAXP];
    Dual() : v(0.0) { for (int i = 0; i < g_np; ++i) d[i] = 0.0; }
    Dual(double x) : v(x) { for (int i = 0; i < g_np; ++i) d[i] = 0.0; }
};
inline Dual operator+(const Dual& a, const Dual& b) { Dual r; r.v = a.v + b.v; for (int i = 0; i < g_np; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
inline Dual operator-(const Dual& a, const Dual& b) { Dual r; r.v = a.v - b.v; for (int i = 0; i < g_np; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
inline Dual operator-(const Dual& a) { Dual r; r.v = -a.v; for (int i = 0; i < g_np; ++i) r.d[i] = -a.d[i]; return r; }
inline Dual operator*(const Dual& a, const Dual& b) { Dual r; r.v = a.v * b.v; for (int i = 0; i < g_np; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
inline Dual operator/(const Dual& a, const Dual& b) { Dual r; r.v = a.v / b.v; double ib = 1.0 / b.v; for (int i = 0; i < g_np; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
inline Dual operator+(const Dual& a, double b) { Dual r = a; r.v += b; return r; }
inline Dual operator+(double a, const Dual& b) { return b + a; }
inline Dual operator-(const Dual& a, double b) { Dual r = a; r.v -= b; return r; }
inline Dual operator-(double a, const Dual& b) { return (-b) + a; }
inline Dual operator*(const Dual& a, double b) { Dual r; r.v = a.v * b; for (int i = 0; i < g_np; ++i) r.d[i] = a.d[i] * b; return r; }
inline Dual operator*(double a, const Dual& b) { return b * a; }
inline Dual operator/(const Dual& a, double b) { return a * (1.0 / b); }
inline Dual operator/(double a, const Dual& b) { return Dual(a) / b; }
inline Dual sqrt(const Dual& a) { Dual r; r.v = std::sqrt(a.v); double s = 0.5 / r.v; for (int i = 0; i < g_np; ++i) r.d[i] = a.d[i] * s; return r; }
// ForwardDiff: abs(d) = signbit(value(d)) ? -d : d   (derivative +1 at +0.0, -1 at -0.0)
inline Dual absd(const Dual& a) { return std::signbit(a.v) ? -a : a; }
inline double absd(double a) { return std::fabs(a); }
inline double value(double a) { return a; }
inline double value(const Dual& a) { return a.v; }
inline double value(const std::complex<double>& a) { return std::abs(a); }  // used only for norms
inline double magnitude(double a) { return std::fabs(a); }
inline double magnitude(const Dual& a) { return std::fabs(a.v); }
inline double magnitude(const std::complex<double>& a) { return std::abs(a); }
using std::sqrt;

// ------------------------------------------------------------------------------------------
// constants: spin.jl:122-125, 142-173, 176-206, 220-247, 301-332
// ------------------------------------------------------------------------------------------
constexpr double PI = 3.14159265358979323846264338327950288;
constexpr double FQ = 1.4e-2;                     // spin.jl:148
constexpr double FBFQ_A = 0.202407;               // spin.jl:157
constexpr double FBFQ_B = 0.5;                    // spin.jl:158
constexpr double AYPIBY2 = 1.25e-1;               // spin.jl:159
constexpr double PFIR_B1 = 0.049922035, PFIR_B2 = -0.095993537, PFIR_B3 = 0.050612699,
                 PFIR_B4 = -0.004408786;          // spin.jl:166-169
constexpr double PFIR_A2 = -2.494956002, PFIR_A3 = 2.017265875, PFIR_A4 = -0.522189400;  // 171-173
constexpr double TTOT_ZPIBY2 = 17.857142857142858;    // spin.jl:557
constexpr double TX_YPIBY2 = 2.1656249366575766;      // spin.jl:608
constexpr double TZ_YPIBY2 = 15.1423305995572655;     // spin.jl:609

// raw T1 times in microseconds (spin.jl:192-196); spin.jl:181-185 is the same table * 1e3 (ns)
static const double T1_ARRAY_REDUCED[18] = {1597.923, 1627.93, 301.86,  269.03,   476.33,   1783.19,
                                            2131.76,  2634.50, 4364.68, 2587.82,  1661.915, 1794.468,
                                            2173.88,  1188.83, 1576.493, 965.183, 560.251,  310.88};
static const double FBFQ_ARRAY[18] = {0.26, 0.28, 0.32,  0.34, 0.36,  0.38,  0.4,   0.42,  0.44,
                                      0.46, 0.465, 0.47, 0.475, 0.48, 0.484, 0.488, 0.492, 0.5};  // 197-201

// NEGI (spin.jl:220-223) times iso(sigma_z), iso(sigma_x)  (spin.jl:231-236), without the pi.
static const double N0[4][4] = {{0, 0, 1, 0}, {0, 0, 0, -1}, {-1, 0, 0, 0}, {0, 1, 0, 0}};
static const double N1[4][4] = {{0, 0, 0, 1}, {0, 0, 1, 0}, {0, -1, 0, 0}, {-1, 0, 0, 0}};

template <class T> struct Mat4 { T a[4][4]; };
template <class T> struct Vec4 { T a[4]; };

template <class T> Mat4<T> matmul(const Mat4<T>& A, const Mat4<T>& B) {
    Mat4<T> C;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            T s = A.a[i][0] * B.a[0][j];
            for (int k = 1; k < 4; ++k) s = s + A.a[i][k] * B.a[k][j];
            C.a[i][j] = s;
        }
    return C;
}
template <class T, class V> void matvec(const Mat4<T>& A, const V* x, V* y) {
    V t[4];
    for (int i = 0; i < 4; ++i) {
        V s = A.a[i][0] * x[0];
        for (int k = 1; k < 4; ++k) s = s + A.a[i][k] * x[k];
        t[i] = s;
    }
    for (int i = 0; i < 4; ++i) y[i] = t[i];
}
template <class T> Mat4<T> eye4() {
    Mat4<T> I;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) I.a[i][j] = T(i == j ? 1.0 : 0.0);
    return I;
}
// c0*I + c1*A2 + c2*A2^2 + ...   (Julia's @evalpoly on a matrix argument, Horner form)
template <class T> Mat4<T> evalpoly(const Mat4<T>& X, const double* c, int n) {
    Mat4<T> R;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) R.a[i][j] = T(i == j ? c[n - 1] : 0.0);
    for (int k = n - 2; k >= 0; --k) {
        R = matmul(X, R);
        for (int i = 0; i < 4; ++i) R.a[i][i] = R.a[i][i] + c[k];
    }
    return R;
}
// X = A \ B by LU with partial pivoting (what StaticArrays falls back to for 4x4).
template <class T> Mat4<T> solve(Mat4<T> A, Mat4<T> B) {
    for (int c = 0; c < 4; ++c) {
        int p = c; double best = magnitude(A.a[c][c]);
        for (int r = c + 1; r < 4; ++r) { double m = magnitude(A.a[r][c]); if (m > best) { best = m; p = r; } }
        if (p != c) for (int j = 0; j < 4; ++j) { std::swap(A.a[c][j], A.a[p][j]); std::swap(B.a[c][j], B.a[p][j]); }
        for (int r = c + 1; r < 4; ++r) {
            T f = A.a[r][c] / A.a[c][c];
            for (int j = c; j < 4; ++j) A.a[r][j] = A.a[r][j] - f * A.a[c][j];
            for (int j = 0; j < 4; ++j) B.a[r][j] = B.a[r][j] - f * B.a[c][j];
        }
    }
    Mat4<T> X;
    for (int j = 0; j < 4; ++j)
        for (int r = 3; r >= 0; --r) {
            T s = B.a[r][j];
            for (int k = r + 1; k < 4; ++k) s = s - A.a[r][k] * X.a[k][j];
            X.a[r][j] = s / A.a[r][r];
        }
    return X;
}

// ------------------------------------------------------------------------------------------
// exp(::SMatrix{4,4}) of StaticArrays 0.12.5 (src/expm.jl), itself adapted from Julia Base
// `exp!` / Higham 2005: no balancing, 1-norm thresholds 0.015 / 0.25 / 0.95 / 2.1, Pade-13 with
// scaling by 2^ceil(log2(nA/5.4)) otherwise.  Un-vendored dependency (Manifest.toml:1410-1414):
// restated from its published algorithm; the degree-13 table is cross-checked against the
// reference's in-tree copy at spin26.jl:25-40.  Reference call sites: spin13.jl:46,
// spin12.jl:168,175-176, spin30.jl:108,113,181, spin15.jl:56, spin.jl:445,478,532.
// ------------------------------------------------------------------------------------------
// degree-13 Pade numerator coefficients (Higham 2005, table 2.3); the reference carries the same table in-tree at
// spin26.jl:25-40, which tests/test_oracle_cpu.py compares against
static const double PADE13[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                  129060195264000.,   10559470521600.,    670442572800.,     33522128640.,
                                  1323241920.,        40840800.,          960960.,           16380.,
                                  182.,               1.};
template <class T> Mat4<T> expm_ref(const Mat4<T>& A_in, int* degree_out = nullptr) {
    Mat4<T> A = A_in;
    double nA = 0.0;
    for (int j = 0; j < 4; ++j) {
        double s = 0.0;
        for (int i = 0; i < 4; ++i) s += magnitude(A.a[i][j]);
        nA = std::max(nA, s);
    }
    Mat4<T> U, V;
    int si = 0, deg;
    if (nA <= 2.1) {
        Mat4<T> A2 = matmul(A, A);
        if (nA > 0.95) {
            static const double cu[5] = {8821612800., 302702400., 2162160., 3960., 1.};
            static const double cv[5] = {17643225600., 2075673600., 30270240., 110880., 90.};
            U = matmul(A, evalpoly(A2, cu, 5)); V = evalpoly(A2, cv, 5); deg = 9;
        } else if (nA > 0.25) {
            static const double cu[4] = {8648640., 277200., 1512., 1.};
            static const double cv[4] = {17297280., 1995840., 25200., 56.};
            U = matmul(A, evalpoly(A2, cu, 4)); V = evalpoly(A2, cv, 4); deg = 7;
        } else if (nA > 0.015) {
            static const double cu[3] = {15120., 420., 1.};
            static const double cv[3] = {30240., 3360., 30.};
            U = matmul(A, evalpoly(A2, cu, 3)); V = evalpoly(A2, cv, 3); deg = 5;
        } else {
            static const double cu[2] = {60., 1.};
            static const double cv[2] = {120., 12.};
            U = matmul(A, evalpoly(A2, cu, 2)); V = evalpoly(A2, cv, 2); deg = 3;
        }
    } else {
        deg = 13;
        double s = std::log2(nA / 5.4);
        if (s > 0) {
            si = (int)std::ceil(s);
            double sc = std::ldexp(1.0, -si);
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) A.a[i][j] = A.a[i][j] * sc;
        }
        const double* C = PADE13;
        Mat4<T> A2 = matmul(A, A), A4 = matmul(A2, A2), A6 = matmul(A2, A4);
        Mat4<T> t1, t2, u, v;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double id = (i == j) ? 1.0 : 0.0;
                t1.a[i][j] = C[13] * A6.a[i][j] + C[11] * A4.a[i][j] + C[9] * A2.a[i][j];
                t2.a[i][j] = C[12] * A6.a[i][j] + C[10] * A4.a[i][j] + C[8] * A2.a[i][j];
                u.a[i][j] = C[7] * A6.a[i][j] + C[5] * A4.a[i][j] + C[3] * A2.a[i][j] + T(C[1] * id);
                v.a[i][j] = C[6] * A6.a[i][j] + C[4] * A4.a[i][j] + C[2] * A2.a[i][j] + T(C[0] * id);
            }
        Mat4<T> a6t1 = matmul(A6, t1), a6t2 = matmul(A6, t2);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { u.a[i][j] = a6t1.a[i][j] + u.a[i][j]; v.a[i][j] = a6t2.a[i][j] + v.a[i][j]; }
        U = matmul(A, u); V = v;
    }
    Mat4<T> P, Q;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { P.a[i][j] = V.a[i][j] + U.a[i][j]; Q.a[i][j] = V.a[i][j] - U.a[i][j]; }
    Mat4<T> E = solve(Q, P);
    for (int t = 0; t < si; ++t) E = matmul(E, E);
    if (degree_out) *degree_out = deg;
    return E;
}

// dt * (f * NEGI_H0_ISO + a * NEGI_H1_ISO)   (spin.jl:234-238)
template <class T> Mat4<T> generator(const T& f, const T& a, const T& dt) {
    Mat4<T> G;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) G.a[i][j] = dt * (f * (PI * N0[i][j]) + a * (PI * N1[i][j]));
    return G;
}
template <class T> Mat4<T> propagator(const T& f, const T& a, const T& dt) { return expm_ref(generator(f, a, dt)); }

// ------------------------------------------------------------------------------------------
// T1(flux) model: spin.jl:203-206 (gridded-linear interpolant, Flat extrapolation), 366, 393, 396
// Interpolations.jl semantics: clamp to the knot range (a clamped dual loses its partials),
// interval by searchsortedlast clamped to [1, n-1], value (1-w) y_i + w y_{i+1}.
// ------------------------------------------------------------------------------------------
template <class T> T interp_t1(const T& fbfq_in, double scale) {
    T x = fbfq_in;
    if (value(x) > FBFQ_ARRAY[17]) x = T(FBFQ_ARRAY[17]);
    else if (value(x) < FBFQ_ARRAY[0]) x = T(FBFQ_ARRAY[0]);
    int i = 0;
    while (i < 16 && FBFQ_ARRAY[i + 1] <= value(x)) ++i;  // searchsortedlast, clamped to n-1
    T w = (x - FBFQ_ARRAY[i]) / (FBFQ_ARRAY[i + 1] - FBFQ_ARRAY[i]);
    return (1.0 - w) * (T1_ARRAY_REDUCED[i] * scale) + w * (T1_ARRAY_REDUCED[i + 1] * scale);
}
template <class T> T amp_fbfq_lo(const T& a) { return -absd(a) * FBFQ_A + FBFQ_B; }           // spin.jl:366
template <class T> T amp_t1_spline(const T& a) { return interp_t1(amp_fbfq_lo(a), 1e3); }     // ns, spin.jl:393
template <class T> T amp_t1_reduced_spline(const T& a) { return interp_t1(amp_fbfq_lo(a), 1.0); }  // us, spin.jl:396

// ------------------------------------------------------------------------------------------
// model dimensions: spin13.jl:23-26,39-40; spin12.jl:26-33; spin15.jl:43-46; spin11.jl:53-55;
// spin17.jl state_dim; spin18.jl:23-28; spin30.jl:27-32,67-70
// ------------------------------------------------------------------------------------------
static int state_dim(const rbq_model_params* p) {
    switch (p->model_id) {
    case RBQ_MODEL_SPIN13: return 11;
    case RBQ_MODEL_SPIN12: case RBQ_MODEL_SPIN18: return 43;
    case RBQ_MODEL_SPIN15: return 11 + (p->decay_aware ? 1 : 0);
    case RBQ_MODEL_SPIN11: case RBQ_MODEL_SPIN17:
        if (p->derivative_order < 0 || p->derivative_order > 2) return RBQ_E_MODEL;
        return 11 + 4 * p->derivative_order;
    case RBQ_MODEL_SPIN30:
        if (p->sample_state_count < 1 || p->sample_state_count > RBQ_MAX_SAMPLE_STATES) return RBQ_E_MODEL;
        return 15 + 41 * p->sample_state_count;
    case RBQ_MODEL_SPIN25:   // spin25.jl:28-33, 73-75
        if (p->sample_state_count < 1 || p->sample_state_count > RBQ_MAX_SAMPLE_STATES) return RBQ_E_MODEL;
        return 17 + 41 * p->sample_state_count;
    default: return RBQ_E_MODEL;
    }
}
static int control_dim(const rbq_model_params* p) {
    if (p->model_id == RBQ_MODEL_SPIN15) return 1 + (p->time_optimal ? 1 : 0);
    return state_dim(p) < 0 ? RBQ_E_MODEL : 1;
}

// ------------------------------------------------------------------------------------------
// discrete dynamics, one function per reference file
// ------------------------------------------------------------------------------------------
// the [int a, a, da] Euler chain shared by every model (spin13.jl:49-51)
template <class T> void control_chain(const T* x, int base, const T& u, const T& dt, T* xn) {
    xn[base] = x[base] + x[base + 1] * dt;
    xn[base + 1] = x[base + 1] + x[base + 2] * dt;
    xn[base + 2] = x[base + 2] + u * dt;
}

// spin13.jl:44-58
template <class T> void step_spin13(const rbq_model_params* p, const T* x, const T* u, const T& dt, T* xn) {
    Mat4<T> E = propagator(T(p->fq), x[9], dt);
    matvec(E, x + 0, xn + 0);
    matvec(E, x + 4, xn + 4);
    control_chain(x, 8, u[0], dt, xn);
}
// spin12.jl:165-192 (the stray `p` at :184 is a typo in the reference and is ignored)
template <class T> void step_spin12(const rbq_model_params* p, const T* x, const T* u, const T& dt, T* xn) {
    Mat4<T> E = propagator(T(p->fq), x[9], dt);
    matvec(E, x + 0, xn + 0);
    matvec(E, x + 4, xn + 4);
    control_chain(x, 8, u[0], dt, xn);
    Mat4<T> Ep = propagator(T(p->fq_samples[0]), x[9], dt);
    Mat4<T> En = propagator(T(p->fq_samples[1]), x[9], dt);
    for (int s = 0; s < 4; ++s) matvec(Ep, x + 11 + 4 * s, xn + 11 + 4 * s);
    for (int s = 4; s < 8; ++s) matvec(En, x + 11 + 4 * s, xn + 11 + 4 * s);
}
// spin18.jl:162-190
template <class T> void step_spin18(const rbq_model_params* p, const T* x, const T* u, const T& dt, T* xn) {
    Mat4<T> E = propagator(T(p->fq), x[9], dt);
    matvec(E, x + 0, xn + 0);
    matvec(E, x + 4, xn + 4);
    control_chain(x, 8, u[0], dt, xn);
    Mat4<T> Ep = propagator(T(p->fq), x[9] + p->namp, dt);
    Mat4<T> En = propagator(T(p->fq), x[9] - p->namp, dt);
    for (int s = 0; s < 4; ++s) matvec(Ep, x + 11 + 4 * s, xn + 11 + 4 * s);
    for (int s = 4; s < 8; ++s) matvec(En, x + 11 + 4 * s, xn + 11 + 4 * s);
}
// spin15.jl:50-73
template <class T> void step_spin15(const rbq_model_params* p, const T* x, const T* u, const T& dt_in, T* xn) {
    T dt = dt_in;
    if (p->time_optimal) dt = u[1] * u[1];                                  // spin15.jl:52-54
    Mat4<T> E = propagator(T(p->fq), x[9], dt);
    matvec(E, x + 0, xn + 0);
    matvec(E, x + 4, xn + 4);
    xn[8] = x[8] + dt * x[9];
    xn[9] = x[9] + dt * x[10];
    xn[10] = x[10] + dt * u[0];
    if (p->decay_aware) xn[11] = x[11] + dt * (1.0 / amp_t1_reduced_spline(x[9]));  // spin15.jl:66-69
}
// spin11.jl:64-117 (which = 0, H0) and spin17.jl:61-114 (which = 1, H1)
template <class T> void step_spin11_17(const rbq_model_params* p, int which, const T* x, const T* u, const T& dt, T* xn) {
    const double (*NH)[4] = which == 0 ? N0 : N1;
    Mat4<T> E = propagator(T(p->fq), x[9], dt);
    matvec(E, x + 0, xn + 0);
    matvec(E, x + 4, xn + 4);
    control_chain(x, 8, u[0], dt, xn);
    if (p->derivative_order >= 1) {                                           // spin11.jl:83-86
        T w[4];
        for (int i = 0; i < 4; ++i) {
            T s = T(0.0);
            for (int k = 0; k < 4; ++k) s = s + dt * (PI * NH[i][k]) * x[k];
            w[i] = x[11 + i] + s;
        }
        matvec(E, w, xn + 11);
    }
    if (p->derivative_order >= 2) {                                           // spin11.jl:101-104
        T w[4];
        for (int i = 0; i < 4; ++i) {
            T s = T(0.0);
            for (int k = 0; k < 4; ++k) s = s + dt * (2.0 * PI * NH[i][k]) * x[11 + k];
            w[i] = x[15 + i] + s;
        }
        matvec(E, w, xn + 15);
    }
}
// generic lower Cholesky, the `cholesky(Symmetric(cov)).L` of spin30.jl:140
template <class T> void cholesky_lower(int n, const T* A, T* L) {
    for (int i = 0; i < n * n; ++i) L[i] = T(0.0);
    for (int j = 0; j < n; ++j) {
        T s = A[j * n + j];
        for (int k = 0; k < j; ++k) s = s - L[j * n + k] * L[j * n + k];
        T ljj = sqrt(s);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            T t = A[i * n + j];
            for (int k = 0; k < j; ++k) t = t - L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / ljj;
        }
    }
}
// unscented_transform of spin30.jl:86-174 / spin25.jl:91-178 for the chunk at `off`: Ep / En propagate points 5 / 10
template <class T>
void unscented_transform(const rbq_model_params* p, const T* x, int off, int nom, const Mat4<T>& E, const Mat4<T>& Ep,
                         const Mat4<T>& En, double joint_cov, T* xn) {
    T s[10][4];
    for (int j = 0; j < 10; ++j) {
        const Mat4<T>& M = (j == 4) ? Ep : (j == 9) ? En : E;
        matvec(M, x + off + 4 * j, s[j]);
    }
    T sm[4];
    for (int i = 0; i < 4; ++i) {                                          // spin30.jl:115-118
        T acc = s[0][i];
        for (int j = 1; j < 10; ++j) acc = acc + s[j][i];
        sm[i] = 0.1 * acc;
    }
    T d[10][4];
    for (int j = 0; j < 10; ++j) for (int i = 0; i < 4; ++i) d[j][i] = s[j][i] - sm[i];
    T cov[25];
    for (int i = 0; i < 25; ++i) cov[i] = T(0.0);
    const double cf = 1.0 / (2.0 * p->alpha * p->alpha);                  // spin30.jl:130
    T scov[4][4];
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 4; ++k) {
            T acc = d[0][i] * d[0][k];
            for (int j = 1; j < 10; ++j) acc = acc + d[j][i] * d[j][k];
            scov[i][k] = cf * acc;
            cov[i * 5 + k] = scov[i][k];
        }
    cov[24] = T(joint_cov);                                                // spin30.jl:138 / spin25.jl:144
    T L[25];
    cholesky_lower(5, cov, L);
    T snew[10][4];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            T col = p->alpha * L[i * 5 + j];                               // spin30.jl:140-145
            snew[j][i] = sm[i] + col;
            snew[5 + j][i] = sm[i] - col;
        }
    for (int i = 0; i < 4; ++i) { snew[4][i] = sm[i]; snew[9][i] = sm[i]; }
    for (int j = 0; j < 10; ++j) {                                         // spin30.jl:157-166
        T nn = snew[j][0] * snew[j][0];
        for (int i = 1; i < 4; ++i) nn = nn + snew[j][i] * snew[j][i];
        T nr = sqrt(nn);
        for (int i = 0; i < 4; ++i) xn[off + 4 * j + i] = snew[j][i] / nr;
    }
    T pen = T(0.0);                                                        // spin30.jl:168-169
    for (int i = 0; i < 4; ++i) pen = pen + scov[i][i] * p->S_diag[i];
    for (int i = 0; i < 4; ++i) {
        T dn = sm[i] - x[nom + i];
        pen = pen + dn * p->S_diag[i] * dn;
    }
    xn[off + 40] = pen;
}
// spin30.jl:177-202
template <class T> void step_spin30(const rbq_model_params* p, const T* x, const T* u, const T& dt, T* xn) {
    const T a = x[13];
    Mat4<T> E = propagator(T(p->fq), a, dt);                                  // spin30.jl:181
    matvec(E, x + 0, xn + 0);
    matvec(E, x + 4, xn + 4);
    matvec(E, x + 8, xn + 8);
    control_chain(x, 12, u[0], dt, xn);
    const double fq_chol5 = std::sqrt(p->fq_cov);                             // spin30.jl:102
    Mat4<T> Ep = propagator(T(p->fq + fq_chol5), a, dt);                      // spin30.jl:108
    Mat4<T> En = propagator(T(p->fq - fq_chol5), a, dt);                      // spin30.jl:113
    for (int c = 0; c < p->sample_state_count; ++c)
        unscented_transform(p, x, 15 + 41 * c, p->nominal_offset[c], E, Ep, En, p->fq_cov, xn);
}
// Int(div(time, dt)) with Julia's floating-point div = round((x - rem(x, y)) / y)   (spin25.jl:196)
static long long knot_point_of(double time, double dt) { return (long long)std::rint((time - std::fmod(time, dt)) / dt); }
// spin25.jl:181-235: amplitude-noise unscented model; the pink filter's taps are part of the state and its start-up
// branch depends on the knot index derived from the TIME argument.  `fq_cov` carries wnamp_cov (include/rbq.h).
template <class T> void step_spin25(const rbq_model_params* p, const T* x, const T* u, double time, const T& dt, T* xn) {
    const T camp = x[9];
    Mat4<T> E = propagator(T(p->fq), camp, dt);
    matvec(E, x + 0, xn + 0);
    matvec(E, x + 4, xn + 4);
    control_chain(x, 8, u[0], dt, xn);
    const double xk = std::sqrt(p->fq_cov);
    const long long kp = knot_point_of(time, value(dt));
    const T xp1 = x[11], xp2 = x[12], xp3 = x[13], yp1 = x[14], yp2 = x[15], yp3 = x[16];
    T yk;
    if (kp == 1) yk = T(PFIR_B1 * xk);
    else if (kp == 2) yk = PFIR_B1 * xk + PFIR_B2 * xp1 - PFIR_A2 * yp1;
    else if (kp == 3) yk = PFIR_B1 * xk + PFIR_B2 * xp1 + PFIR_B3 * xp2 - PFIR_A2 * yp1 - PFIR_A3 * yp2;
    else yk = PFIR_B1 * xk + PFIR_B2 * xp1 + PFIR_B3 * xp2 + PFIR_B4 * xp3 - PFIR_A2 * yp1 - PFIR_A3 * yp2 - PFIR_A4 * yp3;
    xn[11] = T(xk); xn[12] = xp1; xn[13] = xp2; xn[14] = yk; xn[15] = yp1; xn[16] = yp2;
    const T namp = p->namp * yk;
    Mat4<T> Ep = propagator(T(p->fq), camp + namp, dt);                      // spin25.jl:111
    Mat4<T> En = propagator(T(p->fq), camp - namp, dt);                      // spin25.jl:116
    for (int c = 0; c < p->sample_state_count; ++c)
        unscented_transform(p, x, 17 + 41 * c, p->nominal_offset[c], E, Ep, En, p->fq_cov, xn);
}

// ------------------------------------------------------------------------------------------
// The reference's Runge-Kutta variants: RD.dynamics(model, x, u, t) integrated by RobotDynamics' explicit RK2 / RK3 / RK4
// (RobotDynamics 0.2.1 @ SchusterLab/rbqoc-stable, Manifest.toml:1283-1289, un-vendored: restated from its published
// tableaux -- RK2 midpoint; RK3 k3 at x - k1 + 2 k2, weights (1,4,1)/6; classic RK4; IT_RDI at rbqoc.jl:49-54).
// Continuous dynamics: spin15_standalone.jl:263-312 (nominal + decay aware + time optimal, T1 in ns via amp_t1_spline, the
// time-optimal form multiplies every component by u2^2); derivative states as spin19.jl:66-129 / spin21.jl:63-109
// ((d^o psi)' = o H_d d^(o-1) psi + H d^o psi), here with the full generator H = fq H0 + a H1 of the spin13 layout.
// ------------------------------------------------------------------------------------------
template <class T> void mul_generator(double fq, const T& a, const T* x, T* y) {
    for (int i = 0; i < 4; ++i) {
        T s = T(0.0);
        for (int k = 0; k < 4; ++k) s = s + (T(fq * (PI * N0[i][k])) + a * (PI * N1[i][k])) * x[k];
        y[i] = s;
    }
}
template <class T> void continuous_dynamics(const rbq_model_params* p, int n, const T* x, const T* u, T* dx) {
    const T a = x[9];
    mul_generator<T>(p->fq, a, x, dx);
    mul_generator<T>(p->fq, a, x + 4, dx + 4);
    dx[8] = a; dx[9] = x[10]; dx[10] = u[0];
    if (p->model_id == RBQ_MODEL_SPIN15) {
        if (p->decay_aware) dx[11] = 1.0 / amp_t1_spline<T>(a);
        if (p->time_optimal) { const T s = u[1] * u[1]; for (int i = 0; i < n; ++i) dx[i] = dx[i] * s; }
    } else if (p->model_id == RBQ_MODEL_SPIN11 || p->model_id == RBQ_MODEL_SPIN17) {
        const double (*Hd)[4] = p->model_id == RBQ_MODEL_SPIN17 ? N1 : N0;
        for (int o = 1; o <= p->derivative_order && o <= 2; ++o) {
            const T* src = o == 1 ? x : x + 11;
            T* dst = dx + 11 + 4 * (o - 1);
            mul_generator<T>(p->fq, a, x + 11 + 4 * (o - 1), dst);
            for (int i = 0; i < 4; ++i) {
                T g = T(0.0);
                for (int k = 0; k < 4; ++k) g = g + src[k] * (PI * Hd[i][k]);
                dst[i] = dst[i] + g * (double)o;
            }
        }
    }
}
template <class T> int step_rk(const rbq_model_params* p, const T* x, const T* u, const T& dt, T* xn) {
    const int n = state_dim(p);
    if (n < 0 || n > 32) return RBQ_E_MODEL;
    if (!(p->model_id == RBQ_MODEL_SPIN13 || p->model_id == RBQ_MODEL_SPIN15 || p->model_id == RBQ_MODEL_SPIN11 || p->model_id == RBQ_MODEL_SPIN17)) return RBQ_E_MODEL;
    std::vector<T> k1(n), k2(n), k3(n), k4(n), xt(n);
    continuous_dynamics<T>(p, n, x, u, k1.data());
    if (p->integrator == RBQ_INTEGRATOR_RK2) {
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * dt * 0.5;
        continuous_dynamics<T>(p, n, xt.data(), u, k2.data());
        for (int i = 0; i < n; ++i) xn[i] = x[i] + k2[i] * dt;
    } else if (p->integrator == RBQ_INTEGRATOR_RK3) {
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * dt * 0.5;
        continuous_dynamics<T>(p, n, xt.data(), u, k2.data());
        for (int i = 0; i < n; ++i) xt[i] = x[i] - k1[i] * dt + k2[i] * dt * 2.0;
        continuous_dynamics<T>(p, n, xt.data(), u, k3.data());
        for (int i = 0; i < n; ++i) xn[i] = x[i] + (k1[i] * dt + k2[i] * dt * 4.0 + k3[i] * dt) / 6.0;
    } else if (p->integrator == RBQ_INTEGRATOR_RK4) {
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] * dt * 0.5;
        continuous_dynamics<T>(p, n, xt.data(), u, k2.data());
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k2[i] * dt * 0.5;
        continuous_dynamics<T>(p, n, xt.data(), u, k3.data());
        for (int i = 0; i < n; ++i) xt[i] = x[i] + k3[i] * dt;
        continuous_dynamics<T>(p, n, xt.data(), u, k4.data());
        for (int i = 0; i < n; ++i) xn[i] = x[i] + (k1[i] * dt + k2[i] * dt * 2.0 + k3[i] * dt * 2.0 + k4[i] * dt) / 6.0;
    } else return RBQ_E_MODEL;
    return 0;
}
template <class T> int step(const rbq_model_params* p, const T* x, const T* u, double time, const T& dt, T* xn) {
    if (p->integrator != RBQ_INTEGRATOR_EXP) return step_rk<T>(p, x, u, dt, xn);
    switch (p->model_id) {
    case RBQ_MODEL_SPIN25: step_spin25(p, x, u, time, dt, xn); return 0;
    case RBQ_MODEL_SPIN13: step_spin13(p, x, u, dt, xn); return 0;
    case RBQ_MODEL_SPIN12: step_spin12(p, x, u, dt, xn); return 0;
    case RBQ_MODEL_SPIN18: step_spin18(p, x, u, dt, xn); return 0;
    case RBQ_MODEL_SPIN15: step_spin15(p, x, u, dt, xn); return 0;
    case RBQ_MODEL_SPIN11: step_spin11_17(p, 0, x, u, dt, xn); return 0;
    case RBQ_MODEL_SPIN17: step_spin11_17(p, 1, x, u, dt, xn); return 0;
    case RBQ_MODEL_SPIN30: step_spin30(p, x, u, dt, xn); return 0;
    default: return RBQ_E_MODEL;
    }
}

// ------------------------------------------------------------------------------------------
// gates and fidelities: spin.jl:301-323, 1035-1037, 1101-1104, 1198-1240
// ------------------------------------------------------------------------------------------
typedef std::complex<double> cplx;
static bool gate_matrix(int gate, cplx g[2][2]) {
    const double r = 1.0 / std::sqrt(2.0);
    const cplx I(0, 1);
    switch (gate) {
    case RBQ_GATE_ZPIBY2: g[0][0] = (1.0 - I) * r; g[0][1] = 0; g[1][0] = 0; g[1][1] = (1.0 + I) * r; return true;
    case RBQ_GATE_YPIBY2: g[0][0] = r; g[0][1] = -r; g[1][0] = r; g[1][1] = r; return true;
    case RBQ_GATE_XPIBY2: g[0][0] = r; g[0][1] = -I * r; g[1][0] = -I * r; g[1][1] = r; return true;
    case RBQ_GATE_XPI: g[0][0] = 0; g[0][1] = -I; g[1][0] = -I; g[1][1] = 0; return true;
    }
    return false;
}
// get_mat_iso, rbqoc.jl:204-210
static void mat_iso(const cplx g[2][2], double G[4][4]) {
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            G[i][j] = g[i][j].real(); G[i][j + 2] = -g[i][j].imag();
            G[i + 2][j] = g[i][j].imag(); G[i + 2][j + 2] = g[i][j].real();
        }
}
// spin.jl:1035-1037
template <class R> R fidelity_vec_iso2(const R* s1, const R* s2) {
    R a = s1[0] * s2[0] + s1[1] * s2[1] + s1[2] * s2[2] + s1[3] * s2[3];
    R b = s1[0] * s2[2] + s1[1] * s2[3] - s1[2] * s2[0] - s1[3] * s2[1];
    return a * a + b * b;
}

// Hermitian 2x2 eigen-decomposition -> principal square root (complex when an eigenvalue is
// negative, as Julia's sqrt(::Hermitian) does).
static void sqrt_herm2(const cplx m[2][2], cplx out[2][2]) {
    double a = m[0][0].real(), d = m[1][1].real();
    cplx b = 0.5 * (m[0][1] + std::conj(m[1][0]));
    double tr = a + d, df = a - d;
    double disc = std::sqrt(df * df + 4.0 * std::norm(b));
    double l1 = 0.5 * (tr + disc), l2 = 0.5 * (tr - disc);
    cplx v1[2], v2[2];
    if (std::abs(b) > 0) {
        v1[0] = b; v1[1] = l1 - a;
        double n1 = std::sqrt(std::norm(v1[0]) + std::norm(v1[1]));
        v1[0] /= n1; v1[1] /= n1;
        v2[0] = -std::conj(v1[1]); v2[1] = std::conj(v1[0]);
    } else {
        if (a >= d) { v1[0] = 1; v1[1] = 0; v2[0] = 0; v2[1] = 1; }
        else { v1[0] = 0; v1[1] = 1; v2[0] = 1; v2[1] = 0; }
    }
    cplx s1 = l1 >= 0 ? cplx(std::sqrt(l1), 0) : cplx(0, std::sqrt(-l1));
    cplx s2 = l2 >= 0 ? cplx(std::sqrt(l2), 0) : cplx(0, std::sqrt(-l2));
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) out[i][j] = s1 * v1[i] * std::conj(v1[j]) + s2 * v2[i] * std::conj(v2[j]);
}
// general 2x2 principal square root (the outer `sqrt` of spin.jl:1103 acts on a non-Hermitian
// -typed product): sqrt(M) = (M + s I)/t, s = sqrt(det M), t = sqrt(tr M + 2 s).
static cplx tr_sqrt_general2(const cplx m[2][2]) {
    cplx det = m[0][0] * m[1][1] - m[0][1] * m[1][0];
    cplx tr = m[0][0] + m[1][1];
    cplx s = std::sqrt(det);
    cplx t = std::sqrt(tr + 2.0 * s);
    if (std::abs(t) == 0) return 0;
    return (tr + 2.0 * s) / t;
}
// spin.jl:1101-1104: real(tr(sqrt(sqrt(m1) * m2 * sqrt(m1))))   (square-ROOT fidelity)
static double fidelity_mat(const cplx m1[2][2], const cplx m2[2][2]) {
    cplx r[2][2], t[2][2], q[2][2];
    sqrt_herm2(m1, r);
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i][j] = r[i][0] * m2[0][j] + r[i][1] * m2[1][j];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) q[i][j] = t[i][0] * r[0][j] + t[i][1] * r[1][j];
    return tr_sqrt_general2(q).real();
}
// closed form for 2x2 states: F = sqrt(tr(m1 m2) + 2 sqrt(det m1 det m2)); identical in exact
// arithmetic, free of the sqrt(rounding) noise the eigen route has for pure targets.
static double fidelity_mat_closed(const cplx m1[2][2], const cplx m2[2][2]) {
    double trp = 0;
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) trp += (m1[i][j] * m2[j][i]).real();
    double d1 = (m1[0][0] * m1[1][1] - m1[0][1] * m1[1][0]).real();
    double d2 = (m2[0][0] * m2[1][1] - m2[0][1] * m2[1][0]).real();
    double dd = std::max(d1, 0.0) * std::max(d2, 0.0);
    return std::sqrt(std::max(trp + 2.0 * std::sqrt(dd), 0.0));
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  The reference seeds Julia's MersenneTwister
// (spin.jl:1131,1179), which cannot be reproduced outside Julia; the engine and this oracle
// therefore share a counter-based generator keyed by the global sample index.
//   counter = (index_lo, index_hi, j, stream), key = (seed_lo, seed_hi)
//   stream 0: j=0,1 -> 4 uniforms of psi0;  stream 1: j=0 -> normals (fq, da);
//   stream 2: j -> white-noise pair (x_{2j}, x_{2j+1}) of the pink filter.
// ------------------------------------------------------------------------------------------
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u53(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0); }
static void philox_uniform2(uint64_t index, uint64_t seed, uint32_t j, uint32_t stream, double* u0, double* u1) {
    uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), j, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    *u0 = u53(o[0], o[1]);
    *u1 = u53(o[2], o[3]);
}
static void philox_normal2(uint64_t index, uint64_t seed, uint32_t j, uint32_t stream, double* z0, double* z1) {
    double u0, u1;
    philox_uniform2(index, seed, j, stream, &u0, &u1);
    double r = std::sqrt(-2.0 * std::log(1.0 - u0));   // 1-u0 in (0,1]
    double ang = 2.0 * PI * u1;
    *z0 = r * std::cos(ang);
    *z1 = r * std::sin(ang);
}
// gen_rand_state_iso recipe (spin.jl:1132-1136): rand(2) + im*rand(2), normalised, [re; im]
static void sample_psi0(uint64_t index, uint64_t seed, double psi[4]) {
    philox_uniform2(index, seed, 0, 0, &psi[0], &psi[1]);
    philox_uniform2(index, seed, 1, 0, &psi[2], &psi[3]);
    double n = std::sqrt(psi[0] * psi[0] + psi[1] * psi[1] + psi[2] * psi[2] + psi[3] * psi[3]);
    for (int i = 0; i < 4; ++i) psi[i] /= n;
}
static void sample_static(uint64_t index, uint64_t seed, double fq0, double fq_rel_sigma, double fq_rel_clip,
                          double da_sigma, double* fq, double* da) {
    double z0, z1;
    philox_normal2(index, seed, 0, 1, &z0, &z1);
    double dev = fq_rel_sigma * z0;
    if (dev > fq_rel_clip) dev = fq_rel_clip;
    if (dev < -fq_rel_clip) dev = -fq_rel_clip;
    *fq = fq0 * (1.0 + dev);
    *da = da_sigma * z1;
}

// gen_pink_noise, spin.jl:1178-1192, on a given white sequence xs
static void pink_filter(const double* xs, int64_t count, double dt_inv, double* ys) {
    for (int64_t i = 0; i < count; ++i) {
        double y = PFIR_B1 * xs[i];
        if (i >= 1) y += PFIR_B2 * xs[i - 1] - PFIR_A2 * ys[i - 1];
        if (i >= 2) y += PFIR_B3 * xs[i - 2] - PFIR_A3 * ys[i - 2];
        if (i >= 3) y += PFIR_B4 * xs[i - 3] - PFIR_A4 * ys[i - 3];
        ys[i] = y;
    }
    for (int64_t i = 0; i < count; ++i) ys[i] *= dt_inv;
}

// ------------------------------------------------------------------------------------------
// compute_fidelities for iso state vectors, spin.jl:1198-1240 (states are BigFloat in
// run_sim_prop, spin.jl:1581: long double is the widest native type here)
// ------------------------------------------------------------------------------------------
typedef long double xreal;
static void gate_targets(int gate, const xreal* s0, xreal tg[4][4]) {
    cplx g[2][2];
    gate_matrix(gate, g);
    double G1[4][4], G2[4][4], G3[4][4];
    mat_iso(g, G1);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += G1[i][k] * G1[k][j]; G2[i][j] = s; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += G2[i][k] * G1[k][j]; G3[i][j] = s; }
    for (int i = 0; i < 4; ++i) {
        tg[0][i] = s0[i];
        xreal a = 0, b = 0, c = 0;
        for (int k = 0; k < 4; ++k) { a += (xreal)G1[i][k] * s0[k]; b += (xreal)G2[i][k] * s0[k]; c += (xreal)G3[i][k] * s0[k]; }
        tg[1][i] = a; tg[2][i] = b; tg[3][i] = c;
    }
}

// integrate_prop_schroed! (spin.jl:438-452) + compute_fidelities for ONE sample
static void sim_schroed(const double* controls, int64_t nk, double dt, int gate, int64_t gate_count, double fq,
                        double da, const double* psi0, double* fid, double* state_out) {
    Mat4<double> prop = eye4<double>();
    for (int64_t j = 0; j < nk; ++j) {
        Mat4<double> H;                                   // hamiltonian = negi_h0 + controls[j]*NEGI_H1_ISO
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) H.a[r][c] = (fq * (PI * N0[r][c]) + (controls[j] + da) * (PI * N1[r][c])) * dt;
        prop = matmul(expm_ref(H), prop);
    }
    xreal s0[4], st[4], tg[4][4];
    for (int i = 0; i < 4; ++i) { s0[i] = psi0[i]; st[i] = psi0[i]; }
    gate_targets(gate, s0, tg);
    if (fid) fid[0] = (double)fidelity_vec_iso2<xreal>(st, tg[0]);
    for (int64_t g = 1; g <= gate_count; ++g) {
        xreal nx[4];
        for (int i = 0; i < 4; ++i) { xreal s = 0; for (int k = 0; k < 4; ++k) s += (xreal)prop.a[i][k] * st[k]; nx[i] = s; }
        for (int i = 0; i < 4; ++i) st[i] = nx[i];
        if (fid) fid[g] = (double)fidelity_vec_iso2<xreal>(st, tg[g % 4]);
    }
    if (state_out) for (int i = 0; i < 4; ++i) state_out[i] = (double)st[i];
}

// analytic gates: ordered product of exp(segment generator * segment duration)  (spin.jl:562-570, 721-727)
static void sim_schroed_segments(const double* amps, const double* durs, int64_t nseg, int gate, int64_t gate_count, double fq,
                                 double da, const double* psi0, double* fid) {
    Mat4<double> prop = eye4<double>();
    for (int64_t j = 0; j < nseg; ++j) {
        Mat4<double> H;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) H.a[r][c] = (fq * (PI * N0[r][c]) + (amps[j] + da) * (PI * N1[r][c])) * durs[j];
        prop = matmul(expm_ref(H), prop);
    }
    xreal s0[4], st[4], tg[4][4];
    for (int i = 0; i < 4; ++i) { s0[i] = psi0[i]; st[i] = psi0[i]; }
    gate_targets(gate, s0, tg);
    fid[0] = (double)fidelity_vec_iso2<xreal>(st, tg[0]);
    for (int64_t g = 1; g <= gate_count; ++g) {
        xreal nx[4];
        for (int i = 0; i < 4; ++i) { xreal s = 0; for (int k = 0; k < 4; ++k) s += (xreal)prop.a[i][k] * st[k]; nx[i] = s; }
        for (int i = 0; i < 4; ++i) st[i] = nx[i];
        fid[g] = (double)fidelity_vec_iso2<xreal>(st, tg[g % 4]);
    }
}

// integrate_prop_xpiby2da! (spin.jl:819-905) generalised to any segment list: the state is stepped through every segment of
// every gate; segment j reads noise[nidx[j]] -- the SAME indices in every gate, as the reference does (it never advances
// the noise index from one gate to the next)
static void sim_schroedda_segments(const double* amps, const double* durs, const int32_t* nidx, int64_t nseg, int gate,
                                   int64_t gate_count, double fq, const double* noise, const double* psi0, double* fid) {
    xreal s0[4], st[4], tg[4][4];
    for (int i = 0; i < 4; ++i) { s0[i] = psi0[i]; st[i] = psi0[i]; }
    gate_targets(gate, s0, tg);
    fid[0] = (double)fidelity_vec_iso2<xreal>(st, tg[0]);
    for (int64_t g = 1; g <= gate_count; ++g) {
        for (int64_t j = 0; j < nseg; ++j) {
            const double amp = amps[j] + noise[nidx[j]];
            Mat4<double> H;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) H.a[r][c] = durs[j] * (fq * (PI * N0[r][c]) + amp * (PI * N1[r][c]));
            Mat4<double> U = expm_ref(H);
            xreal nx[4];
            for (int i = 0; i < 4; ++i) { xreal s = 0; for (int k = 0; k < 4; ++k) s += (xreal)U.a[i][k] * st[k]; nx[i] = s; }
            for (int i = 0; i < 4; ++i) st[i] = nx[i];
        }
        fid[g] = (double)fidelity_vec_iso2<xreal>(st, tg[g % 4]);
    }
}

// integrate_prop_schroedda! (spin.jl:469-484) for ONE sample
static void sim_schroedda(const double* controls, int64_t nk, double dt, int gate, int64_t gate_count, double fq,
                          const double* noise, int64_t noise_len, double noise_dt_inv, const double* psi0,
                          double* fid, double* state_out) {
    xreal s0[4], st[4], tg[4][4];
    for (int i = 0; i < 4; ++i) { s0[i] = psi0[i]; st[i] = psi0[i]; }
    gate_targets(gate, s0, tg);
    if (fid) fid[0] = (double)fidelity_vec_iso2<xreal>(st, tg[0]);
    double time = 0.0;
    for (int64_t g = 1; g <= gate_count; ++g) {
        for (int64_t j = 0; j < nk; ++j) {
            int64_t idx = (int64_t)std::floor(time * noise_dt_inv);   // 0-based version of spin.jl:476
            if (idx >= noise_len) idx = noise_len - 1;                 // (the reference would throw)
            double delta_a = noise[idx];
            Mat4<double> H;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) H.a[r][c] = (fq * (PI * N0[r][c]) + (controls[j] + delta_a) * (PI * N1[r][c])) * dt;
            Mat4<double> U = expm_ref(H);
            xreal nx[4];
            for (int i = 0; i < 4; ++i) { xreal s = 0; for (int k = 0; k < 4; ++k) s += (xreal)U.a[i][k] * st[k]; nx[i] = s; }
            for (int i = 0; i < 4; ++i) st[i] = nx[i];
            time = time + dt;
        }
        if (fid) fid[g] = (double)fidelity_vec_iso2<xreal>(st, tg[g % 4]);
    }
    if (state_out) for (int i = 0; i < 4; ++i) state_out[i] = (double)st[i];
}

// ------------------------------------------------------------------------------------------
// Lindblad, vec form: spin.jl:240-247, 255-276, 521-539
// ------------------------------------------------------------------------------------------
static void kron2(const cplx A[2][2], const cplx B[2][2], cplx K[4][4]) {
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k) for (int l = 0; l < 2; ++l)
        K[2 * i + k][2 * j + l] = A[i][j] * B[k][l];
}
struct LindbladOps { cplx H0v[4][4], H1v[4][4], LOP[4][4]; };
static LindbladOps make_lindblad_ops(double fq) {
    LindbladOps o;
    const cplx I(0, 1);
    cplx Id[2][2] = {{1, 0}, {0, 1}};
    cplx H0[2][2] = {{fq * -I * PI, 0}, {0, -(fq * -I * PI)}};          // FQ_NEGI_H0_, spin.jl:240
    cplx H1[2][2] = {{0, -I * PI}, {-I * PI, 0}};                        // NEGI_H1_, spin.jl:244
    cplx H0t[2][2] = {{H0[0][0], H0[1][0]}, {H0[0][1], H0[1][1]}};
    cplx H1t[2][2] = {{H1[0][0], H1[1][0]}, {H1[0][1], H1[1][1]}};
    cplx A[4][4], B[4][4];
    kron2(Id, H0, A); kron2(H0t, Id, B);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o.H0v[i][j] = A[i][j] - B[i][j];      // spin.jl:242
    kron2(Id, H1, A); kron2(H1t, Id, B);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o.H1v[i][j] = A[i][j] - B[i][j];      // spin.jl:246
    cplx GG[2][2] = {{1, 0}, {0, 0}}, GE[2][2] = {{0, 1}, {0, 0}}, EG[2][2] = {{0, 0}, {1, 0}}, EE[2][2] = {{0, 0}, {0, 1}};
    cplx k1[4][4], k2[4][4], k3[4][4], k4[4][4], k5[4][4], k6[4][4];
    kron2(GE, GE, k1); kron2(Id, EE, k2); kron2(EE, Id, k3);              // LOPP_VISO_, spin.jl:267
    kron2(EG, EG, k4); kron2(Id, GG, k5); kron2(GG, Id, k6);              // LOPM_VISO_, spin.jl:272
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
        o.LOP[i][j] = (k1[i][j] - 0.5 * k2[i][j] - 0.5 * k3[i][j]) + (k4[i][j] - 0.5 * k5[i][j] - 0.5 * k6[i][j]);
    return o;
}
// integrate_prop_lindbladt1! for one realisation.  rho0/rho_out: vec (column-major) 4 complex.
static void sim_lindbladt1(const double* controls, int64_t nk, double dt, int gate, int64_t gate_count, double fq,
                           double da, const cplx* rho0v, double* fid, double* fid_closed, cplx* rho_out) {
    LindbladOps o = make_lindblad_ops(fq);
    Mat4<cplx> prop = eye4<cplx>();
    for (int64_t j = 0; j < nk; ++j) {
        double c1 = controls[j] + da;
        double gamma1 = 1.0 / amp_t1_spline<double>(c1);                      // spin.jl:529
        Mat4<cplx> L;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) L.a[r][c] = (o.H0v[r][c] + c1 * o.H1v[r][c] + gamma1 * o.LOP[r][c]) * dt;
        prop = matmul(expm_ref(L), prop);
    }
    cplx g[2][2], g2[2][2], g3[2][2];
    gate_matrix(gate, g);
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) g2[i][j] = g[i][0] * g[0][j] + g[i][1] * g[1][j];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) g3[i][j] = g2[i][0] * g[0][j] + g2[i][1] * g[1][j];
    cplx r0[2][2] = {{rho0v[0], rho0v[2]}, {rho0v[1], rho0v[3]}};
    cplx tg[4][2][2];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) tg[0][i][j] = r0[i][j];
    const cplx (*gs[3])[2] = {g, g2, g3};
    for (int q = 0; q < 3; ++q) {                                              // id_k = g_k * id0 * g_k'
        cplx t[2][2];
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i][j] = gs[q][i][0] * r0[0][j] + gs[q][i][1] * r0[1][j];
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) tg[q + 1][i][j] = t[i][0] * std::conj(gs[q][j][0]) + t[i][1] * std::conj(gs[q][j][1]);
    }
    cplx v[4] = {rho0v[0], rho0v[1], rho0v[2], rho0v[3]};
    for (int64_t k = 0; k <= gate_count; ++k) {
        if (k > 0) matvec(prop, v, v);
        cplx m[2][2] = {{v[0], v[2]}, {v[1], v[3]}};
        if (fid) fid[k] = fidelity_mat(m, tg[k % 4]);
        if (fid_closed) fid_closed[k] = fidelity_mat_closed(m, tg[k % 4]);
    }
    if (rho_out) for (int i = 0; i < 4; ++i) rho_out[i] = v[i];
}

}  // namespace orc

// =============================================================================================
// C exports (loaded with ctypes by oracle/oracle.py)
// =============================================================================================
using namespace orc;
extern "C" {

int orc_state_dim(const rbq_model_params* p) { return p ? state_dim(p) : RBQ_E_NULL; }
int orc_control_dim(const rbq_model_params* p) { return p ? control_dim(p) : RBQ_E_NULL; }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// A, E row-major 4x4.  Returns the Pade degree used.
// thread count of the OpenMP loops that take no nthreads argument (orc_jacobians, orc_rollout, ...); n <= 0: all
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}
int orc_expm4(const double* A, double* E) {
    Mat4<double> M; for (int i = 0; i < 16; ++i) M.a[i / 4][i % 4] = A[i];
    int deg; Mat4<double> R = expm_ref(M, &deg);
    for (int i = 0; i < 16; ++i) E[i] = R.a[i / 4][i % 4];
    return deg;
}
// complex row-major 4x4 as interleaved (re,im)
int orc_expm4c(const double* A, double* E) {
    Mat4<cplx> M; for (int i = 0; i < 16; ++i) M.a[i / 4][i % 4] = cplx(A[2 * i], A[2 * i + 1]);
    int deg; Mat4<cplx> R = expm_ref(M, &deg);
    for (int i = 0; i < 16; ++i) { E[2 * i] = R.a[i / 4][i % 4].real(); E[2 * i + 1] = R.a[i / 4][i % 4].imag(); }
    return deg;
}
void orc_pade13_coeffs(double* out) { for (int i = 0; i < 14; ++i) out[i] = PADE13[i]; }
void orc_propagator(double fq, double a, double dt, double* E) {
    Mat4<double> R = propagator<double>(fq, a, dt);
    for (int i = 0; i < 16; ++i) E[i] = R.a[i / 4][i % 4];
}
double orc_t1_ns(double a) { return amp_t1_spline<double>(a); }
double orc_t1_reduced_us(double a) { return amp_t1_reduced_spline<double>(a); }

int orc_step(const rbq_model_params* p, const double* x, const double* u, double t, double dt, double* xn) {
    if (!p || !x || !u || !xn) return RBQ_E_NULL;
    int n = state_dim(p); if (n < 0) return n;
    std::vector<double> tmp(n);
    int rc = step<double>(p, x, u, t, dt, tmp.data());
    if (rc == 0) std::memcpy(xn, tmp.data(), sizeof(double) * n);
    return rc;
}
// ForwardDiff.jacobian(s -> discrete_dynamics(RK3, model, s[ix], s[iu], t, dt), [x;u]),
// written column-major n x (n+m)  (RobotDynamics 0.2.1 default discrete_jacobian!)
int orc_jacobian(const rbq_model_params* p, const double* x, const double* u, double t, double dt, double* AB) {
    if (!p || !x || !u || !AB) return RBQ_E_NULL;
    int n = state_dim(p), m = control_dim(p); if (n < 0 || m < 0) return RBQ_E_MODEL;
    g_np = n + m;
    std::vector<Dual> xd(n), ud(m), xo(n);
    for (int i = 0; i < n; ++i) { xd[i] = Dual(x[i]); xd[i].d[i] = 1.0; }
    for (int j = 0; j < m; ++j) { ud[j] = Dual(u[j]); ud[j].d[n + j] = 1.0; }
    int rc = step<Dual>(p, xd.data(), ud.data(), t, Dual(dt), xo.data());
    if (rc == 0)
        for (int c = 0; c < n + m; ++c) for (int r = 0; r < n; ++r) AB[(size_t)c * n + r] = xo[r].d[c];
    g_np = 0;
    return rc;
}
int orc_jacobians(const rbq_model_params* p, int64_t K, const double* X, const double* U, const double* t, const double* dt, double* AB) {
    int n = state_dim(p), m = control_dim(p); if (n < 0 || m < 0) return RBQ_E_MODEL;
    int rc = 0;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < K; ++k) {
        int r = orc_jacobian(p, X + k * n, U + k * m, t ? t[k] : 0.0, dt[k], AB + (size_t)k * n * (n + m));
        if (r) rc = r;
    }
    return rc;
}
// TO.rollout!: x[k+1] = discrete_dynamics(RK3, model, x[k], u[k], t, dt[k]); B trajectories
int orc_rollout(const rbq_model_params* p, int64_t B, int64_t N, const double* x0, const double* U, const double* dt, double* X) {
    int n = state_dim(p), m = control_dim(p); if (n < 0 || m < 0) return RBQ_E_MODEL;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        double* Xb = X + (size_t)b * N * n;
        std::memcpy(Xb, x0 + b * n, sizeof(double) * n);
        double time = 0.0;                                                  // knot times accumulate like a TO.Traj's
        for (int64_t k = 0; k + 1 < N; ++k) {
            step<double>(p, Xb + k * n, U + ((size_t)b * (N - 1) + k) * m, time, dt[k], Xb + (k + 1) * n);
            time = time + dt[k];
        }
    }
    return 0;
}
// closed-loop rollout of the iLQR forward pass (Altro `rollout!` with gains; external, from its
// published algorithm): du = K (x - xref) + alpha d
int orc_rollout_closedloop(const rbq_model_params* p, int64_t N, const double* Xref, const double* Uref, const double* K,
                           const double* d, const double* dt, int64_t nalpha, const double* alphas, double* X, double* Uout) {
    int n = state_dim(p), m = control_dim(p); if (n < 0 || m < 0) return RBQ_E_MODEL;
    for (int64_t c = 0; c < nalpha; ++c) {
        double* Xc = X + (size_t)c * N * n;
        double* Uc = Uout + (size_t)c * (N - 1) * m;
        std::memcpy(Xc, Xref, sizeof(double) * n);
        double time = 0.0;
        for (int64_t k = 0; k + 1 < N; ++k) {
            for (int j = 0; j < m; ++j) {
                double du = alphas[c] * d[k * m + j];
                for (int i = 0; i < n; ++i) du += K[(size_t)k * m * n + (size_t)i * m + j] * (Xc[k * n + i] - Xref[k * n + i]);
                Uc[k * m + j] = Uref[k * m + j] + du;
            }
            step<double>(p, Xc + k * n, Uc + k * m, time, dt[k], Xc + (k + 1) * n);
            time = time + dt[k];
        }
    }
    return 0;
}

// gate_error_iso2 / jacobian_gate_error_iso2 / hessian_gate_error_iso2, spin.jl:1040-1092 (row-major 4x4 Hessian)
void orc_gate_error_expansion(const double* s1, const double* s2, double* err, double* grad, double* hess) {
    const double a = s1[0] * s2[0] + s1[1] * s2[1] + s1[2] * s2[2] + s1[3] * s2[3];
    const double b = -s1[2] * s2[0] - s1[3] * s2[1] + s1[0] * s2[2] + s1[1] * s2[3];
    if (err) *err = 1 - a * a - b * b;
    if (grad) {
        const double a2 = 2 * a, b2 = 2 * b;
        grad[0] = -a2 * s2[0] - b2 * s2[2];
        grad[1] = -a2 * s2[1] - b2 * s2[3];
        grad[2] = -a2 * s2[2] + b2 * s2[0];
        grad[3] = -a2 * s2[3] + b2 * s2[1];
    }
    if (hess) {
        const double d11 = -2 * s2[0] * s2[0] - 2 * s2[2] * s2[2], d12 = -2 * s2[0] * s2[1] - 2 * s2[2] * s2[3], d13 = 0;
        const double d14 = 2 * s2[1] * s2[2] - 2 * s2[0] * s2[3], d22 = -2 * s2[1] * s2[1] - 2 * s2[3] * s2[3];
        const double d23 = -2 * s2[1] * s2[2] + 2 * s2[0] * s2[3], d24 = 0, d33 = d11, d34 = d12, d44 = d22;
        const double H[16] = {d11, d12, d13, d14, d12, d22, d23, d24, d13, d23, d33, d34, d14, d24, d34, d44};
        for (int i = 0; i < 16; ++i) hess[i] = H[i];
    }
}
double orc_fidelity_vec_iso2(const double* s1, const double* s2) { return fidelity_vec_iso2<double>(s1, s2); }
// m1, m2: 2x2 complex row-major interleaved
double orc_fidelity_mat(const double* m1, const double* m2, int closed) {
    cplx a[2][2], b[2][2];
    for (int i = 0; i < 4; ++i) { a[i / 2][i % 2] = cplx(m1[2 * i], m1[2 * i + 1]); b[i / 2][i % 2] = cplx(m2[2 * i], m2[2 * i + 1]); }
    return closed ? fidelity_mat_closed(a, b) : fidelity_mat(a, b);
}
int orc_gate_iso(int gate, double* G) {
    cplx g[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
    double M[4][4]; mat_iso(g, M);
    for (int i = 0; i < 16; ++i) G[i] = M[i / 4][i % 4];
    return 0;
}

// run_sim_prop(...; dynamics_type=schroed) over S samples (one reference call per sample).
// fid[S*(gate_count+1)], states[S*4] (final state) may be NULL.  Returns 0.
int orc_mc_static(const double* controls, int64_t nk, double dt, int gate, int64_t gate_count, int64_t S,
                  const double* fq, const double* da, const double* psi0, double* fid, double* states, int nthreads) {
    cplx g[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s)
        sim_schroed(controls, nk, dt, gate, gate_count, fq[s], da ? da[s] : 0.0, psi0 + 4 * s,
                    fid ? fid + s * (gate_count + 1) : nullptr, states ? states + 4 * s : nullptr);
    return 0;
}
int orc_mc_noise(const double* controls, int64_t nk, double dt, int gate, int64_t gate_count, int64_t S, double fq,
                 const double* noise, int64_t noise_len, double noise_dt_inv, const double* psi0, double* fid,
                 double* states, int nthreads) {
    cplx g[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s)
        sim_schroedda(controls, nk, dt, gate, gate_count, fq, noise + s * noise_len, noise_len, noise_dt_inv,
                      psi0 + 4 * s, fid ? fid + s * (gate_count + 1) : nullptr, states ? states + 4 * s : nullptr);
    return 0;
}
// rho0[S*8], fid/fid_closed[S*(gate_count+1)], rho_out[S*8]
int orc_lindblad_t1(const double* controls, int64_t nk, double dt, int gate, int64_t gate_count, int64_t S, double fq,
                    const double* da, const double* rho0, double* fid, double* fid_closed, double* rho_out, int nthreads) {
    cplx g[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        cplx r0[4], ro[4];
        for (int i = 0; i < 4; ++i) r0[i] = cplx(rho0[8 * s + 2 * i], rho0[8 * s + 2 * i + 1]);
        sim_lindbladt1(controls, nk, dt, gate, gate_count, fq, da ? da[s] : 0.0, r0,
                       fid ? fid + s * (gate_count + 1) : nullptr, fid_closed ? fid_closed + s * (gate_count + 1) : nullptr, ro);
        if (rho_out) for (int i = 0; i < 4; ++i) { rho_out[8 * s + 2 * i] = ro[i].real(); rho_out[8 * s + 2 * i + 1] = ro[i].imag(); }
    }
    return 0;
}
int orc_mc_static_segments(const double* amps, const double* durs, int64_t nseg, int gate, int64_t gate_count, int64_t S,
                           const double* fq, const double* da, const double* psi0, double* fid) {
    cplx g[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s)
        sim_schroed_segments(amps, durs, nseg, gate, gate_count, fq[s], da ? da[s] : 0.0, psi0 + 4 * s, fid + s * (gate_count + 1));
    return 0;
}
int orc_mc_noise_segments(const double* amps, const double* durs, const int32_t* nidx, int64_t nseg, int gate, int64_t gate_count,
                          int64_t S, double fq, const double* noise, int64_t noise_len, const double* psi0, double* fid) {
    cplx g[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
    for (int64_t j = 0; j < nseg; ++j) if (nidx[j] < 0 || nidx[j] >= noise_len) return RBQ_E_SIZE;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s)
        sim_schroedda_segments(amps, durs, nidx, nseg, gate, gate_count, fq, noise + s * noise_len, psi0 + 4 * s, fid + s * (gate_count + 1));
    return 0;
}
// run_sim_deqjl's ODE (dynamics_schroed_deqjl / dynamics_schroedda_deqjl, spin.jl:424-467) solved exactly segment by segment: the
// right-hand side is constant between breakpoints, so state <- exp(segment generator * duration) * state (generic Pade exp,
// extended-precision state, as in integrate_prop_schroedda!); fidelities at t = 0 and after save_after[g] segments
int orc_mc_timeline(const double* amps, const double* durs, const int32_t* nidx, int64_t nseg, const int64_t* save_after, int64_t nsave,
                    int gate, int64_t S, const double* fq, const double* da, const double* noise, int64_t noise_len,
                    const double* psi0, double* fid, double* states) {
    cplx g[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
    if (noise) for (int64_t j = 0; j < nseg; ++j) if (nidx[j] < 0 || nidx[j] >= noise_len) return RBQ_E_SIZE;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        xreal s0[4], st[4], tg[4][4];
        for (int i = 0; i < 4; ++i) { s0[i] = psi0[4 * s + i]; st[i] = s0[i]; }
        gate_targets(gate, s0, tg);
        double* f = fid + s * (nsave + 1);
        double* sr = states ? states + s * (nsave + 1) * 4 : nullptr;      // the states at the save points (run_sim_fine_deqjl)
        f[0] = (double)fidelity_vec_iso2<xreal>(st, tg[0]);
        if (sr) for (int i = 0; i < 4; ++i) sr[i] = (double)st[i];
        int64_t j = 0;
        for (int64_t q = 1; q <= nsave; ++q) {
            for (; j < save_after[q - 1]; ++j) {
                const double amp = amps[j] + (da ? da[s] : 0.0) + (noise ? noise[s * noise_len + nidx[j]] : 0.0);
                Mat4<double> H;
                for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) H.a[r][c] = durs[j] * (fq[s] * (PI * N0[r][c]) + amp * (PI * N1[r][c]));
                Mat4<double> U = expm_ref(H);
                xreal nx[4];
                for (int i = 0; i < 4; ++i) { xreal t = 0; for (int k = 0; k < 4; ++k) t += (xreal)U.a[i][k] * st[k]; nx[i] = t; }
                for (int i = 0; i < 4; ++i) st[i] = nx[i];
            }
            f[q] = (double)fidelity_vec_iso2<xreal>(st, tg[q % 4]);
            if (sr) for (int i = 0; i < 4; ++i) sr[4 * q + i] = (double)st[i];
        }
    }
    return 0;
}
// integrate_prop_ypiby2t1! / xpiby2t1! / zpiby2t1! generalised to any segment list (spin.jl:582-594, 651-670, 765-785)
int orc_lindblad_segments(const double* amps, const double* durs, int64_t nseg, int gate, int64_t gate_count, int64_t S, double fq,
                          const double* da, const double* rho0, double* fid_closed, double* rho_out) {
    cplx g[2][2], g2[2][2], g3[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) g2[i][j] = g[i][0] * g[0][j] + g[i][1] * g[1][j];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) g3[i][j] = g2[i][0] * g[0][j] + g2[i][1] * g[1][j];
    LindbladOps o = make_lindblad_ops(fq);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        Mat4<cplx> prop = eye4<cplx>();
        for (int64_t j = 0; j < nseg; ++j) {
            const double c1 = amps[j] + (da ? da[s] : 0.0);
            const double gamma1 = 1.0 / amp_t1_spline<double>(c1);
            Mat4<cplx> L;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) L.a[r][c] = (o.H0v[r][c] + c1 * o.H1v[r][c] + gamma1 * o.LOP[r][c]) * durs[j];
            prop = matmul(expm_ref(L), prop);
        }
        cplx v[4], r0[2][2], tg[4][2][2];
        for (int i = 0; i < 4; ++i) v[i] = cplx(rho0[8 * s + 2 * i], rho0[8 * s + 2 * i + 1]);
        r0[0][0] = v[0]; r0[1][0] = v[1]; r0[0][1] = v[2]; r0[1][1] = v[3];
        const cplx (*gs[3])[2] = {g, g2, g3};
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) tg[0][i][j] = r0[i][j];
        for (int q = 0; q < 3; ++q) {
            cplx t[2][2];
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i][j] = gs[q][i][0] * r0[0][j] + gs[q][i][1] * r0[1][j];
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) tg[q + 1][i][j] = t[i][0] * std::conj(gs[q][j][0]) + t[i][1] * std::conj(gs[q][j][1]);
        }
        for (int64_t k = 0; k <= gate_count; ++k) {
            if (k > 0) matvec(prop, v, v);
            cplx m[2][2] = {{v[0], v[2]}, {v[1], v[3]}};
            if (fid_closed) fid_closed[s * (gate_count + 1) + k] = fidelity_mat_closed(m, tg[k % 4]);
        }
        if (rho_out) for (int i = 0; i < 4; ++i) { rho_out[8 * s + 2 * i] = v[i].real(); rho_out[8 * s + 2 * i + 1] = v[i].imag(); }
    }
    return 0;
}
// ------------------------------------------------------------------------------------------
// The density-matrix ODEs the reference hands to DifferentialEquations.Vern9 (reltol = abstol = 1e-12, spin.jl:1277-1380):
//   dynamics_lindbladt1_deqjl (spin.jl:505-518)  and  dynamics_lindbladt2_deqjl (spin.jl:542-554)
// on the 4x4 REAL iso density  rho_iso = [[Re rho, -Im rho], [Im rho, Re rho]]:
//   d rho/dt = negi_h rho - rho negi_h
//            + gamma_1 (GE rho EG + NEG_EE_BY2 rho + rho NEG_EE_BY2) + gamma_1 (EG rho GE + NEG_GG_BY2 rho + rho NEG_GG_BY2)   [T1]
//            + (GAMMAC + 2 gammaf^2 t) NEG_DOP_ISO .* rho                                                                    [T2]
// with negi_h = fq NEGI_H0_ISO + a NEGI_H1_ISO and a piecewise-constant amplitude.  Between amplitude breakpoints the
// right-hand side is L_A rho + (t - t0) L_1 rho (L_1 = the 2 gammaf^2 dephasing mask), so the exact solution has the Taylor
// coefficients  (n+1) c_{n+1} = L_A c_n + L_1 c_{n-1},  rho(t0 + h) = sum c_n h^n,  summed here in long double until the terms
// vanish: the rounding-free solution of the reference's ODE, not an approximation of it by another integrator.
// Both channels may be switched on together (the engine offers T1 + T2; the reference has one function per channel).
// ------------------------------------------------------------------------------------------
typedef long double xr;
struct IsoDensityRhs {
    xr h[4][4];        // negi_h
    xr gamma1, gam0;   // T1 rate; dephasing rate at the segment start
    // T1 dissipator matrices of spin.jl:278-297 and the dephasing mask spin.jl:291-295
    void apply(const xr r[4][4], xr out[4][4]) const {
        static const int GE[4][4] = {{0, 1, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 1}, {0, 0, 0, 0}};
        static const int EG[4][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 1, 0}};
        static const int GGd[4] = {1, 0, 1, 0}, EEd[4] = {0, 1, 0, 1};      // diagonals of -2 NEG_GG_BY2_ISO, -2 NEG_EE_BY2_ISO
        static const int DOP[4][4] = {{0, 1, 0, 1}, {1, 0, 1, 0}, {0, 1, 0, 1}, {1, 0, 1, 0}};
        xr t1[4][4], t2[4][4];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
            xr s = 0;
            for (int k = 0; k < 4; ++k) s += h[i][k] * r[k][j] - r[i][k] * h[k][j];
            out[i][j] = s;
        }
        if (gamma1 != 0) {
            // GE rho EG and EG rho GE
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
                xr a = 0, b = 0;
                for (int k = 0; k < 4; ++k) { a += GE[i][k] * r[k][j]; b += EG[i][k] * r[k][j]; }
                t1[i][j] = a; t2[i][j] = b;
            }
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
                xr a = 0, b = 0;
                for (int k = 0; k < 4; ++k) { a += t1[i][k] * EG[k][j]; b += t2[i][k] * GE[k][j]; }
                out[i][j] += gamma1 * (a - (xr)0.5 * (EEd[i] + EEd[j]) * r[i][j]) + gamma1 * (b - (xr)0.5 * (GGd[i] + GGd[j]) * r[i][j]);
            }
        }
        if (gam0 != 0) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[i][j] -= gam0 * DOP[i][j] * r[i][j];
    }
};
// advance the iso density over one constant-amplitude segment [t0, t0 + dur)
static void iso_density_segment(xr rho[4][4], double fq, double amp, xr t0, xr dur, int channels, double gamma_c, double gamma_f) {
    static const int DOP[4][4] = {{0, 1, 0, 1}, {1, 0, 1, 0}, {0, 1, 0, 1}, {1, 0, 1, 0}};
    IsoDensityRhs f;
    const xr pil = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) f.h[i][j] = pil * ((xr)fq * N0[i][j] + (xr)amp * N1[i][j]);
    f.gamma1 = (channels & RBQ_LINDBLAD_T1) ? (xr)1.0 / (xr)amp_t1_spline<double>(amp) : 0;            // spin.jl:508
    const bool t2 = (channels & RBQ_LINDBLAD_T2) != 0;
    const xr g1coef = t2 ? (xr)2.0 * (xr)gamma_f * (xr)gamma_f : 0;                                      // spin.jl:552
    f.gam0 = t2 ? (xr)gamma_c + g1coef * t0 : 0;
    // sub-steps keep ||L h|| <= ~0.5 so that the series is short and free of cancellation
    const xr nrm = 2 * pil * (std::fabs((xr)fq) + std::fabs((xr)amp)) + 2 * f.gamma1 + f.gam0;
    int sub = (int)std::ceil((double)(nrm * dur / 0.5L)); if (sub < 1) sub = 1;
    const xr hh = dur / sub;
    for (int q = 0; q < sub; ++q) {
        xr cprev[4][4], c[4][4], acc[4][4], nxt[4][4];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { c[i][j] = rho[i][j]; cprev[i][j] = 0; acc[i][j] = rho[i][j]; }
        xr hp = 1;
        for (int n = 0; n < 60; ++n) {
            f.apply(c, nxt);
            xr big = 0;
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
                nxt[i][j] = (nxt[i][j] - g1coef * DOP[i][j] * cprev[i][j]) / (xr)(n + 1);
                cprev[i][j] = c[i][j];
            }
            hp *= hh;
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { c[i][j] = nxt[i][j]; acc[i][j] += c[i][j] * hp; big = std::max(big, std::fabs(c[i][j] * hp)); }
            if (big < 1e-24L && n > 2) break;
        }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) rho[i][j] = acc[i][j];
        f.gam0 += g1coef * hh;
    }
}
// Exact solution of the reference's density ODEs over a segment timeline with save points (the shape run_sim_deqjl /
// run_sim_fine_deqjl integrate, spin.jl:1277-1478).  rho0[S*8] column-major vec (re,im); fid_closed[S*(nsave+1)] Uhlmann
// sqrt-fidelity against the 4-cyclic targets (NULL to skip); rho_saves[S*(nsave+1)*8] (NULL to skip).
int orc_lindblad_ode(const double* amps, const double* durs, int64_t nseg, const int64_t* save_after, int64_t nsave, int gate,
                     int64_t S, double fq, const double* da, const double* rho0, int channels, double gamma_c, double gamma_f,
                     double* fid_closed, double* rho_saves, int nthreads) {
    cplx g[2][2], g2[2][2], g3[2][2]; if (!gate_matrix(gate, g)) return RBQ_E_GATE;
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) g2[i][j] = g[i][0] * g[0][j] + g[i][1] * g[1][j];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) g3[i][j] = g2[i][0] * g[0][j] + g2[i][1] * g[1][j];
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        cplx v[4], r0[2][2], tg[4][2][2];
        for (int i = 0; i < 4; ++i) v[i] = cplx(rho0[8 * s + 2 * i], rho0[8 * s + 2 * i + 1]);
        r0[0][0] = v[0]; r0[1][0] = v[1]; r0[0][1] = v[2]; r0[1][1] = v[3];
        const cplx (*gs[3])[2] = {g, g2, g3};
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) tg[0][i][j] = r0[i][j];
        for (int q = 0; q < 3; ++q) {
            cplx t[2][2];
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i][j] = gs[q][i][0] * r0[0][j] + gs[q][i][1] * r0[1][j];
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) tg[q + 1][i][j] = t[i][0] * std::conj(gs[q][j][0]) + t[i][1] * std::conj(gs[q][j][1]);
        }
        // get_mat_iso of the density (rbqoc.jl:204-210)
        xr rho[4][4];
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
            rho[i][j] = r0[i][j].real(); rho[i][j + 2] = -r0[i][j].imag(); rho[i + 2][j] = r0[i][j].imag(); rho[i + 2][j + 2] = r0[i][j].real();
        }
        auto record = [&](int64_t k) {
            cplx m[2][2];
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) m[i][j] = cplx((double)rho[i][j], (double)rho[i + 2][j]);
            if (fid_closed) fid_closed[s * (nsave + 1) + k] = fidelity_mat_closed(m, tg[k % 4]);
            if (rho_saves) {
                double* o = rho_saves + (s * (nsave + 1) + k) * 8;
                o[0] = m[0][0].real(); o[1] = m[0][0].imag(); o[2] = m[1][0].real(); o[3] = m[1][0].imag();
                o[4] = m[0][1].real(); o[5] = m[0][1].imag(); o[6] = m[1][1].real(); o[7] = m[1][1].imag();
            }
        };
        record(0);
        xr t = 0;
        int64_t j = 0;
        for (int64_t k = 1; k <= nsave; ++k) {
            for (; j < save_after[k - 1] && j < nseg; ++j) {
                iso_density_segment(rho, fq, amps[j] + (da ? da[s] : 0.0), t, (xr)durs[j], channels, gamma_c, gamma_f);
                t += (xr)durs[j];
            }
            record(k);
        }
    }
    return 0;
}
// the 4x4 complex Liouvillian of spin.jl:530-531 (row-major interleaved), for referee checks
void orc_lindblad_generator(double fq, double a, double* L) {
    LindbladOps o = make_lindblad_ops(fq);
    double gamma1 = 1.0 / amp_t1_spline<double>(a);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
        cplx v = o.H0v[r][c] + a * o.H1v[r][c] + gamma1 * o.LOP[r][c];
        L[2 * (4 * r + c)] = v.real(); L[2 * (4 * r + c) + 1] = v.imag();
    }
}

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void orc_philox_samples(int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma, double fq_rel_clip,
                        double da_sigma, double* fq, double* da, double* psi0) {
    for (int64_t s = 0; s < S; ++s) {
        if (psi0) sample_psi0((uint64_t)(first + s), seed, psi0 + 4 * s);
        double f, a; sample_static((uint64_t)(first + s), seed, fq0, fq_rel_sigma, fq_rel_clip, da_sigma, &f, &a);
        if (fq) fq[s] = f;
        if (da) da[s] = a;
    }
}
void orc_pink_filter(const double* xs, int64_t count, double dt_inv, double* ys) { pink_filter(xs, count, dt_inv, ys); }
// noise[S*noise_len] = namp * gen_pink_noise(noise_len, noise_dt_inv, N(0,1); Philox(seed, first+s))
void orc_pink_noise(int64_t first, int64_t S, uint64_t seed, double namp, double noise_dt_inv, int64_t noise_len, double* noise) {
    std::vector<double> xs(noise_len + 1);
    for (int64_t s = 0; s < S; ++s) {
        for (int64_t j = 0; 2 * j < noise_len; ++j) philox_normal2((uint64_t)(first + s), seed, (uint32_t)j, 2, &xs[2 * j], &xs[2 * j + 1]);
        pink_filter(xs.data(), noise_len, noise_dt_inv, noise + s * noise_len);
        for (int64_t i = 0; i < noise_len; ++i) noise[s * noise_len + i] *= namp;
    }
}

// statistics of infidelities e[S] exactly as rbq_stats defines them
void orc_stats(const double* e, int64_t S, rbq_stats* st) {
    std::memset(st, 0, sizeof(*st));
    st->min = INFINITY; st->max = -INFINITY;
    for (int64_t s = 0; s < S; ++s) {
        double v = e[s];
        if (!std::isfinite(v)) { st->nonfinite++; continue; }
        st->count++; st->sum += v; st->sumsq += v * v;
        st->min = std::min(st->min, v); st->max = std::max(st->max, v);
        double l = (std::log10(v > 0 ? v : 0.0) + 16.0) * 16.0;
        int b = v > 0 ? (int)std::floor(l) : 0;
        if (b < 0) b = 0; if (b > RBQ_HIST_BINS - 1) b = RBQ_HIST_BINS - 1;
        st->hist[b]++;
    }
}

}  // extern "C"
