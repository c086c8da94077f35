// rbq_jac_unscented.cu — [A_k B_k] of the unscented models (spin30.jl:86-202, spin25.jl:91-239) assembled from their block
// structure instead of generic forward mode (RD.discrete_jacobian! = ForwardDiff.jacobian over [x;u], external).
//
// The 56 x 57 (spin30) / 58 x 59 (spin25) block is mostly structure:
//   * nominal states:        blkdiag(E, E[, E]) and one column d/da = (dE/da) psi
//   * control chain / taps:  constants (1, dt, filter coefficients)
//   * sigma-point chunk:     41 x 40 dense block d(chunk')/d(chunk) + the column d/da (+ spin25: d/d(filter taps), all multiples of
//                            ONE column d/d(noise amplitude)) + 4 entries of the penalty row against the nominal state
// Only the chunk needs differentiation, and a seed in sigma point j0 reaches the outputs through
//     propagate (a column m of M_j0)  ->  mean (dsm = m / 10)  ->  covariance (dP = cf sym(m d_j0'), because the deviations sum to 0)
//     ->  Cholesky (dL from dP by forward substitution, no square roots: 1 / L_jj is a value)  ->  resample  ->  normalise  ->  penalty,
// ~250 multiply-adds instead of the ~1500 a generic tangent pushes through the whole 56-state step.  So:
//   unscented_values_kernel   one THREAD per knot: the value path once (not once per seed column), the intermediate values
//                             the tangents need written as a 136-double record
//   unscented_tangent_kernel  one 64-thread block per knot (persistent over knots): lanes 0..39 the sigma-point columns, the
//                             rest the structural entries; the dense directions (d/da, spin25: d/d noise amplitude and the
//                             tap columns) are linear combinations of the sigma-point columns (a 41 x 40 mat-vec out of shared
//                             memory); the block is assembled in shared memory (zeros written once: the sparsity pattern never
//                             changes) and leaves by ONE bulk TMA store (cp.async.bulk.global.shared::cta), double buffered
//                             against the next knot's arithmetic; the next knot's record is prefetched into registers.
// Derivatives are exact derivatives of the closed-form step; they equal the reference's dual numbers to ~1e-14 (tests: 1e-10).
// sm_100a only.
#include "rbq_internal.h"
#include "rbq_models.cuh"
#include "rbq_tma.cuh"

namespace rbq {

// record layout (doubles)
constexpr int UR_E = 0;       // E, Ep, En as (c, sp, sq): 9
constexpr int UR_DE = 9;      // d/d(amplitude) of the same: 9
constexpr int UR_D = 18;      // deviations d_j = s_j - mean: 10 x 4
constexpr int UR_L = 58;      // Cholesky factor, packed lower (00 10 11 20 21 22 30 31 32 33)
constexpr int UR_INV = 68;    // 1 / L_jj
constexpr int UR_U = 72;      // the 10 normalised outputs
constexpr int UR_N = 112;     // their 1 / norm before normalisation
constexpr int UR_DN = 122;    // mean - nominal state
constexpr int UR_YK = 126;    // spin25: d yk / d (xp1, xp2, xp3, yp1, yp2, yp3)
constexpr int UR_W = 136;     // w_j = M_j' (dM_j/da) s_j^in: 10 x 4 (the amplitude direction in sigma-point coordinates)
constexpr int UR_ST = 176;    // values of the structural entries, in the order of structural_positions()
constexpr int UR_SIZE = 256;
constexpr int UJ_NST30 = 70, UJ_NST25 = 60;
// destination (column-major index into the n x (n+m) block) of structural entry e; the values kernel writes the values in
// the same order.  spin30: 48 E-block entries, 12 amplitude-column entries of the nominal states, 6 chain entries, 4 penalty
// entries against the nominal state; spin25: 32 + 8 + 6 + 4, then the yk row (6) and the four tap shifts.
__constant__ unsigned short c_stpos[2][3][72];      // [spin25?][nominal block][entry]: every upload of a slice writes the same values
static void structural_positions(bool s25, int nom, unsigned short* pos) {
    const int N = s25 ? 58 : 56, OFF = s25 ? 17 : 15, ACOL = s25 ? 9 : 13, NB = s25 ? 2 : 3, CB = s25 ? 8 : 12;
    int e = 0;
    for (int b = 0; b < NB; ++b) for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) pos[e++] = (unsigned short)((4 * b + c) * N + 4 * b + r);
    for (int b = 0; b < NB; ++b) for (int r = 0; r < 4; ++r) pos[e++] = (unsigned short)(ACOL * N + 4 * b + r);
    for (int r = 0; r < 3; ++r) {
        pos[e++] = (unsigned short)((CB + r) * N + CB + r);                       // 1 on the diagonal
        pos[e++] = (unsigned short)((r == 2 ? N : CB + r + 1) * N + CB + r);      // dt next to it (the last one in the control column)
    }
    for (int i = 0; i < 4; ++i) pos[e++] = (unsigned short)((nom + i) * N + OFF + 40);
    if (s25) {
        for (int q = 0; q < 6; ++q) pos[e++] = (unsigned short)((11 + q) * N + 14);
        const int rr[4] = {12, 13, 15, 16};
        for (int q = 0; q < 4; ++q) pos[e++] = (unsigned short)((rr[q] - 1) * N + rr[q]);
    }
}
#define URTRI(i, k) ((i) * ((i) + 1) / 2 + (k))

// entry (r, c) of the 4x4 iso matrix c I + sp N0 + sq N1 (rows of prop_apply)
RBQ_DEV double iso_entry(const double* e3, int r, int c) {
    const int idx = 4 * r + c;
    const double cc = e3[0], sp = e3[1], sq = e3[2];
    switch (idx) {
    case 0: case 5: case 10: case 15: return cc;
    case 2: case 13: return sp;
    case 7: case 8: return -sp;
    case 3: case 6: return sq;
    case 9: case 12: return -sq;
    default: return 0.0;
    }
}
RBQ_DEV void iso_apply(const double* e3, const double* x, double* y) {
    y[0] = fma(e3[2], x[3], fma(e3[1], x[2], e3[0] * x[0]));
    y[1] = fma(e3[2], x[2], fma(-e3[1], x[3], e3[0] * x[1]));
    y[2] = fma(-e3[2], x[1], fma(-e3[1], x[0], e3[0] * x[2]));
    y[3] = fma(-e3[2], x[0], fma(e3[1], x[1], e3[0] * x[3]));
}

struct UnscentedJacArgs {
    rbq_model_params p;
    int n, m;
    int64_t K;
    const double* X; const double* U; const double* tk; const double* dt;
    double* rec;     // [K][UR_SIZE]
    double* AB;      // [K][n (n+m)]
};

template <int MODEL>
__global__ void __launch_bounds__(64) unscented_values_kernel(const UnscentedJacArgs a) {
    constexpr bool S25 = MODEL == RBQ_MODEL_SPIN25;
    constexpr int OFF = S25 ? 17 : 15, ACOL = S25 ? 9 : 13;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const rbq_model_params& p = a.p;
    const double* x = a.X + k * a.n;
    double* rec = a.rec + k * UR_SIZE;
    const double dt = a.dt[k], amp = x[ACOL];
    // the three propagators and their derivatives with respect to the amplitude they are evaluated at
    double fE = p.fq, fP = p.fq, fN = p.fq, aP = amp, aN = amp;
    if (S25) {
        // pink-filter output of this knot (spin25.jl:194-219) and its dependence on the taps
        const double xk = sqrt(p.fq_cov);
        const long long kp = knot_point_of(a.tk ? a.tk[k] : 0.0, dt);
        const double xp1 = x[11], xp2 = x[12], xp3 = x[13], yp1 = x[14], yp2 = x[15], yp3 = x[16];
        double yk = 0.049922035 * xk;
        double cf6[6] = {0, 0, 0, 0, 0, 0};
        if (kp == 1) {
        } else if (kp == 2) {
            yk = yk + xp1 * (-0.095993537) - yp1 * (-2.494956002);
            cf6[0] = -0.095993537; cf6[3] = 2.494956002;
        } else if (kp == 3) {
            yk = yk + xp1 * (-0.095993537) + xp2 * 0.050612699 - yp1 * (-2.494956002) - yp2 * 2.017265875;
            cf6[0] = -0.095993537; cf6[1] = 0.050612699; cf6[3] = 2.494956002; cf6[4] = -2.017265875;
        } else {
            yk = yk + xp1 * (-0.095993537) + xp2 * 0.050612699 + xp3 * (-0.004408786) - yp1 * (-2.494956002) - yp2 * 2.017265875 -
                 yp3 * (-0.522189400);
            cf6[0] = -0.095993537; cf6[1] = 0.050612699; cf6[2] = -0.004408786; cf6[3] = 2.494956002; cf6[4] = -2.017265875; cf6[5] = 0.522189400;
        }
        const double namp = yk * p.namp;
        aP = amp + namp; aN = amp - namp;
#pragma unroll
        for (int i = 0; i < 6; ++i) rec[UR_YK + i] = cf6[i];     // d yk / d tap (the namp factor is applied by the tangent kernel)
    } else {
        const double sig = sqrt(p.fq_cov);
        fP = p.fq + sig; fN = p.fq - sig;
    }
    double e[9], de[9];
    {
        const double fs[3] = {fE, fP, fN}, as[3] = {amp, aP, aN};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            Tan<1> ta; ta.v = as[q]; ta.d[0] = 1.0;
            const Prop<Tan<1>> E = prop_of<Tan<1>, double>(fs[q], ta, dt);
            e[3 * q] = E.c.v; e[3 * q + 1] = E.sp.v; e[3 * q + 2] = E.sq.v;
            de[3 * q] = E.c.d[0]; de[3 * q + 1] = E.sp.d[0]; de[3 * q + 2] = E.sq.d[0];
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) { rec[UR_E + i] = e[i]; rec[UR_DE + i] = de[i]; }
    // value path of the chunk (same operations as unscented_chunk in rbq_models.cuh)
    double s[10][4], sm[4];
#pragma unroll
    for (int j = 0; j < 10; ++j) iso_apply(e + (j == 4 ? 3 : (j == 9 ? 6 : 0)), x + OFF + 4 * j, s[j]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double acc = s[0][i];
#pragma unroll
        for (int j = 1; j < 10; ++j) acc += s[j][i];
        sm[i] = acc * 0.1;
    }
#pragma unroll
    for (int j = 0; j < 10; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[j][i] -= sm[i]; rec[UR_D + 4 * j + i] = s[j][i]; }
    const double cf = 1.0 / (2.0 * p.alpha * p.alpha);
    double P[10], L[10], inv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c <= i; ++c) {
            double acc = s[0][i] * s[0][c];
#pragma unroll
            for (int j = 1; j < 10; ++j) acc = fma(s[j][i], s[j][c], acc);
            P[URTRI(i, c)] = acc * cf;
        }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double d = P[URTRI(j, j)];
#pragma unroll
        for (int c = 0; c < j; ++c) d = fma(-L[URTRI(j, c)], L[URTRI(j, c)], d);
        const double ljj = sqrt(d);
        L[URTRI(j, j)] = ljj; inv[j] = 1.0 / ljj;
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double t = P[URTRI(i, j)];
#pragma unroll
            for (int c = 0; c < j; ++c) t = fma(-L[URTRI(i, c)], L[URTRI(j, c)], t);
            L[URTRI(i, j)] = t * inv[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) rec[UR_L + i] = L[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) rec[UR_INV + i] = inv[i];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        double vp[4], vm[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double col = (j < 4 && i >= j) ? L[URTRI(i, j)] * p.alpha : 0.0;
            vp[i] = sm[i] + col; vm[i] = sm[i] - col;
        }
        const double np = 1.0 / sqrt(vp[0] * vp[0] + vp[1] * vp[1] + vp[2] * vp[2] + vp[3] * vp[3]);
        const double nm = 1.0 / sqrt(vm[0] * vm[0] + vm[1] * vm[1] + vm[2] * vm[2] + vm[3] * vm[3]);
        rec[UR_N + j] = np; rec[UR_N + j + 5] = nm;
#pragma unroll
        for (int i = 0; i < 4; ++i) { rec[UR_U + 4 * j + i] = vp[i] * np; rec[UR_U + 4 * (j + 5) + i] = vm[i] * nm; }
    }
    const int nom = p.nominal_offset[0];
    double dn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { dn[i] = sm[i] - x[nom + i]; rec[UR_DN + i] = dn[i]; }
    // the amplitude direction in sigma-point coordinates: w_j = M_j' (dM_j/da) s_j^in (spin25: its entries 16..19 and 36..39
    // are also the noise-amplitude direction, with opposite signs)
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const int sel = j == 4 ? 3 : (j == 9 ? 6 : 0);
        double t[4], w[4];
        iso_apply(de + sel, x + OFF + 4 * j, t);
        const double et[3] = {e[sel], -e[sel + 1], -e[sel + 2]};      // the transpose
        iso_apply(et, t, w);
#pragma unroll
        for (int i = 0; i < 4; ++i) rec[UR_W + 4 * j + i] = w[i];
    }
    // structural entries, in the order of structural_positions()
    constexpr int NB = S25 ? 2 : 3;
    double* st = rec + UR_ST;
    int q = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) st[q++] = iso_entry(e, r, c);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        double t[4];
        iso_apply(de, x + 4 * b, t);
#pragma unroll
        for (int r = 0; r < 4; ++r) st[q++] = t[r];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { st[q++] = 1.0; st[q++] = dt; }
#pragma unroll
    for (int i = 0; i < 4; ++i) st[q++] = -2.0 * p.S_diag[i] * dn[i];
    if (S25) {
#pragma unroll
        for (int i = 0; i < 6; ++i) st[q++] = rec[UR_YK + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) st[q++] = 1.0;
    }
}

// Given the tangent of the mean (dsm) and of the packed covariance (dP), the tangents of the chunk's 41 outputs.
// rec: the knot's record in shared memory.  out: 41 consecutive doubles (a column segment of the block).
RBQ_DEV void chunk_tangent_tail(const double* __restrict__ rec, double alpha, const double* Sd, const double (&dsm)[4],
                                const double (&dP)[10], double* __restrict__ out, double scale) {
    const double* L = rec + UR_L;
    const double* inv = rec + UR_INV;
    double dL[10];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double dd = dP[URTRI(j, j)];
#pragma unroll
        for (int c = 0; c < j; ++c) dd = fma(-2.0 * L[URTRI(j, c)], dL[URTRI(j, c)], dd);
        const double dljj = 0.5 * dd * inv[j];
        dL[URTRI(j, j)] = dljj;
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double t = dP[URTRI(i, j)];
#pragma unroll
            for (int c = 0; c < j; ++c) t = fma(-dL[URTRI(i, c)], L[URTRI(j, c)], fma(-L[URTRI(i, c)], dL[URTRI(j, c)], t));
            dL[URTRI(i, j)] = fma(-L[URTRI(i, j)], dljj, t) * inv[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
#pragma unroll
        for (int sgn = 0; sgn < 2; ++sgn) {
            const int jj = j + 5 * sgn;
            double dv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double col = (j < 4 && i >= j) ? alpha * dL[URTRI(i, j)] : 0.0;
                dv[i] = sgn ? dsm[i] - col : dsm[i] + col;
            }
            const double* u = rec + UR_U + 4 * jj;
            const double nn = rec[UR_N + jj] * scale;
            const double ud = fma(u[3], dv[3], fma(u[2], dv[2], fma(u[1], dv[1], u[0] * dv[0])));
#pragma unroll
            for (int i = 0; i < 4; ++i) out[4 * jj + i] = nn * fma(-u[i], ud, dv[i]);
        }
    }
    const double* dn = rec + UR_DN;
    double pen = Sd[0] * dP[0] + Sd[1] * dP[2] + Sd[2] * dP[5] + Sd[3] * dP[9];
#pragma unroll
    for (int i = 0; i < 4; ++i) pen = fma(2.0 * Sd[i] * dn[i], dsm[i], pen);
    out[40] = pen * scale;
}

RBQ_DEV void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
RBQ_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> RBQ_DEV void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
RBQ_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int UJ_THREADS = 64;
// The dense directions need no differentiation of their own.  The chunk's outputs depend on the propagated sigma points s_j
// only, and a seed e_i in sigma point j gives ds_j = M_j e_i (the 40 sparse columns).  A change of the amplitude gives
// ds_j = (dM_j/da) s_j^in = M_j w_j with w_j = M_j' (dM_j/da) s_j^in (M_j is orthogonal), so by linearity
//     column(a) = sum_{j,i} w_j[i] column(j, i):
// a 41 x 40 matrix-vector product over the columns already sitting in the staging buffer, one row per lane.  The same holds
// for spin25's noise amplitude (only points 5 and 10 move, with opposite signs), of which the six tap columns are multiples.
// One staging buffer per block and as many blocks per SM as shared memory holds (8 / 7): the kernel is a chain of short
// dependent sections, and knots in flight are what hides them.
template <int MODEL>
__global__ void __launch_bounds__(UJ_THREADS) unscented_tangent_kernel(const UnscentedJacArgs a) {
    constexpr bool S25 = MODEL == RBQ_MODEL_SPIN25;
    constexpr int OFF = S25 ? 17 : 15, ACOL = S25 ? 9 : 13;
    constexpr int N = S25 ? 58 : 56, NM = N + 1, BLK = N * NM, NST = S25 ? UJ_NST25 : UJ_NST30;
    constexpr int NREC = UR_SIZE / UJ_THREADS;      // record doubles each thread moves (256 / 64)
    extern __shared__ __align__(128) double sm_[];
    double* stage = sm_;                        // [BLK]
    double* rec = sm_ + BLK;                    // [UR_SIZE]
    const int tid = threadIdx.x;
    const rbq_model_params& p = a.p;
    for (int i = tid; i < BLK; i += UJ_THREADS) stage[i] = 0.0;
    const double alpha = p.alpha, cf = 1.0 / (2.0 * p.alpha * p.alpha);
    const double Sd[4] = {p.S_diag[0], p.S_diag[1], p.S_diag[2], p.S_diag[3]};
    const int nomb = p.nominal_offset[0] >> 2;
    // the next knot's record travels in registers while this knot is being worked on
    double nrec[NREC];
    auto fetch = [&](int64_t k) {
        if (k < a.K) {
#pragma unroll
            for (int q = 0; q < NREC; ++q) nrec[q] = a.rec[k * UR_SIZE + tid + q * UJ_THREADS];
        }
    };
    fetch(blockIdx.x);
    for (int64_t k = blockIdx.x; k < a.K; k += gridDim.x) {
#pragma unroll
        for (int q = 0; q < NREC; ++q) rec[tid + q * UJ_THREADS] = nrec[q];
        // the bulk store of the previous knot must have finished READING the staging buffer
        if (tid == 0) bulk_wait_read<0>();
        __syncthreads();
        fetch(k + gridDim.x);
        if (tid < 40) {
            // seed in sigma point j0, component i0: ds = column i0 of M_j0
            const int j0 = tid >> 2, i0 = tid & 3;
            const double* e3 = rec + UR_E + (j0 == 4 ? 3 : (j0 == 9 ? 6 : 0));
            double mcol[4], dsm[4], dP[10];
#pragma unroll
            for (int i = 0; i < 4; ++i) { mcol[i] = iso_entry(e3, i, i0); dsm[i] = 0.1 * mcol[i]; }
            const double* dj = rec + UR_D + 4 * j0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c <= i; ++c) dP[URTRI(i, c)] = cf * fma(mcol[i], dj[c], dj[i] * mcol[c]);
            chunk_tangent_tail(rec, alpha, Sd, dsm, dP, stage + (size_t)(OFF + tid) * N + OFF, 1.0);
        }
        // structural entries: values from the record, destinations from the constant table
        for (int e = tid; e < NST; e += UJ_THREADS) stage[c_stpos[S25 ? 1 : 0][nomb][e]] = rec[UR_ST + e];
        __syncthreads();
        if (tid < 41) {
            // amplitude column of the chunk rows: one row per lane, sum over the 40 sigma-point columns
            const double* rowp = stage + (size_t)OFF * N + OFF + tid;
            const double* wv = rec + UR_W;
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
            for (int c = 0; c < 40; c += 4) {
                acc0 = fma(rowp[(size_t)c * N], wv[c], acc0);
                acc1 = fma(rowp[(size_t)(c + 1) * N], wv[c + 1], acc1);
                acc2 = fma(rowp[(size_t)(c + 2) * N], wv[c + 2], acc2);
                acc3 = fma(rowp[(size_t)(c + 3) * N], wv[c + 3], acc3);
            }
            stage[(size_t)ACOL * N + OFF + tid] = (acc0 + acc1) + (acc2 + acc3);
            if (S25) {
                // noise amplitude: points 5 (+) and 10 (-) only; the six tap columns are namp * (d yk / d tap) times it
                double an = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) an = fma(rowp[(size_t)(16 + i) * N], wv[16 + i], fma(-rowp[(size_t)(36 + i) * N], wv[36 + i], an));
#pragma unroll
                for (int f = 0; f < 6; ++f) stage[(size_t)(11 + f) * N + OFF + tid] = p.namp * rec[UR_YK + f] * an;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(a.AB + (size_t)k * BLK, stage, (uint32_t)(BLK * sizeof(double)));
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_read<0>();
}

size_t unscented_record_bytes(int64_t K) { return sizeof(double) * UR_SIZE * (size_t)K; }

// preconditions (checked by the caller): model spin30 / spin25 with ONE sample state, d_AB 16-byte aligned
int launch_jacobians_unscented(rbq_handle* h, const rbq_model_params& p, int64_t K, const double* d_X, const double* d_U,
                               const double* d_t, const double* d_dt, double* d_rec, double* d_AB) {
    const int n = model_state_dim(p), m = model_control_dim(p);
    const bool s25 = p.model_id == RBQ_MODEL_SPIN25;
    UnscentedJacArgs a{p, n, m, K, d_X, d_U, d_t, d_dt, d_rec, d_AB};
    {   // destinations of the structural entries (they depend on the nominal offset only); 144 bytes, stream ordered
        unsigned short pos[72] = {0};
        structural_positions(s25, p.nominal_offset[0], pos);
        const int nb = p.nominal_offset[0] >> 2;
        if (nb < 0 || nb > 2 || (p.nominal_offset[0] & 3)) return RBQ_E_MODEL;
        cudaError_t e = cudaMemcpyToSymbolAsync(c_stpos, pos, sizeof(pos), ((s25 ? 3 : 0) + nb) * sizeof(pos), cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) return (int)e;
    }
    const unsigned vblocks = (unsigned)((K + 63) / 64);
    const size_t smem = sizeof(double) * ((size_t)n * (n + m) + UR_SIZE);
    const int64_t resident = (int64_t)h->sm_count * (s25 ? 7 : 8);
    const unsigned tblocks = (unsigned)(K < resident ? K : resident);
    cudaError_t e;
    if (!s25) {
        unscented_values_kernel<RBQ_MODEL_SPIN30><<<vblocks, 64, 0, h->stream>>>(a);
        e = cudaFuncSetAttribute(unscented_tangent_kernel<RBQ_MODEL_SPIN30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        unscented_tangent_kernel<RBQ_MODEL_SPIN30><<<tblocks, UJ_THREADS, smem, h->stream>>>(a);
    } else {
        unscented_values_kernel<RBQ_MODEL_SPIN25><<<vblocks, 64, 0, h->stream>>>(a);
        e = cudaFuncSetAttribute(unscented_tangent_kernel<RBQ_MODEL_SPIN25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        unscented_tangent_kernel<RBQ_MODEL_SPIN25><<<tblocks, UJ_THREADS, smem, h->stream>>>(a);
    }
    h->launches += 2;
    return (int)cudaGetLastError();
}

}  // namespace rbq
