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
//   unscented_jac_kernel      one 64-thread block per knot (persistent over knots): the value path ONCE per knot, spread over the
//                             lanes; lanes 0..39 the sigma-point columns; the dense directions (d/da, spin25: d/d noise
//                             amplitude and the tap columns) as linear combinations of the sigma-point columns; the block is
//                             assembled in shared memory and leaves by ONE bulk TMA store (cp.async.bulk.global.shared::cta).
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
// ONE kernel, one 64-thread block per knot (persistent over knots), in three stages:
//   values    the value path of the unscented transform, spread over the lanes in seven short phases (propagators with their
//             amplitude derivatives on 3 lanes, the 10 sigma points on 10 lanes, deviations on 40, covariance on 10, the 4 x 4
//             Cholesky on one lane -- rsqrt, no division --, the 10 normalised outputs on 10 lanes); the intermediate values the
//             tangents need stay in shared memory (the "record")
//   tangents  lanes 0..39: one sigma-point column each through chunk_tangent_tail; every lane: its share of the structural entries
//   dense     The dense directions need no differentiation of their own: the chunk's outputs depend on the propagated sigma
//             points s_j only, a seed e_i in sigma point j gives ds_j = M_j e_i (the sparse columns), and a change of the amplitude
//             gives ds_j = (dM_j/da) s_j^in = M_j w_j with w_j = M_j' (dM_j/da) s_j^in (M_j orthogonal), so
//                 column(a) = sum_{j,i} w_j[i] column(j, i),
//             a 41 x 40 matrix-vector product over the columns already in the staging buffer, one row per lane.  Likewise
//             spin25's noise amplitude (points 5 and 10 only, opposite signs), of which the six tap columns are multiples.
// The block is assembled in shared memory (zeros written once: the sparsity pattern never changes) and leaves by ONE bulk TMA
// store; one staging buffer per block and as many blocks per SM as shared memory holds (8 / 7): the kernel is a chain of
// short dependent sections, and knots in flight are what hides them.  The next knot's state is prefetched into registers.
template <int MODEL>
__global__ void __launch_bounds__(UJ_THREADS) unscented_jac_kernel(const UnscentedJacArgs a) {
    constexpr bool S25 = MODEL == RBQ_MODEL_SPIN25;
    constexpr int OFF = S25 ? 17 : 15, ACOL = S25 ? 9 : 13, NB = S25 ? 2 : 3, CB = S25 ? 8 : 12;
    constexpr int N = S25 ? 58 : 56, NM = N + 1, BLK = N * NM;
    extern __shared__ __align__(128) double sm_[];
    double* stage = sm_;                        // [BLK]
    double* rec = sm_ + BLK;                    // [UR_W + 40]: the record (layout UR_*)
    double* xs = rec + UR_W + 40;               // [N + 2]: x, dt, t
    double* sv = xs + N + 2;                    // [40] propagated sigma points
    double* smv = sv + 40;                      // [4] their mean
    double* Pv = smv + 4;                       // [10] covariance, packed lower
    const int tid = threadIdx.x;
    const rbq_model_params& p = a.p;
    for (int i = tid; i < BLK; i += UJ_THREADS) stage[i] = 0.0;
    const double alpha = p.alpha, cf = 1.0 / (2.0 * p.alpha * p.alpha);
    const double Sd[4] = {p.S_diag[0], p.S_diag[1], p.S_diag[2], p.S_diag[3]};
    const int nom = p.nominal_offset[0];
    double nx = 0.0;
    auto fetch = [&](int64_t k) {
        if (k < a.K) nx = tid < N ? a.X[k * N + tid] : (tid == N ? a.dt[k] : ((tid == N + 1 && a.tk) ? a.tk[k] : 0.0));
    };
    fetch(blockIdx.x);
    for (int64_t k = blockIdx.x; k < a.K; k += gridDim.x) {
        if (tid < N + 2) xs[tid] = nx;
        // the bulk store of the previous knot must have finished READING the staging buffer
        if (tid == 0) bulk_wait_read<0>();
        __syncthreads();
        fetch(k + gridDim.x);
        const double dt = xs[N];
        // ---- values 1: the three propagators and their derivatives with respect to the amplitude they are evaluated at
        if (tid < 3) {
            const double amp = xs[ACOL];
            double f = p.fq, am = amp;
            if (S25) {
                // pink-filter output of this knot (spin25.jl:194-219) and its dependence on the taps
                const double xk = sqrt(p.fq_cov);
                const long long kp = knot_point_of(xs[N + 1], dt);
                const double xp1 = xs[11], xp2 = xs[12], xp3 = xs[13], yp1 = xs[14], yp2 = xs[15], yp3 = xs[16];
                double yk = 0.049922035 * xk;
                double c6[6] = {0, 0, 0, 0, 0, 0};
                if (kp == 1) {
                } else if (kp == 2) {
                    yk = yk + xp1 * (-0.095993537) - yp1 * (-2.494956002);
                    c6[0] = -0.095993537; c6[3] = 2.494956002;
                } else if (kp == 3) {
                    yk = yk + xp1 * (-0.095993537) + xp2 * 0.050612699 - yp1 * (-2.494956002) - yp2 * 2.017265875;
                    c6[0] = -0.095993537; c6[1] = 0.050612699; c6[3] = 2.494956002; c6[4] = -2.017265875;
                } else {
                    yk = yk + xp1 * (-0.095993537) + xp2 * 0.050612699 + xp3 * (-0.004408786) - yp1 * (-2.494956002) - yp2 * 2.017265875 -
                         yp3 * (-0.522189400);
                    c6[0] = -0.095993537; c6[1] = 0.050612699; c6[2] = -0.004408786; c6[3] = 2.494956002; c6[4] = -2.017265875; c6[5] = 0.522189400;
                }
                const double namp = yk * p.namp;
                am = tid == 0 ? amp : (tid == 1 ? amp + namp : amp - namp);
                if (tid == 0) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) rec[UR_YK + i] = c6[i];
                }
            } else {
                const double sig = sqrt(p.fq_cov);      // a deviation WITHOUT alpha, spin30.jl:102,108,113
                f = tid == 0 ? p.fq : (tid == 1 ? p.fq + sig : p.fq - sig);
            }
            Tan<1> ta; ta.v = am; ta.d[0] = 1.0;
            const Prop<Tan<1>> E = prop_of<Tan<1>, double>(f, ta, dt);
            rec[UR_E + 3 * tid] = E.c.v; rec[UR_E + 3 * tid + 1] = E.sp.v; rec[UR_E + 3 * tid + 2] = E.sq.v;
            rec[UR_DE + 3 * tid] = E.c.d[0]; rec[UR_DE + 3 * tid + 1] = E.sp.d[0]; rec[UR_DE + 3 * tid + 2] = E.sq.d[0];
        }
        __syncthreads();
        // ---- values 2: lanes 0..9 propagate the sigma points, lanes 16..25 the amplitude direction w_j, lanes 32.. structure
        if (tid < 10) {
            iso_apply(rec + UR_E + (tid == 4 ? 3 : (tid == 9 ? 6 : 0)), xs + OFF + 4 * tid, sv + 4 * tid);
        } else if (tid >= 16 && tid < 26) {
            const int j = tid - 16, sel = j == 4 ? 3 : (j == 9 ? 6 : 0);
            double t[4];
            iso_apply(rec + UR_DE + sel, xs + OFF + 4 * j, t);
            const double et[3] = {rec[UR_E + sel], -rec[UR_E + sel + 1], -rec[UR_E + sel + 2]};     // the transpose
            iso_apply(et, t, rec + UR_W + 4 * j);
        } else if (tid >= 32) {
            constexpr int E_END = 16 * NB, A_END = E_END + 4 * NB, C_END = A_END + 6, F_END = C_END + (S25 ? 10 : 0);
            for (int e = tid - 32; e < F_END; e += 32) {
                if (e < E_END) {                                   // blkdiag(E, E[, E])
                    const int b = e >> 4, r = (e >> 2) & 3, c = e & 3;
                    stage[(size_t)(4 * b + c) * N + 4 * b + r] = iso_entry(rec + UR_E, r, c);
                } else if (e < A_END) {                            // (dE/da) psi_b in the amplitude column
                    const int q = e - E_END, b = q >> 2, r = q & 3;
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc = fma(iso_entry(rec + UR_DE, r, c), xs[4 * b + c], acc);
                    stage[(size_t)ACOL * N + 4 * b + r] = acc;
                } else if (e < C_END) {                            // [int a, a, da] Euler chain (spin13.jl:49-51) and B = dt e_last
                    const int q = e - A_END, r = q >> 1;
                    if (q & 1) stage[(size_t)(r == 2 ? N : CB + r + 1) * N + CB + r] = dt;
                    else stage[(size_t)(CB + r) * N + CB + r] = 1.0;
                } else if (S25) {                                  // pink-filter taps (spin25.jl:194-219)
                    const int q = e - C_END;
                    if (q < 6) stage[(size_t)(11 + q) * N + 14] = rec[UR_YK + q];                  // yk row
                    else {
                        const int rr = q == 6 ? 12 : (q == 7 ? 13 : (q == 8 ? 15 : 16));             // tap shifts: (row, row - 1) = 1
                        stage[(size_t)(rr - 1) * N + rr] = 1.0;
                    }
                }
            }
        }
        __syncthreads();
        // ---- values 3: mean and deviations (lane = sigma point j, component i)
        if (tid < 40) {
            const int i = tid & 3;
            double acc = sv[i];
#pragma unroll
            for (int j = 1; j < 10; ++j) acc += sv[4 * j + i];
            const double m = acc * 0.1;
            rec[UR_D + tid] = sv[tid] - m;
            if (tid < 4) smv[tid] = m;
        }
        __syncthreads();
        // ---- values 4: covariance, packed lower triangle
        if (tid < 10) {
            const int i = tid < 1 ? 0 : (tid < 3 ? 1 : (tid < 6 ? 2 : 3)), c = tid - i * (i + 1) / 2;
            double acc = rec[UR_D + i] * rec[UR_D + c];
#pragma unroll
            for (int j = 1; j < 10; ++j) acc = fma(rec[UR_D + 4 * j + i], rec[UR_D + 4 * j + c], acc);
            Pv[tid] = acc * cf;
        } else if (tid >= 16 && tid < 20) {
            // penalty row against the nominal state of the INPUT
            const int i = tid - 16;
            const double dn = smv[i] - xs[nom + i];
            rec[UR_DN + i] = dn;
            stage[(size_t)(nom + i) * N + OFF + 40] = -2.0 * p.S_diag[i] * dn;
        }
        __syncthreads();
        // ---- values 5: 4 x 4 Cholesky of blkdiag(P, parameter variance) (the 5th row / column decouples exactly)
        if (tid == 0) {
            double L[10];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double d = Pv[URTRI(j, j)];
#pragma unroll
                for (int c = 0; c < j; ++c) d = fma(-L[URTRI(j, c)], L[URTRI(j, c)], d);
                const double inv = rsqrt(d);
                L[URTRI(j, j)] = d * inv; rec[UR_INV + j] = inv;
#pragma unroll
                for (int i = j + 1; i < 4; ++i) {
                    double t = Pv[URTRI(i, j)];
#pragma unroll
                    for (int c = 0; c < j; ++c) t = fma(-L[URTRI(i, c)], L[URTRI(j, c)], t);
                    L[URTRI(i, j)] = t * inv;
                }
            }
#pragma unroll
            for (int i = 0; i < 10; ++i) rec[UR_L + i] = L[i];
        }
        __syncthreads();
        // ---- values 6: the ten resampled, normalised outputs (lane jj: + column jj for jj < 4, mean for 4 and 9, - column jj-5)
        if (tid < 10) {
            const int j = tid < 5 ? tid : tid - 5;
            const double sgn = tid < 5 ? alpha : -alpha;
            double v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = smv[i] + ((j < 4 && i >= j) ? sgn * rec[UR_L + URTRI(i, j)] : 0.0);
            const double nn = rsqrt(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
            rec[UR_N + tid] = nn;
#pragma unroll
            for (int i = 0; i < 4; ++i) rec[UR_U + 4 * tid + i] = v[i] * nn;
        }
        __syncthreads();
        // ---- tangents: seed in sigma point j0, component i0: ds = column i0 of M_j0
        if (tid < 40) {
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
        __syncthreads();
        // ---- dense directions: one chunk row per lane, sum over the 40 sigma-point columns
        if (tid < 41) {
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

size_t unscented_record_bytes(int64_t) { return 0; }     // the record lives in shared memory

// preconditions (checked by the caller): model spin30 / spin25 with ONE sample state, nominal offset 0 / 4 / 8, d_AB 16-byte aligned
int launch_jacobians_unscented(rbq_handle* h, const rbq_model_params& p, int64_t K, const double* d_X, const double* d_U,
                               const double* d_t, const double* d_dt, double* d_rec, double* d_AB) {
    const int n = model_state_dim(p), m = model_control_dim(p);
    const bool s25 = p.model_id == RBQ_MODEL_SPIN25;
    UnscentedJacArgs a{p, n, m, K, d_X, d_U, d_t, d_dt, d_rec, d_AB};
    const size_t smem = sizeof(double) * ((size_t)n * (n + m) + UR_W + 40 + n + 2 + 40 + 4 + 10 + 2);
    const int64_t resident = (int64_t)h->sm_count * (s25 ? 7 : 8);
    const unsigned blocks = (unsigned)(K < resident ? K : resident);
    cudaError_t e;
    if (!s25) {
        e = cudaFuncSetAttribute(unscented_jac_kernel<RBQ_MODEL_SPIN30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        unscented_jac_kernel<RBQ_MODEL_SPIN30><<<blocks, UJ_THREADS, smem, h->stream>>>(a);
    } else {
        e = cudaFuncSetAttribute(unscented_jac_kernel<RBQ_MODEL_SPIN25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        unscented_jac_kernel<RBQ_MODEL_SPIN25><<<blocks, UJ_THREADS, smem, h->stream>>>(a);
    }
    h->launches += 1;
    return (int)cudaGetLastError();
}

}  // namespace rbq
