// rbq_warp_rollout.cuh — one WARP per trajectory: the augmented state lives in registers, each lane owns one
// 4-real iso block (nominal states, sample states, sigma points, derivative states) and evaluates only the
// propagator that block needs; the control chain is carried redundantly by every lane.  Used for
//   * open-loop rollouts      x[k+1] = discrete_dynamics(RK3, model, x[k], u[k], t, dt[k])          (TO.rollout!)
//   * closed-loop rollouts    u[k] = uref[k] + K[k](x[k] - xref[k]) + alpha d[k]                    (Altro forward pass)
// The single-thread kernels pay a global store->load round trip per knot (~300 cycles); here a step is a short
// dependent chain of FMAs plus, where the model needs them, warp shuffles (dot product with the feedback gain,
// Lawson-Euler source terms, the unscented transform's mean / covariance reductions).
// Reference step maps: spin13.jl:44-58, spin12.jl:165-192, spin18.jl:162-190, spin15.jl:50-73, spin11.jl:64-117,
// spin17.jl:61-114, spin30.jl:86-202.
#pragma once
#include "rbq_models.cuh"

namespace rbq {

constexpr unsigned FULL = 0xffffffffu;
RBQ_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
template <int W> RBQ_DEV double sum_lanes(double v) {   // total over the first W lanes, left on each of them (W = 4, 8, 16, 32)
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
RBQ_DEV double sum16(double v) {   // lanes 0..15 (the xor butterfly leaves the total on each of them)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// The propagator of a rollout lane: the degree-5 series is evaluated unconditionally (it is exact to 1e-17 up to x = 0.04, which
// covers every grid of the reference) and replaced only where x is larger -- no branch in front of the common path, so the
// series overlaps the shuffle latencies of the feedback law and of the unscented transform.  Same operations as prop_of.
RBQ_DEV Prop<double> prop_of_traj(double f, double a, double dt) {
    const double w = RBQ_PI * dt;
    const double p = w * f, q = a * w;
    const double x = p * p + q * q;
    double c, s;
    su2_series<5>(x, c, s);
    if (!(x <= 4.0e-2)) su2_general(x, c, s);
    Prop<double> E; E.c = c; E.sp = s * p; E.sq = s * q;
    return E;
}

struct WarpRolloutArgs {
    rbq_model_params p;
    int n, m;
    int64_t N, count;            // knots per trajectory, number of trajectories / candidates
    double t0;                   // time of knot 0 (knot times accumulate: t[k+1] = t[k] + dt[k])
    const double* x0;            // open loop: [count][n];  closed loop: Xref [N][n]
    const double* U;             // open loop: [count][N-1][m];  closed loop: Uref [N-1][m]
    const double* K;             // closed loop: [N-1] blocks, column-major m x n
    const double* d;             // closed loop: [N-1][m]
    const double* dt;            // [N-1]
    const double* alphas;        // closed loop: [count]
    double* X;                   // [count][N][n]
    double* Uout;                // closed loop: [count][N-1][m]
};

constexpr int CL_TILE = 32;   // knots of (K, Xref, Uref, d) staged per shared-memory tile in the closed-loop kernel
__host__ __device__ inline size_t closedloop_smem_bytes(int n, int m) { return sizeof(double) * CL_TILE * (size_t)(n * m + n + 2 * m); }

// SSC: the number of sigma-point chunks of the unscented models when known at compile time (1 in every run of the
// reference; 0: read it from the parameters) -- the chunk loops and their register arrays fold away
template <int MODEL, bool CLOSED, int SSC = 0>
__global__ void __launch_bounds__(32) warp_rollout_kernel(const WarpRolloutArgs a) {
    extern __shared__ __align__(16) double cl_smem[];
    double* sK = cl_smem;                                    // [CL_TILE][n][m] (column-major m x n per knot)
    double* sX = sK + CL_TILE * a.n * a.m;                   // [CL_TILE][n]
    double* sU = sX + CL_TILE * a.n;                         // [CL_TILE][m]
    double* sD = sU + CL_TILE * a.m;                         // [CL_TILE][m]
    const rbq_model_params& p = a.p;
    const int lane = threadIdx.x;
    const int64_t tr = blockIdx.x;
    // what the model fixes is a constant here: one control for every model but the time-optimal spin15, n of spin13 / spin12
    const int n = MODEL == RBQ_MODEL_SPIN13 ? 11 : (MODEL == RBQ_MODEL_SPIN12 ? 43 : a.n);
    const int m = MODEL == RBQ_MODEL_SPIN15 ? a.m : 1;
    const int64_t N = a.N;
    constexpr bool S25 = (MODEL == RBQ_MODEL_SPIN25);
    constexpr bool U30 = (MODEL == RBQ_MODEL_SPIN30) || S25;   // the unscented models (sigma points on lanes 0..9)
    constexpr int NCH = U30 ? (SSC ? SSC : RBQ_MAX_SAMPLE_STATES) : 1;   // sigma-point chunks a lane may hold
    constexpr int UB = S25 ? 17 : 15;                          // offset of the first sigma-point chunk
    constexpr int NNOM = S25 ? 2 : 3;                          // nominal blocks (lanes 10..)
    const int ssc = U30 ? (SSC ? SSC : p.sample_state_count) : 1;
    const int cb = (U30 && !S25) ? 12 : 8;                     // [int a, a, da] chain base
    // closed loop: the lanes that hold state are 0 .. TAIL (blocks first, then one lane for the scalar tail), so the feedback law's
    // dot product is a butterfly over FBW lanes only -- 2 shuffle stages instead of 5 for spin13 / spin15.  Lanes beyond FBW hold
    // no state and store nothing; the control they carry is not the trajectory's.
    constexpr int TAIL = (MODEL == RBQ_MODEL_SPIN13 || MODEL == RBQ_MODEL_SPIN15) ? 2 : (MODEL == RBQ_MODEL_SPIN11 ? 4 : (MODEL == RBQ_MODEL_SPIN12 ? 10 : 13));
    constexpr int FBW = TAIL < 4 ? 4 : (TAIL < 8 ? 8 : 16);

    // ---- lane roles ---------------------------------------------------------------------------------
    int off = -1;                 // state offset of this lane's block (chunk 0 for sigma lanes)
    double f = p.fq, da_l = 0.0;  // this lane's generator: f NEGI_H0 + (a + da_l) NEGI_H1
    if (MODEL == RBQ_MODEL_SPIN13 || MODEL == RBQ_MODEL_SPIN15) {
        if (lane < 2) off = 4 * lane;
    } else if (MODEL == RBQ_MODEL_SPIN12) {
        if (lane < 2) off = 4 * lane;
        else if (lane < 10) {
            off = 11 + 4 * (lane - 2);
            const int grp = (lane - 2) >> 2;
            if (p.model_id == RBQ_MODEL_SPIN12) f = p.fq_samples[grp];
            else da_l = grp == 0 ? p.namp : -p.namp;
        }
    } else if (MODEL == RBQ_MODEL_SPIN11) {
        if (lane < 2) off = 4 * lane;
        else if (lane - 2 < p.derivative_order) off = 11 + 4 * (lane - 2);
    } else {
        const double sig = sqrt(p.fq_cov);
        if (lane < 10) {
            off = UB + 4 * lane;
            if (!S25) { if (lane == 4) f = p.fq + sig; if (lane == 9) f = p.fq - sig; }   // spin25 shifts the amplitude instead (per knot)
        } else if (lane < 10 + NNOM) off = 4 * (lane - 10);
    }
    const bool sigma_lane = U30 && lane < 10;
    const bool has_block = off >= 0;
    const bool src_h1 = p.model_id == RBQ_MODEL_SPIN17;

    const double* __restrict__ x0 = CLOSED ? a.x0 : a.x0 + tr * n;
    const double* __restrict__ Uin = CLOSED ? a.U : a.U + tr * (N - 1) * m;
    double* __restrict__ X = a.X + tr * N * n;
    double* __restrict__ Uo = CLOSED ? a.Uout + tr * (N - 1) * m : nullptr;
    const double alpha = CLOSED ? a.alphas[tr] : 0.0;

    // ---- state in registers ---------------------------------------------------------------------------
    double psi[NCH][4];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) psi[c][i] = 0.0;
    if (has_block) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            if (c == 0 || (sigma_lane && c < ssc))
#pragma unroll
                for (int i = 0; i < 4; ++i) psi[c][i] = x0[off + 41 * c + i];
    }
    double ci = x0[cb], ca = x0[cb + 1], cd = x0[cb + 2];
    const bool DA = MODEL == RBQ_MODEL_SPIN15 && p.decay_aware;
    const bool TO = MODEL == RBQ_MODEL_SPIN15 && p.time_optimal;
    double gacc = DA ? x0[11] : 0.0;
    double pen[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) pen[c] = (U30 && c < ssc) ? x0[UB + 41 * c + 40] : 0.0;
    // spin25: pink-filter taps xp1..3 (11..13), yp1..3 (14..16), carried by every lane
    double fx1 = 0, fx2 = 0, fx3 = 0, fy1 = 0, fy2 = 0, fy3 = 0;
    if (S25) { fx1 = x0[11]; fx2 = x0[12]; fx3 = x0[13]; fy1 = x0[14]; fy2 = x0[15]; fy3 = x0[16]; }
    double time = a.t0;
    // knot 0
    for (int i = lane; i < n; i += 32) X[i] = x0[i];

    const double cf = U30 ? 1.0 / (2.0 * p.alpha * p.alpha) : 0.0;
    // the next knot's control and dt are fetched one step ahead: with a single warp per SM nothing else hides the load
    double u0_n = 0.0, u1_n = 0.0, dt_n = 0.0;
    if (N > 1) {
        if (!CLOSED) { u0_n = Uin[0]; if (m > 1) u1_n = Uin[1]; }
        dt_n = a.dt[0];
    }

    for (int64_t k = 0; k + 1 < N; ++k) {
        // ---- control ---------------------------------------------------------------------------------
        double u0, u1 = 0.0;
        if (CLOSED) {
            // Gains, reference states and feed-forward terms do not depend on the state: the warp stages CL_TILE knots of
            // them in shared memory with coalesced loads, so the L2 round trip is paid once per tile instead of once per
            // knot on the critical path of the feedback law.
            const int kt = (int)(k % CL_TILE);
            if (kt == 0) {
                __syncwarp();
                const int64_t len = (N - 1 - k) < CL_TILE ? (N - 1 - k) : CL_TILE;
                const double* gK = a.K + k * m * n;
                const double* gX = a.x0 + k * n;
                for (int64_t i = lane; i < len * m * n; i += 32) sK[i] = gK[i];
                for (int64_t i = lane; i < len * n; i += 32) sX[i] = gX[i];
                for (int64_t i = lane; i < len * m; i += 32) { sU[i] = a.U[k * m + i]; sD[i] = a.d[k * m + i]; }
                __syncwarp();
            }
            const double* __restrict__ Kk = sK + kt * m * n;
            const double* __restrict__ xr = sX + kt * n;
            double s0 = 0.0, s1 = 0.0;
            if (!U30) {
                // every lane runs the same four multiply-adds on "its" four state entries -- a block lane its block, the tail
                // lane [int a, a, da(, decay)] -- with the weights of the unused slots selected to zero: no divergent region in
                // front of the butterfly, so the propagator of this knot (which does not depend on u_k) overlaps it
                // (8 candidates x 1001 knots: spin13 0.41 -> 0.31 ms, spin12 0.54 -> 0.45 ms).  The same scheme with per-lane
                // index tables for the unscented models' penalties and filter taps was slower (spin30 1.22 -> 1.48 ms).
                const int fidx = has_block ? off : cb;
                const int fcnt = has_block ? 4 : (lane == TAIL ? (DA ? 4 : 3) : 0);
                const double fv[4] = {has_block ? psi[0][0] : ci, has_block ? psi[0][1] : ca, has_block ? psi[0][2] : cd, has_block ? psi[0][3] : gacc};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = fidx + (i < 3 || has_block || DA ? i : 0);       // (stays inside the state when the tail has 3 entries)
                    const double dx = fv[i] - xr[idx];
                    const double k0 = i < fcnt ? Kk[idx * m] : 0.0;
                    s0 = fma(k0, dx, s0);
                    if (m > 1) { const double k1 = i < fcnt ? Kk[idx * m + 1] : 0.0; s1 = fma(k1, dx, s1); }
                }
            } else if (NCH == 1) {
                // the unscented models with one sigma-point chunk (every reference run): same scheme, the tail lane's fourth entry
                // is the chunk's penalty
                const bool tl = lane == TAIL;
                const double fv[4] = {has_block ? psi[0][0] : ci, has_block ? psi[0][1] : ca, has_block ? psi[0][2] : cd, has_block ? psi[0][3] : pen[0]};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = has_block ? off + i : (i < 3 ? cb + i : UB + 40);
                    const double dx = fv[i] - xr[idx];
                    const double k0 = (has_block || tl) ? Kk[idx * m] : 0.0;
                    s0 = fma(k0, dx, s0);
                }
            } else if (has_block) {
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    if (c == 0 || (sigma_lane && c < ssc))
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int idx = off + 41 * c + i;
                            const double dx = psi[c][i] - xr[idx];
                            s0 = fma(Kk[idx * m], dx, s0);
                            if (m > 1) s1 = fma(Kk[idx * m + 1], dx, s1);
                        }
            }
            if (U30 && NCH == 1 && S25 && lane == TAIL) {   // spin25: the pink-filter taps
                const double taps[6] = {fx1, fx2, fx3, fy1, fy2, fy3};
#pragma unroll
                for (int i = 0; i < 6; ++i) s0 = fma(Kk[(11 + i) * m], taps[i] - xr[11 + i], s0);
            }
            if (U30 && NCH > 1 && lane == TAIL) { // scalar tail of the state: chain, decay accumulator, unscented penalties
                const double tail[3] = {ci - xr[cb], ca - xr[cb + 1], cd - xr[cb + 2]};
#pragma unroll
                for (int i = 0; i < 3; ++i) { s0 = fma(Kk[(cb + i) * m], tail[i], s0); if (m > 1) s1 = fma(Kk[(cb + i) * m + 1], tail[i], s1); }
                if (DA) { const double dx = gacc - xr[11]; s0 = fma(Kk[11 * m], dx, s0); if (m > 1) s1 = fma(Kk[11 * m + 1], dx, s1); }
                if (U30) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        if (c < ssc) { const int idx = UB + 41 * c + 40; s0 = fma(Kk[idx * m], pen[c] - xr[idx], s0); }
                }
                if (S25) {
                    const double taps[6] = {fx1, fx2, fx3, fy1, fy2, fy3};
#pragma unroll
                    for (int i = 0; i < 6; ++i) s0 = fma(Kk[(11 + i) * m], taps[i] - xr[11 + i], s0);
                }
            }
            s0 = sum_lanes<FBW>(s0);
            u0 = sU[kt * m] + fma(alpha, sD[kt * m], s0);
            if (m > 1) { s1 = sum_lanes<FBW>(s1); u1 = sU[kt * m + 1] + fma(alpha, sD[kt * m + 1], s1); }
            if (lane == 0) { Uo[k * m] = u0; if (m > 1) Uo[k * m + 1] = u1; }
        } else {
            u0 = u0_n; u1 = u1_n;
            if (k + 2 < N) { u0_n = Uin[(k + 1) * m]; if (m > 1) u1_n = Uin[(k + 1) * m + 1]; }
        }
        const double dt = TO ? u1 * u1 : dt_n;
        if (k + 2 < N) dt_n = a.dt[k + 1];
        double* Xn = X + (k + 1) * n;

        // ---- spin25: advance the pink filter (spin25.jl:194-219); its output scales the amplitude offset of points 5, 10
        if (S25) {
            const double xk = sqrt(p.fq_cov);
            const long long kp = knot_point_of(time, dt);
            double yk = 0.049922035 * xk;
            if (kp == 1) {
            } else if (kp == 2) {
                yk = yk + fx1 * (-0.095993537) - fy1 * (-2.494956002);
            } else if (kp == 3) {
                yk = yk + fx1 * (-0.095993537) + fx2 * 0.050612699 - fy1 * (-2.494956002) - fy2 * 2.017265875;
            } else {
                yk = yk + fx1 * (-0.095993537) + fx2 * 0.050612699 + fx3 * (-0.004408786) - fy1 * (-2.494956002) - fy2 * 2.017265875 -
                     fy3 * (-0.522189400);
            }
            fx3 = fx2; fx2 = fx1; fx1 = xk; fy3 = fy2; fy2 = fy1; fy1 = yk;
            const double namp = yk * p.namp;
            da_l = lane == 4 ? namp : (lane == 9 ? -namp : 0.0);
            if (lane == 0) { Xn[11] = fx1; Xn[12] = fx2; Xn[13] = fx3; Xn[14] = fy1; Xn[15] = fy2; Xn[16] = fy3; }
            time = time + dt;
        }
        // ---- this lane's propagator and block ---------------------------------------------------------
        const Prop<double> E = prop_of_traj(f, ca + da_l, dt);
        double in0[4] = {psi[0][0], psi[0][1], psi[0][2], psi[0][3]};   // block value at knot k (chunk 0)
        if (!U30) {
            double w[4] = {in0[0], in0[1], in0[2], in0[3]};
            if (MODEL == RBQ_MODEL_SPIN11) {
                // Lawson-Euler source: d psi += dt H psi1 (lane 2 <- lane 0), d2 psi += 2 dt H d psi (lane 3 <- lane 2)
                const int src = lane == 3 ? 2 : 0;
                double s[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) s[i] = __shfl_sync(FULL, in0[i], src);
                if (lane == 2 || lane == 3) {
                    double g[4];
                    if (src_h1) apply_N1<double>(s[0], s[1], s[2], s[3], g[0], g[1], g[2], g[3]);
                    else apply_N0<double>(s[0], s[1], s[2], s[3], g[0], g[1], g[2], g[3]);
                    const double fac = (RBQ_PI * dt) * (double)(lane - 1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[i] = fma(g[i], fac, w[i]);
                }
            }
            if (has_block) {
                prop_apply<double>(E, w[0], w[1], w[2], w[3], psi[0][0], psi[0][1], psi[0][2], psi[0][3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) Xn[off + i] = psi[0][i];
            }
        } else {
            // every lane applies its propagator to its (chunk 0) block -- lanes without one hold zeros -- so that no divergent
            // region sits between the propagator and the reductions; the nominal lanes 10.. keep the result and store it
            double e0[4];
            prop_apply<double>(E, in0[0], in0[1], in0[2], in0[3], e0[0], e0[1], e0[2], e0[3]);
            const bool nominal_lane = lane >= 10 && lane < 10 + NNOM;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (nominal_lane) { psi[0][i] = e0[i]; Xn[off + i] = e0[i]; }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (c >= ssc) break;
                double s[4] = {0.0, 0.0, 0.0, 0.0};
                if (c == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) s[i] = sigma_lane ? e0[i] : 0.0;
                } else if (sigma_lane) prop_apply<double>(E, psi[c][0], psi[c][1], psi[c][2], psi[c][3], s[0], s[1], s[2], s[3]);
                double sm[4];
                double P[10];   // packed lower triangle 00 10 11 20 21 22 30 31 32 33
                if constexpr (!CLOSED) {
                    // mean and covariance in ONE butterfly pass: deviations are taken from a reference point that is known without
                    // a reduction -- the midpoint of the two frequency-shifted points, which were both resampled AT the previous
                    // mean and so sit within O(spread^2) of the new one -- and the first and second moments of those deviations
                    // are summed together (14 independent butterflies): mean = ref + delta, P = cf (M2 - 10 delta delta'),
                    // delta = sum(d') / 10.  The correction term is far below M2, so nothing cancels; the dependent chain is one
                    // shuffle stage + one butterfly instead of two butterflies (open loop, 1001 knots: 0.92 -> 0.77 ms).  In the
                    // closed-loop kernel the same change measured slower (8 candidates: 1.18 -> 1.35 ms), so it keeps two passes.
                    double ref[4], dp[4], dl[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ref[i] = 0.5 * (__shfl_sync(FULL, s[i], 4) + __shfl_sync(FULL, s[i], 9));
                        dp[i] = sigma_lane ? s[i] - ref[i] : 0.0;
                    }
                    {
                        int q = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int kk = 0; kk <= i; ++kk) P[q++] = sum16(dp[i] * dp[kk]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) { dl[i] = 0.1 * sum16(dp[i]); sm[i] = ref[i] + dl[i]; }
                    {
                        int q = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int kk = 0; kk <= i; ++kk) { P[q] = cf * fma(-10.0 * dl[i], dl[kk], P[q]); ++q; }
                    }
                } else {
                    double dv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { sm[i] = 0.1 * sum16(s[i]); dv[i] = sigma_lane ? s[i] - sm[i] : 0.0; }
                    int q = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int kk = 0; kk <= i; ++kk) P[q++] = cf * sum16(dv[i] * dv[kk]);
                }
                double L[10];
#define RBQ_TRI(i, k) ((i) * ((i) + 1) / 2 + (k))
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double dd = P[RBQ_TRI(j, j)];
#pragma unroll
                    for (int kk = 0; kk < j; ++kk) dd = fma(-L[RBQ_TRI(j, kk)], L[RBQ_TRI(j, kk)], dd);
                    const double inv = rsqrt_lean(dd);   // one reciprocal square root instead of sqrt + divide (a failed factorisation still gives NaN)
                    L[RBQ_TRI(j, j)] = dd * inv;
#pragma unroll
                    for (int i = j + 1; i < 4; ++i) {
                        double t = P[RBQ_TRI(i, j)];
#pragma unroll
                        for (int kk = 0; kk < j; ++kk) t = fma(-L[RBQ_TRI(i, kk)], L[RBQ_TRI(j, kk)], t);
                        L[RBQ_TRI(i, j)] = t * inv;
                    }
                }
                // resample: lanes 0..3 sm + alpha L[:,j], lane 4 sm, lanes 5..8 sm - alpha L[:,j-5], lane 9 sm
                const int j = lane < 5 ? lane : lane - 5;
                const double sgn = lane < 5 ? p.alpha : -p.alpha;
                double v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double col = 0.0;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (jj <= i && j == jj) col = L[RBQ_TRI(i, jj)];
                    v[i] = fma(sgn, col, sm[i]);
                }
#undef RBQ_TRI
                const double rn = rsqrt_lean(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double vn = v[i] * rn;
                    if (sigma_lane) { psi[c][i] = vn; Xn[off + 41 * c + i] = vn; }
                }
                // penalty with the nominal state AT KNOT k (lane 10 + nominal block)
                const int nl = 10 + (p.nominal_offset[c] >> 2);
                double pv = P[0] * p.S_diag[0] + P[2] * p.S_diag[1] + P[5] * p.S_diag[2] + P[9] * p.S_diag[3];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double dn = sm[i] - __shfl_sync(FULL, in0[i], nl);
                    pv = fma(dn * dn, p.S_diag[i], pv);
                }
                pen[c] = pv;
                if (lane == 0) Xn[UB + 41 * c + 40] = pv;
            }
        }
        // ---- scalar tail -----------------------------------------------------------------------------------
        if (DA) {
            gacc = fma(1.0 / t1_interp<double>(ca, 1.0), dt, gacc);
            if (lane == 0) Xn[11] = gacc;
        }
        const double ni = fma(ca, dt, ci), na = fma(cd, dt, ca), nd = fma(u0, dt, cd);
        ci = ni; ca = na; cd = nd;
        if (lane == 0) { Xn[cb] = ci; Xn[cb + 1] = ca; Xn[cb + 2] = cd; }
    }
}

// Batches of open-loop rollouts for the models whose 4-real blocks evolve independently (spin13 / spin15: 2 blocks, spin12 /
// spin18: 10): ONE THREAD per (trajectory, block), the control chain carried redundantly.  The warp-per-trajectory kernel
// above keeps 2 (10) of 32 lanes busy while every FP64 instruction still occupies the pipe for the whole warp; here all lanes
// work, the arithmetic per block is the same expression by expression, so the states are bit-identical.
template <int MODEL>
__global__ void __launch_bounds__(128) packed_rollout_kernel(const WarpRolloutArgs a, int nb, int threads) {
    const rbq_model_params& p = a.p;
    const int64_t gid = (int64_t)blockIdx.x * threads + threadIdx.x;
    const int64_t tr = gid / nb;
    const int blk = (int)(gid - tr * nb);
    if (tr >= a.count) return;
    const int n = a.n, m = a.m;
    const int64_t N = a.N;
    int off;
    double f = p.fq, da_l = 0.0;
    if (MODEL == RBQ_MODEL_SPIN12) {
        if (blk < 2) off = 4 * blk;
        else {
            off = 11 + 4 * (blk - 2);
            const int grp = (blk - 2) >> 2;
            if (p.model_id == RBQ_MODEL_SPIN12) f = p.fq_samples[grp];
            else da_l = grp == 0 ? p.namp : -p.namp;
        }
    } else off = 4 * blk;
    const int cb = 8;
    const double* __restrict__ x0 = a.x0 + tr * n;
    const double* __restrict__ Uin = a.U + tr * (N - 1) * m;
    double* __restrict__ X = a.X + tr * N * n;
    double psi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { psi[i] = x0[off + i]; X[off + i] = psi[i]; }
    double ci = x0[cb], ca = x0[cb + 1], cd = x0[cb + 2];
    const bool DA = MODEL == RBQ_MODEL_SPIN15 && p.decay_aware;
    const bool TO = MODEL == RBQ_MODEL_SPIN15 && p.time_optimal;
    double gacc = DA ? x0[11] : 0.0;
    if (blk == 0) { X[cb] = ci; X[cb + 1] = ca; X[cb + 2] = cd; if (DA) X[11] = gacc; }
    // the next knot's control and dt are fetched one step ahead (nothing else hides the load in a small batch)
    double u0_n = 0.0, u1_n = 0.0, dt_n = 0.0;
    if (N > 1) { u0_n = Uin[0]; if (TO) u1_n = Uin[1]; dt_n = a.dt[0]; }
    for (int64_t k = 0; k + 1 < N; ++k) {
        const double u0 = u0_n;
        const double dt = TO ? u1_n * u1_n : dt_n;
        if (k + 2 < N) { u0_n = Uin[(k + 1) * m]; if (TO) u1_n = Uin[(k + 1) * m + 1]; dt_n = a.dt[k + 1]; }
        double* Xn = X + (k + 1) * n;
        const Prop<double> E = prop_of_traj(f, ca + da_l, dt);
        const double w0 = psi[0], w1 = psi[1], w2 = psi[2], w3 = psi[3];
        prop_apply<double>(E, w0, w1, w2, w3, psi[0], psi[1], psi[2], psi[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) Xn[off + i] = psi[i];
        if (DA) {
            gacc = fma(1.0 / t1_interp<double>(ca, 1.0), dt, gacc);
            if (blk == 0) Xn[11] = gacc;
        }
        const double ni = fma(ca, dt, ci), na = fma(cd, dt, ca), nd = fma(u0, dt, cd);
        ci = ni; ca = na; cd = nd;
        if (blk == 0) { Xn[cb] = ci; Xn[cb + 1] = ca; Xn[cb + 2] = cd; }
    }
}

}  // namespace rbq
