// rbq_mc.cu — Monte-Carlo robustness kernels (one thread per sample, FP64-pipe bound):
//   mc_static_kernel   run_sim_prop(...; dynamics_type=schroed)    spin.jl:438-452, 1553-1608
//   mc_static_ctab_kernel + mc_static_finish_kernel: the same sweep with its knot table in constant memory (device-resident calls)
//   mc_noise_kernel    run_sim_prop(...; dynamics_type=schroedda)  spin.jl:469-484
//   lindblad_kernel    run_sim_prop(...; dynamics_type=lindbladt1) spin.jl:521-539 (+ T2 of 542-554)
//   compute_fidelities / fidelity_vec_iso2 / fidelity_mat fused into each (spin.jl:1035-1037,
//   1101-1104, 1198-1240), infidelity statistics reduced with warp shuffles.
// The pulse is staged into shared memory tile by tile with 1-D bulk TMA copies
// (cp.async.bulk + mbarrier, double buffered); all threads of a block walk the pulse in lock step.
// sm_100a only.
#include <mutex>
#include "rbq_internal.h"
#include <stdlib.h>

#include "rbq_math.cuh"
#include "rbq_tanpoly.cuh"
#include "rbq_tma.cuh"

namespace rbq {

constexpr int MC_THREADS = 256;
constexpr int MC_TILE = 2048;   // knots per shared-memory tile (16 KB), two tiles in flight

// Walks a staged pulse (16-byte aligned, zero padded to a multiple of 2 knots) through two
// shared-memory tiles.  `pass` counts tiles since kernel start so the pulse can be replayed once per gate.
// Optionally a companion int32 table (one row of `aux_stride` entries per replay) rides in the same tiles:
// the noise-sample index of every (gate, knot) of integrate_prop_schroedda!.
template <int TILE, int EPK = 1>
struct PulsePipeT {
    double* tile;       // [2][TILE * EPK]   (EPK doubles per knot: 1 = the pulse itself, 4 = the static sweep's knot table)
    int32_t* aux_tile;  // [2][TILE] or nullptr
    uint64_t* bar;      // [2]
    const double* src;
    const int32_t* aux_src;
    int64_t aux_stride;
    int64_t nk;
    int ntiles;
    uint32_t pass;
    RBQ_DEV void init(double* t, uint64_t* b, const double* s, int64_t n, int32_t* at = nullptr, const int32_t* as = nullptr,
                      int64_t astride = 0, int64_t replay0 = 0) {
        tile = t; bar = b; src = s; nk = n; pass = 0; aux_tile = at; aux_src = as; aux_stride = astride;
        ntiles = (int)((n + TILE - 1) / TILE);
        if (threadIdx.x == 0) {
            mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init();
            if (n > 0) issue(0, 0, replay0);     // the first tile is in flight while the block loads its samples
        }
        __syncthreads();
    }
    RBQ_DEV int tile_len(int ti) const { const int64_t r = nk - (int64_t)ti * TILE; return r < TILE ? (int)r : TILE; }
    RBQ_DEV void issue(int ti, uint32_t slot, int64_t replay) {
        const int len = tile_len(ti);
        const uint32_t bytes = (uint32_t)((EPK == 1 ? ((len + 1) & ~1) : len * EPK) * sizeof(double));
        const uint32_t abytes = aux_src ? (uint32_t)(((len + 3) & ~3) * sizeof(int32_t)) : 0u;
        mbar_expect_tx(&bar[slot], bytes + abytes);
        bulk_g2s(tile + slot * (TILE * EPK), src + (int64_t)ti * (TILE * EPK), bytes, &bar[slot]);
        if (aux_src) bulk_g2s(aux_tile + slot * TILE, aux_src + replay * aux_stride + (int64_t)ti * TILE, abytes, &bar[slot]);
    }
    // call with ti = 0..ntiles-1 in order, replay = 0,1,..; every thread of the block must call it
    RBQ_DEV const double* acquire(int ti, bool more_passes, int64_t replay = 0) {
        const uint32_t slot = pass & 1u;
        if (threadIdx.x == 0) {
            // prefetch the next tile (of this replay, or the first tile of the next one) into the other slot;
            // the __syncthreads in release() guarantees nobody still reads it
            if (ti + 1 < ntiles) issue(ti + 1, slot ^ 1u, replay);
            else if (more_passes) issue(0, slot ^ 1u, replay + 1);
        }
        mbar_wait(&bar[slot], (pass >> 1) & 1u);
        return tile + slot * (TILE * EPK);
    }
    RBQ_DEV const int32_t* aux() const { return aux_tile + (pass & 1u) * TILE; }
    // `more`: another acquire() follows (the barrier protects the slot the next prefetch overwrites)
    RBQ_DEV void release(bool more = true) { if (more) __syncthreads(); ++pass; }
};
typedef PulsePipeT<MC_TILE> PulsePipe;
constexpr int NOISE_TILE = 512;   // knots per tile of the noise sweeps: 16 KB of knot-table rows + 2 KB of noise indices
constexpr int ST_TILE = 256;    // knots per tile of the static sweep's knot table (4 doubles per knot: 8 KB)
constexpr int ST_EPK = 4;

// ---- statistics ---------------------------------------------------------------------------------------
struct ThreadStats {
    double sum = 0.0, sumsq = 0.0, mn = INFINITY, mx = -INFINITY;
    long long cnt = 0, nonfinite = 0;
    RBQ_DEV void add(double e, unsigned* hist_s) {
        if (!isfinite(e)) { ++nonfinite; return; }
        ++cnt; sum += e; sumsq = fma(e, e, sumsq);
        mn = fmin(mn, e); mx = fmax(mx, e);
        int b = 0;
        if (e > 0.0) {
            b = (int)floor((log10(e) + 16.0) * 16.0);
            b = b < 0 ? 0 : (b > RBQ_HIST_BINS - 1 ? RBQ_HIST_BINS - 1 : b);
        }
        atomicAdd(&hist_s[b], 1u);
    }
};
// warp shuffle -> shared -> one BlockPartial per block; histogram bins flushed with integer atomics
RBQ_DEV void block_stats_flush(ThreadStats st, unsigned* hist_s, BlockPartial* red_s, BlockPartial* partials,
                               unsigned long long* hist_g) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        st.sum += __shfl_down_sync(0xffffffffu, st.sum, o);
        st.sumsq += __shfl_down_sync(0xffffffffu, st.sumsq, o);
        st.mn = fmin(st.mn, __shfl_down_sync(0xffffffffu, st.mn, o));
        st.mx = fmax(st.mx, __shfl_down_sync(0xffffffffu, st.mx, o));
        st.cnt += __shfl_down_sync(0xffffffffu, st.cnt, o);
        st.nonfinite += __shfl_down_sync(0xffffffffu, st.nonfinite, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red_s[warp] = BlockPartial{st.sum, st.sumsq, st.mn, st.mx, st.cnt, st.nonfinite};
    __syncthreads();
    if (threadIdx.x == 0) {
        BlockPartial r = red_s[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            r.sum += red_s[w].sum; r.sumsq += red_s[w].sumsq;
            r.mn = fmin(r.mn, red_s[w].mn); r.mx = fmax(r.mx, red_s[w].mx);
            r.cnt += red_s[w].cnt; r.nonfinite += red_s[w].nonfinite;
        }
        partials[blockIdx.x] = r;
    }
    for (int b = threadIdx.x; b < RBQ_HIST_BINS; b += blockDim.x)
        if (hist_s[b]) atomicAdd(&hist_g[b], (unsigned long long)hist_s[b]);
}

constexpr int FIN_THREADS = 512;    // one block; fixed thread -> partial assignment and tree => deterministic sums
__global__ void __launch_bounds__(FIN_THREADS) stats_finalize_kernel(const BlockPartial* __restrict__ partials, int nblocks,
                                                             rbq_stats* __restrict__ out) {
    __shared__ BlockPartial red[FIN_THREADS];
    BlockPartial r{0.0, 0.0, INFINITY, -INFINITY, 0, 0};
    for (int i = threadIdx.x; i < nblocks; i += FIN_THREADS) {   // fixed assignment => deterministic sum order
        const BlockPartial q = partials[i];
        r.sum += q.sum; r.sumsq += q.sumsq; r.mn = fmin(r.mn, q.mn); r.mx = fmax(r.mx, q.mx);
        r.cnt += q.cnt; r.nonfinite += q.nonfinite;
    }
    red[threadIdx.x] = r;
    __syncthreads();
    for (int o = FIN_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            BlockPartial& a = red[threadIdx.x];
            const BlockPartial& b = red[threadIdx.x + o];
            a.sum += b.sum; a.sumsq += b.sumsq; a.mn = fmin(a.mn, b.mn); a.mx = fmax(a.mx, b.mx);
            a.cnt += b.cnt; a.nonfinite += b.nonfinite;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out->count += red[0].cnt; out->nonfinite += red[0].nonfinite;
        out->sum += red[0].sum; out->sumsq += red[0].sumsq;
        out->min = fmin(out->min, red[0].mn); out->max = fmax(out->max, red[0].mx);
    }
}

__global__ void stats_reset_kernel(rbq_stats* __restrict__ st) {
    for (int b = threadIdx.x; b < RBQ_HIST_BINS; b += blockDim.x) st->hist[b] = 0;
    if (threadIdx.x == 0) { st->count = 0; st->nonfinite = 0; st->sum = 0.0; st->sumsq = 0.0; st->min = INFINITY; st->max = -INFINITY; }
}

// ---- pulse staging: copy into an aligned zero-padded buffer, max |a| for the series-degree bound ----
__global__ void __launch_bounds__(1024) pulse_prep_kernel(const double* __restrict__ controls, int64_t nk,
                                                          int64_t padded, double* __restrict__ staged,
                                                          double* __restrict__ amax) {
    __shared__ double red[32];
    double m = 0.0;
    for (int64_t j = threadIdx.x; j < padded; j += 1024) {
        const double a = j < nk ? controls[j] : 0.0;
        staged[j] = a;
        m = fmax(m, fabs(a));   // NaN controls propagate through the rollout itself
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) m = fmax(m, red[w]);
        *amax = m;
    }
}

// ---- knot table of the static sweep ------------------------------------------------------------------------------------
// T(x) = tan(sqrt x)/sqrt x = sum t_k x^k (all t_k > 0).  Around the sweep's nominal point (f0, a0) every knot has
// x_j = q_j^2 + p0^2, q_j = w (a_j + a0), p0 = w f0, and a sample (f, da) sees x = x_j + eps with
//     eps = 2 q_j dq + dq^2 + (p^2 - p0^2),   dq = w (da - a0),  p = w f,
// so T(x) = c0_j + eps (c1_j + eps c2_j) + R,  c_i = (i-th derivative of T at x_j)/i!,  0 <= R <= 0.057 eps^3 for x <= 0.03.
// Table row j = (q_j, c0_j, c1_j, c2_j); scal = (max|a|, f0, a0, max|q_j|, table_ok, max x_j).
RBQ_DEV void tan_taylor(double x, double& c0, double& c1, double& c2) {
    constexpr double t[12] = {1.0, 1.0 / 3.0, 2.0 / 15.0, 17.0 / 315.0, 62.0 / 2835.0, 1382.0 / 155925.0, 21844.0 / 6081075.0,
                              929569.0 / 638512875.0, 6404582.0 / 10854718875.0, 443861162.0 / 1856156927625.0,
                              18888466084.0 / 194896477400625.0, 113927491862.0 / 2900518163668125.0};
    double s0 = t[11], s1 = 11.0 * t[11], s2 = 55.0 * t[11];
#pragma unroll
    for (int k = 10; k >= 0; --k) {
        s0 = fma(s0, x, t[k]);
        if (k >= 1) s1 = fma(s1, x, (double)k * t[k]);
        if (k >= 2) s2 = fma(s2, x, (double)(k * (k - 1) / 2) * t[k]);
    }
    c0 = s0; c1 = s1; c2 = s2;
}
constexpr double DELTA_XMAX = 0.03;       // nominal x up to which the 12-term series and the remainder bound are used
constexpr double DELTA_R3 = 0.057;        // >= (third derivative of T)/6 on [0, DELTA_XMAX + 1e-4]
constexpr double DELTA_TOL = 5e-17;       // admitted relative truncation of T per knot
__global__ void __launch_bounds__(1024) static_table_kernel(const double* __restrict__ controls, int64_t nk, double dt,
                                                            double f0_in, double a0_in, double* __restrict__ table,
                                                            double* __restrict__ scal, double* __restrict__ table_m) {
    __shared__ double red[3][32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double f0 = isfinite(f0_in) ? f0_in : 0.0, a0 = isfinite(a0_in) ? a0_in : 0.0;
    const double w = RBQ_PI * dt, p0 = w * f0, p02 = p0 * p0;
    double amax = 0.0, qmax = 0.0, xmax = 0.0;
    for (int64_t j = tid; j < nk; j += 1024) {
        const double a = controls[j];
        const double q = w * (a + a0);
        const double x = fma(q, q, p02);
        double c0, c1, c2;
        tan_taylor(x <= DELTA_XMAX ? x : DELTA_XMAX, c0, c1, c2);   // rows beyond the range are only read for q_j
        table[4 * j] = q; table[4 * j + 1] = c0; table[4 * j + 2] = c1; table[4 * j + 3] = c2;
        // the same quadratic with its leading coefficient factored out, T = c2 (eps (eps + rho) + sigma): the rows of the
        // constant-memory form of the sweep (mc_static_ctab_kernel), in which every instruction takes one table value
        if (table_m) { table_m[4 * j] = q; table_m[4 * j + 1] = c1 / c2; table_m[4 * j + 2] = c0 / c2; table_m[4 * j + 3] = c2; }
        amax = fmax(amax, fabs(a)); qmax = fmax(qmax, fabs(q));
        xmax = x == x ? fmax(xmax, x) : INFINITY;                    // NaN controls switch the table class off
    }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        amax = fmax(amax, __shfl_down_sync(0xffffffffu, amax, o));
        qmax = fmax(qmax, __shfl_down_sync(0xffffffffu, qmax, o));
        xmax = fmax(xmax, __shfl_down_sync(0xffffffffu, xmax, o));
    }
    if (lane == 0) { red[0][warp] = amax; red[1][warp] = qmax; red[2][warp] = xmax; }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < 32; ++i) { amax = fmax(amax, red[0][i]); qmax = fmax(qmax, red[1][i]); xmax = fmax(xmax, red[2][i]); }
        scal[0] = amax; scal[1] = f0; scal[2] = a0; scal[3] = qmax;
        scal[4] = (xmax <= DELTA_XMAX) ? 1.0 : 0.0;
        scal[5] = xmax;
    }
}

// ---- per-sample pieces ----------------------------------------------------------------------------------
// propagator class from a bound on x = theta^2:
//   0 -> tan form, degree-4 polynomial (x <= 1e-3: the dt = 0.01 grids)     1 -> tan form, degree 6 (x <= 0.026: dt = 0.1)
//   2 -> normalised (cos, sinc) Taylor pair, K = 8 (x <= 0.7)                  3 -> sincos
RBQ_DEV int degree_class(double xb) { return xb <= RBQ_TANA_XMAX ? 0 : (xb <= RBQ_TANB_XMAX ? 1 : (xb <= 0.7 ? 2 : 3)); }
// warps only meet at tile boundaries, so the series degree needs to be uniform per warp, not per block
RBQ_DEV int warp_max_class(int c) { return __reduce_max_sync(0xffffffffu, c); }
template <int K> RBQ_DEV void su2_coeffs(double x, double& c, double& s) {
    if (K > 0) su2_series<(K > 0 ? K : 1)>(x, c, s);
    else {
        if (x < 0.5) { su2_series<8>(x, c, s); return; }
        const double th = sqrt(x);
        double sn, cs;
        sincos(th, &sn, &cs);
        c = cs; s = sn / th;
    }
}
// advance SPT states through `len` knots of a table tile (row j starts with q_j = w (a_j + a0)):  q = q_j + dq
// (normalised form, any x)
template <int K, int SPT>
RBQ_DEV void su2_tile(const double* __restrict__ tile, int len, const double (&p)[SPT], const double (&p2)[SPT],
                      const double (&qd)[SPT], double (&v)[SPT][4]) {
#pragma unroll 4
    for (int j = 0; j < len; ++j) {
        const double wa = tile[ST_EPK * j];
#pragma unroll
        for (int i = 0; i < SPT; ++i) {
            const double q = wa + qd[i];
            const double x = fma(q, q, p2[i]);
            double c, s;
            su2_coeffs<K>(x, c, s);
            Prop<double> E{c, s * p[i], s * q};
            prop_apply(E, v[i]);
        }
    }
}
// Tan form: exp(G) = cos(theta) (I + T G), T = tan(theta)/theta a polynomial in x = theta^2.  The scalar cos(theta) only
// rescales the state, and the callers renormalise (the exact evolution is unitary), so a step is
//     v <- v + (T p) N0 v + (T q) N1 v :   1 add + 1 fma + DEG fma + 2 mul + 8 fma.
// U knots x SPT samples are evaluated together so that the Horner chains interleave (U*SPT independent chains).
template <int DEG> RBQ_DEV double tan_poly(double x, const double (&cf)[DEG + 1]) {
    double t = fma(cf[DEG], x, cf[DEG - 1]);
#pragma unroll
    for (int k = DEG - 2; k >= 0; --k) t = fma(t, x, cf[k]);
    return t;
}
RBQ_DEV void tan_apply(double tp, double tq, double* v) {
    const double a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
    v[0] = fma(tq, a3, fma(tp, a2, a0));
    v[1] = fma(tq, a2, fma(-tp, a3, a1));
    v[2] = fma(-tq, a1, fma(-tp, a0, a2));
    v[3] = fma(-tq, a0, fma(tp, a1, a3));
}
template <int DEG, int SPT, int U>
RBQ_DEV void tan_tile(const double* __restrict__ tile, int len, const double (&p)[SPT], const double (&p2)[SPT],
                      const double (&qd)[SPT], double (&v)[SPT][4], const double (&cf)[DEG + 1]) {
    int j = 0;
    for (; j + U <= len; j += U) {
        double q[U][SPT], x[U][SPT], t[U][SPT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double wa = tile[ST_EPK * (j + u)];
#pragma unroll
            for (int i = 0; i < SPT; ++i) { q[u][i] = wa + qd[i]; x[u][i] = fma(q[u][i], q[u][i], p2[i]); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < SPT; ++i) t[u][i] = fma(cf[DEG], x[u][i], cf[DEG - 1]);
#pragma unroll
        for (int k = DEG - 2; k >= 0; --k)
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int i = 0; i < SPT; ++i) t[u][i] = fma(t[u][i], x[u][i], cf[k]);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < SPT; ++i) tan_apply(t[u][i] * p[i], t[u][i] * q[u][i], v[i]);
    }
    for (; j < len; ++j) {
        const double wa = tile[ST_EPK * j];
#pragma unroll
        for (int i = 0; i < SPT; ++i) {
            const double q = wa + qd[i];
            const double t = tan_poly<DEG>(fma(q, q, p2[i]), cf);
            tan_apply(t * p[i], t * q, v[i]);
        }
    }
    // |v| grows by 1/cos(theta) per knot (at most e^27 over a 2048-knot tile in class 1): pull it back to 1 per tile
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
        const double inv = rsqrt(fma(v[i][3], v[i][3], fma(v[i][2], v[i][2], fma(v[i][1], v[i][1], v[i][0] * v[i][0]))));
#pragma unroll
        for (int c = 0; c < 4; ++c) v[i][c] *= inv;
    }
}
// Table class: T from the knot table's local Taylor polynomial in eps = q_j (2 dq) + e0 (see static_table_kernel):
//     1 fma (eps) + 2 fma (T) + 1 add (q) + 2 mul + 8 fma = 14 FP64 instructions per knot, U knots interleaved.
template <int SPT, int U>
RBQ_DEV void delta_tile(const double* __restrict__ tile, int len, const double (&p)[SPT], const double (&dq)[SPT],
                        const double (&dq2)[SPT], const double (&e0)[SPT], double (&v)[SPT][4]) {
    int j = 0;
    for (; j + U <= len; j += U) {
        double q[U][SPT], t[U][SPT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double* __restrict__ r = tile + ST_EPK * (j + u);
#pragma unroll
            for (int i = 0; i < SPT; ++i) {
                const double eps = fma(r[0], dq2[i], e0[i]);
                q[u][i] = r[0] + dq[i];
                t[u][i] = fma(fma(r[3], eps, r[2]), eps, r[1]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < SPT; ++i) tan_apply(t[u][i] * p[i], t[u][i] * q[u][i], v[i]);
    }
    for (; j < len; ++j) {
        const double* __restrict__ r = tile + ST_EPK * j;
#pragma unroll
        for (int i = 0; i < SPT; ++i) {
            const double eps = fma(r[0], dq2[i], e0[i]);
            const double t = fma(fma(r[3], eps, r[2]), eps, r[1]);
            tan_apply(t * p[i], t * (r[0] + dq[i]), v[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
        const double inv = rsqrt(fma(v[i][3], v[i][3], fma(v[i][2], v[i][2], fma(v[i][1], v[i][1], v[i][0] * v[i][0]))));
#pragma unroll
        for (int c = 0; c < 4; ++c) v[i][c] *= inv;
    }
}
// The same step for one sample per thread, laid out for the register file (tools/sass_fp64_cost.py, profiles/README.md): a warp-wide
// FP64 instruction holds the pipe for 2.0 cycles when at most two of its register operands must be fetched, 2.9 with three
// (tools/proto/regbank.cu); an operand is free when the previous instruction carried it in the same slot (operand reuse cache).  So the batch
// runs in two phases that the scheduler cannot interleave (`once` is 1 at run time, unknown at compile time: the second phase is
// a loop body of its own): first the scalars of all U knots level by level, each level's sample-invariant operands in fixed slots
// (dq2 and e0, dq, p), then the U dependent updates, every group of four sharing its multiplier and the first instruction of the
// second group taking its multiplicand from the last one of the first.
#ifndef RBQ_MC_DEFAULT_VARIANT
#define RBQ_MC_DEFAULT_VARIANT 3161
#endif
#ifndef RBQ_DELTA_PHASED
#define RBQ_DELTA_PHASED 1
#endif
#ifndef RBQ_DELTA_PAIR
#define RBQ_DELTA_PAIR 0
#endif
#ifndef RBQ_DELTA_U
#define RBQ_DELTA_U 8
#endif
// FP64 operations as volatile PTX: the compiler keeps volatile statements in program order
RBQ_DEV double vfma(double a, double b, double c) { double d; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c)); return d; }
RBQ_DEV double vadd(double a, double b) { double d; asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b)); return d; }
RBQ_DEV double vmul(double a, double b) { double d; asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b)); return d; }
template <int U>
RBQ_DEV void delta_tile_phased(const double* __restrict__ tile, int len, double p, double dq, double dq2, double e0, double (&v)[4],
                               int once) {
    asm volatile("" : "+d"(dq2));               // (keeps dq2 in its register: the compiler would recompute dq + dq per batch)
    int j = 0;
    for (; j + U <= len; j += U) {
        double tp[U], tq[U];
        {
            double eps[U], q[U], h[U];
            const double* __restrict__ r = tile + ST_EPK * j;
            double r0[U], r1[U], r2[U], r3[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double2 lo = *reinterpret_cast<const double2*>(r + ST_EPK * u), hi = *reinterpret_cast<const double2*>(r + ST_EPK * u + 2);
                r0[u] = lo.x; r1[u] = lo.y; r2[u] = hi.x; r3[u] = hi.y;
            }
#if RBQ_DELTA_PAIR
#pragma unroll
            for (int u = 0; u < U; u += RBQ_DELTA_PAIR) {
#pragma unroll
                for (int w = 0; w < RBQ_DELTA_PAIR; ++w) eps[u + w] = vfma(r0[u + w], dq2, e0);
#pragma unroll
                for (int w = 0; w < RBQ_DELTA_PAIR; ++w) q[u + w] = vadd(r0[u + w], dq);
#pragma unroll
                for (int w = 0; w < RBQ_DELTA_PAIR; ++w) h[u + w] = vfma(r3[u + w], eps[u + w], r2[u + w]);
#pragma unroll
                for (int w = 0; w < RBQ_DELTA_PAIR; ++w) h[u + w] = vfma(h[u + w], eps[u + w], r1[u + w]);
#pragma unroll
                for (int w = 0; w < RBQ_DELTA_PAIR; ++w) tp[u + w] = vmul(h[u + w], p);
#pragma unroll
                for (int w = 0; w < RBQ_DELTA_PAIR; ++w) tq[u + w] = vmul(h[u + w], q[u + w]);
            }
#else
#pragma unroll
            for (int u = 0; u < U; ++u) eps[u] = vfma(r0[u], dq2, e0);
#pragma unroll
            for (int u = 0; u < U; ++u) q[u] = vadd(r0[u], dq);
#pragma unroll
            for (int u = 0; u < U; ++u) h[u] = vfma(r3[u], eps[u], r2[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) h[u] = vfma(h[u], eps[u], r1[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) tp[u] = vmul(h[u], p);
#pragma unroll
            for (int u = 0; u < U; ++u) tq[u] = vmul(h[u], q[u]);
#endif
        }
        for (int rep = 0; rep < once; ++rep) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
                const double x0 = fma(tp[u], a2, a0);
                const double x3 = fma(tp[u], a1, a3);
                const double x2 = fma(-tp[u], a0, a2);
                const double x1 = fma(-tp[u], a3, a1);
                v[0] = fma(tq[u], a3, x0);
                v[2] = fma(-tq[u], a1, x2);
                v[3] = fma(-tq[u], a0, x3);
                v[1] = fma(tq[u], a2, x1);
            }
        }
    }
    for (; j < len; ++j) {
        const double* __restrict__ r = tile + ST_EPK * j;
        const double eps = fma(r[0], dq2, e0);
        const double t = fma(fma(r[3], eps, r[2]), eps, r[1]);
        tan_apply(t * p, t * (r[0] + dq), v);
    }
    const double inv = rsqrt(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] *= inv;
}
// the 4x4 iso matrix whose first column is v (real iso of an SU(2) element)
RBQ_DEV void iso_from_column(const double* v, double* E) {
    E[0] = v[0];  E[1] = -v[1]; E[2] = -v[2];  E[3] = -v[3];
    E[4] = v[1];  E[5] = v[0];  E[6] = -v[3];  E[7] = v[2];
    E[8] = v[2];  E[9] = v[3];  E[10] = v[0];  E[11] = -v[1];
    E[12] = v[3]; E[13] = -v[2]; E[14] = v[1]; E[15] = v[0];
}
RBQ_DEV void mv4_to(const double* E, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = fma(E[4 * i + 3], x[3], fma(E[4 * i + 2], x[2], fma(E[4 * i + 1], x[1], E[4 * i] * x[0])));
}
// compute_fidelities (spin.jl:1198-1240): apply the gate propagator gate_count times, compare with
// g^(k mod 4) psi0 after each; returns the last fidelity
RBQ_DEV double gate_train(const double* E, const double* psi0, const GateIso& G, int64_t gate_count, double* fid_row) {
    double tg[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) tg[0][i] = psi0[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) mv4_to(G.g[k], psi0, tg[k + 1]);
    double st[4] = {psi0[0], psi0[1], psi0[2], psi0[3]};
    double f = fidelity_vec_iso2(st, tg[0]);
    if (gate_count == 1 && fid_row && (reinterpret_cast<uintptr_t>(fid_row) & 15) == 0) {
        // one gate: the (before, after) pair leaves as one 16-byte store, so a warp writes 512 contiguous bytes (this
        // matters when fid_row is mapped host memory: full sectors over PCIe)
        mv4(E, st);
        const double f1 = fidelity_vec_iso2(st, tg[1]);
        *reinterpret_cast<double2*>(fid_row) = make_double2(f, f1);
        return f1;
    }
    if (fid_row) fid_row[0] = f;
    for (int64_t g = 1; g <= gate_count; ++g) {
        mv4(E, st);
        const int k = (int)(g & 3);
        const double* t = k == 0 ? tg[0] : (k == 1 ? tg[1] : (k == 2 ? tg[2] : tg[3]));
        f = fidelity_vec_iso2(st, t);
        if (fid_row) fid_row[g] = f;
    }
    return f;
}

// ---- static-parameter sweep -------------------------------------------------------------------------------
template <int ALGO, int SPT, int MINB, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB) mc_static_kernel(const McStaticArgs a) {
    __shared__ __align__(128) double tile_s[2 * ST_TILE * ST_EPK];
    __shared__ __align__(8) uint64_t bar_s[2];
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[THREADS / 32];
    const int tid = threadIdx.x;
    for (int b = tid; b < RBQ_HIST_BINS; b += THREADS) hist_s[b] = 0;
    PulsePipeT<ST_TILE, ST_EPK> pipe;
    pipe.init(tile_s, bar_s, a.table, a.nk);

    const int64_t base = (int64_t)blockIdx.x * (THREADS * SPT);
    const double w = RBQ_PI * a.dt;
    // q = q_j + qd with q_j = w (a_j + a0) from the table and qd = w (da - a0)
    double p[SPT], p2[SPT], qd[SPT], qd2[SPT], e0[SPT], psi0[SPT][4];
    bool live[SPT];
    int cls = -1;
    const double a0 = a.scal[2], qmax = a.scal[3], p0 = w * a.scal[1];
    const bool table_ok = ALGO == RBQ_ALGO_SU2 && a.scal[4] != 0.0 && !a.no_table_class;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
        const int64_t s = base + (int64_t)i * THREADS + tid;
        live[i] = s < a.S;
        double fq = 0.0, da = 0.0;
        psi0[i][0] = 1.0; psi0[i][1] = psi0[i][2] = psi0[i][3] = 0.0;
        if (live[i]) {
            if (a.philox) {
                sample_static((uint64_t)(a.first + s), a.seed, a.fq0, a.fq_rel_sigma, a.fq_rel_clip, a.da_sigma, fq, da);
                sample_psi0((uint64_t)(a.first + s), a.seed, psi0[i]);
            } else {
                fq = a.fq[s];
                da = a.da ? a.da[s] : 0.0;
#pragma unroll
                for (int c = 0; c < 4; ++c) psi0[i][c] = a.psi0[4 * s + c];
            }
        }
        p[i] = w * fq; p2[i] = p[i] * p[i]; qd[i] = w * (da - a0);
        qd2[i] = 2.0 * qd[i]; e0[i] = fma(qd[i], qd[i], (p[i] - p0) * (p[i] + p0));
        const double epsb = fma(fabs(qd2[i]), qmax, fabs(e0[i]));        // |eps| <= 2 |q_j| |dq| + |e0| at every knot
        const bool delta_ok = table_ok && DELTA_R3 * epsb * epsb * epsb <= DELTA_TOL;   // false for NaN
        const double qb = qmax + fabs(qd[i]);
        const double xb = fma(qb, qb, p2[i]);
        cls = max(cls, delta_ok ? -1 : (xb == xb ? degree_class(xb) : 3));
    }
    cls = warp_max_class(cls);

    double E[SPT][16];
    if (ALGO == RBQ_ALGO_SU2) {
        constexpr int UNR = 8 / SPT;                 // 8 independent Horner chains per thread
        const double cfA[RBQ_TANA_DEG + 1] = RBQ_TANA_COEFS, cfB[RBQ_TANB_DEG + 1] = RBQ_TANB_COEFS;
        double v[SPT][4];
#pragma unroll
        for (int i = 0; i < SPT; ++i) { v[i][0] = 1.0; v[i][1] = v[i][2] = v[i][3] = 0.0; }
        for (int ti = 0; ti < pipe.ntiles; ++ti) {
            const double* tile = pipe.acquire(ti, false);
            const int len = pipe.tile_len(ti);
            if (cls < 0) {
                if constexpr (SPT == 1 && RBQ_DELTA_PHASED) delta_tile_phased<RBQ_DELTA_U>(tile, len, p[0], qd[0], qd2[0], e0[0], v[0], 1 | a.no_table_class);   // = 1, opaque to the compiler
                else delta_tile<SPT, UNR>(tile, len, p, qd, qd2, e0, v);
            }
            else if (cls == 0) tan_tile<RBQ_TANA_DEG, SPT, UNR>(tile, len, p, p2, qd, v, cfA);
            else if (cls == 1) tan_tile<RBQ_TANB_DEG, SPT, UNR>(tile, len, p, p2, qd, v, cfB);
            else if (cls == 2) su2_tile<8, SPT>(tile, len, p, p2, qd, v);
            else su2_tile<0, SPT>(tile, len, p, p2, qd, v);
            pipe.release(ti + 1 < pipe.ntiles);
        }
        // the exact product is a unit quaternion: dividing out |v| removes the radial rounding error, which
        // for piecewise-constant pulses accumulates coherently (the same (cos, sinc) rounding at every knot)
#pragma unroll
        for (int i = 0; i < SPT; ++i) {
            const double inv = 1.0 / sqrt(fma(v[i][3], v[i][3], fma(v[i][2], v[i][2], fma(v[i][1], v[i][1], v[i][0] * v[i][0]))));
#pragma unroll
            for (int c = 0; c < 4; ++c) v[i][c] *= inv;
            iso_from_column(v[i], E[i]);
        }
    } else {
        // reference algorithm (spin.jl:443-447): prop = exp((H0' + a_j H1) dt) * prop with the generic Pade exp
#pragma unroll
        for (int i = 0; i < SPT; ++i)
#pragma unroll
            for (int c = 0; c < 16; ++c) E[i][c] = (c % 5 == 0) ? 1.0 : 0.0;
        for (int ti = 0; ti < pipe.ntiles; ++ti) {
            const double* tile = pipe.acquire(ti, false);
            const int len = pipe.tile_len(ti);
            for (int j = 0; j < len; ++j) {
                const double wa = tile[ST_EPK * j];
#pragma unroll
                for (int i = 0; i < SPT; ++i) {
                    double G[16], X[16], T[16];
                    generator4(p[i], wa + qd[i], G);
                    expm4_pade(G, X);
                    mm4(X, E[i], T);
#pragma unroll
                    for (int c = 0; c < 16; ++c) E[i][c] = T[c];
                }
            }
            pipe.release(ti + 1 < pipe.ntiles);
        }
    }

    ThreadStats st;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
        if (!live[i]) continue;
        const int64_t s = base + (int64_t)i * THREADS + tid;
        double* row = a.fid ? a.fid + s * (a.gate_count + 1) : nullptr;
        const double f = gate_train(E[i], psi0[i], a.gate, a.gate_count, row);
        if (a.partials) st.add(1.0 - f, hist_s);
    }
    if (a.partials) {
        __syncthreads();
        block_stats_flush(st, hist_s, red_s, a.partials, a.hist);
    }
}

// ---- static-parameter sweep, knot table in constant memory -----------------------------------------------------------------
// Why: the FP64 pipe takes one 64-bit register operand per cycle from the register file (tools/proto/regbank.cu: a warp-wide
// DFMA holds the pipe 2.0 cycles with up to two fetched register operands, 2.9 with three), and the LDS.128 that bring the
// table rows into registers write through the same file (tools/proto/ldsprobe.cu).  Table values that arrive as uniform-register
// operands (LDCU from constant memory) cost neither: no instruction of the scalar phase has more than two register operands,
// whatever order the scheduler picks.  The table of one pulse (32 B per knot) lives in g_ctab, up to CT_MAX knots per launch.
constexpr int CT_MAX = 2000;
constexpr int CT_WIN = 7 * ST_TILE;   // knots per launch of a pulse that does not fit
__constant__ double g_ctab[4 * CT_MAX];

// FORM 0: rows (q_j, c0, c1, c2), T = c0 + eps (c1 + eps c2)        (two table values in one instruction: ptxas loads the rows
//         with vector LDC.64, 82 cycles per knot -- kept for A/B only)
// FORM 1: rows (q_j, rho, sigma, c2), T = c2 (eps (eps + rho) + sigma)  (one table value per instruction, 15 FP64 instructions per knot)
// FORM 2: rows (q_j, c0, c1, c2) with c1 read from shared memory (c1_s, one LDS.128 per two knots) and q_j, c0, c2 as uniform
//         operands: Horner again, 14 instructions, the arithmetic -- and the bits -- of mc_static_kernel's table step
template <int U, int FORM>
RBQ_DEV void delta_const_sweep(int nk, double p, double dq, double dq2, double e0, double (&v)[4], int once, const double* __restrict__ c1_s) {
    asm volatile("" : "+d"(dq2));
    for (int t0 = 0; t0 < nk; t0 += ST_TILE) {          // the norm is restored in the rhythm of the shared-memory form: per ST_TILE knots
        const int len = nk - t0 < ST_TILE ? nk - t0 : ST_TILE;
        int j = 0;
        for (; j + U <= len; j += U) {
            double tp[U], tq[U];
            const int jb = t0 + j;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = 4 * (jb + u);
                const double eps = fma(g_ctab[k], dq2, e0);
                const double q = g_ctab[k] + dq;
                double t;
                if (FORM == 0) t = fma(fma(g_ctab[k + 3], eps, g_ctab[k + 2]), eps, g_ctab[k + 1]);
                else if (FORM == 2) t = fma(fma(g_ctab[k + 3], eps, c1_s[jb + u]), eps, g_ctab[k + 1]);
                else t = fma(eps, eps + g_ctab[k + 1], g_ctab[k + 2]) * g_ctab[k + 3];
                tp[u] = t * p; tq[u] = t * q;
            }
            for (int rep = 0; rep < once; ++rep) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
                    const double x0 = fma(tp[u], a2, a0);
                    const double x3 = fma(tp[u], a1, a3);
                    const double x2 = fma(-tp[u], a0, a2);
                    const double x1 = fma(-tp[u], a3, a1);
                    v[0] = fma(tq[u], a3, x0);
                    v[2] = fma(-tq[u], a1, x2);
                    v[3] = fma(-tq[u], a0, x3);
                    v[1] = fma(tq[u], a2, x1);
                }
            }
        }
        for (; j < len; ++j) {
            const int k = 4 * (t0 + j);
            const double eps = fma(g_ctab[k], dq2, e0);
            double t;
            if (FORM == 0) t = fma(fma(g_ctab[k + 3], eps, g_ctab[k + 2]), eps, g_ctab[k + 1]);
            else if (FORM == 2) t = fma(fma(g_ctab[k + 3], eps, c1_s[t0 + j]), eps, g_ctab[k + 1]);
            else t = fma(eps, eps + g_ctab[k + 1], g_ctab[k + 2]) * g_ctab[k + 3];
            tan_apply(t * p, t * (g_ctab[k] + dq), v);
        }
        const double inv = rsqrt(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] *= inv;
    }
}
// One sample per thread.  A block whose warps all take the table step (the verdict of mc_static_kernel, per sample) reads the
// table from constant memory; any other block walks the shared-memory tiles exactly like mc_static_kernel.  The kernel ends with
// the sample's unit quaternion (32 B to a.quat); fidelities and statistics are mc_static_finish_kernel's: with that epilogue in
// the same kernel the compiler loads the table rows into ordinary registers again (ptxas 12.9: a dot product of the quaternion
// with the loaded initial state still keeps the uniform loads, iso_from_column + one matrix-vector product + the fidelity does
// not, whatever the register cap and wherever the initial state is fetched), and the two extra passes over 32 B per sample cost
// 2.5 % of a 1000-knot sweep.
template <int MINB, int THREADS, int U, int FORM>
__global__ void __launch_bounds__(THREADS, MINB) mc_static_ctab_kernel(const McStaticArgs a) {
    __shared__ __align__(128) double tile_s[2 * ST_TILE * ST_EPK];
    __shared__ __align__(8) uint64_t bar_s[2];
    const int tid = threadIdx.x;
    const int64_t s = (int64_t)blockIdx.x * THREADS + tid;
    const double w = RBQ_PI * a.dt;
    const double a0 = a.scal[2], qmax = a.scal[3], p0 = w * a.scal[1];
    const bool table_ok = a.scal[4] != 0.0 && !a.no_table_class;
    const bool live = s < a.S;
    double fq = 0.0, da = 0.0;
    if (live) {
        if (a.philox) sample_static((uint64_t)(a.first + s), a.seed, a.fq0, a.fq_rel_sigma, a.fq_rel_clip, a.da_sigma, fq, da);
        else { fq = a.fq[s]; da = a.da ? a.da[s] : 0.0; }
    }
    const double p = w * fq, p2 = p * p, qd = w * (da - a0), qd2 = 2.0 * qd, e0 = fma(qd, qd, (p - p0) * (p + p0));
    const double epsb = fma(fabs(qd2), qmax, fabs(e0));
    const bool delta_ok = table_ok && DELTA_R3 * epsb * epsb * epsb <= DELTA_TOL;
    const double qb = qmax + fabs(qd);
    const double xb = fma(qb, qb, p2);
    const int cls = warp_max_class(delta_ok ? -1 : (xb == xb ? degree_class(xb) : 3));
    double v[4] = {1.0, 0.0, 0.0, 0.0};
    // (the vote tells the compiler that the branch is warp-uniform: uniform-register loads are only generated in such regions)
    if (__all_sync(0xffffffffu, __syncthreads_and(cls < 0))) {
        // pulses longer than the constant-memory table take one launch per window of knots; the quaternion travels through a.quat
        if (a.win_first > 0 && live) {
            const double4 q = reinterpret_cast<const double4*>(a.quat)[s];
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        }
        static_assert(2 * ST_TILE * ST_EPK >= CT_MAX, "the tile buffer holds one window's c1 column");
        if (FORM == 2) {                                // this window's c1 column: 8 B per knot into the (otherwise idle) tile buffer
            for (int i = tid; i < a.win_knots; i += THREADS) tile_s[i] = a.table[4 * (a.win_first + i) + 2];
            __syncthreads();
        }
        delta_const_sweep<U, FORM>(a.win_knots, p, qd, qd2, e0, v, 1 | a.no_table_class, tile_s);
    } else if (a.win_first > 0) {
        return;                                          // this block walked the whole pulse in the first window's launch
    } else {
        const double cfA[RBQ_TANA_DEG + 1] = RBQ_TANA_COEFS, cfB[RBQ_TANB_DEG + 1] = RBQ_TANB_COEFS;
        PulsePipeT<ST_TILE, ST_EPK> pipe;
        pipe.init(tile_s, bar_s, a.table, a.nk);
        double pp[1] = {p}, pp2[1] = {p2}, pqd[1] = {qd}, vv[1][4] = {{1.0, 0.0, 0.0, 0.0}};
        for (int ti = 0; ti < pipe.ntiles; ++ti) {
            const double* tile = pipe.acquire(ti, false);
            const int len = pipe.tile_len(ti);
            if (cls < 0) delta_tile_phased<RBQ_DELTA_U>(tile, len, p, qd, qd2, e0, vv[0], 1 | a.no_table_class);
            else if (cls == 0) tan_tile<RBQ_TANA_DEG, 1, 8>(tile, len, pp, pp2, pqd, vv, cfA);
            else if (cls == 1) tan_tile<RBQ_TANB_DEG, 1, 8>(tile, len, pp, pp2, pqd, vv, cfB);
            else if (cls == 2) su2_tile<8, 1>(tile, len, pp, pp2, pqd, vv);
            else su2_tile<0, 1>(tile, len, pp, pp2, pqd, vv);
            pipe.release(ti + 1 < pipe.ntiles);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = vv[0][c];
    }
    if (live) reinterpret_cast<double4*>(a.quat)[s] = make_double4(v[0], v[1], v[2], v[3]);
}
// second half of the constant-memory form: quaternion -> gate propagator -> compute_fidelities -> statistics, the same
// arithmetic and the same block geometry (one BlockPartial per THREADS samples) as mc_static_kernel's epilogue
template <int THREADS>
__global__ void __launch_bounds__(THREADS) mc_static_finish_kernel(const McStaticArgs a) {
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[THREADS / 32];
    const int tid = threadIdx.x;
    for (int b = tid; b < RBQ_HIST_BINS; b += THREADS) hist_s[b] = 0;
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * THREADS + tid;
    ThreadStats st;
    if (s < a.S) {
        const double4 q = reinterpret_cast<const double4*>(a.quat)[s];
        double v[4] = {q.x, q.y, q.z, q.w}, psi0[4], E[16];
        if (a.philox) sample_psi0((uint64_t)(a.first + s), a.seed, psi0);
        else {
#pragma unroll
            for (int c = 0; c < 4; ++c) psi0[c] = a.psi0[4 * s + c];
        }
        // the exact product is a unit quaternion: dividing out |v| removes the radial rounding error (see mc_static_kernel)
        const double inv = 1.0 / sqrt(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] *= inv;
        iso_from_column(v, E);
        double* row = a.fid ? a.fid + s * (a.gate_count + 1) : nullptr;
        const double f = gate_train(E, psi0, a.gate, a.gate_count, row);
        if (a.partials) st.add(1.0 - f, hist_s);
    }
    if (a.partials) {
        __syncthreads();
        block_stats_flush(st, hist_s, red_s, a.partials, a.hist);
    }
}

// ---- time-dependent amplitude noise --------------------------------------------------------------------------
// white -> pink filter state of gen_pink_noise (spin.jl:1178-1192), generated on demand
struct PinkGen {
    uint64_t index, seed;
    double x1 = 0, x2 = 0, x3 = 0, y1 = 0, y2 = 0, y3 = 0, spare = 0;
    int64_t produced = 0;   // number of pink samples produced so far
    double scale_dt_inv, namp, last = 0;
    RBQ_DEV double next() {
        double x0;
        if ((produced & 1) == 0) philox_normal2(index, seed, (uint32_t)(produced >> 1), 2u, x0, spare);
        else x0 = spare;
        return filter(x0);
    }
    // one step of the white -> pink filter on a normal deviate drawn elsewhere (pink_filter_kernel)
    RBQ_DEV double filter(double x0) {
        double y = 0.049922035 * x0;
        if (produced >= 1) y += -0.095993537 * x1 - (-2.494956002) * y1;
        if (produced >= 2) y += 0.050612699 * x2 - 2.017265875 * y2;
        if (produced >= 3) y += -0.004408786 * x3 - (-0.522189400) * y3;
        x3 = x2; x2 = x1; x1 = x0; y3 = y2; y2 = y1; y1 = y;
        ++produced;
        last = (y * scale_dt_inv) * namp;
        return last;
    }
    RBQ_DEV double at(int64_t idx) {   // idx is non-decreasing over the rollout
        while (produced <= idx) next();
        return last;
    }
};

// one normalised step with any x (used where a noise sample pushes x outside the polynomial's range)
RBQ_DEV void general_step(double p, double q, double x, double* v) {
    double c, s;
    su2_general(x, c, s);
    Prop<double> E{c, s * p, s * q};
    prop_apply(E, v);
}
// `len` knots of one tile.  idx[j] = noise-sample index of knot j (host-built table, same for every sample), the noise
// amplitude is re-read only when the index changes (a warp-uniform branch).  DEG > 0: tan form in batches of 4 knots.
template <int DEG>
RBQ_DEV void noise_tile(const double* __restrict__ tile, const int32_t* __restrict__ idx, int len, double w, double p, double p2,
                        const double* __restrict__ row, PinkGen* pg, int32_t& cur, double& qd, double* v,
                        const double (&cf)[(DEG > 0 ? DEG : 1) + 1], double xmax) {
    int j = 0;
    if (DEG > 0) {
        for (; j + 4 <= len; j += 4) {
            double q[4], x[4], t[4];
            bool ok = true;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int32_t iu = idx[j + u];
                if (iu != cur) { cur = iu; qd = w * (row ? row[iu] : pg->at(iu)); }
                q[u] = tile[ST_EPK * (j + u)] + qd;
                x[u] = fma(q[u], q[u], p2);
                ok = ok && (x[u] <= xmax);
            }
            if (ok) {
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = fma(cf[DEG], x[u], cf[DEG - 1]);
#pragma unroll
                for (int k = DEG - 2; k >= 0; --k)
#pragma unroll
                    for (int u = 0; u < 4; ++u) t[u] = fma(t[u], x[u], cf[k]);
#pragma unroll
                for (int u = 0; u < 4; ++u) tan_apply(t[u] * p, t[u] * q[u], v);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) general_step(p, q[u], x[u], v);
            }
        }
    }
    for (; j < len; ++j) {
        const int32_t iu = idx[j];
        if (iu != cur) { cur = iu; qd = w * (row ? row[iu] : pg->at(iu)); }
        const double q = tile[ST_EPK * j] + qd;
        general_step(p, q, fma(q, q, p2), v);
    }
}

// Table class of the noise sweeps (all samples share f_q = the table's centre, so eps = q_j (2 dq) + dq^2 with dq = w * noise):
// the noise amplitude and everything derived from it -- dq, 2 dq, dq^2 and the verdict "eps stays inside the table's range" --
// change only when the noise index does (a warp-uniform branch), so a knot costs 14 FP64 instructions and no compare.
RBQ_DEV void noise_tile_table(const double* __restrict__ tile, const int32_t* __restrict__ idx, int len, double w, double p, double p2,
                              double qmax, const double* __restrict__ row, PinkGen* pg, int32_t& cur, double& qd, double& qd2,
                              double& qsq, bool& in_range, double* v) {
    int j = 0;
    constexpr int NU = 4;                                    // knots per batch (independent chains; 8 gains nothing and spills)
    for (; j + NU <= len; j += NU) {
        double q[NU], t[NU];
        bool ok = true;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int32_t iu = idx[j + u];
            if (iu != cur) {
                cur = iu; qd = w * (row ? row[iu] : pg->at(iu)); qd2 = 2.0 * qd; qsq = qd * qd;
                const double eb = fma(fabs(qd2), qmax, qsq);
                in_range = DELTA_R3 * eb * eb * eb <= DELTA_TOL;             // false for NaN
            }
            ok = ok && in_range;
            const double* __restrict__ r = tile + ST_EPK * (j + u);
            const double eps = fma(r[0], qd2, qsq);
            q[u] = r[0] + qd;
            t[u] = fma(fma(r[3], eps, r[2]), eps, r[1]);
        }
        if (ok) {
#pragma unroll
            for (int u = 0; u < NU; ++u) tan_apply(t[u] * p, t[u] * q[u], v);
        } else {
#pragma unroll
            for (int u = 0; u < NU; ++u) general_step(p, q[u], fma(q[u], q[u], p2), v);
        }
    }
    for (; j < len; ++j) {
        const int32_t iu = idx[j];
        if (iu != cur) {
            cur = iu; qd = w * (row ? row[iu] : pg->at(iu)); qd2 = 2.0 * qd; qsq = qd * qd;
            const double eb = fma(fabs(qd2), qmax, qsq);
            in_range = DELTA_R3 * eb * eb * eb <= DELTA_TOL;
        }
        const double q = tile[ST_EPK * j] + qd;
        general_step(p, q, fma(q, q, p2), v);
    }
}

template <int ALGO>
__global__ void __launch_bounds__(MC_THREADS) mc_noise_kernel(const McNoiseArgs a) {
    __shared__ __align__(128) double tile_s[2 * NOISE_TILE * ST_EPK];
    __shared__ __align__(128) int32_t idx_s[2 * NOISE_TILE];
    __shared__ __align__(8) uint64_t bar_s[2];
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[MC_THREADS / 32];
    const int tid = threadIdx.x;
    for (int b = tid; b < RBQ_HIST_BINS; b += MC_THREADS) hist_s[b] = 0;
    PulsePipeT<NOISE_TILE, ST_EPK> pipe;
    pipe.init(tile_s, bar_s, a.table, a.nk, idx_s, a.idx_table, a.idx_stride);

    const int64_t s = (int64_t)blockIdx.x * MC_THREADS + tid;
    const bool live = s < a.S;
    const int64_t sc = live ? s : a.S - 1;   // dead lanes shadow the last sample so the block stays in lock step
    const double w = RBQ_PI * a.dt;
    const double p = w * a.fq, p2 = p * p;
    const double* row = a.noise ? a.noise + sc * a.noise_len : nullptr;   // explicit or materialised rows; NULL: generated on the fly
    PinkGen pg;
    pg.index = (uint64_t)(a.first + sc); pg.seed = a.seed; pg.scale_dt_inv = a.noise_dt_inv; pg.namp = a.namp;
    double psi0[4];
    if (a.psi0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) psi0[c] = a.psi0[4 * sc + c];
    } else sample_psi0((uint64_t)(a.first + sc), a.seed, psi0);
    // propagator class from the noiseless pulse with head-room; noise_tile re-checks x at every knot
    int cls;
    {
        const double qb = a.scal[3] * 1.05;                  // max |q_j| of the noiseless pulse
        const double xb = fma(qb, qb, p2);
        cls = xb == xb ? degree_class(xb) : 3;
    }
    // every sample shares f_q, the centre the knot table was built around: the table step applies wherever the noise is small
    const bool use_table = a.scal[4] != 0.0 && !a.no_table_class;
    const double qmax = a.scal[3];
    double qd2 = 0.0, qsq = 0.0;
    bool in_range = true;
    const double cfA[RBQ_TANA_DEG + 1] = RBQ_TANA_COEFS, cfB[RBQ_TANB_DEG + 1] = RBQ_TANB_COEFS, cf0[2] = {0.0, 0.0};
    double tg[4][4], st[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { tg[0][i] = psi0[i]; st[i] = psi0[i]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) mv4_to(a.gate.g[k], psi0, tg[k + 1]);
    double* frow = (a.fid && live) ? a.fid + s * (a.gate_count + 1) : nullptr;
    double f = fidelity_vec_iso2(st, tg[0]);
    if (frow) frow[0] = f;
    const double norm0 = fma(st[3], st[3], fma(st[2], st[2], fma(st[1], st[1], st[0] * st[0])));
    double qd = 0.0;
    int32_t cur = -1;
    for (int64_t g = 1; g <= a.gate_count; ++g) {
        for (int ti = 0; ti < pipe.ntiles; ++ti) {
            const double* tile = pipe.acquire(ti, g < a.gate_count, g - 1);
            const int32_t* idx = pipe.aux();
            const int len = pipe.tile_len(ti);
            if (ALGO == RBQ_ALGO_SU2) {
                if (use_table) noise_tile_table(tile, idx, len, w, p, p2, qmax, row, &pg, cur, qd, qd2, qsq, in_range, st);
                else if (cls == 0) noise_tile<RBQ_TANA_DEG>(tile, idx, len, w, p, p2, row, &pg, cur, qd, st, cfA, RBQ_TANA_XMAX);
                else if (cls == 1) noise_tile<RBQ_TANB_DEG>(tile, idx, len, w, p, p2, row, &pg, cur, qd, st, cfB, RBQ_TANB_XMAX);
                else noise_tile<0>(tile, idx, len, w, p, p2, row, &pg, cur, qd, st, cf0, 0.0);
                // unitary evolution preserves the norm: shed the tan form's 1/cos growth and the radial rounding
                const double n2 = fma(st[3], st[3], fma(st[2], st[2], fma(st[1], st[1], st[0] * st[0])));
                const double sc2 = sqrt(norm0 / n2);
#pragma unroll
                for (int i = 0; i < 4; ++i) st[i] *= sc2;
            } else {
                for (int j = 0; j < len; ++j) {
                    const int32_t iu = idx[j];
                    if (iu != cur) { cur = iu; qd = w * (row ? row[iu] : pg.at(iu)); }
                    double G[16], X[16];
                    generator4(p, tile[ST_EPK * j] + qd, G);
                    expm4_pade(G, X);
                    mv4(X, st);
                }
            }
            pipe.release(g < a.gate_count || ti + 1 < pipe.ntiles);
        }
        const int k = (int)(g & 3);
        f = fidelity_vec_iso2(st, k == 0 ? tg[0] : (k == 1 ? tg[1] : (k == 2 ? tg[2] : tg[3])));
        if (frow) frow[g] = f;
    }
    if (a.partials) {
        ThreadStats ts;
        if (live) ts.add(1.0 - f, hist_s);
        __syncthreads();
        block_stats_flush(ts, hist_s, red_s, a.partials, a.hist);
    }
}

// Gate-parallel form of the same sweep for small sample counts: the evolution over one gate is an SU(2) element whatever
// the state, so thread (sample s, gate g) composes gate g's propagator as a quaternion (S x G independent units instead of
// S chains of G x N dependent knots), and a second kernel multiplies each sample's G quaternions in order and evaluates the
// fidelities.  The reference's largest workload (figures.jl gen_3b: 1000 seeds x 200 gates x 5680 knots) has S = 1000.
__global__ void __launch_bounds__(MC_THREADS) mc_noise_gate_kernel(const McNoiseArgs a, double* __restrict__ quat) {
    __shared__ __align__(128) double tile_s[2 * NOISE_TILE * ST_EPK];
    __shared__ __align__(128) int32_t idx_s[2 * NOISE_TILE];
    __shared__ __align__(8) uint64_t bar_s[2];
    const int64_t g = blockIdx.y;
    PulsePipeT<NOISE_TILE, ST_EPK> pipe;
    pipe.init(tile_s, bar_s, a.table, a.nk, idx_s, a.idx_table, a.idx_stride, g);
    const int64_t s = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    const bool live = s < a.S;
    const int64_t sc = live ? s : a.S - 1;
    const double w = RBQ_PI * a.dt;
    const double p = w * a.fq, p2 = p * p;
    const double* row = a.noise + sc * a.noise_len;
    int cls;
    {
        const double qb = a.scal[3] * 1.05;                  // max |q_j| of the noiseless pulse
        const double xb = fma(qb, qb, p2);
        cls = xb == xb ? degree_class(xb) : 3;
    }
    // every sample shares f_q, the centre the knot table was built around: the table step applies wherever the noise is small
    const bool use_table = a.scal[4] != 0.0 && !a.no_table_class;
    const double qmax = a.scal[3];
    double qd2 = 0.0, qsq = 0.0;
    bool in_range = true;
    const double cfA[RBQ_TANA_DEG + 1] = RBQ_TANA_COEFS, cfB[RBQ_TANB_DEG + 1] = RBQ_TANB_COEFS, cf0[2] = {0.0, 0.0};
    double v[4] = {1.0, 0.0, 0.0, 0.0}, qd = 0.0;
    int32_t cur = -1;
    for (int ti = 0; ti < pipe.ntiles; ++ti) {
        const double* tile = pipe.acquire(ti, false, g);
        const int32_t* idx = pipe.aux();
        const int len = pipe.tile_len(ti);
        if (use_table) noise_tile_table(tile, idx, len, w, p, p2, qmax, row, nullptr, cur, qd, qd2, qsq, in_range, v);
        else if (cls == 0) noise_tile<RBQ_TANA_DEG>(tile, idx, len, w, p, p2, row, nullptr, cur, qd, v, cfA, RBQ_TANA_XMAX);
        else if (cls == 1) noise_tile<RBQ_TANB_DEG>(tile, idx, len, w, p, p2, row, nullptr, cur, qd, v, cfB, RBQ_TANB_XMAX);
        else noise_tile<0>(tile, idx, len, w, p, p2, row, nullptr, cur, qd, v, cf0, 0.0);
        const double inv = rsqrt(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] *= inv;
        pipe.release(ti + 1 < pipe.ntiles);
    }
    if (live) {
        double* q = quat + (s * a.gate_count + g) * 4;
        q[0] = v[0]; q[1] = v[1]; q[2] = v[2]; q[3] = v[3];
    }
}
__global__ void __launch_bounds__(MC_THREADS) mc_noise_combine_kernel(const McNoiseArgs a, const double* __restrict__ quat) {
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[MC_THREADS / 32];
    const int tid = threadIdx.x;
    for (int b = tid; b < RBQ_HIST_BINS; b += MC_THREADS) hist_s[b] = 0;
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * MC_THREADS + tid;
    const bool live = s < a.S;
    const int64_t sc = live ? s : a.S - 1;
    double psi0[4];
    if (a.psi0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) psi0[c] = a.psi0[4 * sc + c];
    } else sample_psi0((uint64_t)(a.first + sc), a.seed, psi0);
    double tg[4][4], st[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { tg[0][i] = psi0[i]; st[i] = psi0[i]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) mv4_to(a.gate.g[k], psi0, tg[k + 1]);
    double* frow = (a.fid && live) ? a.fid + s * (a.gate_count + 1) : nullptr;
    double f = fidelity_vec_iso2(st, tg[0]);
    if (frow) frow[0] = f;
    const double norm0 = fma(st[3], st[3], fma(st[2], st[2], fma(st[1], st[1], st[0] * st[0])));
    for (int64_t g = 1; g <= a.gate_count; ++g) {
        const double* q = quat + (sc * a.gate_count + (g - 1)) * 4;
        double E[16];
        iso_from_column(q, E);
        mv4(E, st);
        const double n2 = fma(st[3], st[3], fma(st[2], st[2], fma(st[1], st[1], st[0] * st[0])));
        const double sc2 = sqrt(norm0 / n2);
#pragma unroll
        for (int i = 0; i < 4; ++i) st[i] *= sc2;
        const int k = (int)(g & 3);
        f = fidelity_vec_iso2(st, k == 0 ? tg[0] : (k == 1 ? tg[1] : (k == 2 ? tg[2] : tg[3])));
        if (frow) frow[g] = f;
    }
    if (a.partials) {
        ThreadStats ts;
        if (live) ts.add(1.0 - f, hist_s);
        __syncthreads();
        block_stats_flush(ts, hist_s, red_s, a.partials, a.hist);
    }
}

// ---- Lindblad T1 (+T2) rollout in the Bloch picture ------------------------------------------------------------
// rho = (t I + r.sigma)/2.  With equal up/down rates gamma_1 (spin.jl:529-531) the map is unital:
//   dr/dt = Omega x r - diag(g2, g2, g1) r,  Omega = 2 pi (a, 0, fq),  g2 = gamma_1 + Gamma(t), g1 = 2 gamma_1,
// the real 3x3 equivalent of the complex 4x4 vec-Liouvillian the reference exponentiates.  exp(M dt) v is
// summed as a Taylor series whose length comes from a bound on ||M dt||.
// wxa / wxb: the (y,z) and (z,y) couplings.  They are both the rotation rate wx except when the dephasing rate runs with time
// (T2 with gamma_f != 0): Gamma(t) = Gamma(t_mid) + Gamma_1 (t - t_mid) inside a knot does not commute with the rotation
// about x, and the fourth-order Magnus expansion adds the commutator term -(Gamma_1 h^3 / 12) [D, M(t_mid)], D = diag(1,1,0),
// i.e. a SYMMETRIC coupling e = gamma_f^2 h^2 wx / 6 between y and z (wxa = wx - e, wxb = wx + e).  Without it the
// mid-point rule is 1.4e-11 from the exact solution of the reference's ODE after 2000 knots; with it 6e-15.
struct BlochGen { double wz, wxa, wxb, g2, g1; };
// r <- exp(M) r for M = [[-g2, -wz, 0], [wz, -g2, -wxa], [0, wxb, -g1]].  The scalar part is factored out, exp(M) = e^{-g2} exp(N)
// with N = M + g2 I (two rotation rates and the one leftover rate g1 - g2), and the truncated series is summed in Horner form,
//     exp(N) v ~ v + N/1 (v + N/2 (v + ... (v + N/K v))),
// innermost term first: 5 multiply-adds for N u and 3 for v + (N u)/k per term and vector -- no scaling of the generator and
// no separate accumulation (the plain term-by-term sum costs 14).  g2 <= ~1e-6 per knot, so e^{-g2} is its cubic Taylor polynomial.
template <int NV> RBQ_DEV void bloch_advance(const BlochGen& m, int kmax, const double* __restrict__ invk, double (&r)[NV][3]) {
    const double wz = m.wz, wxa = m.wxa, wxb = m.wxb, dl = m.g1 - m.g2, g2 = m.g2;
    double u[NV][3];
#pragma unroll
    for (int c = 0; c < NV; ++c) { u[c][0] = r[c][0]; u[c][1] = r[c][1]; u[c][2] = r[c][2]; }
    for (int k = kmax; k >= 1; --k) {
        const double ik = invk[k];   // 1/k from shared memory
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            const double n0 = -wz * u[c][1];
            const double n1 = fma(wz, u[c][0], -wxa * u[c][2]);
            const double n2 = fma(wxb, u[c][1], -dl * u[c][2]);
            u[c][0] = fma(n0, ik, r[c][0]); u[c][1] = fma(n1, ik, r[c][1]); u[c][2] = fma(n2, ik, r[c][2]);
        }
    }
    // e^{-g2} - 1 = -g2 + g2^2/2 - g2^3/6 (+ g2^4/24 < 1e-17 for g2 < 1e-4; larger decay per knot takes the library expm1),
    // applied as u + (e^{-g2} - 1) u: a factor e^{-g2} rounded to binary64 would carry the same ~1e-16 relative error at every
    // knot of a constant-amplitude segment and bias the norm coherently (1.6e-13 after 6 gates of the 1947-knot Y/2 pulse)
    const double dec = g2 < 1e-4 && g2 > -1e-4 ? g2 * fma(g2, fma(g2, -1.0 / 6.0, 0.5), -1.0) : expm1(-g2);
#pragma unroll
    for (int c = 0; c < NV; ++c) { r[c][0] = fma(dec, u[c][0], u[c][0]); r[c][1] = fma(dec, u[c][1], u[c][1]); r[c][2] = fma(dec, u[c][2], u[c][2]); }
}
// T1 (ns) of amp_t1_spline (spin.jl:366, 393): gridded-linear in fbfq = 0.5 - 0.202407 |a| with flat extrapolation, for plain
// doubles.  The table's knots are multiples of 0.001, so the interval comes from a 241-entry shared-memory index of 0.001-wide
// cells plus a fix-up for arguments within rounding of a knot (same interval choice as t1_interp's search: the last knot <= x).
RBQ_DEV void t1_cells_init(unsigned char* cells, int tid, int nthreads) {
    for (int c = tid; c < 241; c += nthreads) {
        const double x = 0.26 + 0.001 * (double)c + 0.0005;      // cell centre
        int i = 0;
        while (i < 16 && kFbfq[i + 1] <= x) ++i;
        cells[c] = (unsigned char)i;
    }
}
RBQ_DEV double t1_ns_fast(double amplitude, const unsigned char* __restrict__ cells) {
    double x = fma(-fabs(amplitude), 0.202407, 0.5);
    x = x > kFbfq[17] ? kFbfq[17] : (x < kFbfq[0] ? kFbfq[0] : x);                 // NaN falls through and propagates
    int c = (int)((x - 0.26) * 1000.0);
    c = c < 0 ? 0 : (c > 240 ? 240 : c);
    int i = cells[c];
    if (x < kFbfq[i]) --i;                                      // x sits within rounding of the knot that bounds its cell
    else if (i < 16 && kFbfq[i + 1] <= x) ++i;
    const double x0 = kFbfq[i], h = kFbfq[i + 1] - kFbfq[i];
    const double w = (x - x0) * (1.0 / h);
    return (1.0 - w) * (kT1us[i] * 1e3) + w * (kT1us[i + 1] * 1e3);   // the same expression t1_interp evaluates
}
RBQ_DEV int taylor_terms(double b) {   // smallest k with b^(k+1)/(k+1)! < 1e-18 (b >= 0), capped
    double term = b;
    int k = 1;
    while (term > 1e-18 && k < 60) { ++k; term *= b / (double)k; }
    return k;
}
// Uhlmann sqrt-fidelity of 2x2 states in Bloch form: sqrt(tr(r1 r2) + 2 sqrt(det r1 det r2))
RBQ_DEV double fidelity_bloch(double t1, const double* r1, double t2, const double* r2) {
    const double trp = 0.5 * fma(t1, t2, fma(r1[0], r2[0], fma(r1[1], r2[1], r1[2] * r2[2])));
    const double d1 = 0.25 * (t1 * t1 - (r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]));
    const double d2 = 0.25 * (t2 * t2 - (r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]));
    const double dd = fmax(d1, 0.0) * fmax(d2, 0.0);
    return sqrt(fmax(trp + 2.0 * sqrt(dd), 0.0));
}
RBQ_DEV void mv3_to(const double* R, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = fma(R[3 * i + 2], x[2], fma(R[3 * i + 1], x[1], R[3 * i] * x[0]));
}

__global__ void __launch_bounds__(MC_THREADS) lindblad_kernel(const LindbladArgs a, int channels) {
    __shared__ __align__(128) double tile_s[2 * MC_TILE];
    __shared__ __align__(8) uint64_t bar_s[2];
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[MC_THREADS / 32];
    __shared__ int kmax_s, sub_s;
    __shared__ double invk_s[64];
    __shared__ unsigned char t1cells_s[244];
    const int tid = threadIdx.x;
    for (int b = tid; b < RBQ_HIST_BINS; b += MC_THREADS) hist_s[b] = 0;
    if (tid < 64) invk_s[tid] = tid ? 1.0 / (double)tid : 0.0;
    t1_cells_init(t1cells_s, tid, (int)blockDim.x);
    if (tid == 0) { kmax_s = 1; sub_s = 1; }
    PulsePipe pipe;
    if (!a.shared_map) pipe.init(tile_s, bar_s, a.controls, a.nk);   // with a shared one-gate map the pulse is not walked here
    else __syncthreads();

    const int64_t s = (int64_t)blockIdx.x * MC_THREADS + tid;
    const bool live = s < a.S;
    const int64_t sc = live ? s : a.S - 1;
    const double da = a.da ? a.da[sc] : 0.0;
    const bool t1_on = (channels & RBQ_LINDBLAD_T1) != 0, t2_on = (channels & RBQ_LINDBLAD_T2) != 0;
    const bool time_dep = t2_on && a.t2_gamma_f != 0.0;   // Gamma(t) differs from gate to gate
    const double w2 = 2.0 * RBQ_PI * a.dt;
    // rho0: column-major vec, (re,im) pairs: [r00, r10, r01, r11]
    const double* q = a.rho0 + 8 * sc;
    const double tr0 = q[0] + q[6];
    double r0[3] = {q[2] + q[4], q[3] - q[5], q[0] - q[6]};   // x = 2 Re rho10 (Hermitian part), y = 2 Im rho10, z
    // Taylor length from the largest generator norm this block can see
    {
        const double ttot = a.dt * (double)a.nk * (double)a.gate_count;
        const double gmax = (t1_on ? 2.0 / 269.03e3 : 0.0) + (t2_on ? a.t2_gamma_c + 2.0 * a.t2_gamma_f * a.t2_gamma_f * ttot : 0.0);
        const double b = w2 * (fabs(a.fq) + *a.amax + fabs(da)) + gmax * a.dt;
        // coarse grids (||M dt|| > 1): equal sub-steps keep the series short and free of cancellation (its terms peak at
        // ||M||^k/k!); the sub-step count is block uniform like the series length
        const int sub = b == b && b > 1.0 ? (int)fmin(ceil(b), 4096.0) : 1;
        atomicMax(&sub_s, sub);
        __syncthreads();
        atomicMax(&kmax_s, b == b ? taylor_terms(b / (double)sub_s) : 60);
    }
    __syncthreads();
    const int kmax = kmax_s, nsub = sub_s;
    const double hsub = 1.0 / (double)nsub;

    double tg[4][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) tg[0][i] = r0[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) mv3_to(a.gate.r[k], r0, tg[k + 1]);
    double* frow = (a.fid && live) ? a.fid + s * (a.gate_count + 1) : nullptr;
    double f = fidelity_bloch(tr0, r0, tr0, tg[0]);
    if (frow) frow[0] = f;
    double r[3] = {r0[0], r0[1], r0[2]};

    // generator of one (sub-)step of length h = dt / nsub starting at time t0
    auto generator = [&](double aj, double t0, double h) {
        const double c1 = aj + da, w2h = 2.0 * RBQ_PI * h;
        BlochGen m;
        const double wx = w2h * c1;
        m.wz = w2h * a.fq;
        const double e = time_dep ? a.t2_gamma_f * a.t2_gamma_f * h * h * wx * (1.0 / 6.0) : 0.0;   // Magnus commutator term
        m.wxa = wx - e; m.wxb = wx + e;
        const double g1 = t1_on ? h / t1_ns_fast(c1, t1cells_s) : 0.0;   // gamma_1 h, T1 in ns (spin.jl:529)
        const double gphi = t2_on ? h * (a.t2_gamma_c + 2.0 * a.t2_gamma_f * a.t2_gamma_f * (t0 + 0.5 * h)) : 0.0;
        m.g2 = g1 + gphi; m.g1 = 2.0 * g1;
        return m;
    };

    // few gates: carry the state itself (1 vector per knot and gate); many gates: compose the one-gate map once (3 vectors)
    const bool direct = !a.shared_map && (time_dep || a.gate_count <= 3);
    if (!direct) {
        // compose the one-gate map from its action on the three basis vectors, then apply it gate_count times
        double P[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};   // P[c] = image of e_c
        if (a.shared_map) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int i = 0; i < 3; ++i) P[c][i] = a.shared_map[3 * c + i];
        } else {
            for (int ti = 0; ti < pipe.ntiles; ++ti) {
                const double* tile = pipe.acquire(ti, false);
                const int len = pipe.tile_len(ti);
                for (int j = 0; j < len; ++j) {
                    const BlochGen mj = generator(tile[j], 0.0, a.dt * hsub);   // one of nsub equal sub-steps (rates constant here)
                    if (nsub == 1) bloch_advance<3>(mj, kmax, invk_s, P);
                    else for (int q2 = 0; q2 < nsub; ++q2) bloch_advance<3>(mj, kmax, invk_s, P);
                }
                pipe.release(ti + 1 < pipe.ntiles);
            }
        }
        for (int64_t g = 1; g <= a.gate_count; ++g) {
            const double n0 = fma(P[2][0], r[2], fma(P[1][0], r[1], P[0][0] * r[0]));
            const double n1 = fma(P[2][1], r[2], fma(P[1][1], r[1], P[0][1] * r[0]));
            const double n2 = fma(P[2][2], r[2], fma(P[1][2], r[1], P[0][2] * r[0]));
            r[0] = n0; r[1] = n1; r[2] = n2;
            const int k = (int)(g & 3);
            f = fidelity_bloch(tr0, r, tr0, k == 0 ? tg[0] : (k == 1 ? tg[1] : (k == 2 ? tg[2] : tg[3])));
            if (frow) frow[g] = f;
        }
    } else {
        double rr[1][3] = {{r[0], r[1], r[2]}};
        int64_t step = 0;
        for (int64_t g = 1; g <= a.gate_count; ++g) {
            for (int ti = 0; ti < pipe.ntiles; ++ti) {
                const double* tile = pipe.acquire(ti, g < a.gate_count);
                const int len = pipe.tile_len(ti);
                for (int j = 0; j < len; ++j, ++step) {
                    const double t0 = (double)step * a.dt, hh = a.dt * hsub;
                    if (nsub == 1) bloch_advance<1>(generator(tile[j], t0, hh), kmax, invk_s, rr);
                    else for (int q2 = 0; q2 < nsub; ++q2) bloch_advance<1>(generator(tile[j], t0 + (double)q2 * hh, hh), kmax, invk_s, rr);
                }
                pipe.release(g < a.gate_count || ti + 1 < pipe.ntiles);
            }
            const int k = (int)(g & 3);
            f = fidelity_bloch(tr0, rr[0], tr0, k == 0 ? tg[0] : (k == 1 ? tg[1] : (k == 2 ? tg[2] : tg[3])));
            if (frow) frow[g] = f;
        }
        r[0] = rr[0][0]; r[1] = rr[0][1]; r[2] = rr[0][2];
    }
    if (a.rho_out && live) {
        double* o = a.rho_out + 8 * s;
        o[0] = 0.5 * (tr0 + r[2]); o[1] = 0.0;           // rho00
        o[2] = 0.5 * r[0];         o[3] = 0.5 * r[1];    // rho10 = (x + i y)/2
        o[4] = 0.5 * r[0];         o[5] = -0.5 * r[1];   // rho01
        o[6] = 0.5 * (tr0 - r[2]); o[7] = 0.0;           // rho11
    }
    if (a.partials) {
        ThreadStats ts;
        if (live) ts.add(1.0 - f, hist_s);
        __syncthreads();
        block_stats_flush(ts, hist_s, red_s, a.partials, a.hist);
    }
}

// ---- Lindblad sweep with per-realisation amplitude offsets: per-knot Frechet table -------------------------------------------
// A realisation's generator at knot k is M_k(d) = M_k + d D1_k + d^2 D2_k + d^3 D3_k + O(d^4), d = its static offset
// (D1 = the rotation about x plus the slope of gamma_1 = 1/T1(a); D2, D3 = the curvature of 1/T1 inside a linear cell of the
// T1 table), so exp(M_k(d)) = E_k + d L1_k + d^2 L2_k + d^3 L3_k + O(d^4) with 3x3 coefficient matrices that depend on the
// knot only.  lindblad_table_kernel computes them once per sweep (one thread per knot: the Taylor series of exp in
// truncated-polynomial arithmetic, E_k itself in double-double so that its rounding does not repeat coherently along a
// constant-amplitude segment), and the sweep kernel advances a Bloch vector with
//     r <- E_hi r + (E_lo + d (L1 + d (L2 + d L3))) r          27 + 18 = 45 FP64 instructions
// instead of the ~105 of the 13-term series.  Rows with flag != 0 (the offset range straddles a kink of the T1 table or of
// |a|, the curvature terms are not negligible, or ||M dt|| > 1) and realisations with |d| > da_tab take the series step.
constexpr int LT_ROW = 48;      // doubles per table row: E_hi, E_lo, L1, L2, L3 (9 each, row-major), a_k, flag, kappa
constexpr int LT_TILE = 32;     // knots per shared-memory tile (12 KB), two tiles in flight
struct dd { double h, l; };
RBQ_DEV dd dd_two_sum(double a, double b) { const double s = __dadd_rn(a, b), bb = __dadd_rn(s, -a); return {s, __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb))}; }
RBQ_DEV dd dd_add(dd a, dd b) { dd s = dd_two_sum(a.h, b.h); const double e = __dadd_rn(s.l, __dadd_rn(a.l, b.l)); const double h = __dadd_rn(s.h, e); return {h, __dadd_rn(e, -__dadd_rn(h, -s.h))}; }
RBQ_DEV dd dd_mul(dd a, dd b) { const double p = __dmul_rn(a.h, b.h); double e = __fma_rn(a.h, b.h, -p); e = __fma_rn(a.h, b.l, __fma_rn(a.l, b.h, e)); const double h = __dadd_rn(p, e); return {h, __dadd_rn(e, -__dadd_rn(h, -p))}; }
__global__ void __launch_bounds__(128) lindblad_table_kernel(const double* __restrict__ controls, int64_t nk, double dt, double fq,
                                                             int channels, double t2_gamma_c, double da_tab, double* __restrict__ table) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nk) return;
    const bool t1_on = (channels & RBQ_LINDBLAD_T1) != 0, t2_on = (channels & RBQ_LINDBLAD_T2) != 0;
    const double a = controls[k], w2 = 2.0 * RBQ_PI * dt;
    // gamma_1(a + d) dt = dt / (T + T' d) = g1 (1 - rho d + rho^2 d^2 - rho^3 d^3 + ...), rho = T'/T, inside one cell of the table
    double g1 = 0.0, rho = 0.0;
    bool flag = !(fabs(a) > da_tab);                      // |a + d| has its kink inside the offset range (also catches NaN)
    if (t1_on) {
        Tan<1> ta; ta.v = a; ta.d[0] = 1.0;
        const Tan<1> T = t1_interp<Tan<1>>(ta, 1e3);
        g1 = dt / T.v; rho = T.d[0] / T.v;
        // same cell over the whole offset range?  compare the interpolant with its linear model at both ends
        const double Tp = t1_interp<double>(a + da_tab, 1e3), Tm = t1_interp<double>(a - da_tab, 1e3);
        const double lin_p = fma(T.d[0], da_tab, T.v), lin_m = fma(-T.d[0], da_tab, T.v);
        if (fabs(Tp - lin_p) > 1e-9 * T.v || fabs(Tm - lin_m) > 1e-9 * T.v) flag = true;
    }
    // the dropped fourth-order curvature term of gamma_1 dt is g1 (rho d)^4: the sweep kernel compares kappa d_max^4, with
    // d_max the largest offset of the WARP, against 2e-17 (2e-14 after 1000 knots if every knot were at the limit)
    const double kappa = g1 * (rho * rho) * (rho * rho);
    const double gphi = t2_on ? dt * t2_gamma_c : 0.0;
    const double wz = w2 * fq, wx = w2 * a;
    const double nb = w2 * (fabs(fq) + fabs(a) + da_tab) + 2.0 * g1 + gphi;            // bound of ||M(d) dt|| over the offset range
    if (!(nb <= 1.0)) flag = true;                                                      // the sweep kernel sub-steps such knots
    // generator polynomial in d: D[0] = M, D[1] = w2 Lx - g1 rho diag(1,1,2), D[2] = g1 rho^2 (-diag), D[3] = -g1 rho^3 (-diag)
    double D[4][9];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 9; ++i) D[j][i] = 0.0;
    D[0][0] = -(g1 + gphi); D[0][1] = -wz; D[0][3] = wz; D[0][4] = -(g1 + gphi); D[0][5] = -wx; D[0][7] = wx; D[0][8] = -2.0 * g1;
    double cj = g1;
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        cj *= -rho;                                       // g1 (-rho)^j
        D[j][0] = -cj; D[j][4] = -cj; D[j][8] = -2.0 * cj;
    }
    D[1][5] = -w2; D[1][7] = w2;
    // Taylor series of exp(sum_j d^j D_j) truncated at d^3: T_n[j] = sum_{i<=j} T_{n-1}[j-i] D_i / n; order 0 in double-double
    dd T0[9], A0[9];
    double T[3][9], A[3][9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        T0[i] = dd{(i % 4 == 0) ? 1.0 : 0.0, 0.0}; A0[i] = T0[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) { T[j][i] = 0.0; A[j][i] = 0.0; }
    }
    // the series stops when its terms (bounded by nb^n / n!) are below 1e-26: E_hi + E_lo only has to keep the rounding of
    // E_hi from repeating along a segment (1e-26 x 1e4 knots is nothing); 30 terms cover nb = 1
    double term = 1.0;
    for (int n = 1; n <= 30; ++n) {
        const double dn = (double)n, in = 1.0 / dn;
        const dd inv_n{in, fma(-in, dn, 1.0) / dn};        // 1/n to double-double accuracy (1 - in n is exact in one FMA)
        term *= (nb == nb && nb <= 1.0 ? nb : 1.0) * in;
        dd N0[9];
        double N[3][9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                dd acc0{0.0, 0.0};
#pragma unroll
                for (int q = 0; q < 3; ++q) acc0 = dd_add(acc0, dd_mul(T0[3 * r + q], dd{D[0][3 * q + c], 0.0}));
                N0[3 * r + c] = dd_mul(acc0, inv_n);
#pragma unroll
                for (int j = 1; j <= 3; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i <= j; ++i)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const double t = (j - i == 0) ? T0[3 * r + q].h : T[j - i - 1][3 * r + q];
                            acc = fma(t, D[i][3 * q + c], acc);
                        }
                    N[j - 1][3 * r + c] = acc * in;
                }
            }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            T0[i] = N0[i]; A0[i] = dd_add(A0[i], N0[i]);
#pragma unroll
            for (int j = 0; j < 3; ++j) { T[j][i] = N[j][i]; A[j][i] += N[j][i]; }
        }
        if (term < 1e-26) break;
    }
    double* row = table + (size_t)k * LT_ROW;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        row[i] = A0[i].h; row[9 + i] = A0[i].l; row[18 + i] = A[0][i]; row[27 + i] = A[1][i]; row[36 + i] = A[2][i];
    }
    row[45] = a; row[46] = flag ? 1.0 : 0.0; row[47] = kappa;
}

// series step for one (sub-stepped if needed) knot with the realisation's own amplitude: the fallback of the table kernel
template <int NV>
RBQ_DEV void bloch_series_knot(double c1, double fq, double dt, bool t1_on, double gphi_rate, const double* __restrict__ invk,
                               const double* __restrict__ bmax, const unsigned char* __restrict__ cells, double (&r)[NV][3]) {
    const double b = 2.0 * RBQ_PI * dt * (fabs(fq) + fabs(c1)) + dt * ((t1_on ? 2.0 / 269.03e3 : 0.0) + gphi_rate);
    const int nsub = b == b && b > 1.0 ? (int)fmin(ceil(b), 4096.0) : 1;
    const double bs = nsub == 1 ? b : b / (double)nsub;
    // series length from the block's table of bounds (bmax[k] = largest b with b^(k+1)/(k+1)! < 1e-18): no divisions here
    int kmax = 60;
    if (bs == bs) { kmax = 1; while (kmax < 30 && bs > bmax[kmax]) ++kmax; if (bs > bmax[kmax]) kmax = taylor_terms(bs); }
    const double h = nsub == 1 ? dt : dt / (double)nsub, w2h = 2.0 * RBQ_PI * h;
    BlochGen m;
    m.wz = w2h * fq; m.wxa = m.wxb = w2h * c1;
    const double g1 = t1_on ? h / t1_ns_fast(c1, cells) : 0.0;
    m.g2 = g1 + h * gphi_rate; m.g1 = 2.0 * g1;
    for (int q = 0; q < nsub; ++q) bloch_advance<NV>(m, kmax, invk, r);
}

// SPT realisations per thread (they share the table row's registers), NV = 1 (the state itself, gate_count <= 3) or 3 (the
// one-gate map's three columns, applied gate_count times afterwards)
template <int SPT, int NV, int MAXT = 512>
__global__ void __launch_bounds__(MAXT, 1) lindblad_tab_kernel(const LindbladArgs a, int channels, const double* __restrict__ table,
                                                               double da_tab) {
    __shared__ __align__(128) double tile_s[2 * LT_TILE * LT_ROW];
    __shared__ __align__(8) uint64_t bar_s[2];
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[MAXT / 32];
    __shared__ double invk_s[64];
    __shared__ double bmax_s[32];
    __shared__ unsigned char t1cells_s[244];
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int b = tid; b < RBQ_HIST_BINS; b += nth) hist_s[b] = 0;
    if (tid < 64) invk_s[tid] = tid ? 1.0 / (double)tid : 0.0;
    if (tid < 32) {   // bmax[k]: (1e-18 (k+1)!)^(1/(k+1)), a hair inside
        double lf = 0.0;
        for (int i = 2; i <= tid + 1; ++i) lf += log((double)i);
        bmax_s[tid] = 0.999 * exp((log(1e-18) + lf) / (double)(tid + 1));
    }
    t1_cells_init(t1cells_s, tid, nth);
    PulsePipeT<LT_TILE, LT_ROW> pipe;
    pipe.init(tile_s, bar_s, table, a.nk);

    const bool t1_on = (channels & RBQ_LINDBLAD_T1) != 0, t2_on = (channels & RBQ_LINDBLAD_T2) != 0;
    const double gphi_rate = t2_on ? a.t2_gamma_c : 0.0;
    const int64_t base = (int64_t)blockIdx.x * nth * SPT;
    int64_t s[SPT]; bool live[SPT]; double d[SPT], tr0[SPT], r0[SPT][3];
    double st[SPT][NV][3];
    bool in_range = true;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
        s[i] = base + (int64_t)i * nth + tid;
        live[i] = s[i] < a.S;
        const int64_t sc = live[i] ? s[i] : a.S - 1;
        d[i] = a.da[sc];
        in_range = in_range && (fabs(d[i]) <= da_tab);
        const double* q = a.rho0 + 8 * sc;
        tr0[i] = q[0] + q[6];
        r0[i][0] = q[2] + q[4]; r0[i][1] = q[3] - q[5]; r0[i][2] = q[0] - q[6];
#pragma unroll
        for (int c = 0; c < NV; ++c)
#pragma unroll
            for (int x = 0; x < 3; ++x) st[i][c][x] = NV == 1 ? r0[i][x] : (c == x ? 1.0 : 0.0);
    }
    const bool warp_table = __all_sync(0xffffffffu, in_range);
    double dmax4;
    {
        double dm = 0.0;
#pragma unroll
        for (int i = 0; i < SPT; ++i) dm = fmax(dm, fabs(d[i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        dmax4 = (dm * dm) * (dm * dm);
    }
    const int64_t passes = NV == 1 ? a.gate_count : 1;
    for (int64_t g = 0; g < passes; ++g) {
        for (int ti = 0; ti < pipe.ntiles; ++ti) {
            const double* tile = pipe.acquire(ti, g + 1 < passes);
            const int len = pipe.tile_len(ti);
            for (int j = 0; j < len; ++j) {      // (unrolling by 2 was measured: 0.613 against 0.583 ms for 1e5 x 1000)
                const double2* __restrict__ row = reinterpret_cast<const double2*>(tile + j * LT_ROW);
                const double2 af = row[22];                            // (E/L3 tail, a_k) sits in row[22] = doubles 44, 45
                const double2 fk = row[23];                            // (flag, kappa)
                if (warp_table && fk.x == 0.0 && fk.y * dmax4 <= 2e-17) {
                    double tb[46];
#pragma unroll
                    for (int q = 0; q < 23; ++q) { const double2 v = row[q]; tb[2 * q] = v.x; tb[2 * q + 1] = v.y; }
#pragma unroll
                    for (int i = 0; i < SPT; ++i) {
                        double F[9];
#pragma unroll
                        for (int e = 0; e < 9; ++e) F[e] = fma(d[i], fma(d[i], fma(d[i], tb[36 + e], tb[27 + e]), tb[18 + e]), tb[9 + e]);
#pragma unroll
                        for (int c = 0; c < NV; ++c) {
                            const double x0 = st[i][c][0], x1 = st[i][c][1], x2 = st[i][c][2];
#pragma unroll
                            for (int rr = 0; rr < 3; ++rr) {
                                const double lo = fma(F[3 * rr + 2], x2, fma(F[3 * rr + 1], x1, F[3 * rr] * x0));
                                st[i][c][rr] = fma(tb[3 * rr + 2], x2, fma(tb[3 * rr + 1], x1, fma(tb[3 * rr], x0, lo)));
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < SPT; ++i) bloch_series_knot<NV>(af.y + d[i], a.fq, a.dt, t1_on, gphi_rate, invk_s, bmax_s, t1cells_s, st[i]);
                }
            }
            pipe.release(g + 1 < passes || ti + 1 < pipe.ntiles);
        }
        if (NV == 1) {
#pragma unroll
            for (int i = 0; i < SPT; ++i) {
                if (a.fid && live[i]) {
                    double tg[3];
                    const int k = (int)((g + 1) & 3);
                    if (k == 0) { tg[0] = r0[i][0]; tg[1] = r0[i][1]; tg[2] = r0[i][2]; }
                    else mv3_to(a.gate.r[k - 1], r0[i], tg);
                    a.fid[s[i] * (a.gate_count + 1) + g + 1] = fidelity_bloch(tr0[i], st[i][0], tr0[i], tg);
                }
            }
        }
    }
    ThreadStats ts;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
        double r[3] = {r0[i][0], r0[i][1], r0[i][2]};
        double f = fidelity_bloch(tr0[i], r, tr0[i], r);
        double* frow = (a.fid && live[i]) ? a.fid + s[i] * (a.gate_count + 1) : nullptr;
        if (frow) frow[0] = f;
        if (NV == 1) {
            if (a.gate_count > 0) {
                r[0] = st[i][0][0]; r[1] = st[i][0][1]; r[2] = st[i][0][2];
                double tg[3];
                const int k = (int)(a.gate_count & 3);
                if (k == 0) { tg[0] = r0[i][0]; tg[1] = r0[i][1]; tg[2] = r0[i][2]; }
                else mv3_to(a.gate.r[k - 1], r0[i], tg);
                f = fidelity_bloch(tr0[i], r, tr0[i], tg);
            }
        } else {
            for (int64_t g = 1; g <= a.gate_count; ++g) {
                const double n0 = fma(st[i][2][0], r[2], fma(st[i][1][0], r[1], st[i][0][0] * r[0]));
                const double n1 = fma(st[i][2][1], r[2], fma(st[i][1][1], r[1], st[i][0][1] * r[0]));
                const double n2 = fma(st[i][2][2], r[2], fma(st[i][1][2], r[1], st[i][0][2] * r[0]));
                r[0] = n0; r[1] = n1; r[2] = n2;
                double tg[3];
                const int k = (int)(g & 3);
                if (k == 0) { tg[0] = r0[i][0]; tg[1] = r0[i][1]; tg[2] = r0[i][2]; }
                else mv3_to(a.gate.r[k - 1], r0[i], tg);
                f = fidelity_bloch(tr0[i], r, tr0[i], tg);
                if (frow) frow[g] = f;
            }
        }
        if (a.rho_out && live[i]) {
            double* o = a.rho_out + 8 * s[i];
            o[0] = 0.5 * (tr0[i] + r[2]); o[1] = 0.0;
            o[2] = 0.5 * r[0];            o[3] = 0.5 * r[1];
            o[4] = 0.5 * r[0];            o[5] = -0.5 * r[1];
            o[6] = 0.5 * (tr0[i] - r[2]); o[7] = 0.0;
        }
        if (a.partials && live[i]) ts.add(1.0 - f, hist_s);
    }
    if (a.partials) {
        __syncthreads();
        block_stats_flush(ts, hist_s, red_s, a.partials, a.hist);
    }
}

// One-gate Bloch map shared by every realisation (no per-sample amplitude offsets, time-independent rates): what the
// reference recomputes for each of its 1000 seeds (figures.jl gen_1c).  One block: thread k exponentiates knot k's
// generator (3 columns), an ordered tree product in shared memory multiplies 256 knots at a time into the running map.
__global__ void __launch_bounds__(256) lindblad_shared_map_kernel(const LindbladArgs a, int channels, double* __restrict__ map_out) {
    __shared__ double M[256][9];
    __shared__ double invk_s[64];
    __shared__ unsigned char t1cells_s[244];
    __shared__ double run[9];
    const int tid = threadIdx.x;
    if (tid < 64) invk_s[tid] = tid ? 1.0 / (double)tid : 0.0;
    t1_cells_init(t1cells_s, tid, (int)blockDim.x);
    if (tid < 9) run[tid] = (tid % 4 == 0) ? 1.0 : 0.0;
    const bool t1_on = (channels & RBQ_LINDBLAD_T1) != 0, t2_on = (channels & RBQ_LINDBLAD_T2) != 0;
    const double w2 = 2.0 * RBQ_PI * a.dt;
    const double gmax = (t1_on ? 2.0 / 269.03e3 : 0.0) + (t2_on ? a.t2_gamma_c : 0.0);
    const double bnd = w2 * (fabs(a.fq) + *a.amax) + gmax * a.dt;
    const int nsub = bnd == bnd && bnd > 1.0 ? (int)fmin(ceil(bnd), 4096.0) : 1;     // coarse grids: equal sub-steps
    const double hsub = 1.0 / (double)nsub;
    const int kmax = bnd == bnd ? taylor_terms(bnd * hsub) : 60;
    __syncthreads();
    for (int64_t base = 0; base < a.nk; base += 256) {
        const int64_t k = base + tid;
        double P[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        if (k < a.nk) {
            const double c1 = a.controls[k];
            BlochGen m;
            m.wz = w2 * a.fq; m.wxa = m.wxb = w2 * c1;
            const double g1 = t1_on ? a.dt / t1_ns_fast(c1, t1cells_s) : 0.0;
            m.g2 = g1 + (t2_on ? a.dt * a.t2_gamma_c : 0.0); m.g1 = 2.0 * g1;
            if (nsub > 1) { m.wz *= hsub; m.wxa *= hsub; m.wxb *= hsub; m.g2 *= hsub; m.g1 *= hsub; }
            for (int q2 = 0; q2 < nsub; ++q2) bloch_advance<3>(m, kmax, invk_s, P);
        }
        // column-major: M[tid][3c+i] = image of e_c, component i
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < 3; ++i) M[tid][3 * c + i] = P[c][i];
        __syncthreads();
        // ordered product: after the loop M[0] = E_{base+255} ... E_{base+1} E_{base}
        for (int stride = 1; stride < 256; stride <<= 1) {
            double A[9], B[9], Cm[9];
            const bool act = (tid % (2 * stride)) == 0;
            if (act) {
#pragma unroll
                for (int i = 0; i < 9; ++i) { B[i] = M[tid][i]; A[i] = M[tid + stride][i]; }   // later knots multiply from the left
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int i = 0; i < 3; ++i) Cm[3 * c + i] = fma(A[6 + i], B[3 * c + 2], fma(A[3 + i], B[3 * c + 1], A[i] * B[3 * c]));
            }
            __syncthreads();
            if (act) {
#pragma unroll
                for (int i = 0; i < 9; ++i) M[tid][i] = Cm[i];
            }
            __syncthreads();
        }
        if (tid == 0) {
            double A[9], B[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) { A[i] = M[0][i]; B[i] = run[i]; }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int i = 0; i < 3; ++i) run[3 * c + i] = fma(A[6 + i], B[3 * c + 2], fma(A[3 + i], B[3 * c + 1], A[i] * B[3 * c]));
        }
        __syncthreads();
    }
    if (tid < 9) map_out[tid] = run[tid];
}

// ---- analytic gates: a few segments with their own durations (spin.jl:562-594, 651-670, 765-785) -----------------------
__global__ void __launch_bounds__(MC_THREADS) mc_static_segments_kernel(const SegmentArgs a) {
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[MC_THREADS / 32];
    for (int b = threadIdx.x; b < RBQ_HIST_BINS; b += MC_THREADS) hist_s[b] = 0;
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    const bool live = s < a.S;
    const int64_t sc = live ? s : a.S - 1;
    const double fq = a.fq ? a.fq[sc] : a.fq_scalar, da = a.da ? a.da[sc] : 0.0;
    const double* nrow = a.noise ? a.noise + sc * a.noise_len : nullptr;
    double v[4] = {1.0, 0.0, 0.0, 0.0};
    for (int64_t j = 0; j < a.nseg; ++j) {
        const double w = RBQ_PI * a.durs[j];
        const double amp = a.amps[j] + (nrow ? nrow[a.nidx[j]] : da);
        const double p = w * fq, q = w * amp;
        general_step(p, q, fma(q, q, p * p), v);
    }
    const double inv = rsqrt(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] *= inv;
    double E[16], psi0[4];
    iso_from_column(v, E);
#pragma unroll
    for (int c = 0; c < 4; ++c) psi0[c] = a.psi0[4 * sc + c];
    ThreadStats st;
    if (live) {
        const double f = gate_train(E, psi0, a.gate, a.gate_count, a.fid ? a.fid + s * (a.gate_count + 1) : nullptr);
        if (a.partials) st.add(1.0 - f, hist_s);
    }
    if (a.partials) {
        __syncthreads();
        block_stats_flush(st, hist_s, red_s, a.partials, a.hist);
    }
}
// ---- piecewise-constant timeline with save points: the exact solution of the ODE run_sim_deqjl hands to Vern9
//      (dynamics_schroed_deqjl / dynamics_schroedda_deqjl, spin.jl:424-467; run_sim_deqjl, spin.jl:1277-1380).  The right-hand
//      side is constant between the breakpoints of the control grid, the noise grid and the save times, so the state after
//      each segment is exp(segment generator * duration) * state; the host mirror builds the segment list. ----------------------
__global__ void __launch_bounds__(MC_THREADS) mc_timeline_kernel(const SegmentArgs a, const int64_t* __restrict__ save_after, int64_t nsave) {
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[MC_THREADS / 32];
    for (int b = threadIdx.x; b < RBQ_HIST_BINS; b += MC_THREADS) hist_s[b] = 0;
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    const bool live = s < a.S;
    const int64_t sc = live ? s : a.S - 1;
    const double fq = a.fq ? a.fq[sc] : a.fq_scalar, da = a.da ? a.da[sc] : 0.0;
    const double* nrow = a.noise ? a.noise + sc * a.noise_len : nullptr;
    double psi0[4], st[4], tg[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { psi0[c] = a.psi0[4 * sc + c]; st[c] = psi0[c]; tg[0][c] = psi0[c]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) mv4_to(a.gate.g[k], psi0, tg[k + 1]);
    double* frow = (a.fid && live) ? a.fid + s * (nsave + 1) : nullptr;
    double* srow = (a.states_out && live) ? a.states_out + s * (nsave + 1) * 4 : nullptr;     // run_sim_fine_deqjl's "states"
    double f = fidelity_vec_iso2(st, tg[0]);
    if (frow) frow[0] = f;
    if (srow) { srow[0] = st[0]; srow[1] = st[1]; srow[2] = st[2]; srow[3] = st[3]; }
    const double norm0 = fma(st[3], st[3], fma(st[2], st[2], fma(st[1], st[1], st[0] * st[0])));
    int64_t j = 0;
    for (int64_t g = 1; g <= nsave; ++g) {
        const int64_t end = save_after[g - 1];
        for (; j < end; ++j) {
            const double w = RBQ_PI * a.durs[j];
            const double amp = a.amps[j] + da + (nrow ? nrow[a.nidx[j]] : 0.0);
            const double p = w * fq, q = w * amp;
            general_step(p, q, fma(q, q, p * p), st);
        }
        // unitary evolution preserves the norm: shed the radial rounding before the fidelity is taken
        const double n2 = fma(st[3], st[3], fma(st[2], st[2], fma(st[1], st[1], st[0] * st[0])));
        const double sc2 = sqrt(norm0 / n2);
#pragma unroll
        for (int i = 0; i < 4; ++i) st[i] *= sc2;
        const int k = (int)(g & 3);
        f = fidelity_vec_iso2(st, k == 0 ? tg[0] : (k == 1 ? tg[1] : (k == 2 ? tg[2] : tg[3])));
        if (frow) frow[g] = f;
        if (srow) { srow[4 * g] = st[0]; srow[4 * g + 1] = st[1]; srow[4 * g + 2] = st[2]; srow[4 * g + 3] = st[3]; }
    }
    if (a.partials) {
        ThreadStats ts;
        if (live) ts.add(1.0 - f, hist_s);
        __syncthreads();
        block_stats_flush(ts, hist_s, red_s, a.partials, a.hist);
    }
}
// The density-matrix counterpart: the exact solution of the ODEs dynamics_lindbladnodis_deqjl / lindbladt1_deqjl / lindbladt2_deqjl
// (spin.jl:488-518, 542-554) that run_sim_deqjl / run_sim_fine_deqjl (spin.jl:1277-1478) hand to Vern9, over a segment timeline
// with save points.  Bloch picture; one Magnus step per (sub-)segment, exact for the time-independent channels and of fourth
// order in the segment length for the running dephasing rate (see BlochGen).  channels = 0: no dissipation.
__global__ void __launch_bounds__(MC_THREADS) lindblad_timeline_kernel(const SegmentArgs a, const int64_t* __restrict__ save_after, int64_t nsave) {
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[MC_THREADS / 32];
    __shared__ double invk_s[64];
    __shared__ unsigned char t1cells_s[244];
    for (int b = threadIdx.x; b < RBQ_HIST_BINS; b += MC_THREADS) hist_s[b] = 0;
    if (threadIdx.x < 64) invk_s[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    t1_cells_init(t1cells_s, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    const bool live = s < a.S;
    const int64_t sc = live ? s : a.S - 1;
    const double da = a.da ? a.da[sc] : 0.0;
    const bool t1_on = (a.channels & RBQ_LINDBLAD_T1) != 0, t2_on = (a.channels & RBQ_LINDBLAD_T2) != 0;
    const double gf2 = t2_on ? a.t2_gamma_f * a.t2_gamma_f : 0.0;
    const double* q = a.rho0 + 8 * sc;
    const double tr0 = q[0] + q[6];
    double r0[3] = {q[2] + q[4], q[3] - q[5], q[0] - q[6]}, tg[4][3];
    double rr[1][3] = {{r0[0], r0[1], r0[2]}};
#pragma unroll
    for (int i = 0; i < 3; ++i) tg[0][i] = r0[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) mv3_to(a.rot.r[k], r0, tg[k + 1]);
    double* frow = (a.fid && live) ? a.fid + s * (nsave + 1) : nullptr;
    double* drow = (a.rho_saves && live) ? a.rho_saves + s * (nsave + 1) * 8 : nullptr;
    auto record = [&](int64_t g) {
        const int k = (int)(g & 3);
        const double f = fidelity_bloch(tr0, rr[0], tr0, k == 0 ? tg[0] : (k == 1 ? tg[1] : (k == 2 ? tg[2] : tg[3])));
        if (frow) frow[g] = f;
        if (drow) {
            double* o = drow + 8 * g;
            o[0] = 0.5 * (tr0 + rr[0][2]); o[1] = 0.0;
            o[2] = 0.5 * rr[0][0];         o[3] = 0.5 * rr[0][1];
            o[4] = 0.5 * rr[0][0];         o[5] = -0.5 * rr[0][1];
            o[6] = 0.5 * (tr0 - rr[0][2]); o[7] = 0.0;
        }
        return f;
    };
    double f = record(0);
    double t0 = 0.0;
    int64_t j = 0;
    for (int64_t g = 1; g <= nsave; ++g) {
        const int64_t end = save_after[g - 1];
        for (; j < end; ++j) {
            const double dt = a.durs[j], c1 = a.amps[j] + da;
            const double gmax = (t1_on ? 2.0 / 269.03e3 : 0.0) + (t2_on ? a.t2_gamma_c + 2.0 * gf2 * (t0 + dt) : 0.0);
            const double bnd = 2.0 * RBQ_PI * dt * (fabs(a.fq_scalar) + fabs(c1)) + dt * gmax;
            const int sub = bnd == bnd && bnd > 1.0 ? (int)fmin(ceil(bnd), 4096.0) : 1;
            const double h = dt / (double)sub, w2h = 2.0 * RBQ_PI * h;
            const int kmax = bnd == bnd ? taylor_terms(bnd / (double)sub) : 60;
            const double wx = w2h * c1, g1 = t1_on ? h / t1_ns_fast(c1, t1cells_s) : 0.0;
            const double e = gf2 * h * h * wx * (1.0 / 6.0);          // Magnus commutator term of the running rate
            for (int r = 0; r < sub; ++r) {
                BlochGen m;
                m.wz = w2h * a.fq_scalar; m.wxa = wx - e; m.wxb = wx + e;
                const double gphi = t2_on ? h * (a.t2_gamma_c + 2.0 * gf2 * (t0 + ((double)r + 0.5) * h)) : 0.0;
                m.g2 = g1 + gphi; m.g1 = 2.0 * g1;
                bloch_advance<1>(m, kmax, invk_s, rr);
            }
            t0 += dt;
        }
        f = record(g);
    }
    if (a.partials) {
        ThreadStats ts;
        if (live) ts.add(1.0 - f, hist_s);
        __syncthreads();
        block_stats_flush(ts, hist_s, red_s, a.partials, a.hist);
    }
}
__global__ void __launch_bounds__(MC_THREADS) lindblad_segments_kernel(const SegmentArgs a) {
    __shared__ unsigned hist_s[RBQ_HIST_BINS];
    __shared__ BlockPartial red_s[MC_THREADS / 32];
    __shared__ double invk_s[64];
    __shared__ unsigned char t1cells_s[244];
    for (int b = threadIdx.x; b < RBQ_HIST_BINS; b += MC_THREADS) hist_s[b] = 0;
    if (threadIdx.x < 64) invk_s[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    t1_cells_init(t1cells_s, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    const bool live = s < a.S;
    const int64_t sc = live ? s : a.S - 1;
    const double da = a.da ? a.da[sc] : 0.0;
    const bool t1_on = (a.channels & RBQ_LINDBLAD_T1) != 0, t2_on = (a.channels & RBQ_LINDBLAD_T2) != 0;
    double P[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int64_t j = 0; j < a.nseg; ++j) {
        const double dt = a.durs[j], c1 = a.amps[j] + da;
        // long segments: split into equal sub-steps so that the Taylor series stays short and well conditioned
        const double bnd = 2.0 * RBQ_PI * dt * (fabs(a.fq_scalar) + fabs(c1)) + dt * (2.0 / 269.03e3 + a.t2_gamma_c);
        const int sub = bnd == bnd && bnd > 1.0 ? (int)fmin(ceil(bnd), 4096.0) : 1;
        const double h = dt / (double)sub;
        BlochGen m;
        m.wz = 2.0 * RBQ_PI * h * a.fq_scalar; m.wxa = m.wxb = 2.0 * RBQ_PI * h * c1;
        const double g1 = t1_on ? h / t1_ns_fast(c1, t1cells_s) : 0.0;
        m.g2 = g1 + (t2_on ? h * a.t2_gamma_c : 0.0); m.g1 = 2.0 * g1;
        const int kmax = taylor_terms(bnd / (double)sub);
        for (int r = 0; r < sub; ++r) bloch_advance<3>(m, kmax, invk_s, P);
    }
    const double* q = a.rho0 + 8 * sc;
    const double tr0 = q[0] + q[6];
    double r0[3] = {q[2] + q[4], q[3] - q[5], q[0] - q[6]}, r[3], tg[4][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { tg[0][i] = r0[i]; r[i] = r0[i]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) mv3_to(a.rot.r[k], r0, tg[k + 1]);
    double* frow = (a.fid && live) ? a.fid + s * (a.gate_count + 1) : nullptr;
    double f = fidelity_bloch(tr0, r0, tr0, tg[0]);
    if (frow) frow[0] = f;
    for (int64_t g = 1; g <= a.gate_count; ++g) {
        const double n0 = fma(P[2][0], r[2], fma(P[1][0], r[1], P[0][0] * r[0]));
        const double n1 = fma(P[2][1], r[2], fma(P[1][1], r[1], P[0][1] * r[0]));
        const double n2 = fma(P[2][2], r[2], fma(P[1][2], r[1], P[0][2] * r[0]));
        r[0] = n0; r[1] = n1; r[2] = n2;
        const int k = (int)(g & 3);
        f = fidelity_bloch(tr0, r, tr0, k == 0 ? tg[0] : (k == 1 ? tg[1] : (k == 2 ? tg[2] : tg[3])));
        if (frow) frow[g] = f;
    }
    if (a.rho_out && live) {
        double* o = a.rho_out + 8 * s;
        o[0] = 0.5 * (tr0 + r[2]); o[1] = 0.0;
        o[2] = 0.5 * r[0];         o[3] = 0.5 * r[1];
        o[4] = 0.5 * r[0];         o[5] = -0.5 * r[1];
        o[6] = 0.5 * (tr0 - r[2]); o[7] = 0.0;
    }
    if (a.partials) {
        ThreadStats ts;
        if (live) ts.add(1.0 - f, hist_s);
        __syncthreads();
        block_stats_flush(ts, hist_s, red_s, a.partials, a.hist);
    }
}

// ---- sample materialisation (parity aids) -------------------------------------------------------------------------
__global__ void philox_samples_kernel(int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma,
                                      double fq_rel_clip, double da_sigma, double* __restrict__ fq, double* __restrict__ da,
                                      double* __restrict__ psi0) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double f, d, psi[4];
    sample_static((uint64_t)(first + s), seed, fq0, fq_rel_sigma, fq_rel_clip, da_sigma, f, d);
    sample_psi0((uint64_t)(first + s), seed, psi);
    if (fq) fq[s] = f;
    if (da) da[s] = d;
    if (psi0) { psi0[4 * s] = psi[0]; psi0[4 * s + 1] = psi[1]; psi0[4 * s + 2] = psi[2]; psi0[4 * s + 3] = psi[3]; }
}
// Pink-noise rows in two passes.  The normal deviates are independent per (sample, index pair): one thread each, every SM busy
// whatever S is, coalesced stores.  The three-pole filter is sequential in time but cheap: one warp per 32 rows, 32 x 32 tiles
// through shared memory (coalesced loads and stores of 256 B per row, lane r filters row r), in place.  (One thread per sample
// walking its whole row -- Philox, log, sincos and the filter in one loop -- left the GPU idle below a few thousand samples and
// wrote 8 bytes per 32-byte sector: 9.2 ms for 65 536 x 11 361.)
__global__ void __launch_bounds__(256) pink_white_kernel(int64_t first, int64_t S, uint64_t seed, int64_t noise_len, double* __restrict__ noise) {
    const int64_t npairs = (noise_len + 1) / 2, total = S * npairs;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int64_t s = t / npairs, j = t - s * npairs;
        double z0, z1;
        philox_normal2((uint64_t)(first + s), seed, (uint32_t)j, 2u, z0, z1);
        double* row = noise + s * noise_len;
        row[2 * j] = z0;
        if (2 * j + 1 < noise_len) row[2 * j + 1] = z1;
    }
}
__global__ void __launch_bounds__(32) pink_filter_kernel(int64_t S, double namp, double noise_dt_inv, int64_t noise_len, double* __restrict__ noise) {
    __shared__ double tile[32][33];
    const int lane = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    const int rows = (int)(S - row0 < 32 ? S - row0 : 32);
    PinkGen pg;
    pg.index = 0; pg.seed = 0; pg.scale_dt_inv = noise_dt_inv; pg.namp = namp;
    double* base = noise + row0 * noise_len + lane;
    double nxt[32];                                  // the next tile's column of this lane, in flight while the current tile is filtered
#pragma unroll
    for (int r = 0; r < 32; ++r) nxt[r] = (r < rows && lane < noise_len) ? base[(int64_t)r * noise_len] : 0.0;
    for (int64_t c0 = 0; c0 < noise_len; c0 += 32) {
        const int cols = (int)(noise_len - c0 < 32 ? noise_len - c0 : 32);
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[r][lane] = nxt[r];
        __syncwarp();
        if (c0 + 32 < noise_len) {
#pragma unroll
            for (int r = 0; r < 32; ++r) nxt[r] = (r < rows && c0 + 32 + lane < noise_len) ? base[(int64_t)r * noise_len + c0 + 32] : 0.0;
        }
        if (lane < rows)
            for (int c = 0; c < cols; ++c) tile[lane][c] = pg.filter(tile[lane][c]);
        __syncwarp();
        if (lane < cols) {
#pragma unroll
            for (int r = 0; r < 32; ++r)
                if (r < rows) base[(int64_t)r * noise_len + c0] = tile[r][lane];
        }
        __syncwarp();
    }
}

// ---- DFMA saturation loop: the FP64 roofline denominator --------------------------------------------------------------
constexpr int PEAK_CHAINS = 16;
// MODE 0 is the roofline denominator (register-operand DFMA).  The other modes are microbenchmarks of the
// instruction forms the sweep kernels use (selected with RBQ_PEAK_MODE; tools/fp64_micro.py).
template <int MODE>
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double* __restrict__ sink) {
    double acc[PEAK_CHAINS];
    const double m = 1.0 - 1e-9 * (double)(threadIdx.x & 7), c = 1e-9 * (double)(blockIdx.x & 3);
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) acc[i] = (double)i + 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i) {
                if (MODE == 0) acc[i] = fma(acc[i], m, c);
                else if (MODE == 1) acc[i] = acc[i] * m;                                   // DMUL
                else if (MODE == 2) acc[i] = acc[i] + c;                                   // DADD
                else if (MODE == 3) acc[i] = fma(acc[i], m, 0.0013888888888888889);        // immediate/uniform addend
                else if (MODE == 4) acc[i] = (i & 3) == 0 ? acc[i] * m : fma(acc[i], m, c);  // 1 DMUL : 3 DFMA
                else if (MODE == 5) acc[i] = fma(acc[i], acc[(i + 1) % PEAK_CHAINS], acc[(i + 5) % PEAK_CHAINS]);  // 3 distinct regs
                else if (MODE == 6) acc[i] = fma(-acc[i], acc[(i + 1) % PEAK_CHAINS], 0.0013888888888888889);
                else if (MODE == 7) acc[i] = fma(m, acc[(i + 1) % PEAK_CHAINS], acc[i]);     // one operand reusable across instructions
                else acc[i] = fma(acc[(i + 1) % PEAK_CHAINS], acc[(i + 2) % PEAK_CHAINS], acc[i]);
            }
        if (MODE >= 5) {
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i) acc[i] = acc[i] * 1e-300 + 0.5;
        }
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) t += acc[i];
    if (t == 123456.789) sink[0] = t;   // never true; keeps the loop alive
}

// dependent-issue latency probe: one warp per SM, one serial DFMA chain (RBQ_PEAK_MODE=9)
__global__ void __launch_bounds__(32) fp64_latency_kernel(int iters, double* __restrict__ sink) {
    const double m = 1.0 - 1e-9 * (double)(threadIdx.x & 7), c = 1e-9 * (double)(blockIdx.x & 3);
    double acc = 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 128; ++r) acc = fma(acc, m, c);
    }
    if (acc == 123456.789) sink[0] = acc;
}

// ---- launchers -----------------------------------------------------------------------------------------------------------
// (samples per thread, min resident blocks per SM, block size) of the closed-form kernel;
// RBQ_MC_VARIANT=<spt><minb><threads/128> overrides the default for tuning runs
static inline void static_variant(const rbq_handle* h, int64_t S, int algo, int* spt, int* minb, int* threads, int* form = nullptr,
                                  int* knots = nullptr) {
    const int sm_count = h->sm_count;
    *spt = 1; *minb = 1; *threads = 256;
    if (form) *form = 0;
    if (knots) *knots = 8;
    if (algo != RBQ_ALGO_SU2) return;
    // measured best of the compiled variants (tools/tune_mc.py): 1 sample per thread, 128-thread blocks, 6 per SM (80 registers: at
    // 72 the scheduler interleaves the updates of neighbouring knots and loses the operand reuse, 1.169 against 1.070 ms).
    // Digits of the variant: <knots per batch><table form><samples per thread><blocks per SM><threads / 128>; table form 0 = rows
    // through shared memory (mc_static_kernel), 1 / 2 / 3 = rows in constant memory (mc_static_ctab_kernel): Horner with vector
    // loads / factored quadratic / Horner with c1 through shared memory (the default: 14 instructions, the bits of form 0)
    int v = RBQ_MC_DEFAULT_VARIANT;
    if (h->opt.mc_variant > 0) v = h->opt.mc_variant;
    *spt = (v / 100) % 10; *minb = (v / 10) % 10; *threads = 128 * (v % 10);
    int f = (v / 1000) % 10, u = (v / 10000) % 10;
    const bool ok = (*spt == 1 || *spt == 2) && ((*threads == 256 && *minb >= 1 && *minb <= 4) || (*threads == 128 && (*minb == 4 || *minb == 6 || *minb == 7 || *minb == 8)));
    if (!ok) { *spt = 1; *minb = 6; *threads = 128; f = 0; }
    // small sweeps: one sample per thread so that every SM gets blocks
    if (S < (int64_t)sm_count * 4 * *threads * *spt) *spt = 1;
    if (f < 0 || f > 3 || *spt != 1 || *threads != 128 || (*minb != 6 && *minb != 8)) f = 0;
    if (form) *form = f;
    if (knots) *knots = u == 4 ? 4 : 8;
}
int mc_static_blocks(const rbq_handle* h, int64_t S, int algo) {
    int spt, minb, threads;
    static_variant(h, S, algo, &spt, &minb, &threads);
    const int per = threads * spt;
    return (int)((S + per - 1) / per);
}
// samples one full wave of resident blocks of the closed-form kernel processes (chunk sizes of the streamed host
// sweep are multiples of this, so that chunking adds no partially filled waves)
int64_t mc_static_wave_samples(const rbq_handle* h, int algo) {
    int spt, minb, threads;
    static_variant(h, (int64_t)1 << 40, algo, &spt, &minb, &threads);
    if (algo != RBQ_ALGO_SU2) { spt = 1; minb = 1; threads = 256; }
    return (int64_t)h->sm_count * minb * threads * spt;
}
int mc_noise_blocks(const rbq_handle*, int64_t S) { return (int)((S + MC_THREADS - 1) / MC_THREADS); }
// upper bound of the grid of either Lindblad sweep kernel (sizes the per-block partials)
int lindblad_blocks(const rbq_handle* h, int64_t S) { return (int)((S + 31) / 32) + h->sm_count; }

int launch_pulse_prep(rbq_handle* h, const double* d_controls, int64_t nk, double* d_amax) {
    // staged pulse lives in dbuf[0] (sized by the caller), padded to a multiple of 2 knots
    const int64_t padded = (nk + 1) & ~(int64_t)1;
    pulse_prep_kernel<<<1, 1024, 0, h->stream>>>(d_controls, nk, padded, (double*)h->dbuf[0], d_amax);
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_static_table(rbq_handle* h, const double* d_controls, int64_t nk, double dt, double f0, double a0, double* d_table,
                        double* d_scal, double* d_table_m) {
    static_table_kernel<<<1, 1024, 0, h->stream>>>(d_controls, nk, dt, f0, a0, d_table, d_scal, d_table_m);
    h->launches++;
    return (int)cudaGetLastError();
}
// g_ctab is one per device: the upload of a handle's table, the launch that reads it and the event that marks the read are
// issued under one lock, and an upload waits for the last reader of the previous owner's table
constexpr int CT_MAX_DEV = 64;
static std::mutex g_ctab_mu;
static const rbq_handle* g_ctab_owner[CT_MAX_DEV];
static uint64_t g_ctab_gen[CT_MAX_DEV];
static int g_ctab_form[CT_MAX_DEV];
static int64_t g_ctab_first[CT_MAX_DEV];      // first knot of the window that is loaded
static cudaEvent_t g_ctab_ev[CT_MAX_DEV];
void ctab_forget(const rbq_handle* h) {
    std::lock_guard<std::mutex> lk(g_ctab_mu);
    for (int d = 0; d < CT_MAX_DEV; ++d) if (g_ctab_owner[d] == h) g_ctab_owner[d] = nullptr;
}
typedef void (*ctab_sweep_fn)(const McStaticArgs);
// The pair of kernels of the constant-memory form (device-resident samples and fidelities; calls whose buffers are page-locked
// host memory keep the one-kernel form, whose blocks hide their PCIe reads and writes behind the arithmetic of their neighbours:
// cutting the pair into 4 to 16 slices with the finish kernels on a second, highest-priority stream measured 1.21 to 1.25 ms against
// 1.12 ms end to end for 2^20 samples with seeds in, 1.6 to 1.8 against 1.17 ms with sample arrays)
static int launch_mc_static_ctab(rbq_handle* h, const McStaticArgs& a0, int blocks, ctab_sweep_fn sweep, int form) {
    const int d = h->device;
    std::lock_guard<std::mutex> lk(g_ctab_mu);
    McStaticArgs a = a0;
    // one launch per window of knots (the whole pulse when it fits); windows are multiples of ST_TILE, the norm-restoring rhythm
    const int64_t win = a.nk <= CT_MAX ? a.nk : CT_WIN;
    cudaError_t e = cudaSuccess;
    for (int64_t k0 = 0; k0 < a.nk && e == cudaSuccess; k0 += win) {
        a.win_first = k0; a.win_knots = (int)(a.nk - k0 < win ? a.nk - k0 : win);
        if (g_ctab_owner[d] != h || g_ctab_gen[d] != a.table_gen || g_ctab_form[d] != form || g_ctab_first[d] != k0) {
            if (!g_ctab_ev[d]) e = cudaEventCreateWithFlags(&g_ctab_ev[d], cudaEventDisableTiming);
            else e = cudaStreamWaitEvent(h->stream, g_ctab_ev[d], 0);
            if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(g_ctab, (form == 2 ? a.table_m : a.table) + 4 * k0, sizeof(double) * 4 * (size_t)a.win_knots,
                                                              0, cudaMemcpyDeviceToDevice, h->stream);
            if (e != cudaSuccess) { g_ctab_owner[d] = nullptr; return (int)e; }
            g_ctab_owner[d] = h; g_ctab_gen[d] = a.table_gen; g_ctab_form[d] = form; g_ctab_first[d] = k0;
        }
        sweep<<<blocks, 128, 0, h->stream>>>(a);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(g_ctab_ev[d], h->stream);      // the table's last reader
        h->launches++;
    }
    if (e != cudaSuccess) return (int)e;
    mc_static_finish_kernel<128><<<blocks, 128, 0, h->stream>>>(a);
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_mc_static(rbq_handle* h, const McStaticArgs& a, int algo, int* nblocks_out) {
    int spt, minb, threads, form, knots;
    static_variant(h, a.S, algo, &spt, &minb, &threads, &form, &knots);
    const int blocks = mc_static_blocks(h, a.S, algo);
    if (nblocks_out) *nblocks_out = blocks;
#define RBQ_GO(SPT_, MINB_, T_) mc_static_kernel<RBQ_ALGO_SU2, SPT_, MINB_, T_><<<blocks, T_, 0, h->stream>>>(a)
    if (algo == RBQ_ALGO_SU2 && form > 0 && !a.host_io && a.table_m && a.quat && h->device >= 0 && h->device < CT_MAX_DEV) {
        ctab_sweep_fn fn;
        switch (knots * 100 + form * 10 + minb) {
        case 836: fn = mc_static_ctab_kernel<6, 128, 8, 2>; break;
        case 838: fn = mc_static_ctab_kernel<8, 128, 8, 2>; break;
        case 436: fn = mc_static_ctab_kernel<6, 128, 4, 2>; break;
        case 816: fn = mc_static_ctab_kernel<6, 128, 8, 0>; break;
        case 826: fn = mc_static_ctab_kernel<6, 128, 8, 1>; break;
        case 818: fn = mc_static_ctab_kernel<8, 128, 8, 0>; break;
        case 828: fn = mc_static_ctab_kernel<8, 128, 8, 1>; break;
        case 416: fn = mc_static_ctab_kernel<6, 128, 4, 0>; break;
        case 426: fn = mc_static_ctab_kernel<6, 128, 4, 1>; break;
        case 418: fn = mc_static_ctab_kernel<8, 128, 4, 0>; break;
        default: fn = mc_static_ctab_kernel<8, 128, 4, 1>; break;
        }
        return launch_mc_static_ctab(h, a, blocks, fn, form);
    } else if (algo == RBQ_ALGO_SU2) {
        switch (spt * 100 + minb * 10 + threads / 128) {
        case 112: RBQ_GO(1, 1, 256); break; case 122: RBQ_GO(1, 2, 256); break; case 132: RBQ_GO(1, 3, 256); break; case 142: RBQ_GO(1, 4, 256); break;
        case 212: RBQ_GO(2, 1, 256); break; case 222: RBQ_GO(2, 2, 256); break; case 232: RBQ_GO(2, 3, 256); break; case 242: RBQ_GO(2, 4, 256); break;
        case 141: RBQ_GO(1, 4, 128); break; case 161: RBQ_GO(1, 6, 128); break; case 171: RBQ_GO(1, 7, 128); break; case 181: RBQ_GO(1, 8, 128); break;
        case 241: RBQ_GO(2, 4, 128); break; case 261: RBQ_GO(2, 6, 128); break; case 281: RBQ_GO(2, 8, 128); break;
        default: RBQ_GO(1, 6, 128); break;
        }
    } else if (algo == RBQ_ALGO_PADE) {
        mc_static_kernel<RBQ_ALGO_PADE, 1, 1, 256><<<blocks, 256, 0, h->stream>>>(a);
    } else return RBQ_E_ALGO;
#undef RBQ_GO
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_mc_noise(rbq_handle* h, const McNoiseArgs& a, int algo, int* nblocks_out) {
    const int blocks = mc_noise_blocks(h, a.S);
    if (nblocks_out) *nblocks_out = blocks;
    if (algo == RBQ_ALGO_SU2) mc_noise_kernel<RBQ_ALGO_SU2><<<blocks, MC_THREADS, 0, h->stream>>>(a);
    else if (algo == RBQ_ALGO_PADE) mc_noise_kernel<RBQ_ALGO_PADE><<<blocks, MC_THREADS, 0, h->stream>>>(a);
    else return RBQ_E_ALGO;
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_mc_noise_gate_parallel(rbq_handle* h, const McNoiseArgs& a, double* d_quat, int* nblocks_out) {
    const int blocks = mc_noise_blocks(h, a.S);
    if (nblocks_out) *nblocks_out = blocks;
    if (a.gate_count > 65535) return RBQ_E_SIZE;
    mc_noise_gate_kernel<<<dim3((unsigned)blocks, (unsigned)a.gate_count), MC_THREADS, 0, h->stream>>>(a, d_quat);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    mc_noise_combine_kernel<<<blocks, MC_THREADS, 0, h->stream>>>(a, d_quat);
    h->launches += 2;
    return (int)cudaGetLastError();
}
int launch_segments(rbq_handle* h, const SegmentArgs& a, bool lindblad, int* nblocks_out) {
    const int blocks = (int)((a.S + MC_THREADS - 1) / MC_THREADS);
    if (nblocks_out) *nblocks_out = blocks;
    if (lindblad) lindblad_segments_kernel<<<blocks, MC_THREADS, 0, h->stream>>>(a);
    else mc_static_segments_kernel<<<blocks, MC_THREADS, 0, h->stream>>>(a);
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_timeline(rbq_handle* h, const SegmentArgs& a, const int64_t* d_save_after, int64_t nsave, int* nblocks_out) {
    const int blocks = (int)((a.S + MC_THREADS - 1) / MC_THREADS);
    if (nblocks_out) *nblocks_out = blocks;
    if (a.rho0) lindblad_timeline_kernel<<<blocks, MC_THREADS, 0, h->stream>>>(a, d_save_after, nsave);
    else mc_timeline_kernel<<<blocks, MC_THREADS, 0, h->stream>>>(a, d_save_after, nsave);
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_lindblad_shared_map(rbq_handle* h, const LindbladArgs& a, int channels, double* d_map) {
    lindblad_shared_map_kernel<<<1, 256, 0, h->stream>>>(a, channels, d_map);
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_lindblad(rbq_handle* h, const LindbladArgs& a, int channels, int* nblocks_out) {
    const int blocks = (int)((a.S + MC_THREADS - 1) / MC_THREADS);
    if (nblocks_out) *nblocks_out = blocks;
    lindblad_kernel<<<blocks, MC_THREADS, 0, h->stream>>>(a, channels);
    h->launches++;
    return (int)cudaGetLastError();
}
// largest offset the Frechet table serves: the dropped term (w2 d)^4 / 24 of the rotation stays below 5e-18
double lindblad_table_range(double dt) { return 1.05e-4 / (2.0 * RBQ_PI * fabs(dt)); }
int launch_lindblad_table(rbq_handle* h, const LindbladArgs& a, int channels, const double* d_controls, double* d_table,
                          int* nblocks_out) {
    const double da_tab = lindblad_table_range(a.dt);
    lindblad_table_kernel<<<(unsigned)((a.nk + 127) / 128), 128, 0, h->stream>>>(d_controls, a.nk, a.dt, a.fq, channels, a.t2_gamma_c, da_tab, d_table);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    // Whole waves: the kernel is FP64-pipe bound, so an SM with one warp more than its neighbours finishes later by that warp's
    // whole run.  Spread the warps evenly: `waves` blocks per SM, every block the same number of warps (<= 16: 128 registers).
    const bool compose = a.gate_count > 3;
    const int64_t warps1 = (a.S + 31) / 32;
    // one realisation per thread and up to 24 warps per SM (85 registers), or two per thread sharing a row's loads and up to 16
    // warps (102 registers): RBQ_LINDBLAD_SPT picks, default by measurement (profiles/README.md)
    int spt = (!compose && warps1 >= (int64_t)h->sm_count * 8) ? 2 : 1;
    if (!compose && h->opt.lindblad_spt > 0) spt = h->opt.lindblad_spt == 2 ? 2 : 1;
    const int maxw = (spt == 1 && !compose) ? 24 : 16;
    const int64_t warps = (a.S + 32 * spt - 1) / (32 * spt);
    const int64_t per_sm = (warps + h->sm_count - 1) / h->sm_count;
    const int64_t waves = (per_sm + maxw - 1) / maxw;
    int wpb = (int)((warps + h->sm_count * waves - 1) / (h->sm_count * waves));
    if (wpb < 1) wpb = 1;
    const int threads = 32 * wpb;
    const int blocks = (int)((a.S + (int64_t)threads * spt - 1) / ((int64_t)threads * spt));
    if (nblocks_out) *nblocks_out = blocks;
    if (compose) lindblad_tab_kernel<1, 3><<<blocks, threads, 0, h->stream>>>(a, channels, d_table, da_tab);
    else if (spt == 2) lindblad_tab_kernel<2, 1><<<blocks, threads, 0, h->stream>>>(a, channels, d_table, da_tab);
    else lindblad_tab_kernel<1, 1, 768><<<blocks, threads, 0, h->stream>>>(a, channels, d_table, da_tab);
    h->launches += 2;
    return (int)cudaGetLastError();
}
int launch_stats_finalize(rbq_handle* h, const BlockPartial* partials, int nblocks, rbq_stats* d_stats) {
    stats_finalize_kernel<<<1, FIN_THREADS, 0, h->stream>>>(partials, nblocks, d_stats);
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_stats_reset(rbq_handle* h, rbq_stats* d_stats) {
    stats_reset_kernel<<<1, 256, 0, h->stream>>>(d_stats);
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_philox_samples(rbq_handle* h, int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma,
                          double fq_rel_clip, double da_sigma, double* d_fq, double* d_da, double* d_psi0) {
    philox_samples_kernel<<<(unsigned)((S + 255) / 256), 256, 0, h->stream>>>(first, S, seed, fq0, fq_rel_sigma, fq_rel_clip,
                                                                            da_sigma, d_fq, d_da, d_psi0);
    h->launches++;
    return (int)cudaGetLastError();
}
int launch_pink_noise(rbq_handle* h, int64_t first, int64_t S, uint64_t seed, double namp, double noise_dt_inv,
                      int64_t noise_len, double* d_noise) {
    const int64_t pairs = S * ((noise_len + 1) / 2);
    const int64_t want = (pairs + 255) / 256, cap = (int64_t)h->sm_count * 64;
    pink_white_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, h->stream>>>(first, S, seed, noise_len, d_noise);
    pink_filter_kernel<<<(unsigned)((S + 31) / 32), 32, 0, h->stream>>>(S, namp, noise_dt_inv, noise_len, d_noise);
    h->launches += 2;
    return (int)cudaGetLastError();
}
int launch_fp64_peak(rbq_handle* h, int iters, double* d_sink, int* nthreads_out, int* flops_per_thread_iter) {
    const int blocks = h->sm_count * 8;
    const int mode = h->opt.peak_mode;
    if (mode == 9) {
        fp64_latency_kernel<<<h->sm_count, 32, 0, h->stream>>>(iters, d_sink);
        h->launches++;
        if (nthreads_out) *nthreads_out = h->sm_count * 32;
        if (flops_per_thread_iter) *flops_per_thread_iter = 128 * 2;
        return (int)cudaGetLastError();
    }
    switch (mode) {
    case 1: fp64_peak_kernel<1><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    case 2: fp64_peak_kernel<2><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    case 3: fp64_peak_kernel<3><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    case 4: fp64_peak_kernel<4><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    case 5: fp64_peak_kernel<5><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    case 6: fp64_peak_kernel<6><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    case 7: fp64_peak_kernel<7><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    case 8: fp64_peak_kernel<8><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    default: fp64_peak_kernel<0><<<blocks, 256, 0, h->stream>>>(iters, d_sink); break;
    }
    h->launches++;
    if (nthreads_out) *nthreads_out = blocks * 256;
    // (modes 5 and up rescale their chains once per iteration: 16 more FP64 instructions on top of the 128)
    if (flops_per_thread_iter) *flops_per_thread_iter = (8 * PEAK_CHAINS + (mode >= 5 && mode <= 8 ? PEAK_CHAINS : 0)) * 2;
    return (int)cudaGetLastError();
}

}  // namespace rbq
