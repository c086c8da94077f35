// rbq_scan_rollout.cuh — open-loop rollout of ONE trajectory as a parallel prefix product (TO.rollout! for the models whose
// iso blocks evolve linearly: spin13.jl:44-58, spin15.jl:50-73, spin12.jl:165-192, spin18.jl:162-190).
//
// In an open-loop rollout the control chain [int a, a, da] is driven by u alone (spin13.jl:49-51), so every knot's propagator
// E_k = exp(dt_k (f H0 + a_k H1)) is known before any state is: psi_k = (E_{k-1} ... E_0) psi_0.  The E_k are iso images of
// SU(2) elements, closed under multiplication and determined by their first column (a quaternion), so the whole trajectory is
//   1. the control chain as three block-wide prefix sums (da from u, a from da, the integral from a),
//   2. all propagators in parallel (one knot-segment per thread),
//   3. an inclusive scan of quaternion products over the block (warp shuffles, then the 32 warp totals),
//   4. psi_k = Q_k psi_0 for every block of the state, written knot by knot.
// The warp-per-trajectory kernel walks the knots one after the other (0.19 us per knot: 190 us for 1001 knots); this one
// takes a few microseconds beyond its launch.  Product order differs from the sequential chain by rounding only
// (~1e-16 per level of the tree); closed-loop rollouts cannot use it (the feedback K dx couples psi into a).
#pragma once
#include "rbq_warp_rollout.cuh"

namespace rbq {

struct Quat { double v[4]; };
// first column of M(b) M(a): a acts first, then b
RBQ_DEV Quat qmul(const Quat& b, const Quat& a) {
    Quat r;
    r.v[0] = fma(-b.v[3], a.v[3], fma(-b.v[2], a.v[2], fma(-b.v[1], a.v[1], b.v[0] * a.v[0])));
    r.v[1] = fma(b.v[2], a.v[3], fma(-b.v[3], a.v[2], fma(b.v[0], a.v[1], b.v[1] * a.v[0])));
    r.v[2] = fma(-b.v[1], a.v[3], fma(b.v[0], a.v[2], fma(b.v[3], a.v[1], b.v[2] * a.v[0])));
    r.v[3] = fma(b.v[0], a.v[3], fma(b.v[1], a.v[2], fma(-b.v[2], a.v[1], b.v[3] * a.v[0])));
    return r;
}
RBQ_DEV Quat quat_of(const Prop<double>& E) { return Quat{{E.c, 0.0, -E.sp, -E.sq}}; }
RBQ_DEV Quat quat_shfl_up(const Quat& q, int o) {
    Quat r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = __shfl_up_sync(FULL, q.v[i], o);
    return r;
}

constexpr int SCAN_T = 512;          // threads per trajectory (128 registers each: the 3-chain models spill at 1024 x 64)
constexpr int SCAN_MAX_PER = 16;     // knots per thread: trajectories of up to 8192 steps
__host__ __device__ inline size_t scan_rollout_smem(int64_t steps) { return sizeof(double) * (3 * (size_t)steps + 3 * 32 * 4 + 32 + 8); }

// MODEL: RBQ_MODEL_SPIN13 (also spin15, decay aware / time optimal included) or RBQ_MODEL_SPIN12 (also spin18)
template <int MODEL>
__global__ void __launch_bounds__(SCAN_T) scan_rollout_kernel(const WarpRolloutArgs a, int per) {
    constexpr int NC = MODEL == RBQ_MODEL_SPIN12 ? 3 : 1;       // distinct propagator chains
    constexpr int NBLK = MODEL == RBQ_MODEL_SPIN12 ? 10 : 2;    // iso blocks of the state
    extern __shared__ __align__(16) double sh[];
    const rbq_model_params& p = a.p;
    const int n = a.n, m = a.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t N = a.N, NS = N - 1, tr = blockIdx.x;
    double* u_s = sh;                 // [NS] u_1 (the second derivative of the amplitude)
    double* dt_s = sh + NS;           // [NS]
    double* a_s = sh + 2 * NS;        // [NS] amplitude at knot k
    double* wtot = sh + 3 * NS;       // [NC][32][4] warp totals
    double* gsum = wtot + 3 * 32 * 4; // [32] warp totals of the decay integral
    const double* __restrict__ x0 = a.x0 + tr * n;
    const double* __restrict__ Uin = a.U + tr * NS * m;
    double* __restrict__ X = a.X + tr * N * n;
    const bool DA = MODEL == RBQ_MODEL_SPIN13 && p.model_id == RBQ_MODEL_SPIN15 && p.decay_aware;
    const bool TO = MODEL == RBQ_MODEL_SPIN13 && p.model_id == RBQ_MODEL_SPIN15 && p.time_optimal;
    const int cb = 8;
    for (int64_t k = tid; k < NS; k += SCAN_T) {
        u_s[k] = Uin[k * m];
        const double u2 = TO ? Uin[k * m + 1] : 0.0;
        dt_s[k] = TO ? u2 * u2 : a.dt[k];
    }
    for (int i = tid; i < n; i += SCAN_T) X[i] = x0[i];
    __syncthreads();
    // 1. control chain [int a, a, da] (spin13.jl:49-51): three first-order recurrences v_{k+1} = fma(src_k, dt_k, v_k), each a
    //    block-wide prefix sum (a thread walks its own knots with the same multiply-adds the step kernels use; only the value
    //    it starts from is summed in tree order, so the chain differs from a sequential walk by ~1e-16 relative).  da first
    //    (driven by u), then a (driven by da), then the integral (driven by a).
    {
        const int64_t c0 = (int64_t)tid * per, c1 = c0 + per < NS ? c0 + per : NS;
        auto exclusive = [&](double v) {          // exclusive prefix sum of v over the block's threads
            double inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) gsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                double w = lane < SCAN_T / 32 ? gsum[lane] : 0.0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(FULL, w, o); if (lane >= o) w += t; }
                gsum[lane] = w;
            }
            __syncthreads();
            double ex = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) ex = 0.0;
            if (warp > 0) ex += gsum[warp - 1];
            __syncthreads();                      // gsum is reused by the next scan
            return ex;
        };
        // da: u_s[k] is replaced by da_{k+1}
        double tot = 0.0;
        for (int64_t k = c0; k < c1; ++k) tot = fma(u_s[k], dt_s[k], tot);
        double v = x0[cb + 2] + exclusive(tot);
        double da_first = v;                      // da at this thread's first knot
        for (int64_t k = c0; k < c1; ++k) { v = fma(u_s[k], dt_s[k], v); u_s[k] = v; X[(k + 1) * n + cb + 2] = v; }
        // a: a_s[k] = amplitude at knot k (k = 0..NS), driven by da_k = (k ? u_s[k-1] : da_0)
        tot = 0.0;
        {
            double dk = da_first;
            for (int64_t k = c0; k < c1; ++k) { tot = fma(dk, dt_s[k], tot); dk = u_s[k]; }
        }
        v = x0[cb + 1] + exclusive(tot);
        {
            double dk = da_first;
            for (int64_t k = c0; k < c1; ++k) { a_s[k] = v; v = fma(dk, dt_s[k], v); dk = u_s[k]; X[(k + 1) * n + cb + 1] = v; }
        }
        // integral of a
        tot = 0.0;
        for (int64_t k = c0; k < c1; ++k) tot = fma(a_s[k], dt_s[k], tot);
        v = x0[cb] + exclusive(tot);
        for (int64_t k = c0; k < c1; ++k) { v = fma(a_s[k], dt_s[k], v); X[(k + 1) * n + cb] = v; }
    }
    __syncthreads();
    // chain c: generator f_c H0 + (a + da_c) H1
    double fc[NC], dac[NC];
    fc[0] = p.fq; dac[0] = 0.0;
    if (NC == 3) {
        if (p.model_id == RBQ_MODEL_SPIN12) { fc[1] = p.fq_samples[0]; fc[2] = p.fq_samples[1]; dac[1] = dac[2] = 0.0; }
        else { fc[1] = fc[2] = p.fq; dac[1] = p.namp; dac[2] = -p.namp; }
    }
    // 2. this thread's knots [k0, k1): local products, latest knot on the left
    const int64_t k0 = (int64_t)tid * per, k1 = k0 + per < NS ? k0 + per : NS;
    Quat loc[NC];
    double gloc = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) loc[c] = Quat{{1.0, 0.0, 0.0, 0.0}};
    for (int64_t k = k0; k < k1; ++k) {
#pragma unroll
        for (int c = 0; c < NC; ++c) loc[c] = qmul(quat_of(prop_of<double, double>(fc[c], a_s[k] + dac[c], dt_s[k])), loc[c]);
        if (DA) gloc = fma(1.0 / t1_interp<double>(a_s[k], 1.0), dt_s[k], gloc);
    }
    // 3. inclusive scan over the block: within the warp by shuffles, then over the 32 warp totals
    Quat inc[NC];
    double ginc = gloc;
#pragma unroll
    for (int c = 0; c < NC; ++c) inc[c] = loc[c];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const Quat prev = quat_shfl_up(inc[c], o);
            if (lane >= o) inc[c] = qmul(inc[c], prev);
        }
        const double gp = __shfl_up_sync(FULL, ginc, o);
        if (lane >= o) ginc += gp;
    }
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) wtot[(c * 32 + warp) * 4 + i] = inc[c].v[i];
        gsum[warp] = ginc;
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = SCAN_T / 32;
        double gw = lane < NW ? gsum[lane] : 0.0;
        Quat w[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) w[c].v[i] = lane < NW ? wtot[(c * 32 + lane) * 4 + i] : (i == 0 ? 1.0 : 0.0);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const Quat prev = quat_shfl_up(w[c], o);
                if (lane >= o) w[c] = qmul(w[c], prev);
            }
            const double gp = __shfl_up_sync(FULL, gw, o);
            if (lane >= o) gw += gp;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) wtot[(c * 32 + lane) * 4 + i] = w[c].v[i];     // inclusive over warps
        gsum[lane] = gw;
    }
    __syncthreads();
    // exclusive prefix of this thread: (warps before) then (lanes before in this warp)
    Quat run[NC];
    double grun = DA ? x0[11] : 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        Quat ex = quat_shfl_up(inc[c], 1);
        if (lane == 0) ex = Quat{{1.0, 0.0, 0.0, 0.0}};
        if (warp > 0) {
            Quat wp;
#pragma unroll
            for (int i = 0; i < 4; ++i) wp.v[i] = wtot[(c * 32 + warp - 1) * 4 + i];
            ex = qmul(ex, wp);
        }
        run[c] = ex;
    }
    {
        double gex = __shfl_up_sync(FULL, ginc, 1);
        if (lane == 0) gex = 0.0;
        if (warp > 0) gex += gsum[warp - 1];
        grun += gex;
    }
    // 4. the states of this thread's knots
    double psi0[NBLK][4];
#pragma unroll
    for (int b = 0; b < NBLK; ++b)
#pragma unroll
        for (int i = 0; i < 4; ++i) psi0[b][i] = x0[(b < 2 ? 4 * b : 11 + 4 * (b - 2)) + i];
    for (int64_t k = k0; k < k1; ++k) {
        double* Xn = X + (k + 1) * n;
#pragma unroll
        for (int c = 0; c < NC; ++c) run[c] = qmul(quat_of(prop_of<double, double>(fc[c], a_s[k] + dac[c], dt_s[k])), run[c]);
#pragma unroll
        for (int b = 0; b < NBLK; ++b) {
            const int c = b < 2 ? 0 : (b < 6 ? 1 : 2);
            const Quat q = run[NC == 1 ? 0 : c];
            Quat x;
#pragma unroll
            for (int i = 0; i < 4; ++i) x.v[i] = psi0[b][i];
            const Quat y = qmul(q, x);
            const int off = b < 2 ? 4 * b : 11 + 4 * (b - 2);
#pragma unroll
            for (int i = 0; i < 4; ++i) Xn[off + i] = y.v[i];
        }
        if (DA) {
            grun = fma(1.0 / t1_interp<double>(a_s[k], 1.0), dt_s[k], grun);
            Xn[11] = grun;
        }
    }
}

}  // namespace rbq
