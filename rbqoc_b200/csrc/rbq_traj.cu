// rbq_traj.cu — trajectory-level kernels behind the RobotDynamics model API:
//   rollout           x[k+1] = discrete_dynamics(RK3, model, x[k], u[k], t, dt[k])   (TO.rollout!)
//   closed-loop       the iLQR line-search rollouts, one thread per step-size candidate
//   jacobians         [A_k B_k] for every knot in one launch; forward-mode tangents written by hand
//                     (Tan<NT>), replacing RD's default ForwardDiff.jacobian over [x;u]
// Reference step maps: see rbq_models.cuh.  sm_100a only.
#include <stdlib.h>

#include "rbq_internal.h"
#include "rbq_models.cuh"
#include "rbq_warp_rollout.cuh"
#include "rbq_scan_rollout.cuh"

namespace rbq {

// ---- IO policies --------------------------------------------------------------------------------
struct ValueIO {
    const double* __restrict__ xi;
    const double* __restrict__ ui;
    double* __restrict__ xo;
    RBQ_DEV double x(int i) const { return xi[i]; }
    RBQ_DEV double u(int j) const { return ui[j]; }
    RBQ_DEV void out(int i, double v) { xo[i] = v; }
};
// lane `c0/NT` of a knot carries the seed columns [c0, c0+NT) of [x;u]
template <int NT> struct TangentIO {
    const double* __restrict__ xi;
    const double* __restrict__ ui;
    double* __restrict__ ab;   // column-major n x (n+m)
    int n, ncol, c0;
    RBQ_DEV Tan<NT> seed(double v, int idx) const {
        Tan<NT> r; r.v = v;
#pragma unroll
        for (int j = 0; j < NT; ++j) r.d[j] = (idx == c0 + j) ? 1.0 : 0.0;
        return r;
    }
    RBQ_DEV Tan<NT> x(int i) const { return seed(xi[i], i); }
    RBQ_DEV Tan<NT> u(int j) const { return seed(ui[j], n + j); }
    RBQ_DEV void out(int i, const Tan<NT>& v) {
#pragma unroll
        for (int j = 0; j < NT; ++j)
            if (c0 + j < ncol) ab[(size_t)(c0 + j) * n + i] = v.d[j];
    }
};

// ---- kernels ------------------------------------------------------------------------------------
template <int MODEL, bool RK = false>
__global__ void __launch_bounds__(128) rollout_kernel(rbq_model_params p, int n, int m, int64_t B, int64_t N, double t0,
                                                      const double* __restrict__ x0, const double* __restrict__ U,
                                                      const double* __restrict__ dt, double* __restrict__ X) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double* Xb = X + (size_t)b * N * n;
    for (int i = 0; i < n; ++i) Xb[i] = x0[(size_t)b * n + i];
    const double* Ub = U + (size_t)b * (N - 1) * m;
    double time = t0;
    for (int64_t k = 0; k + 1 < N; ++k) {
        ValueIO io{Xb + k * n, Ub + k * m, Xb + (k + 1) * n};
        model_step<MODEL, double, ValueIO, RK>(p, io, time, dt[k]);
        time = time + dt[k];
    }
}

template <int MODEL, int NT, bool RK = false>
__global__ void __launch_bounds__(128) jacobians_kernel(rbq_model_params p, int n, int m, int lanes_per_knot,
                                                        int64_t K, const double* __restrict__ X,
                                                        const double* __restrict__ U, const double* __restrict__ tk,
                                                        const double* __restrict__ dt, double* __restrict__ AB) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = t / lanes_per_knot;
    if (k >= K) return;
    const int lane = (int)(t - k * lanes_per_knot);
    TangentIO<NT> io{X + k * n, U + k * m, AB + (size_t)k * n * (n + m), n, n + m, lane * NT};
    model_step<MODEL, Tan<NT>, TangentIO<NT>, RK>(p, io, tk ? tk[k] : 0.0, dt[k]);
}

// Same computation, but a block stages the [A_k B_k] blocks of G consecutive knots in shared memory and writes them out
// with fully coalesced stores (the direct kernel's lanes each scatter 8-byte elements n*NT*8 bytes apart).
template <int MODEL, int NT>
__global__ void __launch_bounds__(256) jacobians_staged_kernel(rbq_model_params p, int n, int m, int lanes_per_knot, int G,
                                                               int64_t K, const double* __restrict__ X,
                                                               const double* __restrict__ U, const double* __restrict__ tk,
                                                               const double* __restrict__ dt, double* __restrict__ AB) {
    extern __shared__ __align__(16) double stage[];
    const int blk = n * (n + m);
    const int64_t k0 = (int64_t)blockIdx.x * G;
    const int g = threadIdx.x / lanes_per_knot;
    const int lane = threadIdx.x - g * lanes_per_knot;
    const int64_t k = k0 + g;
    if (g < G && k < K) {
        TangentIO<NT> io{X + k * n, U + k * m, stage + (size_t)g * blk, n, n + m, lane * NT};
        model_step<MODEL, Tan<NT>>(p, io, tk ? tk[k] : 0.0, dt[k]);
    }
    __syncthreads();
    const int64_t here = (K - k0) < G ? (K - k0) : G;
    const int64_t total = here * blk;
    double* __restrict__ dst = AB + k0 * blk;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {   // 16-byte aligned destination (tested on the pointer: d_AB may be any view): 2 doubles per store
        const int64_t pairs = total >> 1;
        const double2* __restrict__ s2 = reinterpret_cast<const double2*>(stage);
        double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
        for (int64_t i = threadIdx.x; i < pairs; i += blockDim.x) d2[i] = s2[i];
        if ((total & 1) && threadIdx.x == 0) dst[total - 1] = stage[total - 1];
    } else {
        for (int64_t i = threadIdx.x; i < total; i += blockDim.x) dst[i] = stage[i];
    }
}

// gate_error_iso2 and its analytic gradient / Hessian (spin.jl:1040-1092), one thread per state
__global__ void __launch_bounds__(256) gate_error_kernel(int64_t K, const double* __restrict__ s1, int64_t st1,
                                                         const double* __restrict__ s2, int64_t st2, double* __restrict__ err,
                                                         double* __restrict__ grad, double* __restrict__ hess) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const double* x = s1 + k * st1;
    const double* y = s2 + k * st2;
    const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], y0 = y[0], y1 = y[1], y2 = y[2], y3 = y[3];
    const double a = fma(x3, y3, fma(x2, y2, fma(x1, y1, x0 * y0)));          // gate_error_iso2a
    const double b = fma(x1, y3, fma(x0, y2, fma(-x3, y1, -x2 * y0)));        // gate_error_iso2b
    if (err) err[k] = 1.0 - a * a - b * b;
    if (grad) {
        const double a2 = 2.0 * a, b2 = 2.0 * b;
        double* g = grad + 4 * k;
        g[0] = -a2 * y0 - b2 * y2;
        g[1] = -a2 * y1 - b2 * y3;
        g[2] = -a2 * y2 + b2 * y0;
        g[3] = -a2 * y3 + b2 * y1;
    }
    if (hess) {
        const double d11 = -2.0 * y0 * y0 - 2.0 * y2 * y2, d12 = -2.0 * y0 * y1 - 2.0 * y2 * y3;
        const double d14 = 2.0 * y1 * y2 - 2.0 * y0 * y3, d22 = -2.0 * y1 * y1 - 2.0 * y3 * y3;
        const double d23 = -2.0 * y1 * y2 + 2.0 * y0 * y3;
        double* H = hess + 16 * k;
        H[0] = d11; H[1] = d12; H[2] = 0.0; H[3] = d14;
        H[4] = d12; H[5] = d22; H[6] = d23; H[7] = 0.0;
        H[8] = 0.0; H[9] = d23; H[10] = d11; H[11] = d12;
        H[12] = d14; H[13] = 0.0; H[14] = d12; H[15] = d22;
    }
}
int launch_gate_error(rbq_handle* h, int64_t K, const double* d_s1, int64_t st1, const double* d_s2, int64_t st2, double* d_err,
                      double* d_grad, double* d_hess) {
    gate_error_kernel<<<(unsigned)((K + 255) / 256), 256, 0, h->stream>>>(K, d_s1, st1, d_s2, st2, d_err, d_grad, d_hess);
    h->launches++;
    return (int)cudaGetLastError();
}

// stage cost + gradient of the sampling objective (spin12.jl:102-140), one thread per knot
struct SamplingCostArgs { double Q[43], xf[43], R, tg[16], qss[4]; };
__global__ void __launch_bounds__(128) sampling_cost_kernel(int64_t K, const double* __restrict__ X, const double* __restrict__ U,
                                                            SamplingCostArgs a, double* __restrict__ J, double* __restrict__ gx,
                                                            double* __restrict__ gu) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const double* x = X + k * 43;
    double cost = 0.0;
    // 0.5 x'Qx + q'x + c with q = -Q xf, c = 0.5 xf'Q xf  ==  evaluated term by term exactly as the reference sums them
    double quad = 0.0, lin = 0.0, c0 = 0.0;
    for (int i = 0; i < 43; ++i) {
        quad = fma(x[i] * a.Q[i], x[i], quad);
        lin = fma(-a.Q[i] * a.xf[i], x[i], lin);
        c0 = fma(a.xf[i] * a.Q[i], a.xf[i], c0);
        if (gx) gx[k * 43 + i] = a.Q[i] * x[i] - a.Q[i] * a.xf[i];
    }
    cost = 0.5 * quad + lin + 0.5 * c0;
    for (int s = 0; s < 8; ++s) {
        const double* y = a.tg + 4 * (s & 3);
        const double* z = x + 11 + 4 * s;
        const double w = a.qss[s & 3];
        const double aa = fma(z[3], y[3], fma(z[2], y[2], fma(z[1], y[1], z[0] * y[0])));
        const double bb = fma(z[1], y[3], fma(z[0], y[2], fma(-z[3], y[1], -z[2] * y[0])));
        cost = fma(w, 1.0 - aa * aa - bb * bb, cost);
        if (gx) {
            const double a2 = 2.0 * aa, b2 = 2.0 * bb;
            double* g = gx + k * 43 + 11 + 4 * s;
            g[0] += w * (-a2 * y[0] - b2 * y[2]);
            g[1] += w * (-a2 * y[1] - b2 * y[3]);
            g[2] += w * (-a2 * y[2] + b2 * y[0]);
            g[3] += w * (-a2 * y[3] + b2 * y[1]);
        }
    }
    if (U) {
        cost = fma(0.5 * U[k] * a.R, U[k], cost);
        if (gu) gu[k] = a.R * U[k];
    }
    if (J) J[k] = cost;
}
int launch_sampling_cost(rbq_handle* h, int64_t K, const double* d_X, const double* d_U, const double* Qdiag, double R,
                         const double* xf, const double* targets, const double* q_ss, double* d_J, double* d_gx, double* d_gu) {
    SamplingCostArgs a;
    for (int i = 0; i < 43; ++i) { a.Q[i] = Qdiag[i]; a.xf[i] = xf[i]; }
    for (int i = 0; i < 16; ++i) a.tg[i] = targets[i];
    for (int i = 0; i < 4; ++i) a.qss[i] = q_ss[i];
    a.R = R;
    sampling_cost_kernel<<<(unsigned)((K + 127) / 128), 128, 0, h->stream>>>(K, d_X, d_U, a, d_J, d_gx, d_gu);
    h->launches++;
    return (int)cudaGetLastError();
}

// ---- launchers ------------------------------------------------------------------------------------
// launch of the warp-per-trajectory kernel; the unscented models with one sigma-point chunk get the SSC = 1 instantiation
template <int M, bool CLOSED>
static cudaError_t launch_warp_rollout(const WarpRolloutArgs& a, unsigned grid, size_t smem, cudaStream_t stream) {
    auto go = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, 32, smem, stream>>>(a);
        return cudaSuccess;
    };
    if constexpr (M == RBQ_MODEL_SPIN30 || M == RBQ_MODEL_SPIN25) {
        if (a.p.sample_state_count == 1) return go(warp_rollout_kernel<M, CLOSED, 1>);
    }
    return go(warp_rollout_kernel<M, CLOSED, 0>);
}

static inline int family(int model_id) {
    switch (model_id) {
    case RBQ_MODEL_SPIN13: return RBQ_MODEL_SPIN13;
    case RBQ_MODEL_SPIN12: case RBQ_MODEL_SPIN18: return RBQ_MODEL_SPIN12;
    case RBQ_MODEL_SPIN15: return RBQ_MODEL_SPIN15;
    case RBQ_MODEL_SPIN11: case RBQ_MODEL_SPIN17: return RBQ_MODEL_SPIN11;
    case RBQ_MODEL_SPIN30: return RBQ_MODEL_SPIN30;
    case RBQ_MODEL_SPIN25: return RBQ_MODEL_SPIN25;
    }
    return -1;
}
#define RBQ_DISPATCH_FAMILY(fam, CALL)                                   \
    switch (fam) {                                                       \
    case RBQ_MODEL_SPIN13: { CALL(RBQ_MODEL_SPIN13); break; }            \
    case RBQ_MODEL_SPIN12: { CALL(RBQ_MODEL_SPIN12); break; }            \
    case RBQ_MODEL_SPIN15: { CALL(RBQ_MODEL_SPIN15); break; }            \
    case RBQ_MODEL_SPIN11: { CALL(RBQ_MODEL_SPIN11); break; }            \
    case RBQ_MODEL_SPIN30: { CALL(RBQ_MODEL_SPIN30); break; }            \
    case RBQ_MODEL_SPIN25: { CALL(RBQ_MODEL_SPIN25); break; }            \
    default: return RBQ_E_MODEL;                                         \
    }

// one warp per trajectory up to this many trajectories; beyond it one thread per trajectory has the better throughput
static const int64_t WARP_ROLLOUT_MAX = 1 << 16;

int launch_rollout(rbq_handle* h, const rbq_model_params& p, int64_t B, int64_t N, double t0, const double* d_x0,
                   const double* d_U, const double* d_dt, double* d_X) {
    const int n = model_state_dim(p), m = model_control_dim(p);
    if (n < 0 || m < 0) return RBQ_E_MODEL;
    const int fam0 = family(p.model_id);
    // the block-parallel / warp / scan kernels implement the exponential step; Runge-Kutta steps go through model_step (one thread
    // per trajectory)
    const bool rk = p.integrator != RBQ_INTEGRATOR_EXP;
    const bool independent_blocks = !rk && (fam0 == RBQ_MODEL_SPIN13 || fam0 == RBQ_MODEL_SPIN15 || fam0 == RBQ_MODEL_SPIN12);
    // measured (tools/proto/rollout_batch.py, 1001 knots): spin13 4096 / 32768 trajectories 0.33 / 2.3 ms packed against
    // 0.60 / 4.2 ms one warp each; spin12 0.50 against 0.55 ms at 4096 but 6.6 against 3.9 ms at 32768 (43 doubles per
    // knot leave as scattered 32-byte stores), so the 10-block models switch back above 8192
    const bool packed_default = B >= 16 && (fam0 != RBQ_MODEL_SPIN12 || B <= 8192);
    // few trajectories: each as a prefix product over its knots (rbq_scan_rollout.cuh) instead of a walk along them.
    // Open loop only; the blocks of these families evolve linearly.  RBQ_ROLLOUT_SCAN=0/1 forces the choice.
    const int64_t steps = N - 1;
    const bool scan_ok = independent_blocks && steps >= 1 && steps <= (int64_t)SCAN_T * SCAN_MAX_PER && B <= 65535;
    const bool scan_default = B < 16 && steps >= 64;
    if (scan_ok && (h->opt.rollout_scan >= 0 ? h->opt.rollout_scan != 0 : scan_default)) {
        WarpRolloutArgs a{p, n, m, N, B, t0, d_x0, d_U, nullptr, nullptr, d_dt, nullptr, d_X, nullptr};
        const int per = (int)((steps + SCAN_T - 1) / SCAN_T);
        const size_t smem = scan_rollout_smem(steps);
        cudaError_t e;
        if (fam0 == RBQ_MODEL_SPIN12) {
            e = cudaFuncSetAttribute(scan_rollout_kernel<RBQ_MODEL_SPIN12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            scan_rollout_kernel<RBQ_MODEL_SPIN12><<<(unsigned)B, SCAN_T, smem, h->stream>>>(a, per);
        } else {
            e = cudaFuncSetAttribute(scan_rollout_kernel<RBQ_MODEL_SPIN13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            scan_rollout_kernel<RBQ_MODEL_SPIN13><<<(unsigned)B, SCAN_T, smem, h->stream>>>(a, per);
        }
    } else if (independent_blocks && (h->opt.rollout_packed >= 0 ? h->opt.rollout_packed != 0 : packed_default)) {
        // one thread per (trajectory, block): all lanes busy (see packed_rollout_kernel); 32-thread blocks spread small
        // batches over the SMs
        WarpRolloutArgs a{p, n, m, N, B, t0, d_x0, d_U, nullptr, nullptr, d_dt, nullptr, d_X, nullptr};
        const int nb = fam0 == RBQ_MODEL_SPIN12 ? 10 : 2;
        const int pt = 32;
        const unsigned blocks = (unsigned)((B * nb + pt - 1) / pt);
        if (fam0 == RBQ_MODEL_SPIN13) packed_rollout_kernel<RBQ_MODEL_SPIN13><<<blocks, pt, 0, h->stream>>>(a, nb, pt);
        else if (fam0 == RBQ_MODEL_SPIN15) packed_rollout_kernel<RBQ_MODEL_SPIN15><<<blocks, pt, 0, h->stream>>>(a, nb, pt);
        else packed_rollout_kernel<RBQ_MODEL_SPIN12><<<blocks, pt, 0, h->stream>>>(a, nb, pt);
    } else if (B <= WARP_ROLLOUT_MAX && !rk) {
        WarpRolloutArgs a{p, n, m, N, B, t0, d_x0, d_U, nullptr, nullptr, d_dt, nullptr, d_X, nullptr};
#define CALL(M) { cudaError_t e_ = launch_warp_rollout<M, false>(a, (unsigned)B, 0, h->stream); if (e_ != cudaSuccess) return (int)e_; }
        RBQ_DISPATCH_FAMILY(family(p.model_id), CALL)
#undef CALL
    } else if (rk) {
        // Runge-Kutta steps of the continuous dynamics: one thread per trajectory through model_step<.., RK = true>
        const int threads = 32;
        const unsigned blocks = (unsigned)((B + threads - 1) / threads);
        switch (fam0) {
        case RBQ_MODEL_SPIN13: rollout_kernel<RBQ_MODEL_SPIN13, true><<<blocks, threads, 0, h->stream>>>(p, n, m, B, N, t0, d_x0, d_U, d_dt, d_X); break;
        case RBQ_MODEL_SPIN15: rollout_kernel<RBQ_MODEL_SPIN15, true><<<blocks, threads, 0, h->stream>>>(p, n, m, B, N, t0, d_x0, d_U, d_dt, d_X); break;
        case RBQ_MODEL_SPIN11: rollout_kernel<RBQ_MODEL_SPIN11, true><<<blocks, threads, 0, h->stream>>>(p, n, m, B, N, t0, d_x0, d_U, d_dt, d_X); break;
        default: return RBQ_E_MODEL;
        }
    } else {
        const int threads = 128;
        const unsigned blocks = (unsigned)((B + threads - 1) / threads);
#define CALL(M) rollout_kernel<M><<<blocks, threads, 0, h->stream>>>(p, n, m, B, N, t0, d_x0, d_U, d_dt, d_X)
        RBQ_DISPATCH_FAMILY(family(p.model_id), CALL)
#undef CALL
    }
    h->launches++;
    return (int)cudaGetLastError();
}

int launch_closedloop(rbq_handle* h, const rbq_model_params& p, int64_t N, const double* d_Xref, const double* d_Uref,
                      const double* d_K, const double* d_d, const double* d_dt, int64_t nalpha, const double* d_alphas,
                      double* d_X, double* d_Uout) {
    const int n = model_state_dim(p), m = model_control_dim(p);
    if (n < 0 || m < 0) return RBQ_E_MODEL;
    if (p.integrator != RBQ_INTEGRATOR_EXP) return RBQ_E_MODEL;     // closed-loop kernels: exponential step only (include/rbq.h)
    // one step-size candidate per warp, one warp per block so that candidates spread over SMs
    WarpRolloutArgs a{p, n, m, N, nalpha, 0.0, d_Xref, d_Uref, d_K, d_d, d_dt, d_alphas, d_X, d_Uout};
    const size_t smem = closedloop_smem_bytes(n, m);
#define CALL(M) { cudaError_t e_ = launch_warp_rollout<M, true>(a, (unsigned)nalpha, smem, h->stream); if (e_ != cudaSuccess) return (int)e_; }
    RBQ_DISPATCH_FAMILY(family(p.model_id), CALL)
#undef CALL
    h->launches++;
    return (int)cudaGetLastError();
}

static const int JAC_STAGE_BYTES = 200 * 1024;   // shared memory a staging block may use (227 KB per SM on sm_100)
static const int64_t JAC_STAGED_MIN_KNOTS = 4096;  // below this the launch is latency bound and the direct kernel is as fast

int launch_jacobians(rbq_handle* h, const rbq_model_params& p, int64_t K, const double* d_X, const double* d_U,
                     const double* d_t, const double* d_dt, double* d_AB) {
    const int n = model_state_dim(p), m = model_control_dim(p);
    if (n < 0 || m < 0) return RBQ_E_MODEL;
    const int fam = family(p.model_id);
    // columns per lane: 4 for the small models, 2 for the 43/56+-state ones (register budget)
    int nt = (fam == RBQ_MODEL_SPIN30 || fam == RBQ_MODEL_SPIN25 || fam == RBQ_MODEL_SPIN12) ? 2 : 4;
    if (p.integrator != RBQ_INTEGRATOR_EXP) {
        // Runge-Kutta variants: generic forward mode through step_rk, 4 seed columns per lane, direct stores
        const int lpk4 = (n + m + 3) / 4, threads = 128;
        const unsigned blocks = (unsigned)((K * lpk4 + threads - 1) / threads);
        switch (fam) {
        case RBQ_MODEL_SPIN13: jacobians_kernel<RBQ_MODEL_SPIN13, 4, true><<<blocks, threads, 0, h->stream>>>(p, n, m, lpk4, K, d_X, d_U, d_t, d_dt, d_AB); break;
        case RBQ_MODEL_SPIN15: jacobians_kernel<RBQ_MODEL_SPIN15, 4, true><<<blocks, threads, 0, h->stream>>>(p, n, m, lpk4, K, d_X, d_U, d_t, d_dt, d_AB); break;
        case RBQ_MODEL_SPIN11: jacobians_kernel<RBQ_MODEL_SPIN11, 4, true><<<blocks, threads, 0, h->stream>>>(p, n, m, lpk4, K, d_X, d_U, d_t, d_dt, d_AB); break;
        default: return RBQ_E_MODEL;
        }
        h->launches++;
        return (int)cudaGetLastError();
    }
    const bool unscented = fam == RBQ_MODEL_SPIN30 || fam == RBQ_MODEL_SPIN25;
    // unscented models with one sample state (every run of the reference): the block is assembled from its structure, the value
    // path evaluated once per knot instead of once per seed column (rbq_jac_unscented.cu); RBQ_JAC_STRUCTURED=0 keeps the
    // generic forward-mode kernel below (tests compare the two)
    const bool nom_ok = (p.nominal_offset[0] & 3) == 0 && p.nominal_offset[0] >= 0 && p.nominal_offset[0] <= (fam == RBQ_MODEL_SPIN25 ? 4 : 8);
    if (unscented && p.sample_state_count == 1 && nom_ok && h->opt.jac_structured && (reinterpret_cast<uintptr_t>(d_AB) & 15) == 0) {
        void* rec = nullptr;
        const int rc = rbq_scratch(h, RBQ_SCR_JACREC, unscented_record_bytes(K), &rec);
        if (rc) return rc;
        return launch_jacobians_unscented(h, p, K, d_X, d_U, d_t, d_dt, (double*)rec, d_AB);
    }
    // the unscented step is register hungry: one tangent per lane (254 registers, no spills) beats two (255 + 868 B of spills)
    // by 1.4x on batches although every lane repeats the value path; RBQ_JAC_NT_UNSCENTED=2 selects the old split
    if (unscented) nt = h->opt.jac_nt_unscented == 2 ? 2 : 1;
    const int lpk = (n + m + nt - 1) / nt;
    const int blk_bytes = n * (n + m) * (int)sizeof(double);
    // Staging budget per block, measured (tools/proto/jac_sweep.py, 1e6 / 2.6e5 / 1.3e5 knots): small stages leave room for many
    // resident blocks, whose compute and write-out phases then overlap -- 20 KB: 4.75 TB/s written (spin13) and 4.8 TB/s
    // (spin12, one knot per block) against 3.7 / 3.1 TB/s with 200 KB; the unscented models are compute bound (1.7 TB/s), one knot per block is best there too
    int stage_bytes = blk_bytes > 20 * 1024 ? 30 * 1024 : 20 * 1024;
    if (h->opt.jac_stage_kb > 0) stage_bytes = h->opt.jac_stage_kb * 1024;
    int G = stage_bytes / blk_bytes;
    if (G < 1 && blk_bytes <= JAC_STAGE_BYTES) G = 1;
    if (G > 256 / lpk) G = 256 / lpk;
    const bool staged = G >= 1 && (h->opt.jac_staged >= 0 ? h->opt.jac_staged != 0 : K >= JAC_STAGED_MIN_KNOTS);
    if (staged) {
        const int threads = ((G * lpk + 31) / 32) * 32;
        const unsigned blocks = (unsigned)((K + G - 1) / G);
        const size_t smem = (size_t)G * blk_bytes;
#define CALLS(M, NT_)                                                                                                          \
    {                                                                                                                          \
        cudaError_t e = cudaFuncSetAttribute(jacobians_staged_kernel<M, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                             JAC_STAGE_BYTES);                                                                 \
        if (e != cudaSuccess) return (int)e;                                                                                   \
        jacobians_staged_kernel<M, NT_><<<blocks, threads, smem, h->stream>>>(p, n, m, lpk, G, K, d_X, d_U, d_t, d_dt, d_AB);       \
    }
        switch (fam) {
        case RBQ_MODEL_SPIN13: CALLS(RBQ_MODEL_SPIN13, 4); break;
        case RBQ_MODEL_SPIN15: CALLS(RBQ_MODEL_SPIN15, 4); break;
        case RBQ_MODEL_SPIN11: CALLS(RBQ_MODEL_SPIN11, 4); break;
        case RBQ_MODEL_SPIN12: CALLS(RBQ_MODEL_SPIN12, 2); break;
        case RBQ_MODEL_SPIN30: if (nt == 1) CALLS(RBQ_MODEL_SPIN30, 1) else CALLS(RBQ_MODEL_SPIN30, 2) break;
        case RBQ_MODEL_SPIN25: if (nt == 1) CALLS(RBQ_MODEL_SPIN25, 1) else CALLS(RBQ_MODEL_SPIN25, 2) break;
        default: return RBQ_E_MODEL;
        }
#undef CALLS
    } else {
        const int threads = 128;
        const int64_t total = K * lpk;
        const unsigned blocks = (unsigned)((total + threads - 1) / threads);
#define CALL1(M) jacobians_kernel<M, 1><<<blocks, threads, 0, h->stream>>>(p, n, m, lpk, K, d_X, d_U, d_t, d_dt, d_AB)
#define CALL2(M) jacobians_kernel<M, 2><<<blocks, threads, 0, h->stream>>>(p, n, m, lpk, K, d_X, d_U, d_t, d_dt, d_AB)
#define CALL4(M) jacobians_kernel<M, 4><<<blocks, threads, 0, h->stream>>>(p, n, m, lpk, K, d_X, d_U, d_t, d_dt, d_AB)
        switch (fam) {
        case RBQ_MODEL_SPIN13: CALL4(RBQ_MODEL_SPIN13); break;
        case RBQ_MODEL_SPIN15: CALL4(RBQ_MODEL_SPIN15); break;
        case RBQ_MODEL_SPIN11: CALL4(RBQ_MODEL_SPIN11); break;
        case RBQ_MODEL_SPIN12: CALL2(RBQ_MODEL_SPIN12); break;
        case RBQ_MODEL_SPIN30: if (nt == 1) CALL1(RBQ_MODEL_SPIN30); else CALL2(RBQ_MODEL_SPIN30); break;
        case RBQ_MODEL_SPIN25: if (nt == 1) CALL1(RBQ_MODEL_SPIN25); else CALL2(RBQ_MODEL_SPIN25); break;
        default: return RBQ_E_MODEL;
        }
#undef CALL1
#undef CALL2
#undef CALL4
    }
    h->launches++;
    return (int)cudaGetLastError();
}

}  // namespace rbq
