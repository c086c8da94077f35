// rbq_cost.cu — augmented-Lagrangian cost expansion of one ALTRO inner iteration, all knots in one launch.
//
// Replaces, for the problem class the reference sets up (spin13.jl:122-153 and its siblings: diagonal LQRObjective,
// BoundConstraint on the augmented state, GoalConstraint at the last knot, NormConstraint on the iso states), what
// Altro / TrajectoryOptimization do per knot on the host: cost_expansion!(E, ::ALObjective, Z), i.e. the objective's
// gradient and Hessian plus, for every constraint value c with multipliers lambda, penalties mu and active set a,
//     J   += lambda'c + 1/2 c' I_mu c,       I_mu = diag(a .* mu)
//     Exx += cx' I_mu cx,   Ex += cx'(lambda + I_mu c)      (first-order expansion of the constraint: no second-derivative term)
// with a_j = 1 for equalities and (c_j >= 0) | (lambda_j > 0) for inequalities.  The solver is external and un-vendored
// (Manifest.toml:26-32, 1523-1529); this restates its published algorithm (see oracle/al_cost.py, "parity unpinned").
// The blocks are written where rbq_backward_pass_dev reads them, so an iteration never leaves HBM.
//
// One block per knot: the n x n block is zero-filled and written once (the kernel is bound by that write: n^2 * 8 B per knot).
#include "rbq_internal.h"
#include "rbq_math.cuh"

namespace rbq {

constexpr int AL_THREADS = 128;

struct AlArgs {
    int n, m; int64_t N, B, dual_stride;
    const double* X; const double* U;
    const double* Qd; const double* Rd; const double* Qfd; const double* xf; const double* w;
    int ncon; const rbq_constraint* cons; const int32_t* idx_pool; const double* val_pool;
    const double* lambda; const double* mu;
    double* Exx; double* Ex; double* Euu; double* Eux; double* Eu; double* c; double* J;
};

__global__ void __launch_bounds__(AL_THREADS) al_cost_kernel(const AlArgs a) {
    const int n = a.n, m = a.m, tid = threadIdx.x;
    const int64_t k = blockIdx.x, b = blockIdx.y;              // knot, trajectory of the batch
    const bool term = k == a.N - 1;
    const double* x = a.X + (b * a.N + k) * n;
    const double* u = term ? nullptr : a.U + (b * (a.N - 1) + k) * m;
    const double wk = a.w ? a.w[k] : 1.0;
    const bool expand = a.Exx != nullptr;
    double* Exx = expand ? a.Exx + (b * a.N + k) * (int64_t)(n * n) : nullptr;
    double* Ex = expand ? a.Ex + (b * a.N + k) * (int64_t)n : nullptr;
    double* Euu = expand && !term ? a.Euu + (b * (a.N - 1) + k) * (int64_t)(m * m) : nullptr;
    double* Eu = expand && !term ? a.Eu + (b * (a.N - 1) + k) * (int64_t)m : nullptr;
    double Jp = 0.0;                                             // this thread's share of the knot's cost
    // c of trajectory b starts at b * (total dual length); the table's rows are in dual_off order, the last one ends it
    int64_t dual_size = 0;
    if (a.ncon > 0) { const rbq_constraint last = a.cons[a.ncon - 1]; dual_size = last.dual_off + (last.k1 - last.k0) * last.p; }

    // ---- objective: 1/2 (x - xf)'Q(x - xf) + 1/2 u'Ru, Q -> Qf at the last knot, scaled by w[k] ----
    if (expand) {
        for (int i = tid; i < n * n; i += AL_THREADS) Exx[i] = 0.0;
        if (!term) {
            for (int i = tid; i < m * n; i += AL_THREADS) a.Eux[(b * (a.N - 1) + k) * (int64_t)(m * n) + i] = 0.0;
            for (int i = tid; i < m * m; i += AL_THREADS) Euu[i] = 0.0;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += AL_THREADS) {
        const double q = wk * (term ? a.Qfd[i] : a.Qd[i]), dx = x[i] - a.xf[i];
        Jp = fma(0.5 * q * dx, dx, Jp);
        if (expand) { Exx[i * n + i] = q; Ex[i] = q * dx; }
    }
    if (!term)
        for (int i = tid; i < m; i += AL_THREADS) {
            const double r = wk * a.Rd[i];
            Jp = fma(0.5 * r * u[i], u[i], Jp);
            if (expand) { Euu[i * m + i] = r; Eu[i] = r * u[i]; }
        }

    // ---- constraints that cover this knot, one after the other (they may touch the same entries) ----
    for (int ci = 0; ci < a.ncon; ++ci) {
        const rbq_constraint cn = a.cons[ci];
        if (k < cn.k0 || k >= cn.k1) continue;                   // uniform over the block
        __syncthreads();
        const int64_t doff = cn.dual_off + (k - cn.k0) * cn.p;
        const double* lam = a.lambda + b * a.dual_stride + doff;
        const double* mu = a.mu + b * a.dual_stride + doff;
        double* cv = a.c ? a.c + b * dual_size + doff : nullptr;
        if (cn.kind == RBQ_CON_BOUND) {
            // p = 2(n+m): z <= zmax (entries 0..n+m-1) then zmin <= z; thread j owns both entries of z_j
            const int nz = n + m;
            const double* zmax = a.val_pool + cn.val_off;
            const double* zmin = zmax + nz;
            for (int j = tid; j < nz; j += AL_THREADS) {
                const bool isu = j >= n;
                double cu = -INFINITY, cl = -INFINITY;
                if (!(isu && term)) {
                    const double z = isu ? u[j - n] : x[j];
                    cu = z - zmax[j]; cl = zmin[j] - z;
                }
                if (cv) { cv[j] = cu; cv[nz + j] = cl; }
                double h = 0.0, gsum = 0.0;
                if (cu > -INFINITY && (cu >= 0.0 || lam[j] > 0.0)) { h += mu[j]; gsum += fma(mu[j], cu, lam[j]); Jp += fma(0.5 * mu[j] * cu, cu, lam[j] * cu); }
                else if (cu > -INFINITY) { gsum += lam[j]; Jp += lam[j] * cu; }
                if (cl > -INFINITY && (cl >= 0.0 || lam[nz + j] > 0.0)) { h += mu[nz + j]; gsum -= fma(mu[nz + j], cl, lam[nz + j]); Jp += fma(0.5 * mu[nz + j] * cl, cl, lam[nz + j] * cl); }
                else if (cl > -INFINITY) { gsum -= lam[nz + j]; Jp += lam[nz + j] * cl; }
                if (expand && !(isu && term)) {
                    if (isu) { Euu[(j - n) * m + (j - n)] += h; Eu[j - n] += gsum; }
                    else { Exx[j * n + j] += h; Ex[j] += gsum; }
                }
            }
        } else if (cn.kind == RBQ_CON_GOAL) {
            // equality x[idx_j] = target_j (distinct indices)
            const int32_t* idx = a.idx_pool + cn.idx_off;
            const double* tgt = a.val_pool + cn.val_off;
            for (int j = tid; j < cn.nidx; j += AL_THREADS) {
                const int i = idx[j];
                const double cj = x[i] - tgt[j];
                if (cv) cv[j] = cj;
                Jp += fma(0.5 * mu[j] * cj, cj, lam[j] * cj);
                if (expand) { Exx[i * n + i] += mu[j]; Ex[i] += fma(mu[j], cj, lam[j]); }
            }
        } else if (cn.kind == RBQ_CON_NORM) {
            // equality sum_j x[idx_j]^2 = val; gradient 2 x[idx], Gauss-Newton block 4 mu x x'
            const int32_t* idx = a.idx_pool + cn.idx_off;
            const int q = cn.nidx;
            double cc = -a.val_pool[cn.val_off];
            for (int j = 0; j < q; ++j) cc = fma(x[idx[j]], x[idx[j]], cc);    // every thread, same order
            if (tid == 0) { if (cv) cv[0] = cc; Jp += fma(0.5 * mu[0] * cc, cc, lam[0] * cc); }
            if (expand) {
                const double gs = 2.0 * fma(mu[0], cc, lam[0]);
                for (int e = tid; e < q * q; e += AL_THREADS) {
                    const int ia = idx[e % q], ib = idx[e / q];
                    Exx[ib * n + ia] += 4.0 * mu[0] * x[ia] * x[ib];
                }
                for (int j = tid; j < q; j += AL_THREADS) Ex[idx[j]] += gs * x[idx[j]];
            }
        }
    }

    // ---- knot cost: fixed-order tree over the block ----
    __shared__ double red[AL_THREADS];
    red[tid] = Jp;
    __syncthreads();
    for (int s = AL_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    if (tid == 0 && a.J) a.J[b * a.N + k] = red[0];
}

int launch_al_cost(rbq_handle* h, int n, int m, int64_t N, int64_t B, int64_t dual_stride, const double* d_X, const double* d_U, const double* d_Qd,
                   const double* d_Rd, const double* d_Qfd, const double* d_xf, const double* d_w, int ncon,
                   const rbq_constraint* d_cons, const int32_t* d_idx_pool, const double* d_val_pool, const double* d_lambda,
                   const double* d_mu, double* d_Exx, double* d_Ex, double* d_Euu, double* d_Eux, double* d_Eu, double* d_c,
                   double* d_J) {
    AlArgs a{n, m, N, B, dual_stride, d_X, d_U, d_Qd, d_Rd, d_Qfd, d_xf, d_w, ncon, d_cons, d_idx_pool, d_val_pool, d_lambda, d_mu,
             d_Exx, d_Ex, d_Euu, d_Eux, d_Eu, d_c, d_J};
    al_cost_kernel<<<dim3((unsigned)N, (unsigned)B), AL_THREADS, 0, h->stream>>>(a);
    h->launches++;
    return (int)cudaGetLastError();
}

// ---- outer-loop updates of the augmented Lagrangian (Altro dual_update! / penalty_update! / max_violation) ----
// One block row per constraint, blockIdx.y walks its entries in chunks.  lambda <- clamp(lambda + mu c, lo, lambda_max) with
// lo = 0 for inequalities (BoundConstraint) and -lambda_max for equalities; then mu <- min(phi mu, mu_max).  Entries whose c
// is -inf (absent bounds) are left alone.  The largest violation (|c| for equalities, max(c, 0) for inequalities) is
// reduced per block and folded into viol[0] with an atomic max on the bit pattern (non-negative doubles order like integers).
constexpr int DU_THREADS = 256, DU_CHUNK = 4096;
__global__ void __launch_bounds__(DU_THREADS) al_dual_update_kernel(int ncon, const rbq_constraint* __restrict__ cons, int64_t B,
                                                                    int64_t dual_stride, const double* __restrict__ c,
                                                                    double* __restrict__ lambda, double* __restrict__ mu, double phi,
                                                                    double mu_max, double lambda_max, int update,
                                                                    unsigned long long* __restrict__ viol) {
    const rbq_constraint cn = cons[blockIdx.x];
    const rbq_constraint last = cons[ncon - 1];
    const int64_t dual_size = last.dual_off + (last.k1 - last.k0) * last.p;
    const int64_t len = (cn.k1 - cn.k0) * cn.p;
    const bool ineq = cn.kind == RBQ_CON_BOUND;
    double vmax = 0.0;
    for (int64_t b = 0; b < B; ++b) {
        const int64_t base = (int64_t)blockIdx.y * DU_CHUNK;
        for (int64_t e = base + threadIdx.x; e < len && e < base + DU_CHUNK; e += DU_THREADS) {
            const double cv = c[b * dual_size + cn.dual_off + e];
            if (!(cv > -INFINITY)) continue;                                  // absent bound (NaN would be a caller error: skipped too)
            vmax = fmax(vmax, ineq ? fmax(cv, 0.0) : fabs(cv));
            if (update && (dual_stride != 0 || b == 0)) {                     // shared multipliers: updated from trajectory 0 only
                const int64_t i = b * dual_stride + cn.dual_off + e;
                const double l = fma(mu[i], cv, lambda[i]);
                lambda[i] = fmin(fmax(l, ineq ? 0.0 : -lambda_max), lambda_max);
                mu[i] = fmin(phi * mu[i], mu_max);
            }
        }
    }
    __shared__ double red[DU_THREADS];
    red[threadIdx.x] = vmax;
    __syncthreads();
    for (int s = DU_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0 && red[0] > 0.0) atomicMax(viol, (unsigned long long)__double_as_longlong(red[0]));
}

int launch_al_dual_update(rbq_handle* h, int ncon, const rbq_constraint* d_cons, int64_t max_len, int64_t B, int64_t dual_stride,
                          const double* d_c, double* d_lambda, double* d_mu, double phi, double mu_max, double lambda_max,
                          int update, double* d_viol) {
    cudaError_t e = cudaMemsetAsync(d_viol, 0, sizeof(double), h->stream);
    if (e != cudaSuccess) return (int)e;
    const unsigned chunks = (unsigned)((max_len + DU_CHUNK - 1) / DU_CHUNK);
    al_dual_update_kernel<<<dim3((unsigned)ncon, chunks ? chunks : 1), DU_THREADS, 0, h->stream>>>(
        ncon, d_cons, B, dual_stride, d_c, d_lambda, d_mu, phi, mu_max, lambda_max, update, reinterpret_cast<unsigned long long*>(d_viol));
    h->launches++;
    return (int)cudaGetLastError();
}

}  // namespace rbq
