// rbq_riccati.cu — iLQR backward pass (time-varying LQR Riccati recursion) for batches of problems, one CTA per problem.
//
// Replaces Altro.jl 0.2.0 `backwardpass!` (external, SchusterLab/rbqoc-stable fork, Manifest.toml:26-32; un-vendored, restated
// from its published algorithm, see oracle/riccati.py).  It consumes the knot-point Jacobians exactly as rbq_jacobians
// leaves them in HBM (column-major n x (n+m) blocks), so an ALTRO iteration never moves them over PCIe:
//
//   S_N = Exx_N, s_N = Ex_N;  for k = N-1 .. 1:
//     Qxx = Exx + A'SA   Quu = Euu + B'SB   Qux = Eux + B'SA   Qx = Ex + A's   Qu = Eu + B's
//     Quu_reg = Quu + rho I (control)  |  Quu + rho B'B, Qux_reg = Qux + rho B'A (state)
//     K = -Quu_reg^-1 Qux_reg   d = -Quu_reg^-1 Qu
//     S = sym(Qxx + K'Quu K + K'Qux + Qux'K)   s = Qx + K'Quu d + K'Qu + Qux'd   dV += (d'Qu, d'Quu d / 2)
//
// Per knot the work is two dense products, T = S [A B] (n x n x (n+m)) and G = [A B]' T (lower 8x8 tiles only), on the FP64
// tensor cores (mma.sync.m8n8k4.f64, SASS DMMA; tcgen05 has no FP64 kind) out of shared memory; everything else is
// O(n^2 m^2).  DMMA-issue bound for n >= ~40, barrier/latency bound below.  The next knot's Jacobian and cost blocks arrive by
// bulk TMA copies (cp.async below n = 28) while the gains and the value update of this knot are computed.
#include "rbq_internal.h"

#include "rbq_math.cuh"
#include "rbq_tma.cuh"

namespace rbq {

constexpr int BP_WARPS_LARGE = 8;    // warps per problem: 8 for n > 16, 1 (no block-wide barriers at all) for the small models
constexpr int BP_WARPS_SMALL = 1;
constexpr int BP_MAX_M = 4;

RBQ_DEV void cp_async8(double* smem_dst, const double* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
RBQ_DEV void cp_async16(double* smem_dst, const double* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
RBQ_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
RBQ_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// thread-block cluster primitives (CL > 1: one problem split over the column strips of CL CTAs, see backward_pass_kernel)
RBQ_DEV uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
RBQ_DEV uint32_t cluster_map(const void* local, uint32_t rank) {          // address of the same variable in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
RBQ_DEV void st_cluster(uint32_t addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
RBQ_DEV void st_cluster(uint32_t addr, int v) { asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// bulk copy from this CTA's shared memory into a peer's, completing bytes on the peer's mbarrier (both cluster addresses)
RBQ_DEV void bulk_s2peer(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
                 "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
RBQ_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
RBQ_DEV void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
RBQ_DEV void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// phase timing for tuning (make EXTRA=-DRBQ_BP_PROFILE; read with rbq_debug_bp_prof): cycles of thread 0 between barriers
#ifdef RBQ_BP_PROFILE
__device__ long long g_bp_prof[32];       // [0, 16): the first CTA of a problem, [16, 32): the CTA that computes the gains (cluster form)
#define BP_PROF(i) if (tid == 0 && (rank == 0 || rank == owner)) { long long c_ = clock64(); g_bp_prof[(rank == 0 ? 0 : 16) + i] += c_ - tprev; tprev = c_; }
#else
#define BP_PROF(i)
#endif
struct BackwardArgs {
    int n, m;
    int64_t N, B;
    const double* AB; const double* Exx; const double* Ex; const double* Euu; const double* Eux; const double* Eu;
    double rho; int reg_type;
    double* K; double* d; double* dV; long long* fail;
};

// staging modes of backward_pass_kernel (template parameter MODE)
constexpr int BP_TMA_J = 1, BP_TMA_C = 2, BP_DENSE_J = 4, BP_REG_J = 8, BP_REG_C = 16, BP_REG = BP_REG_J | BP_REG_C;
// BP_REG_J / BP_REG_C: the Jacobian block / the cost blocks travel through registers.  Every thread loads its few elements of the
// next knot's blocks at the top of a knot (plain coalesced loads: nothing waits on them) and stores them into shared memory where
// the existing barriers already separate a buffer's readers from its next writer -- the Jacobian block after the gains, the cost
// blocks after the top barrier of the next knot.  A cp.async instruction holds its warp ~270 cycles whatever its width and a
// small problem issues five per warp and knot (n = 11: 1.21 -> 0.82 us per knot with BP_REG); the per-column bulk copies of an
// even-n Jacobian block occupy every warp for ~700 cycles (n = 56: BP_REG_J | BP_TMA_C).
__host__ __device__ constexpr int bp_reg_count(int elems, int threads) { return (elems + threads - 1) / threads; }

// Shared-memory layout (doubles), all column-major.  Rows are padded with zeros to a multiple of 8 (the DMMA tile), the
// contraction length to a multiple of 4; the leading dimensions ld = n8 + 4 and ldg = q8 + 4 are = 4 (mod 8), which makes every
// fragment load below conflict-free (the four 32-byte pieces a half-warp reads fall into different quarters of the bank window):
//   S[ld*n4] sv[n8] AB[ld*q8] T[ld*q8] G[ldg*q8] Kk[4*n8] QuuK[4*n8] dk[4]      (q8 = n+m+1 rounded up: one spare column)
RBQ_DEV int up4(int x) { return (x + 3) & ~3; }
RBQ_DEV int up8(int x) { return (x + 7) & ~7; }

// FP64 tensor-core tile product D(8x8) += A(8x4) B(4x8).  Fragments (PTX ISA, mma.m8n8k4 .f64): lane l holds
// A[l/4][l%4], B[l%4][l/4], C[l/4][2(l%4)] and C[l/4][2(l%4)+1].
RBQ_DEV void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int BP_MAX_TILES = 9;        // 8-row tiles of a column strip: (64 + 4 + 7) / 8
constexpr int BP_GSLOTS = 8;           // tiles of the second product per warp: 53 over 8 warps at n = 64, 8 for one warp at n = 16

// barrier among the warps that share a problem: a real one for the 8-warp form, a warp barrier for the 1-warp form
template <int NW_> RBQ_DEV void bp_sync() { if (NW_ == 1) __syncwarp(); else __syncthreads(); }
template <int NW_> RBQ_DEV bool bp_sync_or(bool pred) { return NW_ == 1 ? __any_sync(0xffffffffu, pred) != 0 : __syncthreads_or(pred) != 0; }

// tiles of the second product one warp of a CL-CTA cluster has to run (the largest share over the ranks): CTA `rank` owns the
// column strips rank, rank + CL, ...; a strip jt has its lower tiles (ctiles - jt) plus, for the strip that holds the gradient
// column, the jtc tiles above the diagonal
__host__ __device__ constexpr int bp_cluster_gslots(int n, int m, int cl, int nw) {
    const int ctiles = ((n + m + 8) & ~7) / 8, jtc = (n + m) >> 3;
    int worst = 0;
    for (int r = 0; r < cl; ++r) {
        int cnt = 0;
        for (int jt = r; jt < ctiles; jt += cl) cnt += ctiles - jt + (jt == jtc ? jtc : 0);
        worst = cnt > worst ? cnt : worst;
    }
    return (worst + nw - 1) / nw;
}

template <int NW, int MODE, int MM, int NN, int CL = 1>
// two problems per SM where the shared memory allows it (n <= 44: under 100 KB each): batches of mid-size problems then run
// two per SM at a time, which needs the kernel within 128 registers
__global__ void __launch_bounds__(32 * NW, NW == 1 ? 8 : (NN != 0 && NN <= 44 ? 2 : 1)) backward_pass_kernel(const BackwardArgs a) {
    constexpr int BP_THREADS = 32 * NW;
    extern __shared__ __align__(16) double sm[];
    // MM: the number of controls when known at compile time (1 for every reference model but the time-optimal spin15, which
    // has 2; 0: any) -- the m-loops below fold to MM iterations
    // NN: the state dimension itself for the reference models' shapes (0: any) -- strides and tile counts become constants
    const int n = NN ? NN : a.n, m = MM ? MM : a.m, nm = n + m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n4 = up4(n), n8 = up8(n), q8 = up8(nm + 1), ld = n8 + 4, ldg = q8 + 4;
    const int cs = nm;                 // spare column: G(:, cs) = [A B]' s + e, i.e. the gradient g rides in the second product
    // CL > 1 (shape-specialised, one control): the problem is split over a cluster of CL CTAs.  Every CTA holds all of S, s and
    // [A B]; CTA `rank` computes the column strips rank, rank + CL, ... of T and of G (the lower tiles; the strip with the
    // gradient column also above the diagonal), the owner of that strip computes the gains and sends K, Quu K and Qxu to the
    // others, every CTA updates the entries of S below its strips' diagonal and writes them, mirrored, into every copy of S
    // through distributed shared memory.  Two cluster barriers per knot.
    static_assert(CL == 1 || (NN != 0 && MM == 1 && NW > 1), "cluster form: specialised shapes with one control");
    const uint32_t rank = CL > 1 ? cluster_rank() : 0u;
    const int64_t N = a.N, b = blockIdx.x / CL;
    double* S = sm;                    // S(r, kk) at kk * ld + r   (symmetric)
    double* sv = S + ld * n4;
    double* AB = sv + n8;              // [A B](kk, c) at c * ld + kk
    double* T = AB + ld * q8;          // T(r, c) at c * ld + r
    double* G = T + ld * q8;           // G(i, j) at j * ldg + i
    double* g = G + cs * ldg;          // Qx, Qu: column cs of G
    double* Kk = G + ldg * q8;         // K(p, j) at j * m + p
    double* QuuK = Kk + BP_MAX_M * n8;
    double* dk = QuuK + BP_MAX_M * n8;
    double* Hq = dk + BP_MAX_M;        // Qxu as received from the CTA that computed it (cluster form)
    double* Qd = Hq + n8;              // dense landing zone of the cost Hessian block: n*n + 2 doubles, 16-byte aligned
    __shared__ double dv_s[2];
    __shared__ __align__(8) uint64_t bar_s[2];       // bulk-copy arrival: [0] Jacobian block, [1] cost Hessian block
    __shared__ int fail_s;                           // cluster form: the owner's "Quu not positive definite" verdict
    __shared__ __align__(8) uint64_t bar_x;          // cluster form: arrival of the other CTAs' parts of S (and of s)

    const double* ABg = a.AB + b * (N - 1) * (int64_t)(n * nm);
    const double* Exxg = a.Exx + b * N * (int64_t)(n * n);
    const double* Exg = a.Ex + b * N * (int64_t)n;
    const double* Euug = a.Euu + b * (N - 1) * (int64_t)(m * m);
    const double* Euxg = a.Eux + b * (N - 1) * (int64_t)(m * n);
    const double* Eug = a.Eu + b * (N - 1) * (int64_t)m;
    double* Kg = a.K + b * (N - 1) * (int64_t)(m * n);
    double* dg = a.d + b * (N - 1) * (int64_t)m;

    // zero everything once (the padding is never written again), then the terminal value function
    const int total = ld * n4 + n8 + 2 * ld * q8 + ldg * q8;
    for (int i = tid; i < total; i += BP_THREADS) sm[i] = 0.0;
    if (tid == BP_THREADS - 1) { fail_s = 0; dv_s[0] = 0.0; dv_s[1] = 0.0; mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1); mbar_init(&bar_x, 1); mbar_fence_init(); }
    bp_sync<NW>();
    for (int j = wid; j < n; j += NW)
        for (int i = lane; i < n; i += 32)
            S[j * ld + i] = Exxg[(N - 1) * (int64_t)(n * n) + (i >= j ? j * n + i : i * n + j)];   // lower triangle, mirrored
    for (int i = tid; i < n; i += BP_THREADS) sv[i] = Exg[(N - 1) * (int64_t)n + i];
    // Staging cost, measured (make EXTRA=-DRBQ_BP_PROFILE): a warp waits ~270 cycles on every cp.async instruction whatever its
    // width, so a 25 KB block staged by cp.async costs each warp ~2.3 k cycles in which it issues no DMMA.  Even n (every column
    // 16-byte aligned in HBM and in the padded layout): one bulk TMA copy per column instead, the columns dealt over the lanes
    // of all warps (57 copies from a single warp serialise: 3.8 k cycles; spread out, ~0.9 k for the block, which looks like the
    // TMA unit's request rate), and the small cost blocks through registers (load before the first product, store after it).
    // That takes the n = 56 knot from 10.0 to 9.3 us.  RBQ_BP_TMA=0 forces the cp.async path (bit 0: Jacobian, bit 1: cost).
    // A single bulk copy per block needs the unpadded layout (bank conflicts in every fragment load) or a padded-box tensor
    // map.  Issuing the next Jacobian's copies from the warps the gains leave idle changes nothing (9.31 us), issuing the cost
    // blocks a knot ahead moves time between phases without shortening the knot, and issuing cp.async
    // copies one column per contraction step inside the tensor-core loops (double-buffered [A B]) makes it worse (11.5 us).
    // The staging mode is a template parameter chosen by the launcher (bp_stage_mode): with the choice made at run time
    // inside one kernel, the mere presence of the unused paths cost 0.3-0.8 us per knot (2.63 -> 3.1 us at n = 11).
    // BP_DENSE_J: Jacobian block dense in shared memory (leading dimension n) where that keeps the fragment loads off each
    // other's banks (odd n, n = 4 mod 8): the whole block is then ONE bulk copy.  The rows a tile reads past n belong to the
    // next column; they meet zero rows / columns of S and T, so they only need to be finite.
    // BP_TMA_J: padded layout, one bulk copy per column (even n).
    // BP_TMA_C: cost Hessian block as ONE bulk copy into the dense buffer Qd (the second product's accumulators start from
    // it; its few loads do not care about banks); a block that starts on an odd double is copied from the double before it.
    // Small blocks stay on cp.async (a copy or two per thread costs less than the bulk copy's latency, which the short first
    // product cannot hide: 3.0 against 3.4 us per knot at n = 11).
    constexpr bool dense_j = (MODE & BP_DENSE_J) != 0, tma_j = (MODE & BP_TMA_J) != 0, tma_c = (MODE & BP_TMA_C) != 0;
    const int lda = dense_j ? n : ld;
    const double* Qk = Qd;
    auto stage_jacobian = [&](int64_t k) {
        const double* src = ABg + k * (int64_t)(n * nm);
        if (dense_j) {
            if (tid == 0) {
                mbar_expect_tx(&bar_s[0], (uint32_t)(n * nm * sizeof(double)));
                bulk_g2s(AB, src, (uint32_t)(n * nm * sizeof(double)), &bar_s[0]);
            }
        } else if (tma_j) {
            if (tid == 0) mbar_expect_tx(&bar_s[0], (uint32_t)(n * nm * sizeof(double)));
            for (int j = wid + NW * lane; j < nm; j += NW * 32)
                bulk_g2s(AB + j * ld, src + j * n, (uint32_t)(n * sizeof(double)), &bar_s[0]);
        } else if ((n & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {          // every column starts 16-byte aligned
            for (int j = wid; j < nm; j += NW)
                for (int i = 2 * lane; i < n; i += 64) cp_async16(AB + j * ld + i, src + j * n + i);
        } else {
            for (int j = wid; j < nm; j += NW)
                for (int i = lane; i < n; i += 32) cp_async8(AB + j * ld + i, src + j * n + i);
        }
        cp_async_commit();
    };
    // cost blocks of knot k -> G (both triangles from the lower one) and g, in flight behind the first product
    auto stage_cost = [&](int64_t k) {
        const double* Exxk = Exxg + k * (int64_t)(n * n);
        const double* Euuk = Euug + k * (int64_t)(m * m);
        const double* Euxk = Euxg + k * (int64_t)(m * n);
        if (tma_c) {
            const int shift = (int)((reinterpret_cast<uintptr_t>(Exxk) >> 3) & 1);
            Qk = Qd + shift;
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)(((n * n + shift + 1) & ~1) * sizeof(double));
                mbar_expect_tx(&bar_s[1], bytes);
                bulk_g2s(Qd, Exxk - shift, bytes, &bar_s[1]);
            }
            return;                                                 // the small blocks travel through registers: extra_src/_store
        }
        for (int j = wid; j < nm; j += NW)
            for (int i = lane; i < nm; i += 32) {
                const int lo = i < j ? i : j, hi = i < j ? j : i;                     // (hi, lo) in the lower triangle
                const double* src = hi < n ? Exxk + lo * n + hi : (lo < n ? Euxk + lo * m + (hi - n) : Euuk + (lo - n) * m + (hi - n));
                cp_async8(G + j * ldg + i, src);
            }
        for (int i = tid; i < nm; i += BP_THREADS) cp_async8(g + i, i < n ? Exg + k * (int64_t)n + i : Eug + k * (int64_t)m + (i - n));
        cp_async_commit();
    };
    // the small cost blocks (Eux, Euu, Ex, Eu: m n + m^2 + n + m <= 2 per thread) as plain loads issued before the first
    // product and stored after it: no cp.async instruction for a warp to wait on
    const int n_extra = m * n + m * m + nm;
    auto extra_src = [&](int e, int64_t k) -> const double* {
        if (e < m * n) return Euxg + k * (int64_t)(m * n) + e;
        e -= m * n;
        if (e < m * m) { const int p = e / m, q = e % m; return Euug + k * (int64_t)(m * m) + (p < q ? p * m + q : q * m + p); }
        e -= m * m;
        return e < n ? Exg + k * (int64_t)n + e : Eug + k * (int64_t)m + (e - n);
    };
    auto extra_store = [&](int e, double v) {
        if (e < m * n) { const int x = e / m, u = e % m; G[(n + u) * ldg + x] = v; G[x * ldg + n + u] = v; return; }
        e -= m * n;
        if (e < m * m) { G[(n + e / m) * ldg + n + e % m] = v; return; }
        g[e - m * m] = v;
    };
    if (N > 1 && !(MODE & BP_REG_J)) stage_jacobian(N - 2);

    const int rtiles = n8 / 8, ctiles = q8 / 8, ksteps = n4 / 4;
    const int fr = lane >> 2, fc = lane & 3;          // fragment coordinates of this lane
    // tiles of the second product dealt to this warp: the lower triangle row by row, then the tiles above the diagonal in
    // the strip jtc that holds the gradient column cs
    const int jtc = cs >> 3, nlow = ctiles * (ctiles + 1) / 2;
    const uint32_t owner = (uint32_t)jtc % CL;        // the CTA whose strips hold Qxu and the gradient
    int ntile = nlow + jtc;
    if (CL > 1) {
        ntile = 0;
        for (int jt = (int)rank; jt < ctiles; jt += CL) ntile += ctiles - jt + (jt == jtc ? jtc : 0);
    }
    // every warp runs GC = ceil(ntile / NW) slots without predicates (GC is a constant in the shape-specialised kernels); a
    // warp that was dealt one tile fewer repeats its last one (same values stored twice), a warp without tiles sits out
    const int gt0 = ntile * wid / NW, gcnt = ntile * (wid + 1) / NW - gt0;
    constexpr int GC_CL = bp_cluster_gslots(NN, MM, CL, NW);
    const int GC = CL > 1 ? GC_CL : (ntile + NW - 1) / NW;
    int gt_r[BP_GSLOTS], gt_c[BP_GSLOTS];
    const double* gt_a[BP_GSLOTS];
    const double* gt_b[BP_GSLOTS];
#pragma unroll
    for (int s_ = 0; s_ < BP_GSLOTS; ++s_) {
        int idx = gt0 + (s_ < gcnt ? s_ : gcnt - 1), t = 0, jt = 0;
        if (idx < 0) idx = 0;
        if (CL > 1) {                                   // own strips in turn: the lower tiles, then (strip jtc) the upper ones
            for (jt = (int)rank; jt < ctiles; jt += CL) {
                const int low = ctiles - jt, cnt = low + (jt == jtc ? jtc : 0);
                if (idx < cnt) { t = idx < low ? jt + idx : idx - low; break; }
                idx -= cnt;
            }
            if (jt >= ctiles) { jt = (int)rank; t = jt; }
        } else if (idx >= nlow) { t = idx - nlow; jt = jtc; }
        else { while (idx > t) { idx -= t + 1; ++t; } jt = idx; }
        gt_r[s_] = t; gt_c[s_] = jt;
        gt_a[s_] = AB + (8 * t + fr) * lda + fc;
        gt_b[s_] = (8 * jt + fr == cs) ? sv + fc : T + (8 * jt + fr) * ld + fc;    // column cs of the right factor is s
    }
    long long bad = -1;
#ifdef RBQ_BP_PROFILE
    long long tprev = clock64();
#endif
    // cluster form: where S, s, K, Quu K, Qxu, d and the verdict live in the other CTAs
    uint32_t rS[CL], rsv[CL], rKk[CL], rQuuK[CL], rHq[CL], rfail[CL], rbar[CL];
    uint32_t inbound = 0, xphase = 0;                               // bytes of S and s the other CTAs send per knot
    if (CL > 1) {
#pragma unroll
        for (int c = 0; c < CL; ++c) {
            rS[c] = cluster_map(S, c); rsv[c] = cluster_map(sv, c); rKk[c] = cluster_map(Kk, c); rQuuK[c] = cluster_map(QuuK, c);
            rHq[c] = cluster_map(Hq, c); rfail[c] = cluster_map(&fail_s, c); rbar[c] = cluster_map(&bar_x, c);
        }
        for (int j = 0; j < n; ++j)
            if ((uint32_t)(j >> 3) % CL != rank) inbound += (uint32_t)((n8 - (j & ~1)) * sizeof(double));
        if (rank != owner) inbound += (uint32_t)(n8 * sizeof(double));
        cluster_arrive(); cluster_wait();                           // every CTA's shared memory is initialised
    }
    // BP_REG: this thread's elements of a knot's Jacobian block (pa) and cost blocks (pc): where they come from (element of knot
    // 0 + a per-knot stride) and where they go in shared memory
    constexpr bool reg_j = (MODE & BP_REG_J) != 0, reg_c = (MODE & BP_REG_C) != 0;
    constexpr int PA = !reg_j ? 1 : (NN ? bp_reg_count(NN * (NN + (MM ? MM : BP_MAX_M)), BP_THREADS) : 3);
    constexpr int PC = !reg_c ? 1 : (NN ? bp_reg_count((NN + (MM ? MM : BP_MAX_M)) * (NN + (MM ? MM : BP_MAX_M) + 1), BP_THREADS) : 4);
    double pa[PA], pc[PC];
    int pa_dst[PA], pc_dst[PC], pc_str[PC];
    const double* pc_src[PC];
#pragma unroll
    for (int r = 0; r < PA; ++r) {
        const int idx = tid + r * BP_THREADS;
        pa_dst[r] = reg_j && idx < n * nm ? (idx / n) * lda + idx % n : -1;
        pa[r] = 0.0;
    }
    {
#pragma unroll
        for (int r = 0; r < PC; ++r) {
            const int e = tid + r * BP_THREADS;
            pc[r] = 0.0; pc_dst[r] = -1; pc_str[r] = 0; pc_src[r] = Exg;
            if (!reg_c) continue;
            if (e < nm * nm) {                                            // G(i, j), both triangles from the stored lower one
                const int i = e % nm, j = e / nm, lo = i < j ? i : j, hi = i < j ? j : i;
                pc_dst[r] = j * ldg + i;
                if (hi < n) { pc_src[r] = Exxg + lo * n + hi; pc_str[r] = n * n; }
                else if (lo < n) { pc_src[r] = Euxg + lo * m + (hi - n); pc_str[r] = m * n; }
                else { pc_src[r] = Euug + (lo - n) * m + (hi - n); pc_str[r] = m * m; }
            } else if (e < nm * nm + nm) {                                // the gradient column
                const int i = e - nm * nm;
                pc_dst[r] = cs * ldg + i;
                if (i < n) { pc_src[r] = Exg + i; pc_str[r] = n; } else { pc_src[r] = Eug + (i - n); pc_str[r] = m; }
            }
        }
    }
    auto reg_fetch = [&](int64_t kk) {
        const double* src = ABg + kk * (int64_t)(n * nm);
#pragma unroll
        for (int r = 0; r < PA; ++r) if (pa_dst[r] >= 0) pa[r] = src[tid + r * BP_THREADS];
#pragma unroll
        for (int r = 0; r < PC; ++r) if (pc_dst[r] >= 0) pc[r] = pc_src[r][kk * (int64_t)pc_str[r]];
    };
    auto reg_commit_jacobian = [&]() {
#pragma unroll
        for (int r = 0; r < PA; ++r) if (pa_dst[r] >= 0) AB[pa_dst[r]] = pa[r];
    };
    auto reg_commit_cost = [&]() {
#pragma unroll
        for (int r = 0; r < PC; ++r) if (pc_dst[r] >= 0) G[pc_dst[r]] = pc[r];
    };
    if ((reg_j || reg_c) && N > 1) { reg_fetch(N - 2); reg_commit_jacobian(); }   // (the cost blocks wait in pc for the first knot's barrier)
    // cluster form: the tiles of S this warp mirrors at the top of a knot (source and destination offsets of its two elements)
    constexpr int MIR_RT = ((NN + 7) & ~7) / 8, MIR_NT = MIR_RT * (MIR_RT + 1) / 2, MIR_PER = CL > 1 ? (MIR_NT + NW - 1) / NW : 0;
    int mir_src[MIR_PER > 0 ? MIR_PER : 1], mir_dst[MIR_PER > 0 ? MIR_PER : 1][2];
#pragma unroll
    for (int u = 0; u < MIR_PER; ++u) {
        int idx = wid + NW * u, ti = 0;
        if (idx >= MIR_NT) idx = MIR_NT - 1;                        // (a repeated tile rewrites the same values)
        while (idx > ti) { idx -= ti + 1; ++ti; }                   // tile row ti >= tile column idx
        mir_src[u] = (8 * idx + (lane >> 3)) * ld + 8 * ti + (lane & 7);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int r = 8 * ti + (lane & 7), c = 8 * idx + (lane >> 3) + 4 * h2;
            mir_dst[u][h2] = (r > c && r < n) ? r * ld + c : -1;    // (S has n4 columns: nothing beyond row n is mirrored)
        }
    }
    uint32_t phase = 0;
    for (int64_t k = N - 2; k >= 0; --k, phase ^= 1u) {
        cp_async_wait_all();
        if (tma_j || dense_j) mbar_wait(&bar_s[0], phase);
        if constexpr (CL > 1) {
            // the CTAs exchange the lower triangle of S column by column (bulk copies between shared memories, landing on
            // bar_x) and each mirrors it into the upper one itself, tile by tile (lane -> 8 rows x 4 columns: 2-way conflicts)
            if (k < N - 2) { mbar_wait(&bar_x, xphase); xphase ^= 1u; }
            BP_PROF(11)
            double mv[MIR_PER > 0 ? MIR_PER : 1][2];
#pragma unroll
            for (int u = 0; u < MIR_PER; ++u)
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) mv[u][h2] = S[mir_src[u] + 4 * h2 * ld];
#pragma unroll
            for (int u = 0; u < MIR_PER; ++u)
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
                    if (mir_dst[u][h2] >= 0) S[mir_dst[u][h2]] = mv[u][h2];
            __syncthreads();
        }
        else bp_sync<NW>();                                        // AB_k, S, sv visible; G and g are free again
        BP_PROF(0)
        if (reg_c) reg_commit_cost(); else stage_cost(k);
        if ((reg_j || reg_c) && k > 0) reg_fetch(k - 1);
        double ex0 = 0.0, ex1 = 0.0;
        if (tma_c) {
            if (tid < n_extra) ex0 = *extra_src(tid, k);
            if (tid + BP_THREADS < n_extra) ex1 = *extra_src(tid + BP_THREADS, k);
        }
        // ---- T = S [A B]: warp -> one 8-column strip, all row tiles; A fragments from S (symmetric), B fragments from [A B] ----
        if constexpr (CL > 1) {
            // own strips only; the warps left over split a strip's row tiles between them
            constexpr int CT = ((NN + MM + 8) & ~7) / 8, RT = ((NN + 7) & ~7) / 8, OWN = (CT + CL - 1) / CL;
            constexpr int RG = NW / OWN >= 1 ? NW / OWN : 1, TPI = (RT + RG - 1) / RG;
            for (int item = wid; item < OWN * RG; item += NW) {
                const int ct = (int)rank + CL * (item / RG), t0 = (item % RG) * TPI;
                if (ct >= CT) continue;
                // a warp has only TPI tiles: two accumulator chains per tile (even and odd contraction steps), or the
                // dependent DMMAs of one chain would set the pace
                double acc[2][TPI][2];
#pragma unroll
                for (int t = 0; t < TPI; ++t) { acc[0][t][0] = 0.0; acc[0][t][1] = 0.0; acc[1][t][0] = 0.0; acc[1][t][1] = 0.0; }
                const double* bp = AB + (8 * ct + fr) * lda + fc;
                const double* ap = S + fc * ld + fr + 8 * t0;
                for (int ks = 0; ks < ksteps; ks += 2) {
                    const double bf0 = bp[4 * ks];
#pragma unroll
                    for (int t = 0; t < TPI; ++t)
                        if (t0 + t < RT) dmma884(acc[0][t][0], acc[0][t][1], ap[4 * ks * ld + 8 * t], bf0);
                    if (ks + 1 < ksteps) {
                        const double bf1 = bp[4 * ks + 4];
#pragma unroll
                        for (int t = 0; t < TPI; ++t)
                            if (t0 + t < RT) dmma884(acc[1][t][0], acc[1][t][1], ap[(4 * ks + 4) * ld + 8 * t], bf1);
                    }
                }
#pragma unroll
                for (int t = 0; t < TPI; ++t)
                    if (t0 + t < RT) {
                        T[(8 * ct + 2 * fc) * ld + 8 * (t0 + t) + fr] = acc[0][t][0] + acc[1][t][0];
                        T[(8 * ct + 2 * fc + 1) * ld + 8 * (t0 + t) + fr] = acc[0][t][1] + acc[1][t][1];
                    }
            }
        } else
        for (int ct = wid; ct < ctiles; ct += NW) {
            double acc[BP_MAX_TILES - 1][2];
#pragma unroll
            for (int t = 0; t < BP_MAX_TILES - 1; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
            const double* bp = AB + (8 * ct + fr) * lda + fc;
            const double* ap = S + fc * ld + fr;
            for (int ks = 0; ks < ksteps; ++ks) {
                const double bf = bp[4 * ks];
#pragma unroll
                for (int t = 0; t < BP_MAX_TILES - 1; ++t)
                    if (t < rtiles) dmma884(acc[t][0], acc[t][1], ap[4 * ks * ld + 8 * t], bf);
            }
#pragma unroll
            for (int t = 0; t < BP_MAX_TILES - 1; ++t)
                if (t < rtiles) {
                    T[(8 * ct + 2 * fc) * ld + 8 * t + fr] = acc[t][0];
                    T[(8 * ct + 2 * fc + 1) * ld + 8 * t + fr] = acc[t][1];
                }
        }
        cp_async_wait_all();                                       // this thread's share of the cost blocks has landed ...
        if (tma_c) {
            if (tid < n_extra) extra_store(tid, ex0);
            if (tid + BP_THREADS < n_extra) extra_store(tid + BP_THREADS, ex1);
            mbar_wait(&bar_s[1], phase);
        }
        bp_sync<NW>();                                             // ... and everybody's, together with T
        BP_PROF(1)
        // ---- G += [A B]' T: G is symmetric, so only its lower tiles are computed and mirrored into the upper half (bitwise
        // symmetric there), plus the upper tiles of the strip that holds the gradient column; the tiles were dealt evenly
        // over the warps before the loop (gt_r / gt_c).  A fragments are columns of [A B] read along the contraction. ----
        if (gcnt > 0) {
            double acc[BP_GSLOTS][2];
#pragma unroll
            for (int s_ = 0; s_ < BP_GSLOTS; ++s_)
                if (s_ < GC) {
                    const int r = 8 * gt_r[s_] + fr, c = 8 * gt_c[s_] + 2 * fc;    // one load each, the address selected
                    const bool dq = tma_c && r < n;
                    const double* s0 = dq && c < n ? Qk + c * n + r : G + c * ldg + r;
                    const double* s1 = dq && c + 1 < n ? Qk + (c + 1) * n + r : G + (c + 1) * ldg + r;
                    acc[s_][0] = *s0; acc[s_][1] = *s1;
                }
            if (CL > 1) {
                double acc2[BP_GSLOTS][2];                                         // second chain: the odd contraction steps
#pragma unroll
                for (int s_ = 0; s_ < BP_GSLOTS; ++s_) { acc2[s_][0] = 0.0; acc2[s_][1] = 0.0; }
                for (int ks = 0; ks < ksteps; ks += 2) {
#pragma unroll
                    for (int s_ = 0; s_ < BP_GSLOTS; ++s_)
                        if (s_ < GC) dmma884(acc[s_][0], acc[s_][1], gt_a[s_][4 * ks], gt_b[s_][4 * ks]);
                    if (ks + 1 < ksteps) {
#pragma unroll
                        for (int s_ = 0; s_ < BP_GSLOTS; ++s_)
                            if (s_ < GC) dmma884(acc2[s_][0], acc2[s_][1], gt_a[s_][4 * ks + 4], gt_b[s_][4 * ks + 4]);
                    }
                }
#pragma unroll
                for (int s_ = 0; s_ < BP_GSLOTS; ++s_)
                    if (s_ < GC) { acc[s_][0] += acc2[s_][0]; acc[s_][1] += acc2[s_][1]; }
            } else
            for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
                for (int s_ = 0; s_ < BP_GSLOTS; ++s_)
                    if (s_ < GC) dmma884(acc[s_][0], acc[s_][1], gt_a[s_][4 * ks], gt_b[s_][4 * ks]);
            }
#pragma unroll
            for (int s_ = 0; s_ < BP_GSLOTS; ++s_)
                if (s_ < GC) {
                    const int r = 8 * gt_r[s_] + fr, c = 8 * gt_c[s_] + 2 * fc;
                    G[c * ldg + r] = acc[s_][0];
                    G[(c + 1) * ldg + r] = acc[s_][1];
                    if (gt_r[s_] > gt_c[s_] && gt_r[s_] != jtc)                     // mirror (c even, ldg a multiple of 4: 16-byte aligned)
                        *reinterpret_cast<double2*>(G + r * ldg + c) = make_double2(acc[s_][0], acc[s_][1]);
                }
        }
        BP_PROF(2)
        BP_PROF(3)
        if (CL == 1 || rank == owner) bp_sync<NW>();               // (cluster form: only the owner's gains read G here)
        // cluster form: the CTAs that wait for the gains ask for the next Jacobian block meanwhile (the owner after the
        // barrier: issuing the copies before its gains delays the barrier for everybody, 3.67 -> 3.84 us per knot at n = 56)
        const bool early = CL > 1 && rank != owner;
        BP_PROF(4)
        // ---- gains: every thread factors the (tiny) regularised Quu itself, thread j solves for column j of K, thread n for d ----
        bool ok = true;
        if (CL > 1 && rank != owner) {
            // the gains arrive from the owner; meanwhile fetch the next Jacobian block (this CTA is done reading the current one)
            __syncthreads();
            if (reg_j) { if (k > 0) reg_commit_jacobian(); }
            else if (k > 0) stage_jacobian(k - 1); else cp_async_commit();
        } else if (m == 1 && tid <= n) {
            // one control (every reference model but the time-optimal spin15): Quu_reg is a scalar, its "factorisation" one
            // reciprocal -- a much shorter dependent chain than the general path's sqrt, divisions and guarded loops
            const int j = tid;
            const double quu = G[n * ldg + n];
            double quu_reg = quu + a.rho, bb = 0.0;
            if (a.reg_type == RBQ_REG_STATE) {
                for (int r = 0; r < n; ++r) bb = fma(AB[n * lda + r], AB[n * lda + r], bb);
                quu_reg = fma(a.rho, bb, quu);
            }
            ok = quu_reg > 0.0;                                                     // also catches NaN
            double rhs = j < n ? G[n * ldg + j] : g[n];                             // Qxu(j): the coalesced half of the symmetric G
            if (j < n && a.reg_type == RBQ_REG_STATE) {
                double ba = 0.0;
                for (int r = 0; r < n; ++r) ba = fma(AB[n * lda + r], AB[j * lda + r], ba);
                rhs = fma(a.rho, ba, rhs);
            }
            // the reference solves through the Cholesky factor: x = (rhs / l) / l with l = sqrt(Quu_reg); one division by
            // Quu_reg differs from that by rounding only
            const double x = -rhs / quu_reg;
            if (ok) {
                if (j < n) { Kk[j] = x; Kg[k * (int64_t)n + j] = x; QuuK[j] = quu * x; }
                else { dk[0] = x; dg[k] = x; }
            }
            if (CL > 1) {
                const double qxu = j < n ? G[n * ldg + j] : 0.0;
#pragma unroll
                for (int c = 0; c < CL; ++c) {
                    if (c == (int)rank) continue;
                    if (!ok) st_cluster(rfail[c], 1);
                    else if (j < n) { st_cluster(rKk[c] + 8 * j, x); st_cluster(rQuuK[c] + 8 * j, quu * x); st_cluster(rHq[c] + 8 * j, qxu); }
                }
                if (!ok) fail_s = 1;
            }
        } else if (tid <= n) {
            const int j = tid;
            double L[BP_MAX_M][BP_MAX_M], rhs[BP_MAX_M];
#pragma unroll
            for (int p = 0; p < BP_MAX_M; ++p)
#pragma unroll
                for (int q = 0; q < BP_MAX_M; ++q) {
                    double v = 0.0;
                    if (p < m && q <= p) {
                        v = G[(n + q) * ldg + (n + p)];
                        if (a.reg_type == RBQ_REG_STATE) {
                            double bb = 0.0;
                            for (int r = 0; r < n; ++r) bb = fma(AB[(n + p) * lda + r], AB[(n + q) * lda + r], bb);
                            v = fma(a.rho, bb, v);
                        } else if (p == q) v += a.rho;
                    }
                    L[p][q] = v;
                }
#pragma unroll
            for (int c = 0; c < BP_MAX_M; ++c) {                                   // in-place Cholesky (lower), reciprocal pivots
                if (c >= m) continue;
                double dj = L[c][c];
#pragma unroll
                for (int r = 0; r < BP_MAX_M; ++r) if (r < c) dj -= L[c][r] * L[c][r];
                if (!(dj > 0.0)) ok = false;                                        // also catches NaN
                L[c][c] = 1.0 / sqrt(dj);
#pragma unroll
                for (int i = 0; i < BP_MAX_M; ++i) {
                    if (i <= c || i >= m) continue;
                    double v = L[i][c];
#pragma unroll
                    for (int r = 0; r < BP_MAX_M; ++r) if (r < c) v -= L[i][r] * L[c][r];
                    L[i][c] = v * L[c][c];
                }
            }
#pragma unroll
            for (int p = 0; p < BP_MAX_M; ++p) {
                double v = 0.0;
                if (p < m) {
                    v = j < n ? G[(n + p) * ldg + j] : g[n + p];     // Qxu(j, p): the coalesced half of the symmetric G
                    if (j < n && a.reg_type == RBQ_REG_STATE) {
                        double ba = 0.0;
                        for (int r = 0; r < n; ++r) ba = fma(AB[(n + p) * lda + r], AB[j * lda + r], ba);
                        v = fma(a.rho, ba, v);
                    }
                }
                rhs[p] = v;
            }
#pragma unroll
            for (int p = 0; p < BP_MAX_M; ++p) {                                  // L y = rhs
                if (p >= m) continue;
                double v = rhs[p];
#pragma unroll
                for (int r = 0; r < BP_MAX_M; ++r) if (r < p) v -= L[p][r] * rhs[r];
                rhs[p] = v * L[p][p];
            }
#pragma unroll
            for (int p = BP_MAX_M - 1; p >= 0; --p) {                             // L' x = y
                if (p >= m) continue;
                double v = rhs[p];
#pragma unroll
                for (int r = 0; r < BP_MAX_M; ++r) if (r > p && r < m) v -= L[r][p] * rhs[r];
                rhs[p] = v * L[p][p];
            }
            if (ok) {
#pragma unroll
                for (int p = 0; p < BP_MAX_M; ++p) {
                    if (p >= m) continue;
                    if (j < n) { Kk[j * m + p] = -rhs[p]; Kg[k * (int64_t)(m * n) + j * m + p] = -rhs[p]; }
                    else { dk[p] = -rhs[p]; dg[k * (int64_t)m + p] = -rhs[p]; }
                }
                if (j < n) {                                                       // Quu K (unregularised Quu), column j
#pragma unroll
                    for (int p = 0; p < BP_MAX_M; ++p) {
                        if (p >= m) continue;
                        double v = 0.0;
#pragma unroll
                        for (int q = 0; q < BP_MAX_M; ++q) if (q < m) v = fma(G[(n + q) * ldg + (n + p)], -rhs[q], v);
                        QuuK[j * m + p] = v;
                    }
                }
            }
        }
        BP_PROF(5)
        if (CL > 1) {
            cluster_arrive(); cluster_wait();                      // gains (or the verdict) visible in every CTA; G complete
            if (fail_s) {
                bad = k;
                // (the other CTAs already asked for the next Jacobian block: let it land before leaving)
                if (early && k > 0 && (tma_j || dense_j)) mbar_wait(&bar_s[0], phase ^ 1u);
                break;
            }
        } else
        if (bp_sync_or<NW>(!ok)) { bad = k; break; }            // Quu_reg not positive definite: stop here (uniform)
        BP_PROF(6)
        // the Jacobian block is no longer needed: fetch the next one behind the value update
        if (reg_j) { if (k > 0 && !early) reg_commit_jacobian(); }
        else if (!early) { if (k > 0) stage_jacobian(k - 1); else cp_async_commit(); }
        BP_PROF(8)
        // ---- value update: V = Qxx + sym(K'Quu K + K'Qux + Qux'K) into the scratch, then S = (V + V')/2 ----
        // S = sym(Qxx) + sym(K'Quu K + K'Qux + Qux'K), written in one pass.  The symmetrisation is not cosmetic: the recursion
        // contracts symmetric errors through the closed loop (A - BK) but maps an antisymmetric part as A' . A, which grows
        // whenever rho(A) > 1 -- without (S + S')/2 the rounding-level asymmetry of Qxx wrecks a 180-knot pass.  Lane -> row i (its K, Quu K and Qux entries stay in registers), the
        // warp walks the columns j (whose entries are warp-uniform broadcasts).
        if constexpr (CL > 1) {
            // entries on and below the diagonal of the own column strips: lane -> row, warp -> column.  They are sent to the other
            // CTAs below; the upper triangle is mirrored by each CTA at the top of the next knot
            const double* hq = rank == owner ? G + n * ldg : Hq;
            constexpr int CT = ((NN + MM + 8) & ~7) / 8, OWN = (CT + CL - 1) / CL;
            double ki[2], qi[2], hi[2];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int i = lane + 32 * h2;
                const bool on = i < n;
                ki[h2] = on ? Kk[i] : 0.0; qi[h2] = on ? 0.5 * QuuK[i] : 0.0; hi[h2] = on ? hq[i] : 0.0;
            }
            for (int jj = wid; jj < 8 * OWN; jj += NW) {
                const int j = 8 * ((int)rank + CL * (jj >> 3)) + (jj & 7);
                if (j >= n) continue;
                const double cj = fma(0.5, QuuK[j], hq[j]), kj = Kk[j];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int i = lane + 32 * h2;
                    if (i < j || i >= n) continue;
                    // inside a diagonal tile the two triangles were computed separately: average them; below it G(i, j) is the
                    // value both triangles share (its mirror need not exist here: the strip of the gradient column is not mirrored)
                    double v = G[j * ldg + i];
                    if ((i >> 3) == (j >> 3)) v = 0.5 * (v + G[i * ldg + j]);
                    v = fma(ki[h2], cj, v);                                    // K'(Quu K)/2 + K'Qux
                    v = fma(kj, qi[h2] + hi[h2], v);                           // (Quu K)'K/2 + Qux'K
                    S[j * ld + i] = v;
                    if ((i >> 3) == (j >> 3)) S[i * ld + j] = v;               // (a column is sent from an even row: keep the diagonal tile whole)
                }
            }
        } else
        for (int i = lane; i < n; i += 32) {
            double ki[BP_MAX_M], qi[BP_MAX_M], hi[BP_MAX_M];
#pragma unroll
            for (int p = 0; p < BP_MAX_M; ++p) {
                const bool on = p < m;
                ki[p] = on ? Kk[i * m + p] : 0.0; qi[p] = on ? 0.5 * QuuK[i * m + p] : 0.0; hi[p] = on ? G[(n + p) * ldg + i] : 0.0;
            }
#pragma unroll 2
            for (int j = wid; j < n; j += NW) {
                double v = 0.5 * (G[j * ldg + i] + G[i * ldg + j]);      // the transposed read costs a 4-way bank conflict, no barrier
#pragma unroll
                for (int p = 0; p < BP_MAX_M; ++p) {
                    if (p >= m) continue;
                    const double kj = Kk[j * m + p], hj = G[(n + p) * ldg + j];
                    v = fma(ki[p], fma(0.5, QuuK[j * m + p], hj), v);          // K'(Quu K)/2 + K'Qux
                    v = fma(kj, qi[p] + hi[p], v);                             // (Quu K)'K/2 + Qux'K
                }
                S[j * ld + i] = v;
            }
        }
        BP_PROF(9)
        if (CL == 1 || rank == owner)
        for (int i = BP_THREADS - 1 - tid; i < n; i += BP_THREADS) {   // dealt from the last thread down: the first warps carry the tail of the pass above
            double v = g[i];
            for (int p = 0; p < m; ++p) {
                double qd = g[n + p];                                      // Qu + Quu d
                for (int q = 0; q < m; ++q) qd = fma(G[(n + q) * ldg + (n + p)], dk[q], qd);
                v = fma(Kk[i * m + p], qd, v);
                v = fma(G[(n + p) * ldg + i], dk[p], v);                   // Qux' d
            }
            sv[i] = v;                                                    // sv is not read again before the next knot
        }
        if (tid == BP_THREADS - 1 && (CL == 1 || rank == owner)) {
            double d0 = 0.0, d1 = 0.0;
            for (int p = 0; p < m; ++p) {
                d0 = fma(dk[p], g[n + p], d0);
                double qd = 0.0;
                for (int q = 0; q < m; ++q) qd = fma(G[(n + q) * ldg + (n + p)], dk[q], qd);
                d1 = fma(dk[p], qd, d1);
            }
            dv_s[0] += d0; dv_s[1] += 0.5 * d1;
        }
        BP_PROF(7)
        // the barrier at the top of the loop orders these writes against the next knot's reads
        if (CL > 1 && k > 0) {
            // send the own columns of S (from the even row at or above the diagonal down to the padding) and, from the owner,
            // s to every other CTA: one bulk copy each, dealt over the lanes of all warps; the receivers wait on their bar_x
            fence_async_smem();
            __syncthreads();
            if (tid == 0) mbar_expect_tx(&bar_x, inbound);
            constexpr int CT = ((NN + MM + 8) & ~7) / 8, OWN = (CT + CL - 1) / CL;
            const int ncopy = (8 * OWN + (rank == owner ? 1 : 0)) * (CL - 1);
            for (int cidx = wid + NW * lane; cidx < ncopy; cidx += NW * 32) {
                const int jj = cidx / (CL - 1), pc = cidx % (CL - 1), peer = pc + (pc >= (int)rank ? 1 : 0);
                if (jj == 8 * OWN) { bulk_s2peer(rsv[peer], sv, (uint32_t)(n8 * sizeof(double)), rbar[peer]); continue; }
                const int j = 8 * ((int)rank + CL * (jj >> 3)) + (jj & 7);
                if (j >= n) continue;
                const int off = j * ld + (j & ~1);
                bulk_s2peer(rS[peer] + 8 * off, S + off, (uint32_t)((n8 - (j & ~1)) * sizeof(double)), rbar[peer]);
            }
        }
    }
    cp_async_wait_all();
    bp_sync<NW>();
    if (tid == BP_THREADS - 1 && (CL == 1 || rank == owner)) {
        a.dV[2 * b] = dv_s[0]; a.dV[2 * b + 1] = dv_s[1];
        a.fail[b] = bad;
    }
}

size_t backward_pass_smem(int n, int m) {
    const size_t n4 = (size_t)((n + 3) & ~3), n8 = (size_t)((n + 7) & ~7), q8 = (size_t)((n + m + 8) & ~7), ld = n8 + 4, ldg = q8 + 4;
    return sizeof(double) * (ld * n4 + n8 + 2 * ld * q8 + ldg * q8 + 2 * BP_MAX_M * n8 + BP_MAX_M + n8 + (size_t)n * n + 2);
}

// staging mode for a problem shape and its buffers (see backward_pass_kernel); RBQ_BP_TMA masks the bits (0 = cp.async only)
static int bp_stage_mode(const rbq_handle* h, int n, int m, const double* d_AB, const double* d_Exx) {
    const int nm = n + m;
    int mode = 0;
    if (n >= 28) {
        const bool ab16 = (reinterpret_cast<uintptr_t>(d_AB) & 15) == 0, exx16 = (reinterpret_cast<uintptr_t>(d_Exx) & 15) == 0;
        if (ab16 && ((n * nm) & 1) == 0 && ((n & 1) || (n & 7) == 4)) mode |= BP_DENSE_J;   // every block 16-byte aligned
        // (BP_REG_J for the even shapes n = 56 / 58 was measured: 5.75 -> 5.81 us per knot on one CTA, 3.67 -> 3.76 on four: the
        // 13 loads + 13 stores per thread cost what the 57 bulk copies cost)
        else if (ab16 && (n & 1) == 0) mode |= BP_TMA_J;                                      // every column 16-byte aligned
        if (exx16 && m * n + m * m + n + m <= 2 * 32 * BP_WARPS_LARGE) mode |= BP_TMA_C;
    }
    else if (n * nm <= 3 * 32 * BP_WARPS_LARGE && nm * (nm + 1) <= 4 * 32 * BP_WARPS_LARGE) mode = BP_REG;
    if (h->opt.bp_tma >= 0) mode &= h->opt.bp_tma;
    return mode;
}

template <int NW, int MODE, int MM, int NN, int CL = 1>
static cudaError_t bp_launch2(const BackwardArgs& a, int64_t B, size_t smem, cudaStream_t stream) {
    auto kern = backward_pass_kernel<NW, MODE, MM, NN, CL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if constexpr (CL == 1) {
        kern<<<(unsigned)B, 32 * NW, smem, stream>>>(a);
        return cudaGetLastError();
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * CL)); cfg.blockDim = dim3(32 * NW); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, a);
    }
}
// one problem (or a few) of a reference shape with one control: the cluster form, CL CTAs per problem
template <int CL>
static cudaError_t bp_launch_cluster(const BackwardArgs& a, int64_t B, size_t smem, cudaStream_t stream, int mode, bool* done) {
    *done = true;
    if (mode == (BP_DENSE_J | BP_TMA_C) && a.n == 43) return bp_launch2<BP_WARPS_LARGE, BP_DENSE_J | BP_TMA_C, 1, 43, CL>(a, B, smem, stream);
    if (mode == (BP_TMA_J | BP_TMA_C) && a.n == 56) return bp_launch2<BP_WARPS_LARGE, BP_TMA_J | BP_TMA_C, 1, 56, CL>(a, B, smem, stream);
    if (mode == (BP_TMA_J | BP_TMA_C) && a.n == 58) return bp_launch2<BP_WARPS_LARGE, BP_TMA_J | BP_TMA_C, 1, 58, CL>(a, B, smem, stream);
    *done = false;
    return cudaSuccess;
}
// the shapes of the reference models (n = 11, 12, 15, 19, 43, 56, 58 with one control) get kernels with n known at compile time: tile
// loops fully unrolled, constant strides.  RBQ_BP_GENERIC=1 forces the any-shape kernels (tests compare the two).
template <int NW, int MODE>
static cudaError_t bp_launch(const BackwardArgs& a, int64_t B, size_t smem, cudaStream_t stream, bool specialised) {
    if (a.m == 2 && specialised) {                                      // spin15 time-optimal: u = (d2a/dt2, sqrt(dt))
        if constexpr (MODE == 0 || MODE == BP_REG) {
            if (a.n == 11) return bp_launch2<NW, MODE, 2, 11>(a, B, smem, stream);
            if (a.n == 12) return bp_launch2<NW, MODE, 2, 12>(a, B, smem, stream);
        }
    }
    if (a.m != 1) {
        if constexpr (NW == BP_WARPS_SMALL && MODE == BP_REG) return bp_launch2<NW, 0, 0, 0>(a, B, smem, stream);
        else return bp_launch2<NW, MODE, 0, 0>(a, B, smem, stream);
    }
    if (specialised) {
        if constexpr (MODE == 0 || MODE == BP_REG) {
            if (a.n == 11) return bp_launch2<NW, MODE, 1, 11>(a, B, smem, stream);
            if (a.n == 12) return bp_launch2<NW, MODE, 1, 12>(a, B, smem, stream);      // spin15, decay aware
            if (a.n == 15) return bp_launch2<NW, MODE, 1, 15>(a, B, smem, stream);      // spin11 / spin17, first order
            if constexpr (NW == BP_WARPS_LARGE) {
                if (a.n == 19) return bp_launch2<NW, MODE, 1, 19>(a, B, smem, stream);  // spin11 / spin17, second order
            }
        }
        if constexpr (NW == BP_WARPS_LARGE && MODE == (BP_DENSE_J | BP_TMA_C)) {
            if (a.n == 43) return bp_launch2<NW, MODE, 1, 43>(a, B, smem, stream);
        }
        if constexpr (NW == BP_WARPS_LARGE && MODE == (BP_TMA_J | BP_TMA_C)) {
            if (a.n == 56) return bp_launch2<NW, MODE, 1, 56>(a, B, smem, stream);
            if (a.n == 58) return bp_launch2<NW, MODE, 1, 58>(a, B, smem, stream);
        }
    }
    if constexpr (NW == BP_WARPS_SMALL && MODE == BP_REG) return bp_launch2<NW, 0, 1, 0>(a, B, smem, stream);   // (register staging of the one-warp form: specialised shapes only)
    else return bp_launch2<NW, MODE, 1, 0>(a, B, smem, stream);
}

int launch_backward_pass(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* d_AB, const double* d_Exx,
                         const double* d_Ex, const double* d_Euu, const double* d_Eux, const double* d_Eu, double rho,
                         int reg_type, double* d_K, double* d_d, double* d_dV, long long* d_fail) {
    const size_t smem = backward_pass_smem(n, m);
    if (smem > 227 * 1024) return RBQ_E_SIZE;
    BackwardArgs a{n, m, N, B, d_AB, d_Exx, d_Ex, d_Euu, d_Eux, d_Eu, rho, reg_type, d_K, d_d, d_dV, d_fail};
    // n <= 16 (spin13, spin15, spin11 up to first order) in large batches: one warp per problem, warp barriers only, eight
    // problems resident per SM.  Measured at n = 11, 1001 knots (tools/proto/bp_batch.py): one problem takes 1.47 ms as one
    // warp against 1.20 ms as eight; 296 problems 1.80 against 1.77 ms; 370 problems 1.81 against 3.41 ms (the eight-warp
    // form runs two problems per SM at a time); 1184 problems 2.05 against 7.0 ms -- so the warp form is used above two
    // problems per SM.  RBQ_BP_WARPS=1 / 8 forces a form (A/B measurements, tests).
    bool small = n <= 16 && B > 2 * (int64_t)h->sm_count;
    if (h->opt.bp_warps > 0) small = n <= 16 && h->opt.bp_warps == 1;
    cudaError_t e;
    // few large problems: split each over a cluster of CTAs (RBQ_BP_CLUSTER = 1 / 2 / 4 forces the size; default: as many
    // CTAs per problem as keeps every problem resident at once, up to 4)
    int cl = 1;
    if (!small && m == 1 && !h->opt.bp_generic && (n == 43 || n == 56 || n == 58)) {
        // (clusters of 4 fit 132 of the 148 SMs: 33 problems at a time)
        cl = h->opt.bp_cluster > 0 ? h->opt.bp_cluster : (4 * B <= h->sm_count - 20 ? 4 : (2 * B <= h->sm_count ? 2 : 1));
    }
    bool done = false;
    if (cl == 2) e = bp_launch_cluster<2>(a, B, smem, h->stream, bp_stage_mode(h, n, m, d_AB, d_Exx), &done);
    else if (cl == 4) e = bp_launch_cluster<4>(a, B, smem, h->stream, bp_stage_mode(h, n, m, d_AB, d_Exx), &done);
    if (done && e != cudaSuccess) { cudaGetLastError(); done = false; }   // a cluster that cannot be placed: one CTA per problem instead
    if (done) {
    } else if (small) {
        if (bp_stage_mode(h, n, m, d_AB, d_Exx) == BP_REG) e = bp_launch<BP_WARPS_SMALL, BP_REG>(a, B, smem, h->stream, !h->opt.bp_generic);
        else e = bp_launch<BP_WARPS_SMALL, 0>(a, B, smem, h->stream, !h->opt.bp_generic);
    } else {
        switch (bp_stage_mode(h, n, m, d_AB, d_Exx)) {
            case BP_DENSE_J | BP_TMA_C: e = bp_launch<BP_WARPS_LARGE, BP_DENSE_J | BP_TMA_C>(a, B, smem, h->stream, !h->opt.bp_generic); break;
            case BP_TMA_J | BP_TMA_C: e = bp_launch<BP_WARPS_LARGE, BP_TMA_J | BP_TMA_C>(a, B, smem, h->stream, !h->opt.bp_generic); break;
            case BP_TMA_C: e = bp_launch<BP_WARPS_LARGE, BP_TMA_C>(a, B, smem, h->stream, !h->opt.bp_generic); break;
            case BP_REG: e = bp_launch<BP_WARPS_LARGE, BP_REG>(a, B, smem, h->stream, !h->opt.bp_generic); break;
            default: e = bp_launch<BP_WARPS_LARGE, 0>(a, B, smem, h->stream, !h->opt.bp_generic); break;
        }
    }
    h->launches++;
    return (int)e;
}

}  // namespace rbq
#ifdef RBQ_BP_PROFILE
extern "C" int rbq_debug_bp_prof(long long* out, int reset) {
    cudaMemcpyFromSymbol(out, rbq::g_bp_prof, sizeof(long long) * 32);
    if (reset) { long long z[32] = {0}; cudaMemcpyToSymbol(rbq::g_bp_prof, z, sizeof(z)); }
    return 0;
}
#endif
