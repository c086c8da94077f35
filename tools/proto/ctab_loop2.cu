// Standalone replica of the constant-memory table step (mc_static_ctab_kernel's hot loop) for scheduling experiments.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DKU=8 -DMODE=0 -o ctab_loop2 ctab_loop2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define RBQ_DEV __device__ __forceinline__
constexpr int CT_MAX = 2000;
constexpr int ST_TILE = 256;
__constant__ double g_ctab[4 * CT_MAX];
#ifndef KU
#define KU 8
#endif
#ifndef MODE
#define MODE 0
#endif
#ifndef MINB
#define MINB 6
#endif
#ifndef DSMEM
#define DSMEM 0      // dynamic shared memory per block: limits the resident blocks per SM (occupancy experiments)
#endif
RBQ_DEV void updates(const double (&tp)[KU], const double (&tq)[KU], double (&v)[4]) {
#pragma unroll
    for (int u = 0; u < KU; ++u) {
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
RBQ_DEV void sweep(int nk, double p, double dq, double dq2, double e0, double (&v)[4], int once, const double* __restrict__ c1_s) {
    asm volatile("" : "+d"(dq2));
    for (int t0 = 0; t0 < nk; t0 += ST_TILE) {
        const int len = nk - t0 < ST_TILE ? nk - t0 : ST_TILE;
        int j = 0;
#if MODE == 0
        for (; j + KU <= len; j += KU) {
            double tp[KU], tq[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const int k = 4 * (t0 + j + u);
                const double eps = fma(g_ctab[k], dq2, e0);
                const double q = g_ctab[k] + dq;
                const double t = fma(eps, eps + g_ctab[k + 1], g_ctab[k + 2]) * g_ctab[k + 3];
                tp[u] = t * p; tq[u] = t * q;
            }
            for (int rep = 0; rep < once; ++rep) updates(tp, tq, v);
        }
#elif MODE == 3
        // Horner form, 14 instructions: q_j, c0, c2 as uniform operands from constant memory, c1 through shared memory (LDS.64)
        for (; j + KU <= len; j += KU) {
            double tp[KU], tq[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const int k = 4 * (t0 + j + u);
                const double eps = fma(g_ctab[k], dq2, e0);
                const double q = g_ctab[k] + dq;
                const double t = fma(fma(g_ctab[k + 3], eps, c1_s[t0 + j + u]), eps, g_ctab[k + 1]);
                tp[u] = t * p; tq[u] = t * q;
            }
            for (int rep = 0; rep < once; ++rep) updates(tp, tq, v);
        }
#elif MODE == 1
        // rows of the next batch are fetched before the updates of this one
        double r[KU][4];
#pragma unroll
        for (int u = 0; u < KU; ++u)
#pragma unroll
            for (int c = 0; c < 4; ++c) r[u][c] = g_ctab[4 * (t0 + u) + c];
        for (; j + KU <= len; j += KU) {
            double tp[KU], tq[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const double eps = fma(r[u][0], dq2, e0);
                const double q = r[u][0] + dq;
                const double t = fma(eps, eps + r[u][1], r[u][2]) * r[u][3];
                tp[u] = t * p; tq[u] = t * q;
            }
            const int nx = t0 + j + KU < nk ? t0 + j + KU : t0 + j;       // (the last batch fetches itself again)
#pragma unroll
            for (int u = 0; u < KU; ++u)
#pragma unroll
                for (int c = 0; c < 4; ++c) r[u][c] = g_ctab[4 * (nx + u) + c];
            for (int rep = 0; rep < once; ++rep) updates(tp, tq, v);
        }
#elif MODE == 2
        // scalars of batch b+1 computed before the updates of batch b are issued (software pipeline over batches)
        double tp[KU], tq[KU];
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            const int k = 4 * (t0 + u);
            const double eps = fma(g_ctab[k], dq2, e0);
            const double q = g_ctab[k] + dq;
            const double t = fma(eps, eps + g_ctab[k + 1], g_ctab[k + 2]) * g_ctab[k + 3];
            tp[u] = t * p; tq[u] = t * q;
        }
        for (; j + KU <= len; j += KU) {
            double np_[KU], nq_[KU];
            const int nx = t0 + j + KU < nk ? t0 + j + KU : t0 + j;
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const int k = 4 * (nx + u);
                const double eps = fma(g_ctab[k], dq2, e0);
                const double q = g_ctab[k] + dq;
                const double t = fma(eps, eps + g_ctab[k + 1], g_ctab[k + 2]) * g_ctab[k + 3];
                np_[u] = t * p; nq_[u] = t * q;
            }
            for (int rep = 0; rep < once; ++rep) updates(tp, tq, v);
#pragma unroll
            for (int u = 0; u < KU; ++u) { tp[u] = np_[u]; tq[u] = nq_[u]; }
        }
#endif
        const double inv = rsqrt(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] *= inv;
    }
}
#ifndef PERSIST
#define PERSIST 0      // 1: one block per resident slot walks several groups of 128 samples (the c1 copy and its barrier once per block)
#endif
__global__ void __launch_bounds__(128, MINB) k(int nk, int once, const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ c1g, int ngroups) {
    __shared__ double c1_s[CT_MAX];
#if MODE == 3
    for (int i = threadIdx.x; i < nk; i += 128) c1_s[i] = c1g[i];
    __syncthreads();
#endif
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int s = grp * 128 + threadIdx.x;
        double v[4] = {1.0, 0.0, 0.0, 0.0};
        sweep(nk, in[4 * s], in[4 * s + 1], in[4 * s + 2], in[4 * s + 3], v, once, c1_s);
        out[s] = v[0] + v[1] + v[2] + v[3];
    }
}
int main() {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, DSMEM);
    const int S = 1 << 20, nk = 1000;
    double *in, *out; cudaMalloc(&in, S * 32); cudaMalloc(&out, S * 8);
    double* h = (double*)malloc(S * 32);
    for (int s = 0; s < S; ++s) { h[4 * s] = 0.0044 * (1 + 1e-3 * (s % 17)); h[4 * s + 1] = 1e-6 * (s % 13); h[4 * s + 2] = 2e-6 * (s % 13); h[4 * s + 3] = 1e-9 * (s % 7); }
    cudaMemcpy(in, h, S * 32, cudaMemcpyHostToDevice);
    static double tab[4 * CT_MAX];
    for (int j = 0; j < CT_MAX; ++j) { tab[4 * j] = 0.01 * (j % 100) / 100; tab[4 * j + 1] = 2.5; tab[4 * j + 2] = 7.5; tab[4 * j + 3] = 0.13333; }
    cudaMemcpyToSymbol(g_ctab, tab, sizeof(tab));
    double* c1g; cudaMalloc(&c1g, CT_MAX * 8);
    { static double c1h[CT_MAX]; for (int j = 0; j < CT_MAX; ++j) c1h[j] = 0.3333; cudaMemcpy(c1g, c1h, sizeof(c1h), cudaMemcpyHostToDevice); }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        k<<<PERSIST ? 148 * PERSIST : S / 128, 128, DSMEM>>>(nk, 1, in, out, c1g, S / 128);
        cudaEventRecord(a);
        for (int i = 0; i < 10; ++i) k<<<PERSIST ? 148 * PERSIST : S / 128, 128, DSMEM>>>(nk, 1, in, out, c1g, S / 128);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10; best = ms < best ? ms : best;
    }
    printf("U=%d MODE=%d MINB=%d DSMEM=%d PERSIST=%d: %.4f ms per 2^20 x %d sweep, %.2f cycles per knot per warp at 1965 MHz\n", KU, MODE, MINB, DSMEM, PERSIST, best, nk,
           best * 1.965e6 / ((double)S / 32 * nk / (148 * 4)));
    return 0;
}
