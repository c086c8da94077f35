#include <cstdio>
#include <cuda_runtime.h>
#define RBQ_DEV __device__ __forceinline__
constexpr int CT_MAX = 2000;
constexpr int ST_TILE = 256;
__constant__ double g_ctab[4 * CT_MAX];
#ifndef KU
#define KU 8
#endif
#ifndef FORM
#define FORM 1
#endif
RBQ_DEV void sweep(int nk, double p, double dq, double dq2, double e0, double (&v)[4], int once) {
    asm volatile("" : "+d"(dq2));
    for (int t0 = 0; t0 < nk; t0 += ST_TILE) {
        const int len = nk - t0 < ST_TILE ? nk - t0 : ST_TILE;
        int j = 0;
        for (; j + KU <= len; j += KU) {
            double tp[KU], tq[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const int k = 4 * (t0 + j + u);
                const double eps = fma(g_ctab[k], dq2, e0);
                const double q = g_ctab[k] + dq;
                double t;
                if (FORM == 0) t = fma(fma(g_ctab[k + 3], eps, g_ctab[k + 2]), eps, g_ctab[k + 1]);
                else t = fma(eps, eps + g_ctab[k + 1], g_ctab[k + 2]) * g_ctab[k + 3];
                tp[u] = t * p; tq[u] = t * q;
            }
            for (int rep = 0; rep < once; ++rep) {
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
        }
        const double inv = rsqrt(fma(v[3], v[3], fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]))));
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] *= inv;
    }
}
__global__ void __launch_bounds__(128, 6) k(int nk, int once, const double* __restrict__ in, double* __restrict__ out) {
    const int s = blockIdx.x * 128 + threadIdx.x;
    double v[4] = {1.0, 0.0, 0.0, 0.0};
    sweep(nk, in[4 * s], in[4 * s + 1], in[4 * s + 2], in[4 * s + 3], v, once);
    out[s] = v[0] + v[1] + v[2] + v[3];
}
int main() {
    const int S = 1 << 20, nk = 1000;
    double *in, *out; cudaMalloc(&in, S * 32); cudaMalloc(&out, S * 8);
    double* h = (double*)malloc(S * 32);
    for (int s = 0; s < S; ++s) { h[4 * s] = 0.0044 * (1 + 1e-3 * (s % 17)); h[4 * s + 1] = 1e-6 * (s % 13); h[4 * s + 2] = 2e-6 * (s % 13); h[4 * s + 3] = 1e-9 * (s % 7); }
    cudaMemcpy(in, h, S * 32, cudaMemcpyHostToDevice);
    static double tab[4 * CT_MAX];
    for (int j = 0; j < CT_MAX; ++j) { tab[4 * j] = 0.01 * (j % 100) / 100; tab[4 * j + 1] = FORM ? 2.5 : 1.0; tab[4 * j + 2] = FORM ? 7.5 : 0.3333; tab[4 * j + 3] = 0.13333; }
    cudaMemcpyToSymbol(g_ctab, tab, sizeof(tab));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
        k<<<S / 128, 128>>>(nk, 1, in, out);
        cudaEventRecord(a);
        for (int i = 0; i < 10; ++i) k<<<S / 128, 128>>>(nk, 1, in, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
        // warp-steps per sub-partition: S/32 warps * nk / (148*4)
        printf("U=%d FORM=%d: %.4f ms per 2^20 x %d sweep, %.2f cycles per knot per warp at 1965 MHz (6 warps per sub-partition)\n", KU, FORM, ms, nk,
               ms * 1.965e6 / ((double)S / 32 * nk / (148 * 4)));
    }
    return 0;
}
