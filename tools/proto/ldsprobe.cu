// Does shared-memory load traffic into the register file slow the FP64 pipe's operand delivery?  (B200)
// Stream of two-fetched-operand DFMAs (full rate on their own); one 16-byte warp-uniform load per 7 DFMAs, either LDS.128 or a
// uniform-register load from constant memory, its values used as an operand of two of the DFMAs.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NC = 14;
__constant__ double ctab[4096];
template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, const double* __restrict__ gtab, double* __restrict__ sink) {
    __shared__ __align__(16) double stab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) stab[i] = gtab[i];
    __syncthreads();
    double acc[NC];
    const double c = 1e-9 * (double)(blockIdx.x & 3) + 1e-9 * (double)(threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = (double)i + 0.5 + c;
    for (int it = 0; it < iters; ++it) {
        const int base = (it & 127) * 32;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double t0, t1, t2, t3;
            if (MODE == 1) {
                const double2 a = *reinterpret_cast<const double2*>(&stab[base + 4 * r]);
                const double2 b = *reinterpret_cast<const double2*>(&stab[base + 4 * r + 2]);
                t0 = a.x; t1 = a.y; t2 = b.x; t3 = b.y;
            } else if (MODE == 2) {
                t0 = ctab[base + 4 * r]; t1 = ctab[base + 4 * r + 1]; t2 = ctab[base + 4 * r + 2]; t3 = ctab[base + 4 * r + 3];
            } else { t0 = 0.125; t1 = 0.25; t2 = 0.375; t3 = 0.0625; }
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                if (i == 0) acc[i] = fma(acc[i], acc[(i + 1) % NC], t0);
                else if (i == 3) acc[i] = fma(acc[i], acc[(i + 1) % NC], t1);
                else if (i == 7) acc[i] = fma(acc[i], acc[(i + 1) % NC], t2);
                else if (i == 10) acc[i] = fma(acc[i], acc[(i + 1) % NC], t3);
                else acc[i] = fma(acc[i], acc[(i + 1) % NC], 0.0013888888888888889);
            }
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) acc[i] = acc[i] * 1e-300 + 0.5;
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) t += acc[i];
    if (t == 123456.789) sink[0] = t;
}
template <int MODE> void run(int iters, int bps, int threads, const double* gtab, double* sink, const char* what) {
    const int blocks = 148 * bps;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(iters / 4, gtab, sink);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(iters, gtab, sink);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double wps = bps * (threads / 32) / 4.0;
    const double instr = (double)iters * (8 * NC + NC) * wps;
    printf("mode %d %4.1f warps/SMSP  %.3f cycles at 1965 MHz per FP64 instruction   %s\n", MODE, wps, ms * 1.965e6 / instr, what);
}
int main() {
    double* sink; cudaMalloc(&sink, 64);
    double* gtab; cudaMalloc(&gtab, 4096 * 8);
    static double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1e-3 * (i % 97);
    cudaMemcpy(gtab, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(ctab, h, sizeof(h));
    const int cfgs[2][2] = {{6, 128}, {8, 256}};
    for (int c = 0; c < 2; ++c) {
        run<0>(4000, cfgs[c][0], cfgs[c][1], gtab, sink, "immediates only");
        run<1>(4000, cfgs[c][0], cfgs[c][1], gtab, sink, "two LDS.128 per 14 DFMA (4 doubles used as addends)");
        run<2>(4000, cfgs[c][0], cfgs[c][1], gtab, sink, "constant-memory table (uniform registers)");
    }
    return 0;
}
