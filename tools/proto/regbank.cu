// Register-operand delivery probe for FP64 instructions (B200): cycles a warp-wide DFMA holds the FP64 pipe, by how many of its
// register operands have to be fetched.  Timed with clock64 inside the kernel (no assumption about the clock rate).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o regbank regbank.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NC = 16;
template <int MODE>
__global__ void __launch_bounds__(1024) k(int iters, double* __restrict__ sink, long long* __restrict__ cyc) {
    double acc[NC];
    const double m = 1.0 - 1e-9 * (double)(threadIdx.x & 7), c = 1e-9 * (double)(blockIdx.x & 3);
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = (double)i + 0.5 + c;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                if (MODE == 0) acc[i] = fma(acc[i], m, 0.0013888888888888889);            // one fetched operand (m reused, immediate addend)
                else if (MODE == 1) acc[i] = fma(acc[i], acc[(i + 1) % NC], 0.0013888888888888889);   // two fetched, immediate addend
                else if (MODE == 2) acc[i] = fma(m, acc[(i + 1) % NC], acc[i]);            // two fetched + one reusable
                else if (MODE == 3) acc[i] = fma(acc[(i + 1) % NC], acc[(i + 2) % NC], acc[i]);       // three fetched
                else if (MODE == 4) acc[i] = fma(acc[i], m, c);                            // one fetched + two reusable
                else if (MODE == 5) acc[i] = acc[i] * acc[(i + 1) % NC];                   // DMUL two fetched
                else if (MODE == 6) acc[i] = fma(acc[i], acc[i], c);                       // same register in two slots
                else if (MODE == 7) acc[i] = fma(acc[(i + 2) % NC], acc[(i + 4) % NC], acc[i]);       // three fetched, register pairs 4 apart
                else if (MODE == 8) acc[i] = fma(acc[(i + 1) % NC], acc[(i + 3) % NC], acc[i]);
                else if (MODE == 9) acc[i] = fma(acc[(i + 4) % NC], acc[(i + 8) % NC], acc[i]);
                else if (MODE == 10) acc[i] = fma(acc[(i + 2) % NC], acc[(i + 3) % NC], acc[i]);
                else if (MODE == 11) acc[i] = fma(acc[i], acc[(i + 2) % NC], 0.0013888888888888889);  // two fetched 4 apart
                else if (MODE == 12) acc[i] = fma(acc[i], acc[(i + 4) % NC], 0.0013888888888888889);
            }
#pragma unroll
        for (int i = 0; i < NC; ++i) acc[i] = acc[i] * 1e-300 + 0.5;
    }
    const long long t1 = clock64();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) t += acc[i];
    if (t == 123456.789) sink[0] = t;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(int iters, int blocks_per_sm, int threads, double* sink, long long* cyc, const char* what) {
    const int blocks = 148 * blocks_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(iters / 4, sink, cyc);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(iters, sink, cyc);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    static long long h[148 * 16];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += (double)h[i]; mean /= blocks;
    const double warps_per_smsp = blocks_per_sm * (threads / 32) / 4.0;
    const double instr = (double)iters * (8 * NC + NC) * warps_per_smsp;      // FP64 instructions one sub-partition issues while a block runs
    printf("form %2d  %4.1f warps per sub-partition  %.3f cycles at 1965 MHz per FP64 instruction   %s\n", MODE, warps_per_smsp, ms * 1.965e6 / instr, what);
    (void)mean;
}
int main() {
    double* sink; cudaMalloc(&sink, 64);
    long long* cyc; cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
    const int iters = 2000;
    const int cfgs[4][2] = {{1, 128}, {6, 128}, {8, 256}, {4, 1024}};      // 1, 6, 16 warps per sub-partition; 2 x 32 in two rounds
    for (int cfg = 0; cfg < 4; ++cfg) {
        const int bps = cfgs[cfg][0], th = cfgs[cfg][1];
        run<0>(iters, bps, th, sink, cyc, "1 fetched, immediate addend");
        run<4>(iters, bps, th, sink, cyc, "1 fetched + 2 reusable");
        run<6>(iters, bps, th, sink, cyc, "same register in two slots + 1 reusable");
        run<1>(iters, bps, th, sink, cyc, "2 fetched (pairs i, i+1), immediate addend");
        run<11>(iters, bps, th, sink, cyc, "2 fetched (i, i+2)");
        run<12>(iters, bps, th, sink, cyc, "2 fetched (i, i+4)");
        run<5>(iters, bps, th, sink, cyc, "DMUL, 2 fetched");
        run<2>(iters, bps, th, sink, cyc, "2 fetched + 1 reusable (ptxas drops .reuse on every 6th: those fetch 3)");
        run<3>(iters, bps, th, sink, cyc, "3 fetched (i+1, i+2, i)");
        run<7>(iters, bps, th, sink, cyc, "3 fetched (i+2, i+4, i)");
        run<9>(iters, bps, th, sink, cyc, "3 fetched (i+4, i+8, i)");
    }
    return 0;
}
