// Follow-up to fp64_regs.cu: what does the third register operand of a DFMA cost, and do constant-bank operands avoid it?
//   mode 0: x_i = fma(y_i, z_j, x_i)            three varying registers, many different (i, j) combinations
//   mode 1: x_i = fma(y_i, ctab[i], x_i)        constant-bank operand, compile-time offset
//   mode 2: x_i = fma(y_i, ctab[8*u + i], x_i)  constant-bank operand, warp-uniform run-time offset (u = loop counter & 63)
//   mode 3: x_i = fma(y_i, s, x_i) with s loaded from shared memory (broadcast) right before: the K2T pattern
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_regs2 fp64_regs2.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double ctab[1024];

template <int MODE>
__global__ void k(double* out, int iters, double seed, long long* clk) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 1.0 + 1e-9 * i;
    double x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = seed + i + threadIdx.x, y[i] = 1.0 + 1e-9 * (i + threadIdx.x);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int u = it & 63;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) x[i] = fma(y[i], y[(5 * i + 3) & 15], x[i]);
            if (MODE == 1) x[i] = fma(y[i], ctab[i], x[i]);
            if (MODE == 2) x[i] = fma(y[i], ctab[16 * u + i], x[i]);
            if (MODE == 3) x[i] = fma(y[i], sh[16 * u + i], x[i]);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int warps_per_sm, int sms) {
    const int iters = 4096;
    double* out;
    long long* clk;
    cudaMalloc(&out, sizeof(double) * sms * warps_per_sm * 32);
    cudaMalloc(&clk, sizeof(long long) * sms);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 1.0, clk);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 1.0, clk);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mode %d warps/SM %2d: %5.2f clk per DFMA and sub-partition\n", MODE, warps_per_sm, (double)h / iters / 16.0 / (warps_per_sm / 4.0));
    cudaFree(out), cudaFree(clk);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    double h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 1.0 + 1e-9 * i;
    cudaMemcpyToSymbol(ctab, h, sizeof(h));
    const int sms = p.multiProcessorCount;
    for (int w : {4, 8}) run<0>(w, sms), run<1>(w, sms), run<2>(w, sms), run<3>(w, sms);
    return 0;
}
