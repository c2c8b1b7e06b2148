// Does DFMA throughput on B200 depend on how many distinct 64-bit register operands it reads?
//   mode 0: x = fma(x, a, b)      a, b loop-invariant registers (operand reuse possible)
//   mode 1: x = fma(y_i, a, x)    one more varying register source
//   mode 2: x = fma(y_i, z_i, x)  three distinct varying register sources (the K2T quad pattern)
//   mode 3: x = y_i * z_i         DMUL, two varying sources
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_regs fp64_regs.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, int iters, double seed, long long* clk) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = seed + i + threadIdx.x, y[i] = 1.0 + 1e-9 * (i + threadIdx.x), z[i] = 1.0 - 1e-9 * (i + 2 * threadIdx.x);
    const double a = 1.0000001, b = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fma(x[i], a, b);
            if (MODE == 1) x[i] = fma(y[i], a, x[i]);
            if (MODE == 2) x[i] = fma(y[i], z[(i + 3) & 7], x[i]);
            if (MODE == 3) x[i] = y[i] * z[(i + 3) & 7] + 0.0 * x[i];
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("" : "+d"(x[i]), "+d"(y[i]));
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
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
    printf("mode %d warps/SM %2d: %6.2f clk per 8 FP64 instructions per warp slot  (%4.2f clk per instruction and sub-partition)\n", MODE,
           warps_per_sm, (double)h / iters / (warps_per_sm / 4.0), (double)h / iters / 8.0 / (warps_per_sm / 4.0));
    cudaFree(out), cudaFree(clk);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    for (int w : {4, 8, 16}) run<0>(w, sms), run<1>(w, sms), run<2>(w, sms);
    return 0;
}
