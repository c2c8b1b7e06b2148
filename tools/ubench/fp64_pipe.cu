// FP64 pipe microbenchmark (B200): DFMA throughput per SM vs warps/SMSP and ILP, and dependent-issue latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double seed, long long* clk) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed + i + threadIdx.x;
    const double a = 1.0000001, b = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, int sms) {
    const int iters = 4096;
    double* out;
    long long* clk;
    cudaMalloc(&out, sizeof(double) * sms * warps_per_sm * 32);
    cudaMalloc(&clk, sizeof(long long) * sms);
    dfma_kernel<ILP><<<sms, warps_per_sm * 32>>>(out, iters, 1.0, clk);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<sms, warps_per_sm * 32>>>(out, iters, 1.0, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h;
    cudaMemcpy(&h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double dfma_per_sm = (double)iters * ILP * warps_per_sm * 32;
    printf("warps/SM %2d ILP %2d: %8lld clk  -> %6.2f DFMA/clk/SM, %5.2f clk per dependent DFMA;  %7.2f TFLOP/s chip (event timed)\n",
           warps_per_sm, ILP, h, dfma_per_sm / h, (double)h / iters, 2.0 * dfma_per_sm * sms / (ms * 1e-3) / 1e12);
    cudaFree(out), cudaFree(clk);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    const int sms = p.multiProcessorCount;
    run<1>(1, sms), run<2>(1, sms), run<4>(1, sms), run<8>(1, sms), run<16>(1, sms);
    run<1>(4, sms), run<2>(4, sms), run<4>(4, sms), run<8>(4, sms);
    run<1>(8, sms), run<2>(8, sms), run<4>(8, sms), run<8>(8, sms);
    run<4>(16, sms), run<8>(16, sms), run<4>(32, sms), run<8>(32, sms);
    return 0;
}
