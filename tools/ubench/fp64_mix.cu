// Does the SM sub-partition issue other instructions in the shadow of the half-rate FP64 pipe (B200)?
// Per loop iteration: 8 independent DFMA + K independent integer (LOP3/IADD) or shared-load instructions.
// If the shadow is usable the iteration costs max(16, 8 + K) clk per warp slot, otherwise 16 + K.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool LDS>
__global__ void mix_kernel(double* out, int iters, double seed, long long* clk, unsigned salt) {
    __shared__ unsigned sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = i * 2654435761u;
    double x[8];
    unsigned y[K > 0 ? K : 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < (K > 0 ? K : 1); ++i) y[i] = salt + i * 77u + threadIdx.x;
    const double a = 1.0000001, b = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = fma(x[i], a, b);
            if (i < K) {
                if (LDS) y[i] = sh[(y[i] + it) & 1023u];
                else y[i] = (y[i] ^ (unsigned)it) + salt;
            }
        }
#pragma unroll
        for (int i = 8; i < K; ++i) {
            if (LDS) y[i] = sh[(y[i] + it) & 1023u];
            else y[i] = (y[i] ^ (unsigned)it) + salt;
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    unsigned u = 0;
#pragma unroll
    for (int i = 0; i < (K > 0 ? K : 1); ++i) u ^= y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)u;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int K, bool LDS>
void run(int warps_per_sm, int sms) {
    const int iters = 4096;
    double* out;
    long long* clk;
    cudaMalloc(&out, sizeof(double) * sms * warps_per_sm * 32);
    cudaMalloc(&clk, sizeof(long long) * sms);
    mix_kernel<K, LDS><<<sms, warps_per_sm * 32>>>(out, iters, 1.0, clk, 12345u);
    mix_kernel<K, LDS><<<sms, warps_per_sm * 32>>>(out, iters, 1.0, clk, 12345u);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  8 DFMA + %2d %s: %6.2f clk per iteration per SM sub-partition  (%5.2f per warp)\n", warps_per_sm, K,
           LDS ? "LDS      " : "LOP3+IADD", (double)h / iters, (double)h / iters / (warps_per_sm / 4.0));
    cudaFree(out), cudaFree(clk);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    for (int w : {4, 8}) {
        run<0, false>(w, sms), run<2, false>(w, sms), run<4, false>(w, sms), run<8, false>(w, sms), run<12, false>(w, sms), run<16, false>(w, sms);
        run<2, true>(w, sms), run<4, true>(w, sms), run<8, true>(w, sms);
    }
    return 0;
}
