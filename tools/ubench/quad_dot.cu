// Isolated K2T quad record: each thread holds a 32-amplitude sub-cube in registers and evaluates NREC records
// (header + 16 weights in shared memory) with the kernel's dot_pairs; clk per record per SM sub-partition vs the
// FP64-pipe bound (52 FP64 instructions x 2 issue clk = 104 clk per warp-record; 2 warps per SMSP).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o quad_dot quad_dot.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int HB>
__device__ __forceinline__ constexpr int pair_index(int k) { return ((k / HB) * 2 * HB) | (k % HB); }

template <int XR, int VARIANT>
__device__ __forceinline__ double dot_pairs(const double2 (&a)[32], const double2* __restrict__ tab) {
    constexpr int HB = XR >= 16 ? 16 : XR >= 8 ? 8 : XR >= 4 ? 4 : XR >= 2 ? 2 : 1;
    if (VARIANT == 0) {  // as in expectation_tiled.cu: two batches of 8
        double w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            double2 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = tab[4 * b + j];
            double pr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int p = pair_index<HB>(8 * b + j); pr[j] = a[p].x * a[p ^ XR].x; }
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int p = pair_index<HB>(8 * b + j); pr[j] = fma(a[p].y, a[p ^ XR].y, pr[j]); }
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j & 3] = fma(pr[j], (j & 1) ? t[j >> 1].y : t[j >> 1].x, w[j & 3]);
        }
        return (w[0] + w[1]) + (w[2] + w[3]);
    } else if (VARIANT == 1) {  // weights folded first: u = t*ax, v = t*ay; w1 += u*bx; w2 += v*by  (64 ops, chains of 2)
        double w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int p = pair_index<HB>(k);
            const double2 tt = tab[k >> 1];
            const double t = (k & 1) ? tt.y : tt.x;
            w[k & 3] = fma(t * a[p].x, a[p ^ XR].x, w[k & 3]);
            w[4 + (k & 3)] = fma(t * a[p].y, a[p ^ XR].y, w[4 + (k & 3)]);
        }
        return ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
    } else {  // 8 accumulators, single batch of 16
        double w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double pr[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int p = pair_index<HB>(k); pr[k] = a[p].x * a[p ^ XR].x; }
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int p = pair_index<HB>(k); pr[k] = fma(a[p].y, a[p ^ XR].y, pr[k]); }
#pragma unroll
        for (int k = 0; k < 16; ++k) { const double2 tt = tab[k >> 1]; w[k & 7] = fma(pr[k], (k & 1) ? tt.y : tt.x, w[k & 7]); }
        return ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
    }
}

template <int VARIANT, bool BRANCHY>
__global__ void __launch_bounds__(128, 2) quad_kernel(const double2* __restrict__ src, double* out, int nrec, int reps, long long* clk, unsigned mask) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint4* recs = reinterpret_cast<uint4*>(smem);
    for (int u = threadIdx.x; u < nrec * 9; u += 128) {
        uint4 v;
        v.x = u * 2654435761u; v.y = 0x3ff00000u | (u & 0xffff); v.z = u * 40503u; v.w = 0x3ff00000u | ((u * 7) & 0xffff);
        recs[u] = v;  // doubles near 1.0
    }
    double2 a[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) a[p] = src[(threadIdx.x * 32 + p) & 4095];
    __syncthreads();
    double acc0 = 0, acc1 = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const uint4* rp = recs;
        for (int q = 0; q < nrec; q += 5) {
#pragma unroll
            for (int p = 0; p < 32; ++p) asm volatile("" : "+d"(a[p].x), "+d"(a[p].y));
#define QUAD(J)                                                                                           \
    if (!BRANCHY || (mask & (1u << J))) {                                                                 \
        const uint4 hd = *rp;                                                                             \
        const double w = dot_pairs<31 ^ (1 << J), VARIANT>(a, reinterpret_cast<const double2*>(rp + 1)); \
        const unsigned par = __popc(hd.x ^ (hd.z & threadIdx.x));                                         \
        const double sw = __longlong_as_double(__double_as_longlong(w) ^ ((long long)(par & 1u) << 63));  \
        if (J & 1) acc1 += sw; else acc0 += sw;                                                           \
        rp += 9;                                                                                          \
    }
            QUAD(0) QUAD(1) QUAD(2) QUAD(3) QUAD(4)
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 128 + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int VARIANT, bool BRANCHY>
void run(const char* name) {
    const int sms = 148, nrec = 250, reps = 40;
    double2* src; double* out; long long* clk;
    cudaMalloc(&src, 4096 * 16); cudaMalloc(&out, sms * 2 * 128 * 8); cudaMalloc(&clk, sms * 2 * 8);
    cudaMemset(src, 0x3c, 4096 * 16);
    const size_t smem = nrec * 144;
    cudaFuncSetAttribute(quad_kernel<VARIANT, BRANCHY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) quad_kernel<VARIANT, BRANCHY><<<sms * 2, 128, smem>>>(src, out, nrec, reps, clk, 31u);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, quad_kernel<VARIANT, BRANCHY>);
    // per SMSP: 2 warps (one per CTA), each nrec*reps records
    printf("%-28s regs %3d: %6.1f clk per warp-record (2 warps/SMSP => %5.1f clk per record per SMSP; FP64 bound 104)  %s\n", name,
           fa.numRegs, (double)h / (nrec * reps), (double)h / (nrec * reps) / 2.0, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<0, false>("two batches of 8, straight");
    run<0, true>("two batches of 8, branch per quad");
    run<1, false>("weights folded first (64 ops)");
    run<2, false>("one batch of 16, straight");
    run<2, true>("one batch of 16, branch per quad");
    return 0;
}
