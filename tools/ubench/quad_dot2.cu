// K2T inner loop in isolation WITH the sub-cube gathers: per "cube" 32 x LDS.128 from a 64 KB tile, then Q quad
// records whose tables come from (a) shared memory or (b) __constant__ memory.  clk per record per SMSP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o quad_dot2 quad_dot2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ uint4 c_recs[2816];

template <int HB>
__device__ __forceinline__ constexpr int pair_index(int k) { return ((k / HB) * 2 * HB) | (k % HB); }

template <int XR>
__device__ __forceinline__ double dot_pairs(const double2 (&a)[32], const double2* __restrict__ tab) {
    constexpr int HB = XR >= 16 ? 16 : XR >= 8 ? 8 : XR >= 4 ? 4 : XR >= 2 ? 2 : 1;
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
}

__device__ __forceinline__ uint32_t swz(uint32_t t) { return t ^ ((t >> 3) & 7u) ^ ((t >> 6) & 7u) ^ ((t >> 9) & 7u); }

template <bool CONST_TABLES, int Q, bool BRANCHY>
__global__ void __launch_bounds__(128, 2) k(const double2* __restrict__ src, double* out, int ncubes, int reps, long long* clk, const unsigned* masks) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* tile = reinterpret_cast<double2*>(smem);
    uint4* recs = reinterpret_cast<uint4*>(smem + 65536);
    for (int u = threadIdx.x; u < 4096; u += 128) tile[u] = src[u];
    if (!CONST_TABLES)
        for (int u = threadIdx.x; u < 2816; u += 128) {
            uint4 v; v.x = u * 2654435761u; v.y = 0x3ff00000u | (u & 0xffff); v.z = u * 40503u; v.w = 0x3ff00000u | ((u * 7) & 0xffff);
            recs[u] = v;
        }
    __syncthreads();
    double acc0 = 0, acc1 = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const uint4* rp = CONST_TABLES ? c_recs : recs;
        for (int c = 0; c < ncubes; ++c) {
            // gather: origin from the thread id, cube bits vary with c
            const uint32_t sb0 = swz(1u << (c % 3)), sb1 = swz(1u << (3 + c % 2)), sb2 = swz(1u << (5 + c % 2)), sb3 = swz(1u << (7 + c % 2)), sb4 = swz(1u << (9 + c % 3));
            uint32_t org = 0;
            {
                uint32_t used = (1u << (c % 3)) | (1u << (3 + c % 2)) | (1u << (5 + c % 2)) | (1u << (7 + c % 2)) | (1u << (9 + c % 3));
                int j = 0;
                for (int b = 0; b < 12; ++b)
                    if (!((used >> b) & 1u)) { org |= ((threadIdx.x >> j) & 1u) << b; ++j; }
            }
            const uint32_t osw = swz(org);
            double2 a[32];
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                uint32_t off = osw;
                if (p & 1) off ^= sb0;
                if (p & 2) off ^= sb1;
                if (p & 4) off ^= sb2;
                if (p & 8) off ^= sb3;
                if (p & 16) off ^= sb4;
                a[p] = tile[off];
            }
            const unsigned qm = BRANCHY ? masks[c] : 31u;
#define QUAD(J)                                                                                          \
    if (J < Q && (qm & (1u << J))) {                                                                                       \
        const uint4 hd = *rp;                                                                            \
        const double w = dot_pairs<31 ^ (1 << J)>(a, reinterpret_cast<const double2*>(rp + 1));          \
        const unsigned par = __popc(hd.x ^ (hd.z & org));                                                \
        const double sw = __longlong_as_double(__double_as_longlong(w) ^ ((long long)(par & 1u) << 63)); \
        if (J & 1) acc1 += sw; else acc0 += sw;                                                          \
        rp += 9;                                                                                         \
    }
            QUAD(0) QUAD(1) QUAD(2) QUAD(3) QUAD(4)
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 128 + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <bool CONST_TABLES, int Q, bool BRANCHY>
void run(const char* name, unsigned mask_value) {
    const int sms = 148, ncubes = 60, reps = 40;
    double2* src; double* out; long long* clk;
    cudaMalloc(&src, 4096 * 16); cudaMalloc(&out, sms * 2 * 128 * 8); cudaMalloc(&clk, sms * 2 * 8);
    cudaMemset(src, 0x3c, 4096 * 16);
    unsigned hm[64]; for (int i = 0; i < 64; ++i) hm[i] = mask_value; unsigned* dmasks; cudaMalloc(&dmasks, sizeof(hm)); cudaMemcpy(dmasks, hm, sizeof(hm), cudaMemcpyHostToDevice);
    if (CONST_TABLES) {
        static uint4 h[2816];
        for (int u = 0; u < 2816; ++u) { h[u].x = u * 2654435761u; h[u].y = 0x3ff00000u | (u & 0xffff); h[u].z = u * 40503u; h[u].w = 0x3ff00000u | ((u * 7) & 0xffff); }
        cudaMemcpyToSymbol(c_recs, h, sizeof(h));
    }
    const size_t smem = 65536 + 2816 * 16;
    cudaFuncSetAttribute(k<CONST_TABLES, Q, BRANCHY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) k<CONST_TABLES, Q, BRANCHY><<<sms * 2, 128, smem>>>(src, out, ncubes, reps, clk, dmasks);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<CONST_TABLES, Q, BRANCHY>);
    const double per_cube = (double)h / (ncubes * reps) / 2.0;  // per SMSP (2 warps)
    printf("%-34s regs %3d: %7.1f clk per cube per SMSP = %6.1f per quad (gather amortised over %d quads)  %s\n", name, fa.numRegs,
           per_cube, per_cube / __builtin_popcount(mask_value & ((1u << Q) - 1u)), __builtin_popcount(mask_value & ((1u << Q) - 1u)), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<false, 5, false>("smem tables, 5 quads straight", 31u);
    run<false, 5, true>("smem tables, 5 quads, branch each", 31u);
    run<false, 5, true>("smem tables, mask 0b10111 (4 quads)", 23u);
    run<false, 5, true>("smem tables, mask 0b01101 (3 quads)", 13u);
    run<false, 3, false>("smem tables, 3 quads straight", 31u);
    run<true, 3, false>("const tables, 3 quads straight", 31u);
    return 0;
}
