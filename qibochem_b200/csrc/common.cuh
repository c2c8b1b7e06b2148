// Shared declarations of the qibochem_b200 native library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler is attached
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/qibochem_b200.h"

namespace qc {

void set_error(const std::string& msg);

#define QC_CHECK(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            qc::set_error(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                          ":" + std::to_string(__LINE__) + ")");                                         \
            return 1;                                                                                    \
        }                                                                                                \
    } while (0)

#define QC_REQUIRE(cond, msg)                 \
    do {                                      \
        if (!(cond)) {                        \
            qc::set_error(std::string(msg));  \
            return 2;                         \
        }                                     \
    } while (0)

// Pinned host + device arena pair used to ship small per-call programs (descriptors, tables, term lists) with one
// asynchronous copy.  Grown on demand; reused across calls.
struct Arena {
    void* host = nullptr;
    void* dev = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
};

// Kernel classes for the launch counter / live profiler (qc_profile_*)
// NVTX range names of the kernel classes (tracing), same order as KernelClass
constexpr const char* kClassNames[] = {"qc.K0.init", "qc.K1.rotation", "qc.K2.expectation_sweep", "qc.reduce", "qc.K5.sampling",
                                       "qc.K3.shot_statistics", "qc.K4.exchange", "qc.K2T.expectation_tiled", "qc.K6.gradient"};
enum KernelClass { kK0Init = 0, kK1Rotation = 1, kK2Expectation = 2, kReduce = 3, kK5Sampling = 4, kK3Shots = 5, kK4Exchange = 6, kK2Tiled = 7, kGradient = 8, kNumClasses = 9 };

struct TiledPlan;  // expectation_tiled.cu
void free_tiled_plan(TiledPlan* p);
struct HpPlan;     // gradient_tiled.cu
void free_hp_plan(HpPlan* p);
struct SweepPlan;  // expectation.cu
void free_sweep_plan(SweepPlan* p);

struct Profiler {
    bool enabled = false;
    struct Rec {
        int cls;
        cudaEvent_t a, b;
    };
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    double ms[kNumClasses] = {0};
    double bytes[kNumClasses] = {0};       // algorithmic bytes, counted always
    uint64_t launches[kNumClasses] = {0};  // counted always
    cudaEvent_t get_event();
    void collect();  // caller has synchronised the stream
    void release();
};

constexpr int kThreads = 256;       // threads per CTA for the sweep kernels
constexpr int kCtasPerSm = 8;       // 8 x 256 = 2048 resident threads per SM

}  // namespace qc

struct qc_state {
    int device = 0;
    int n = 0;              // local qubits
    uint64_t dim = 0;       // 2^n
    double2* psi = nullptr; // complex128[dim]
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 148;
    qc::Arena arena;        // per-call programs
    double* d_partials = nullptr;  // reduction scratch
    size_t partials_cap = 0;       // in doubles
    double* d_out = nullptr;       // small result buffer (device)
    size_t out_cap = 0;
    double* h_out = nullptr;       // pinned mirror of d_out
    double2* scratch_psi = nullptr;  // scratch copy of the shard (sampling), allocated lazily
    qc::Profiler prof;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;  // host<->device bytes moved by the product path since the last reset
    std::vector<qc::TiledPlan*> k2plans;  // cached plans of the tiled expectation (keyed by a hash of the term list; LRU)
    qc::SweepPlan* sweep_plan = nullptr;  // grouped device program of the last term list the sweep kernel ran
    qc::HpPlan* hp_plan = nullptr;        // cached plan of the tiled H|psi> (adjoint gradient), last term list only
    int k1_mode = 0;                  // 0 auto, 1 one HBM pass per fused run, 2 shared-memory tiles whenever a run fits
    int k2_real = 0;                  // 1: tiled expectation through the real-amplitude kernels when the shard is exactly real
    int last_k2_real = 0;             // whether the last tiled expectation took them
    int k2_mode = 0;                  // 0 auto, 1 force sweeps (one HBM pass per x-mask), 2 force tiles
    int ensure_partials(size_t doubles);
    int ensure_out(size_t doubles);
};

namespace qc {

// ---- device helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t insert_zero_bit(uint64_t p, int b) {
    const uint64_t lo = p & ((1ull << b) - 1ull);
    return ((p >> b) << (b + 1)) | lo;
}
// (-1)^{popcount(i & m)} applied to v by flipping the IEEE sign bit
__device__ __forceinline__ double flip_sign(double v, uint64_t i, uint64_t m) {
    const long long par = (long long)(__popcll(i & m) & 1);
    return __longlong_as_double(__double_as_longlong(v) ^ (par << 63));
}
// software pext for small masks (<= ~12 set bits)
__device__ __forceinline__ uint32_t gather_bits(uint64_t i, uint64_t mask) {
    uint32_t out = 0;
    int k = 0;
    while (mask) {
        const int b = __ffsll((long long)mask) - 1;
        out |= (uint32_t)((i >> b) & 1ull) << k;
        ++k;
        mask &= mask - 1;
    }
    return out;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum over the CTA (blockDim.x == kThreads); result valid in thread 0.  Fixed order => deterministic.
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[kThreads / 32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) warp_part[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) t += warp_part[k];
    }
    return t;
}

// RAII bracket around one kernel launch: counts it and, when profiling is on, times it with CUDA events recorded
// on the handle's stream.
struct LaunchScope {
    qc_state* h;
    int cls;
    cudaEvent_t stop = nullptr;
    LaunchScope(qc_state* h_, int cls_, double algorithmic_bytes) : h(h_), cls(cls_) {
        nvtxRangePushA(kClassNames[cls]);
        h->prof.launches[cls] += 1;
        h->prof.bytes[cls] += algorithmic_bytes;
        if (h->prof.enabled) {
            cudaEvent_t a = h->prof.get_event();
            stop = h->prof.get_event();
            cudaEventRecord(a, h->stream);
            h->prof.pending.push_back({cls, a, stop});
        }
    }
    ~LaunchScope() {
        if (stop) cudaEventRecord(stop, h->stream);
        nvtxRangePop();
    }
};

// out[g] = sum_b partials[g*stride + b] (one CTA per g, fixed order); defined in state.cu
int fold_partials(qc_state* h, const double* partials, int ngroups, int nb, int stride, double* out);

inline int grid_for(uint64_t work_items, int per_thread, int sm_count) {
    const uint64_t per_cta = (uint64_t)kThreads * per_thread;
    uint64_t need = (work_items + per_cta - 1) / per_cta;
    const uint64_t cap = (uint64_t)sm_count * kCtasPerSm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace qc
