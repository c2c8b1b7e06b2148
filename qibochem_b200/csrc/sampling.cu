// K5: sampling in a product Pauli basis, K3: integer shot statistics (parity sums, frequencies).
//
// Reference behaviour replaced (src/qibochem/measurement/result.py:105-117): per measurement group the reference
// copies the circuit, appends gates.M(q, basis=X|Y|Z), RE-SIMULATES it with nshots, builds a Counter of bit strings
// and then evaluates every Z-only term with Python loops over keys (qibo expectation_from_samples, reached from
// result.py:54).  Here psi is computed once; a group costs a scratch copy + one basis-change pass per rotated qubit
// + one probability pass + an inverse-CDF draw per shot, and the term statistics are exact integer sums
// S_j = sum_shots (-1)^{popc(s & m_j)}.
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace qc {

// ---------------------------------------------------------------------------------------------------------------
// single-qubit basis change on a scratch copy: [a;b] <- M [a;b] on index bit `bit`
// ---------------------------------------------------------------------------------------------------------------
struct Mat2 {
    double2 m00, m01, m10, m11;
};
__device__ __forceinline__ double2 cfma2(double2 m, double2 v, double2 acc) {
    return make_double2(acc.x + m.x * v.x - m.y * v.y, acc.y + m.x * v.y + m.y * v.x);
}

template <int U>
__global__ void __launch_bounds__(kThreads) gate1q_pairs_kernel(double2* __restrict__ psi, uint64_t npairs, int bit,
                                                                const Mat2 m) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    const uint64_t bm = 1ull << bit;
    for (uint64_t p0 = (uint64_t)blockIdx.x * kThreads + threadIdx.x; p0 < npairs; p0 += stride * U) {
        uint64_t idx[U];
        double2 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t p = p0 + (uint64_t)u * stride;
            idx[u] = insert_zero_bit(p, bit);
            if (p < npairs) {
                a[u] = psi[idx[u]];
                b[u] = psi[idx[u] | bm];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t p = p0 + (uint64_t)u * stride;
            if (p < npairs) {
                const double2 z = make_double2(0.0, 0.0);
                psi[idx[u]] = cfma2(m.m01, b[u], cfma2(m.m00, a[u], z));
                psi[idx[u] | bm] = cfma2(m.m11, b[u], cfma2(m.m10, a[u], z));
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) gate1q_low_kernel(double2* __restrict__ psi, uint64_t dim, int bit,
                                                              const Mat2 m) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    const int lane = threadIdx.x & 31;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i - lane < dim; i += stride) {
        const bool valid = i < dim;
        const double2 a = valid ? psi[i] : make_double2(0.0, 0.0);
        double2 b;
        b.x = __shfl_xor_sync(0xffffffffu, a.x, 1 << bit);
        b.y = __shfl_xor_sync(0xffffffffu, a.y, 1 << bit);
        if (valid) {
            const double2 z = make_double2(0.0, 0.0);
            const bool hi = (i >> bit) & 1ull;
            // this thread owns `a`; the partner value is `b`
            psi[i] = hi ? cfma2(m.m11, a, cfma2(m.m10, b, z)) : cfma2(m.m01, b, cfma2(m.m00, a, z));
        }
    }
}

static int apply_gate1q(qc_state* h, double2* psi, int bit, const Mat2& m) {
    LaunchScope ls(h, kK5Sampling, 32.0 * (double)h->dim);
    if (bit < 5) {
        const int nb = grid_for(h->dim, 4, h->sm_count);
        gate1q_low_kernel<<<nb, kThreads, 0, h->stream>>>(psi, h->dim, bit, m);
    } else {
        const uint64_t npairs = h->dim >> 1;
        const int nb = grid_for(npairs, 4, h->sm_count);
        gate1q_pairs_kernel<4><<<nb, kThreads, 0, h->stream>>>(psi, npairs, bit, m);
    }
    QC_CHECK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// probabilities: per-chunk sums, prefix sums over chunks, inverse-CDF draws
// ---------------------------------------------------------------------------------------------------------------
// A shot costs a binary search over the chunk prefix sums plus a scan of (on average) half a chunk, so chunks are kept
// small: 128 amplitudes (2 KB) up to 25 qubits, growing so that the single-CTA prefix sum handles <= 2^18 chunk sums.
constexpr int kMinChunkLog2 = 7;
constexpr int kMaxChunksLog2 = 18;
inline int chunk_log2_for(int n) { return std::min(n, std::max(kMinChunkLog2, n - kMaxChunksLog2)); }

// small chunks (<= 1024 amplitudes): one warp per chunk, lane-strided, fixed-order warp reduction
__global__ void __launch_bounds__(kThreads) chunk_prob_warp_kernel(const double2* __restrict__ psi, uint64_t dim,
                                                                   int chunk_log2, double* __restrict__ chunk_sum) {
    const uint64_t nchunks = dim >> chunk_log2;
    const uint64_t csize = 1ull << chunk_log2;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * kThreads) >> 5;
    for (uint64_t c = warp; c < nchunks; c += nwarps) {
        const double2* p = psi + (c << chunk_log2);
        double acc = 0.0;
        for (uint64_t k = lane; k < csize; k += 32) {
            const double2 a = p[k];
            acc += a.x * a.x + a.y * a.y;
        }
        acc = warp_sum(acc);
        if (lane == 0) chunk_sum[c] = acc;
    }
}

__global__ void __launch_bounds__(kThreads) chunk_prob_kernel(const double2* __restrict__ psi, uint64_t dim,
                                                              int chunk_log2, double* __restrict__ chunk_sum) {
    const uint64_t nchunks = dim >> chunk_log2;
    const uint64_t csize = 1ull << chunk_log2;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const double2* p = psi + (c << chunk_log2);
        double acc = 0.0;
        for (uint64_t k = threadIdx.x; k < csize; k += kThreads) {
            const double2 a = p[k];
            acc += a.x * a.x + a.y * a.y;
        }
        const double t = block_sum(acc);
        if (threadIdx.x == 0) chunk_sum[c] = t;
        __syncthreads();
    }
}

// inclusive prefix sum of v[0..n) in place, single CTA of 1024 threads, sequential over tiles (deterministic)
__global__ void __launch_bounds__(1024) inclusive_scan_kernel(double* __restrict__ v, uint64_t n) {
    __shared__ double wsum[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        double x = i < n ? v[i] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            double s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        const double carry = carry_s;
        const double pre = (w > 0 ? wsum[w - 1] : 0.0) + carry;
        if (i < n) v[i] = x + pre;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = x + pre;
        __syncthreads();
    }
}

// Philox4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t& r0, uint32_t& r1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    r0 = c0;
    r1 = c1;
}

__device__ __forceinline__ double philox_uniform53(uint64_t shot, uint64_t seed, uint32_t stream_id) {
    uint32_t r0, r1;
    philox4x32_10((uint32_t)shot, (uint32_t)(shot >> 32), stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r0, r1);
    const uint64_t bits = (((uint64_t)r0 << 21) ^ ((uint64_t)r1 >> 11)) & ((1ull << 53) - 1ull);
    return (double)bits * 0x1.0p-53;
}

// one warp per shot: binary search over the chunk prefix sums, then a cooperative scan inside the chunk
__global__ void __launch_bounds__(kThreads) sample_kernel(const double2* __restrict__ psi, int chunk_log2,
                                                          const double* __restrict__ cum, uint64_t nchunks,
                                                          uint64_t n_shots, uint64_t seed, uint64_t shot_offset,
                                                          uint64_t* __restrict__ samples) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * kThreads) >> 5;
    const double total = cum[nchunks - 1];
    const uint64_t csize = 1ull << chunk_log2;
    for (uint64_t s = warp; s < n_shots; s += nwarps) {
        const double target = philox_uniform53(s + shot_offset, seed, 0u) * total;
        // first chunk c with cum[c] > target
        uint64_t lo = 0, hi = nchunks - 1;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (cum[mid] > target) hi = mid; else lo = mid + 1;
        }
        const uint64_t c = lo;
        double run = c ? cum[c - 1] : 0.0;
        const double2* p = psi + (c << chunk_log2);
        uint64_t found = csize - 1;  // rounding guard: fall back to the chunk's last amplitude
        bool done = false;
        for (uint64_t k0 = 0; k0 < csize && !done; k0 += 32) {
            const uint64_t k = k0 + lane;
            double x = 0.0;
            if (k < csize) {
                const double2 a = p[k];
                x = a.x * a.x + a.y * a.y;
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, (run + x > target) && (k < csize));
            if (hit) {
                found = k0 + (__ffs(hit) - 1);
                done = true;
            }
            run += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) samples[s] = (c << chunk_log2) + found;
    }
}

// ---- small registers (<= kCdfMaxQubits): the FULL cumulative distribution fits L2 (<= 512 KB), so a shot is ONE
// THREAD doing a binary search over it instead of one warp scanning a chunk: 10^6 shots of a 14-qubit register take
// ~20 us instead of ~450 us (the BeH2-sized sampled energy of BASELINE config 3 draws 942 x 10^6 of them).  Same
// Philox stream and the same rule (first index whose cumulative probability exceeds u * total) as sample_kernel.
constexpr int kCdfMaxQubits = 16;

__global__ void __launch_bounds__(kThreads) prob_kernel(const double2* __restrict__ psi, uint64_t dim, double* __restrict__ p) {
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < dim; i += (uint64_t)gridDim.x * kThreads) {
        const double2 a = psi[i];
        p[i] = a.x * a.x + a.y * a.y;
    }
}

// |<i| B |psi>|^2 for the product basis change B = H on the bits of xm, H S^dagger on the bits of ym (<= kFusedBasisBits
// of them), straight from psi: amplitude_i = 2^{-m/2} sum_s (-1)^{popc(i & s)} (-i)^{popc(s & ym)} psi_{i with the rotated
// bits replaced by s}.  Replaces the scratch copy + one gate pass per rotated qubit + the |.|^2 pass of a small register
// by ONE launch (a QWC group of a molecular Hamiltonian rotates <= 4 qubits).
constexpr int kFusedBasisBits = 6;
struct BasisBits {
    int m;
    int pos[kFusedBasisBits];
};

__global__ void __launch_bounds__(kThreads) rotated_prob_kernel(const double2* __restrict__ psi, uint32_t dim, uint32_t xm,
                                                                uint32_t ym, const BasisBits bb, double* __restrict__ p) {
    const uint32_t rot = xm | ym;
    const double scale = 1.0 / (double)(1u << bb.m);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < dim; i += gridDim.x * kThreads) {
        const uint32_t base = i & ~rot;
        double re = 0.0, im = 0.0;
        for (uint32_t k = 0; k < (1u << bb.m); ++k) {
            uint32_t sbits = 0;
            for (int j = 0; j < bb.m; ++j) sbits |= ((k >> j) & 1u) << bb.pos[j];
            const double2 a = psi[base | sbits];
            const bool neg = __popc(i & sbits) & 1;
            double vr = a.x, vi = a.y;
            switch (__popc(sbits & ym) & 3) {  // times (-i)^{nY}
                case 1: vr = a.y, vi = -a.x; break;
                case 2: vr = -a.x, vi = -a.y; break;
                case 3: vr = -a.y, vi = a.x; break;
                default: break;
            }
            re += neg ? -vr : vr;
            im += neg ? -vi : vi;
        }
        p[i] = (re * re + im * im) * scale;
    }
}

__global__ void __launch_bounds__(kThreads) sample_cdf_kernel(const double* __restrict__ cdf, uint32_t dim, uint64_t n_shots,
                                                              uint64_t seed, uint64_t shot_offset, uint64_t* __restrict__ samples) {
    const double total = cdf[dim - 1];
    for (uint64_t s = (uint64_t)blockIdx.x * kThreads + threadIdx.x; s < n_shots; s += (uint64_t)gridDim.x * kThreads) {
        const double target = philox_uniform53(s + shot_offset, seed, 0u) * total;
        uint32_t lo = 0, hi = dim - 1;  // first index with cdf > target; the last one if rounding leaves none
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&cdf[mid]) > target) hi = mid; else lo = mid + 1;
        }
        samples[s] = lo;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: parity sums and frequencies
// ---------------------------------------------------------------------------------------------------------------
constexpr int kParityMaskTile = 256;

__global__ void __launch_bounds__(kThreads) parity_sums_kernel(const uint64_t* __restrict__ samples,
                                                               const int64_t* __restrict__ counts, uint64_t n,
                                                               const uint64_t* __restrict__ masks, uint32_t J,
                                                               unsigned long long* __restrict__ S) {
    __shared__ unsigned long long sacc[kParityMaskTile];
    __shared__ uint64_t smask[kParityMaskTile];
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    const int lane = threadIdx.x & 31;
    for (uint32_t j0 = 0; j0 < J; j0 += kParityMaskTile) {
        const uint32_t jn = min((uint32_t)kParityMaskTile, J - j0);
        for (uint32_t t = threadIdx.x; t < jn; t += kThreads) {
            sacc[t] = 0ull;
            smask[t] = masks[j0 + t];
        }
        __syncthreads();
        for (uint64_t s0 = (uint64_t)blockIdx.x * kThreads + threadIdx.x; s0 - lane < n; s0 += stride) {
            const bool valid = s0 < n;
            const uint64_t smp = valid ? samples[s0] : 0ull;
            const long long w = valid ? (counts ? (long long)counts[s0] : 1ll) : 0ll;
            for (uint32_t t = 0; t < jn; ++t) {
                const uint64_t v = smp & smask[t];
                const unsigned par = __popc((unsigned)v ^ (unsigned)(v >> 32)) & 1u;
                long long c = par ? -w : w;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0) atomicAdd(&sacc[t], (unsigned long long)c);
            }
        }
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < jn; t += kThreads)
            if (sacc[t]) atomicAdd(&S[j0 + t], sacc[t]);
        __syncthreads();
    }
}

__device__ __forceinline__ uint64_t pext64(uint64_t v, uint64_t mask) {
    uint64_t out = 0;
    int k = 0;
    while (mask) {
        const int b = __ffsll((long long)mask) - 1;
        out |= ((v >> b) & 1ull) << k;
        ++k;
        mask &= mask - 1;
    }
    return out;
}

__global__ void __launch_bounds__(kThreads) histogram_direct_kernel(const uint64_t* __restrict__ samples, uint64_t n,
                                                                    uint64_t keep_mask,
                                                                    unsigned long long* __restrict__ bins) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t s = (uint64_t)blockIdx.x * kThreads + threadIdx.x; s < n; s += stride)
        atomicAdd(&bins[pext64(samples[s], keep_mask)], 1ull);
}

constexpr unsigned long long kEmptyKey = ~0ull;
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(kThreads) histogram_hash_kernel(const uint64_t* __restrict__ samples, uint64_t n,
                                                                  uint64_t keep_mask, unsigned long long* __restrict__ keys,
                                                                  unsigned long long* __restrict__ cnts, uint64_t cap_mask) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t s = (uint64_t)blockIdx.x * kThreads + threadIdx.x; s < n; s += stride) {
        const unsigned long long key = pext64(samples[s], keep_mask);
        uint64_t slot = mix64(key) & cap_mask;
        while (true) {
            const unsigned long long prev = atomicCAS(&keys[slot], kEmptyKey, key);
            if (prev == kEmptyKey || prev == key) {
                atomicAdd(&cnts[slot], 1ull);
                break;
            }
            slot = (slot + 1) & cap_mask;
        }
    }
}

// Per distinct outcome u: v_u = sum_j c_j (-1)^{popc(key_u & m_j)};  partials[b] = sum_u (v_u - mean)^2,
// partials[nb + b] = sum_u count_u (v_u - mean)^2 over CTA b's outcomes (fixed-order reduction: deterministic).
constexpr int kVarTerms = 1024;
__global__ void __launch_bounds__(kThreads) outcome_variance_kernel(const uint64_t* __restrict__ keys,
                                                                    const int64_t* __restrict__ counts, size_t n_unique,
                                                                    const uint64_t* __restrict__ masks,
                                                                    const double* __restrict__ coeffs, uint32_t J,
                                                                    double mean, double* __restrict__ partials) {
    __shared__ uint64_t smask[kVarTerms];
    __shared__ double scoef[kVarTerms];
    const size_t stride = (size_t)gridDim.x * kThreads;
    double acc_u = 0.0, acc_w = 0.0;
    for (size_t u0 = (size_t)blockIdx.x * kThreads; u0 < n_unique; u0 += stride) {
        const size_t u = u0 + threadIdx.x;
        const uint64_t key = u < n_unique ? keys[u] : 0ull;
        double v = 0.0;
        for (uint32_t j0 = 0; j0 < J; j0 += kVarTerms) {
            const uint32_t jn = min((uint32_t)kVarTerms, J - j0);
            __syncthreads();
            for (uint32_t j = threadIdx.x; j < jn; j += kThreads) smask[j] = masks[j0 + j], scoef[j] = coeffs[j0 + j];
            __syncthreads();
            for (uint32_t j = 0; j < jn; ++j) v += flip_sign(scoef[j], key, smask[j]);
        }
        if (u < n_unique) {
            const double d = v - mean;
            acc_u += d * d;
            acc_w += (double)counts[u] * d * d;
        }
    }
    const double tu = block_sum(acc_u);
    const double tw = block_sum(acc_w);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = tu;
        partials[gridDim.x + blockIdx.x] = tw;
    }
}

}  // namespace qc



extern "C" {

int qc_sample(qc_state* h, uint64_t x_basis_mask, uint64_t y_basis_mask, size_t n_shots, uint64_t seed,
              uint64_t shot_offset, uint64_t* samples_out, int samples_on_device, double* total_prob_out) {
    using namespace qc;
    QC_REQUIRE(h, "qc_sample: null handle");
    QC_REQUIRE((x_basis_mask & y_basis_mask) == 0, "qc_sample: a qubit cannot be measured in both X and Y");
    QC_REQUIRE((x_basis_mask | y_basis_mask) < h->dim, "qc_sample: basis mask has bits outside the shard");
    QC_REQUIRE(n_shots == 0 || samples_out, "qc_sample: null output");
    QC_CHECK(cudaSetDevice(h->device));
    const double2* src = h->psi;
    const int n_rot = __builtin_popcountll(x_basis_mask | y_basis_mask);
    const bool fused_small = h->n <= kCdfMaxQubits && n_rot <= kFusedBasisBits;
    if ((x_basis_mask | y_basis_mask) && !fused_small) {
        if (!h->scratch_psi) QC_CHECK(cudaMalloc(&h->scratch_psi, h->dim * sizeof(double2)));
        QC_CHECK(cudaMemcpyAsync(h->scratch_psi, h->psi, h->dim * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream));
        const double r = 0.70710678118654752440;
        Mat2 mh, mhs;  // H and H * S^dagger
        mh.m00 = make_double2(r, 0); mh.m01 = make_double2(r, 0); mh.m10 = make_double2(r, 0); mh.m11 = make_double2(-r, 0);
        mhs.m00 = make_double2(r, 0); mhs.m01 = make_double2(0, -r); mhs.m10 = make_double2(r, 0); mhs.m11 = make_double2(0, r);
        for (int b = 0; b < h->n; ++b) {
            if ((x_basis_mask >> b) & 1ull) { if (apply_gate1q(h, h->scratch_psi, b, mh)) return 1; }
            if ((y_basis_mask >> b) & 1ull) { if (apply_gate1q(h, h->scratch_psi, b, mhs)) return 1; }
        }
        src = h->scratch_psi;
    }
    if (h->n <= kCdfMaxQubits) {
        // small register: full cumulative distribution + one thread per shot
        if (h->ensure_partials(h->dim)) return 1;
        {
            LaunchScope ls(h, kK5Sampling, 16.0 * (double)h->dim);
            const int nb = (int)std::min<uint64_t>((h->dim + kThreads - 1) / kThreads, (uint64_t)h->sm_count * kCtasPerSm);
            if (fused_small && n_rot > 0) {
                BasisBits bb;
                bb.m = 0;
                for (int b = 0; b < h->n; ++b)
                    if (((x_basis_mask | y_basis_mask) >> b) & 1ull) bb.pos[bb.m++] = b;
                rotated_prob_kernel<<<nb, kThreads, 0, h->stream>>>(h->psi, (uint32_t)h->dim, (uint32_t)x_basis_mask,
                                                                    (uint32_t)y_basis_mask, bb, h->d_partials);
            } else {
                prob_kernel<<<nb, kThreads, 0, h->stream>>>(src, h->dim, h->d_partials);
            }
        }
        QC_CHECK(cudaGetLastError());
        {
            LaunchScope ls(h, kK5Sampling, 0.0);
            inclusive_scan_kernel<<<1, 1024, 0, h->stream>>>(h->d_partials, h->dim);
        }
        QC_CHECK(cudaGetLastError());
        if (total_prob_out) {
            if (h->ensure_out(1)) return 1;
            QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_partials + (h->dim - 1), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        if (n_shots) {
            uint64_t* dsamples = samples_out;
            if (!samples_on_device) QC_CHECK(cudaMalloc(&dsamples, n_shots * sizeof(uint64_t)));
            const uint64_t ctas = std::min<uint64_t>((n_shots + kThreads - 1) / kThreads, (uint64_t)h->sm_count * kCtasPerSm);
            {
                LaunchScope ls(h, kK5Sampling, 8.0 * (double)n_shots);
                sample_cdf_kernel<<<(int)ctas, kThreads, 0, h->stream>>>(h->d_partials, (uint32_t)h->dim, n_shots, seed, shot_offset, dsamples);
            }
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess && !samples_on_device)
                e = cudaMemcpyAsync(samples_out, dsamples, n_shots * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (!samples_on_device) cudaFree(dsamples);
            QC_CHECK(e);
        } else {
            QC_CHECK(cudaStreamSynchronize(h->stream));
        }
        if (total_prob_out) {
            QC_CHECK(cudaStreamSynchronize(h->stream));
            *total_prob_out = h->h_out[0];
        }
        return 0;
    }
    const int chunk_log2 = chunk_log2_for(h->n);
    const uint64_t nchunks = h->dim >> chunk_log2;
    if (h->ensure_partials(nchunks)) return 1;
    {
        LaunchScope ls(h, kK5Sampling, 16.0 * (double)h->dim);
        if (chunk_log2 <= 10) {
            const int nb = (int)std::min<uint64_t>((nchunks * 32 + kThreads - 1) / kThreads, (uint64_t)h->sm_count * kCtasPerSm);
            chunk_prob_warp_kernel<<<nb, kThreads, 0, h->stream>>>(src, h->dim, chunk_log2, h->d_partials);
        } else {
            const int nb = (int)std::min<uint64_t>(nchunks, (uint64_t)h->sm_count * kCtasPerSm);
            chunk_prob_kernel<<<nb, kThreads, 0, h->stream>>>(src, h->dim, chunk_log2, h->d_partials);
        }
    }
    QC_CHECK(cudaGetLastError());
    {
        LaunchScope ls(h, kK5Sampling, 0.0);
        inclusive_scan_kernel<<<1, 1024, 0, h->stream>>>(h->d_partials, nchunks);
    }
    QC_CHECK(cudaGetLastError());
    if (total_prob_out) {
        if (h->ensure_out(1)) return 1;
        QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_partials + (nchunks - 1), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (n_shots) {
        uint64_t* dsamples = samples_out;
        if (!samples_on_device) QC_CHECK(cudaMalloc(&dsamples, n_shots * sizeof(uint64_t)));
        const uint64_t warps_needed = n_shots;
        const uint64_t ctas = std::min<uint64_t>((warps_needed * 32 + kThreads - 1) / kThreads, (uint64_t)h->sm_count * kCtasPerSm);
        LaunchScope* ls = new LaunchScope(h, kK5Sampling, 0.0);
        sample_kernel<<<(int)ctas, kThreads, 0, h->stream>>>(src, chunk_log2, h->d_partials, nchunks, n_shots, seed,
                                                             shot_offset, dsamples);
        delete ls;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess && !samples_on_device)
            e = cudaMemcpyAsync(samples_out, dsamples, n_shots * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (!samples_on_device) cudaFree(dsamples);
        QC_CHECK(e);
    } else {
        QC_CHECK(cudaStreamSynchronize(h->stream));
    }
    if (total_prob_out) {
        QC_CHECK(cudaStreamSynchronize(h->stream));
        *total_prob_out = h->h_out[0];
    }
    return 0;
}

int qc_parity_sums(qc_state* h, const uint64_t* samples, const int64_t* counts, size_t n, int on_device,
                   const uint64_t* zmask, size_t J, int64_t* S_out) {
    using namespace qc;
    QC_REQUIRE(h && S_out, "qc_parity_sums: null argument");
    for (size_t j = 0; j < J; ++j) S_out[j] = 0;
    if (J == 0 || n == 0) return 0;
    QC_REQUIRE(samples && zmask, "qc_parity_sums: null array");
    QC_REQUIRE(J < (1ull << 31), "qc_parity_sums: too many masks");
    QC_CHECK(cudaSetDevice(h->device));
    const uint64_t* dsamples = samples;
    const int64_t* dcounts = counts;
    void* staged = nullptr;
    // device staging: [masks J][S J][samples n][counts n]
    const size_t bytes_small = 2 * J * sizeof(uint64_t);
    const size_t bytes_data = on_device ? 0 : n * sizeof(uint64_t) * (counts ? 2 : 1);
    QC_CHECK(cudaMalloc(&staged, bytes_small + bytes_data));
    uint64_t* dmasks = (uint64_t*)staged;
    unsigned long long* dS = (unsigned long long*)staged + J;
    cudaError_t e = cudaMemcpyAsync(dmasks, zmask, J * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dS, 0, J * sizeof(uint64_t), h->stream);
    if (e == cudaSuccess && !on_device) {
        uint64_t* ds = (uint64_t*)staged + 2 * J;
        e = cudaMemcpyAsync(ds, samples, n * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream);
        dsamples = ds;
        if (e == cudaSuccess && counts) {
            e = cudaMemcpyAsync(ds + n, counts, n * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream);
            dcounts = (const int64_t*)(ds + n);
        }
    }
    if (e == cudaSuccess) {
        const int nb = grid_for(n, 8, h->sm_count);
        LaunchScope ls(h, kK3Shots, 8.0 * (double)n);
        parity_sums_kernel<<<nb, kThreads, 0, h->stream>>>(dsamples, dcounts, n, dmasks, (uint32_t)J, dS);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(S_out, dS, J * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(staged);
    QC_CHECK(e);
    return 0;
}

int qc_frequencies(qc_state* h, const uint64_t* samples, size_t n, int on_device, uint64_t keep_mask,
                   uint64_t* keys_out, int64_t* counts_out, size_t cap, size_t* n_unique_out) {
    using namespace qc;
    QC_REQUIRE(h && n_unique_out, "qc_frequencies: null argument");
    *n_unique_out = 0;
    if (n == 0) return 0;
    QC_REQUIRE(samples && keys_out && counts_out, "qc_frequencies: null array");
    QC_CHECK(cudaSetDevice(h->device));
    const int m = __builtin_popcountll(keep_mask);
    const uint64_t* dsamples = samples;
    void* dstage = nullptr;
    if (!on_device) {
        QC_CHECK(cudaMalloc(&dstage, n * sizeof(uint64_t)));
        cudaError_t e = cudaMemcpyAsync(dstage, samples, n * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) { cudaFree(dstage); QC_CHECK(e); }
        dsamples = (const uint64_t*)dstage;
    }
    const int nb = grid_for(n, 4, h->sm_count);
    std::vector<std::pair<uint64_t, int64_t>> found;
    cudaError_t e = cudaSuccess;
    if (m <= 24) {
        const size_t nbins = (size_t)1 << m;
        unsigned long long* dbins = nullptr;
        e = cudaMalloc(&dbins, nbins * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemsetAsync(dbins, 0, nbins * sizeof(unsigned long long), h->stream);
        if (e == cudaSuccess) {
            LaunchScope ls(h, kK3Shots, 8.0 * (double)n);
            histogram_direct_kernel<<<nb, kThreads, 0, h->stream>>>(dsamples, n, keep_mask, dbins);
            e = cudaGetLastError();
        }
        std::vector<unsigned long long> hb(nbins);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hb.data(), dbins, nbins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (dbins) cudaFree(dbins);
        if (e == cudaSuccess)
            for (size_t k = 0; k < nbins; ++k)
                if (hb[k]) found.emplace_back((uint64_t)k, (int64_t)hb[k]);
    } else {
        uint64_t capacity = 1024;
        while (capacity < 2 * n) capacity <<= 1;
        unsigned long long *dk = nullptr, *dc = nullptr;
        e = cudaMalloc(&dk, capacity * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&dc, capacity * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemsetAsync(dk, 0xff, capacity * sizeof(unsigned long long), h->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(dc, 0, capacity * sizeof(unsigned long long), h->stream);
        if (e == cudaSuccess) {
            LaunchScope ls(h, kK3Shots, 8.0 * (double)n);
            histogram_hash_kernel<<<nb, kThreads, 0, h->stream>>>(dsamples, n, keep_mask, dk, dc, capacity - 1);
            e = cudaGetLastError();
        }
        std::vector<unsigned long long> hk(capacity), hc(capacity);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hk.data(), dk, capacity * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hc.data(), dc, capacity * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (dk) cudaFree(dk);
        if (dc) cudaFree(dc);
        if (e == cudaSuccess) {
            for (uint64_t k = 0; k < capacity; ++k)
                if (hc[k]) found.emplace_back((uint64_t)hk[k], (int64_t)hc[k]);
            std::sort(found.begin(), found.end());  // ordering of the output dictionary only
        }
    }
    if (dstage) cudaFree(dstage);
    QC_CHECK(e);
    QC_REQUIRE(found.size() <= cap, "qc_frequencies: output capacity too small");
    for (size_t k = 0; k < found.size(); ++k) {
        keys_out[k] = found[k].first;
        counts_out[k] = found[k].second;
    }
    *n_unique_out = found.size();
    return 0;
}

int qc_outcome_variance(qc_state* h, const uint64_t* keys, const int64_t* counts, size_t n_unique,
                        const uint64_t* zmask, const double* coeff, size_t J, double mean, double* sumsq_unweighted_out,
                        double* sumsq_weighted_out) {
    using namespace qc;
    QC_REQUIRE(h && sumsq_unweighted_out && sumsq_weighted_out, "qc_outcome_variance: null argument");
    *sumsq_unweighted_out = *sumsq_weighted_out = 0.0;
    if (n_unique == 0) return 0;
    QC_REQUIRE(keys && counts && (J == 0 || (zmask && coeff)), "qc_outcome_variance: null array");
    QC_REQUIRE(J < (1ull << 31), "qc_outcome_variance: too many terms");
    QC_CHECK(cudaSetDevice(h->device));
    const int nb = (int)std::min<size_t>((n_unique + kThreads - 1) / kThreads, (size_t)h->sm_count * 4);
    if (h->ensure_partials((size_t)nb * 2) || h->ensure_out(2)) return 1;
    void* staged = nullptr;
    const size_t bytes = n_unique * 16 + J * 16;
    QC_CHECK(cudaMalloc(&staged, std::max<size_t>(bytes, 16)));
    uint64_t* dkeys = (uint64_t*)staged;
    int64_t* dcounts = (int64_t*)(dkeys + n_unique);
    uint64_t* dmasks = (uint64_t*)(dcounts + n_unique);
    double* dcoef = (double*)(dmasks + J);
    cudaError_t e = cudaMemcpyAsync(dkeys, keys, n_unique * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dcounts, counts, n_unique * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && J) e = cudaMemcpyAsync(dmasks, zmask, J * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && J) e = cudaMemcpyAsync(dcoef, coeff, J * 8, cudaMemcpyHostToDevice, h->stream);
    h->h2d_bytes += bytes;
    if (e == cudaSuccess) {
        LaunchScope ls(h, kK3Shots, 16.0 * (double)n_unique);
        outcome_variance_kernel<<<nb, kThreads, 0, h->stream>>>(dkeys, dcounts, n_unique, dmasks, dcoef, (uint32_t)J, mean,
                                                                h->d_partials);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && fold_partials(h, h->d_partials, 2, nb, nb, h->d_out)) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->h_out, h->d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(staged);
    QC_CHECK(e);
    *sumsq_unweighted_out = h->h_out[0];
    *sumsq_weighted_out = h->h_out[1];
    return 0;
}

}  // extern "C"
