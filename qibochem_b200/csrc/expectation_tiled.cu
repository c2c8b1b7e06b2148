// K2T: tiled expectation.  One HBM read of the shard serves EVERY x-mask whose support lies inside the tile's bit set.
//
// The sweep kernel (expectation.cu) reads the whole shard once per distinct x-mask; a molecular Hamiltonian of 10^5
// terms has ~2*10^4 x-masks, so the sweeps are 98 % of an energy evaluation.  Here a PASS fixes a set S of 12 index
// bits (bit 0 + 11 chosen by a greedy cover of the x-masks, k2_plan.cpp); a CTA stages one tile (the 4096 amplitudes
// that differ only in the S bits, 64 KB) in shared memory with cp.async and evaluates every x-group of the pass from
// it.  Inside a pass the groups are bucketed into SUB-CUBES: 5 tile bits R with x inside R.  Each of the 128 threads
// gathers the 32 amplitudes of its R-sub-cube into registers ONCE per sub-cube (32 x LDS.128, bank-conflict free
// through an XOR swizzle) and evaluates all groups of the sub-cube from registers: the pair products
// q = conj(a_p) a_{p^x} need no further shared-memory traffic.
//
// Weights.  For a group (x; terms z_k, d_k) and the canonical element p of a pair inside the sub-cube,
// (-1)^{popc(i & z_k)} factors into  sign(tile origin & z_k outside S) * sign(thread's rest bits & z_k) *
// sign(p & z_k inside R).  Terms that agree outside R form a SUBGROUP whose last factor is tabulated on the host:
// tab[pair] = 2 sum_k d_k (-1)^{popc(p & zR_k)} (no factor 2 on the diagonal) -- a DOT record.  A quad group of a JW
// Hamiltonian (XXYY, YYXX, XYYX, YXXY) is one DOT record (48 DFMA per thread).  A hopping group (X..X / Y..Y decorated
// with one Z_r, ~29 subgroups) shares two sign patterns between its subgroups: the pattern sums A_j = sum_p s_j(p) q_p
// are formed once (DOT records in SET mode) and every subgroup costs one LIN entry (4 DFMA).
//
// Bound: shared-memory bandwidth of the gathers (64 KB per sub-cube per tile) and FP64 issue, not HBM -- one pass
// reads the shard once (16*2^n bytes) and typically serves ~90 x-masks.
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "k2_plan.h"


namespace qc {

struct TiledPlan {
    uint64_t hash = 0;
    HostPlan* host = nullptr;
    uint4* d_units = nullptr;
    // the term list the plan was built from: the plan bakes the coefficients in, so a cache hit needs the lists to be
    // IDENTICAL, not merely to hash alike (H and -H must never share a plan)
    std::vector<uint64_t> key_x, key_z;
    std::vector<double> key_c;
    bool matches(size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs) const {
        return key_x.size() == T && memcmp(key_x.data(), xs, T * 8) == 0 && memcmp(key_z.data(), zs, T * 8) == 0 &&
               memcmp(key_c.data(), cs, T * 8) == 0;
    }
};

void free_tiled_plan(TiledPlan* p) {
    if (!p) return;
    if (p->d_units) cudaFree(p->d_units);
    delete p->host;
    delete p;
}

// ---------------------------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ double apply_sign(double v, unsigned par) {
    return __longlong_as_double(__double_as_longlong(v) ^ ((long long)(par & 1u) << 63));
}

// Amplitude arithmetic for the two register representations of a sub-cube: complex128 (double2) and REAL (double: the
// "k2_real" fast path, taken only when every imaginary part of the shard is exactly zero -- a HF determinant evolved by
// UCC rotations, whose strings all carry an odd number of Y, stays real).  For a real state Re(conj(a) b) is ONE
// multiplication, Im(conj(a) b) vanishes (odd-nY Hamiltonian terms contribute nothing), a sub-cube is 64 registers
// instead of 128 and a tile 32 KB instead of 64 KB.
__device__ __forceinline__ double re0(const double2& a, const double2& b) { return a.x * b.x; }
__device__ __forceinline__ double re1(const double2& a, const double2& b, double pr) { return fma(a.y, b.y, pr); }
__device__ __forceinline__ double im0(const double2& a, const double2& b) { return a.x * b.y; }
__device__ __forceinline__ double im1(const double2& a, const double2& b, double pr) { return fma(-a.y, b.x, pr); }
__device__ __forceinline__ double abs2(const double2& a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ void keep(double2& a) { asm volatile("" : "+d"(a.x), "+d"(a.y)); }
__device__ __forceinline__ double re0(double a, double b) { return a * b; }
__device__ __forceinline__ double re1(double, double, double pr) { return pr; }
__device__ __forceinline__ double im0(double, double) { return 0.0; }
__device__ __forceinline__ double im1(double, double, double pr) { return pr; }
__device__ __forceinline__ double abs2(double a) { return a * a; }
__device__ __forceinline__ void keep(double& a) { asm volatile("" : "+d"(a)); }

// w = sum_k tab[k] * prod_k over the 16 canonical pairs {p, p^XR} of the sub-cube (p ascending, highest bit of XR
// clear); prod = Re or Im of conj(a_p) a_{p^XR}.  Everything is static register indexing; 4 independent FMA chains.
// k-th canonical pair index (p has the highest bit HB of XR clear; ascending)
template <int HB>
__device__ __forceinline__ constexpr int pair_index(int k) {
    return ((k / HB) * 2 * HB) | (k % HB);
}

template <int XR, bool IM, typename A>
__device__ __forceinline__ double dot_pairs(const A (&a)[32], const double2* __restrict__ tab) {
    constexpr int HB = XR >= 16 ? 16 : XR >= 8 ? 8 : XR >= 4 ? 4 : XR >= 2 ? 2 : 1;
    double w[4] = {0.0, 0.0, 0.0, 0.0};
    // two batches of 8 pairs; inside a batch the three dependent steps of a pair are 8 instructions apart
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        double2 t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = tab[4 * b + j];
        double pr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = pair_index<HB>(8 * b + j);
            pr[j] = IM ? im0(a[p], a[p ^ XR]) : re0(a[p], a[p ^ XR]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = pair_index<HB>(8 * b + j);
            pr[j] = IM ? im1(a[p], a[p ^ XR], pr[j]) : re1(a[p], a[p ^ XR], pr[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j & 3] = fma(pr[j], (j & 1) ? t[j >> 1].y : t[j >> 1].x, w[j & 3]);
    }
    return (w[0] + w[1]) + (w[2] + w[3]);
}

// The same as dot_pairs<XR, false> but leaves the four partial sums to the caller, which folds them one record later
// (the dependent tail of a record then overlaps the products of the next one).
template <int XR, typename A>
__device__ __forceinline__ void dot_pairs4(const A (&a)[32], const double2* __restrict__ tab, double (&w)[4]) {
    constexpr int HB = XR >= 16 ? 16 : XR >= 8 ? 8 : XR >= 4 ? 4 : XR >= 2 ? 2 : 1;
    w[0] = w[1] = w[2] = w[3] = 0.0;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        double2 t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = tab[4 * b + j];
        double pr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = pair_index<HB>(8 * b + j);
            pr[j] = re0(a[p], a[p ^ XR]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = pair_index<HB>(8 * b + j);
            pr[j] = re1(a[p], a[p ^ XR], pr[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j & 3] = fma(pr[j], (j & 1) ? t[j >> 1].y : t[j >> 1].x, w[j & 3]);
    }
}

template <int HALF, typename A>
__device__ __forceinline__ double dot_diag(const A (&a)[32], const double2* __restrict__ tab) {
    double2 t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = tab[j];
    double w[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const double pr = abs2(a[k + 16 * HALF]);
        const double tk = (k & 1) ? t[k >> 1].y : t[k >> 1].x;
        w[k & 3] = fma(pr, tk, w[k & 3]);
    }
    return (w[0] + w[1]) + (w[2] + w[3]);
}

// HOP record (k2_plan.h): three tables weight the same 16 products; two accumulators per table => six independent chains
template <int XR, typename A>
__device__ __forceinline__ void dot3(const A (&a)[32], const double2* __restrict__ tab_acc,
                                     const double2* __restrict__ tab_pat, double& w_acc, double& A0, double& A1) {
    constexpr int HB = XR >= 16 ? 16 : XR >= 8 ? 8 : XR >= 4 ? 4 : XR >= 2 ? 2 : 1;
    double w[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        double pr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = pair_index<HB>(8 * b + j);
            pr[j] = re0(a[p], a[p ^ XR]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = pair_index<HB>(8 * b + j);
            pr[j] = re1(a[p], a[p ^ XR], pr[j]);
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            double2 tt[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) tt[j] = t == 0 ? tab_acc[4 * b + j] : tab_pat[8 * (t - 1) + 4 * b + j];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[t][j & 1] = fma(pr[j], (j & 1) ? tt[j >> 1].y : tt[j >> 1].x, w[t][j & 1]);
        }
    }
    w_acc = w[0][0] + w[0][1];
    A0 = w[1][0] + w[1][1];
    A1 = w[2][0] + w[2][1];
}

// sel = xr | im << 5 for DOT records; 0 / 32 = lower / upper half of a diagonal record.  All 64 values are cases, so
// the switch compiles to ONE jump table.
template <typename A>
__device__ __forceinline__ double dot_dispatch(uint32_t sel, const A (&a)[32], const double2* __restrict__ tab) {
#define QC_CASE(X)                              \
    case X: return dot_pairs<X, false>(a, tab); \
    case 32 + X: return dot_pairs<X, true>(a, tab);
    switch (sel & 63u) {
        case 0: return dot_diag<0>(a, tab);
        case 32: return dot_diag<1>(a, tab);
        QC_CASE(1) QC_CASE(2) QC_CASE(3) QC_CASE(4) QC_CASE(5) QC_CASE(6) QC_CASE(7) QC_CASE(8) QC_CASE(9) QC_CASE(10)
        QC_CASE(11) QC_CASE(12) QC_CASE(13) QC_CASE(14) QC_CASE(15) QC_CASE(16) QC_CASE(17) QC_CASE(18) QC_CASE(19)
        QC_CASE(20) QC_CASE(21) QC_CASE(22) QC_CASE(23) QC_CASE(24) QC_CASE(25) QC_CASE(26) QC_CASE(27) QC_CASE(28)
        QC_CASE(29) QC_CASE(30) QC_CASE(31)
    }
#undef QC_CASE
    return 0.0;
}

// (-1)^{popc(base & z_outer) + popc(org & z_rest)}: parity is linear over XOR, so one POPC does it
__device__ __forceinline__ unsigned sign_parity(uint64_t base, uint32_t org, const uint4& hd) {
    const uint32_t lo = (uint32_t)base & hd.x, hi = (uint32_t)(base >> 32) & hd.y;
    return (unsigned)__popc(lo ^ hi ^ (org & hd.z));
}

// One HOP record: ACC table + two pattern sums combined by 32-byte entries; returns the next record
template <int XR, typename A>
__device__ __forceinline__ const uint4* hop_record(const A (&a)[32], const uint4* __restrict__ rp, uint64_t base,
                                                   uint32_t org, uint32_t rho, double& acc0, double& acc1) {
    const uint4 hd = *rp;
    double w, A0, A1;
    const uint32_t two8 = (hd.w & kMetaTwoTables) ? 8u : 0u;  // the ACC table for rho = 1 follows the pattern tables
    dot3<XR>(a, reinterpret_cast<const double2*>((two8 && rho) ? rp + kHopHeadUnits : rp + 1),
             reinterpret_cast<const double2*>(rp + 9), w, A0, A1);
    if ((hd.w & kMetaHopFlip0) && rho) A0 = -A0;
    if ((hd.w & kMetaHopFlip1) && rho) A1 = -A1;
    acc0 += apply_sign(w, sign_parity(base, org, hd));
    const uint32_t cnt = meta_count(hd.w);
    const uint4* ep = rp + kHopHeadUnits + two8;
#pragma unroll 2
    for (uint32_t e = 0; e < cnt; ++e, ep += kHopEntryUnits) {
        const uint4 eh = ep[0];
        const double2 c = *reinterpret_cast<const double2*>(ep + 1);
        const double v = fma(c.x, A0, c.y * A1);
        if (e & 1u)
            acc1 += apply_sign(v, sign_parity(base, org, eh));
        else
            acc0 += apply_sign(v, sign_parity(base, org, eh));
    }
    return ep;
}

// One CTA = 128 threads, 64 KB tile + the pass's blob (sub-cube headers + records, <= 44 KB) in shared memory;
// 2 CTAs per SM.  Everything the inner loops read besides the amplitudes comes from shared memory (broadcast LDS)
// or the constant bank (TPass), so no global-load latency sits on the critical path.
// GENERIC = false: the pass holds simple quads only (the common case) and the record loop is compiled out.
// REAL = true: the "k2_real" fast path (see the amplitude helpers above): the tile holds real parts only (32 KB), a
// sub-cube is 64 registers, so THREE CTAs fit an SM (170 registers per thread, 32 KB + the pass's own blob each).
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

template <bool GENERIC, bool REAL>
__global__ void __launch_bounds__(kTileThreads, REAL ? 3 : 2)
    expectation_tiled_kernel(const double2* __restrict__ psi, int n, const TPass pass, const uint4* __restrict__ units,
                             double* __restrict__ partials) {
    using A = typename std::conditional<REAL, double, double2>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A* tile = reinterpret_cast<A*>(smem_raw);
    const uint4* blob = reinterpret_cast<const uint4*>(smem_raw + (size_t)kTileSize * sizeof(A));
    {
        uint4* dst = reinterpret_cast<uint4*>(smem_raw + (size_t)kTileSize * sizeof(A));
        const uint4* src = units + pass.blob_off;
        for (uint32_t u = threadIdx.x; u < pass.blob_units; u += kTileThreads) dst[u] = __ldg(&src[u]);
    }
    const uint64_t n_tiles = 1ull << (n - kTileBits);
    // amplitude offset of tile slot t = k * 128 + tid: the low 7 tile bits come from tid, the high 5 from k
    uint64_t off_lo = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j) off_lo |= (uint64_t)((threadIdx.x >> j) & 1u) << pass.spos[j];
    double acc0 = 0.0, acc1 = 0.0;

    for (uint64_t tau = blockIdx.x; tau < n_tiles; tau += gridDim.x) {
        // tile origin: deposit tau into the index bits outside S
        uint64_t base = tau;
        {
            uint64_t m = pass.s_mask;
            while (m) {
                const int b = __ffsll((long long)m) - 1;
                base = insert_zero_bit(base, b);
                m &= m - 1;
            }
        }
        __syncthreads();  // previous tile fully consumed (and, the first time, the blob staged)
        const double2* src = psi + base + off_lo;
#pragma unroll
        for (int k = 0; k < kTileSize / kTileThreads; ++k) {
            uint64_t o = 0;
#pragma unroll
            for (int j = 0; j < 5; ++j)
                if (k & (1 << j)) o |= 1ull << pass.spos[7 + j];
            if (REAL)
                cp_async8(&tile[k2_swizzle((uint32_t)k * kTileThreads + threadIdx.x)], &src[o].x);
            else
                cp_async16(&tile[k2_swizzle((uint32_t)k * kTileThreads + threadIdx.x)], src + o);
        }
        cp_async_wait_all();
        __syncthreads();

        for (uint32_t rb = 0; rb < pass.n_rblocks; ++rb) {
            const uint4 b0 = blob[rb * kRBlockUnits];      // sbit[0..4], quad_mask
            const uint4 b1 = blob[rb * kRBlockUnits + 1];  // rec_off, n_recs, n_groups, hop_mask
            // tile-local origin of this thread's sub-cube (and its swizzled slot) from the visit's two origin tables
            const uint32_t* otab = reinterpret_cast<const uint32_t*>(blob + rb * kRBlockUnits + 2);
            const uint32_t packed = otab[threadIdx.x & 15u] ^ otab[16u + (threadIdx.x >> 4)];
            const uint32_t org = packed >> 16, org_sw = packed & 0xffffu;
            const uint32_t sb0 = b0.x & 0xffffu, sb1 = b0.x >> 16, sb2 = b0.y & 0xffffu, sb3 = b0.y >> 16, sb4 = b0.z & 0xffffu;
            A a[32];
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                uint32_t off = org_sw;
                if (p & 1) off ^= sb0;
                if (p & 2) off ^= sb1;
                if (p & 4) off ^= sb2;
                if (p & 8) off ^= sb3;
                if (p & 16) off ^= sb4;
                a[p] = tile[off];
            }
            const uint4* rp = blob + b1.x;
            const uint32_t rho8 = ((threadIdx.x >> 6) & 1u) * 8u;  // two-table records: units to skip for rho = 1
            // ---- simple quads of this cube (coordinate vectors of weight 4 and 3): straight-line code, no dispatch
            const uint32_t quad_mask = b0.z >> 16;
            // software-pipelined: a slot leaves its four partial sums and its sign pending and the next slot folds
            // them, so the dependent tail of a record overlaps the products of the next one; the table address of a
            // record comes from the PREVIOUS header (kMetaNextTwoTables), so its loads do not wait for its own header
            double pw[4] = {0.0, 0.0, 0.0, 0.0};
            unsigned psign = 0u;
            uint32_t two = b0.w & 1u;
#define QC_QUAD(S, XR)                                                                          \
    if (quad_mask & (1u << S)) {                                                                \
        double w4[4];                                                                           \
        dot_pairs4<XR>(a, reinterpret_cast<const double2*>(rp + 1 + two * rho8), w4);           \
        const uint4 hd = *rp;                                                                   \
        acc0 += apply_sign((pw[0] + pw[1]) + (pw[2] + pw[3]), psign);                           \
        pw[0] = w4[0], pw[1] = w4[1], pw[2] = w4[2], pw[3] = w4[3];                             \
        psign = sign_parity(base, org, hd);                                                     \
        rp += kDotUnits + two * 8u;                                                             \
        two = (hd.w >> 13) & 1u;                                                                \
    }
            QC_QUAD(0, 30) QC_QUAD(1, 29) QC_QUAD(2, 27) QC_QUAD(3, 23) QC_QUAD(4, 15)
            QC_QUAD(5, 7) QC_QUAD(6, 11) QC_QUAD(7, 13) QC_QUAD(8, 14) QC_QUAD(9, 19)
            QC_QUAD(10, 21) QC_QUAD(11, 22) QC_QUAD(12, 25) QC_QUAD(13, 26) QC_QUAD(14, 28)
            acc1 += apply_sign((pw[0] + pw[1]) + (pw[2] + pw[3]), psign);
#undef QC_QUAD
            // ---- HOP records (weight-2 x-masks): straight-line code per coordinate vector
            if (GENERIC && b1.w) {
                const uint32_t hop_mask = b1.w, rho = (threadIdx.x >> 6) & 1u;
#define QC_HOP(S, XR) \
    if (hop_mask & (1u << S)) rp = hop_record<XR>(a, rp, base, org, rho, acc0, acc1);
                QC_HOP(0, 1) QC_HOP(1, 2) QC_HOP(2, 4) QC_HOP(3, 8) QC_HOP(4, 16)
                QC_HOP(5, 3) QC_HOP(6, 5) QC_HOP(7, 6) QC_HOP(8, 9) QC_HOP(9, 10)
                QC_HOP(10, 12) QC_HOP(11, 17) QC_HOP(12, 18) QC_HOP(13, 20) QC_HOP(14, 24)
#undef QC_HOP
            }
            // ---- generic records
            double A0 = 0.0, A1 = 0.0, A2 = 0.0, A3 = 0.0;
            for (uint32_t r = 0; GENERIC && r < b1.y; ++r) {
                const uint4 hd = *rp;
                const uint32_t meta = hd.w;
                const uint32_t kind = meta_kind(meta);
                if (kind == kRecLin) {
                    const uint32_t cnt = meta_count(meta);
                    const uint4* ep = rp + 1;
#pragma unroll 2
                    for (uint32_t e = 0; e < cnt; ++e, ep += kLinEntryUnits) {
                        const uint4 eh = ep[0];
                        const double2 c01 = *reinterpret_cast<const double2*>(ep + 1);
                        const double2 c23 = *reinterpret_cast<const double2*>(ep + 2);
                        const double v = fma(c01.x, A0, c01.y * A1) + fma(c23.x, A2, c23.y * A3);
                        if (e & 1u)
                            acc1 += apply_sign(v, sign_parity(base, org, eh));
                        else
                            acc0 += apply_sign(v, sign_parity(base, org, eh));
                    }
                    rp = ep;
                    continue;
                }
                // The products below depend only on a[], which is invariant in this loop: without this barrier the
                // compiler hoists the products of EVERY switch case out of the loop (LICM) and spills.
#pragma unroll
                for (int p = 0; p < 32; ++p) keep(a[p]);
                const uint32_t sel = kind == kRecDot ? (meta & 63u) : (kind == kRecDiagHi ? 32u : 0u);
                const uint32_t two = (meta >> 12) & 1u;
                const double w = dot_dispatch(sel, a, reinterpret_cast<const double2*>(rp + 1 + two * rho8));
                rp += kDotUnits + two * 8u;
                const uint32_t mode = meta_mode(meta);
                if (mode == kModeAcc) {
                    acc0 += apply_sign(w, sign_parity(base, org, hd));
                } else {
                    const uint32_t sl = meta_slot(meta);
                    const double prev = mode == kModeAdd ? (sl == 0 ? A0 : sl == 1 ? A1 : sl == 2 ? A2 : A3) : 0.0;
                    const double nv = prev + w;
                    if (sl == 0) A0 = nv;
                    else if (sl == 1) A1 = nv;
                    else if (sl == 2) A2 = nv;
                    else A3 = nv;
                }
            }
        }
    }
    // CTA reduction (kTileThreads = 128 threads), fixed order => deterministic
    __shared__ double wpart[kTileThreads / 32];
    double acc = warp_sum(acc0 + acc1);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kTileThreads / 32; ++k) t += wpart[k];
        partials[blockIdx.x] = t;
    }
}

// sum_i |Im psi_i| over the shard (exactly 0 <=> the state is real): partials[b] per CTA
__global__ void __launch_bounds__(kThreads) imag_abs_sum_kernel(const double2* __restrict__ psi, uint64_t dim,
                                                                double* __restrict__ partials) {
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < dim; i += (uint64_t)gridDim.x * kThreads)
        acc += fabs(psi[i].y);
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

static int imag_abs_sum(qc_state* h, double* out) {
    const int nb = grid_for(h->dim, 4, h->sm_count);
    if (h->ensure_partials((size_t)nb) || h->ensure_out(1)) return 1;
    {
        LaunchScope ls(h, kReduce, 16.0 * (double)h->dim);
        imag_abs_sum_kernel<<<nb, kThreads, 0, h->stream>>>(h->psi, h->dim, h->d_partials);
    }
    QC_CHECK(cudaGetLastError());
    if (fold_partials(h, h->d_partials, 1, nb, nb, h->d_out)) return 1;
    QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->d2h_bytes += sizeof(double);
    QC_CHECK(cudaStreamSynchronize(h->stream));
    *out = h->h_out[0];
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
// word-wise hash with a full finalizer per word (splitmix64), so that high input bits -- the sign and exponent of a
// coefficient -- diffuse into every bit of the hash; a hit is confirmed by comparing the term lists themselves
static uint64_t hash_words(const void* data, size_t bytes, uint64_t h) {
    const uint64_t* w = (const uint64_t*)data;
    const size_t nw = bytes / 8;
    for (size_t i = 0; i < nw; ++i) {
        uint64_t v = w[i] + 0x9e3779b97f4a7c15ull + h;
        v = (v ^ (v >> 30)) * 0xbf58476d1ce4e5b9ull;
        v = (v ^ (v >> 27)) * 0x94d049bb133111ebull;
        h = (h << 7 | h >> 57) ^ v ^ (v >> 31);
    }
    return h;
}

static TiledPlan* upload_plan(HostPlan* hp, uint64_t hash) {
    TiledPlan* plan = new TiledPlan();
    plan->hash = hash;
    plan->host = hp;
    const size_t bytes = std::max<size_t>(hp->units.size(), 1) * sizeof(Unit16);
    if (cudaMalloc(&plan->d_units, bytes) != cudaSuccess ||
        cudaMemcpy(plan->d_units, hp->units.data(), hp->units.size() * sizeof(Unit16), cudaMemcpyHostToDevice) != cudaSuccess) {
        free_tiled_plan(plan);
        return nullptr;
    }
    return plan;
}

// Tiled evaluation of sum_k c_k Re<psi|P_k|psi>.  The device-side sum of all passes lands in h->d_out[slot].
// Leftover terms (not tileable) are returned through the plan for the sweep kernel.
int run_tiled_expectation(qc_state* h, size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs, int out_slot,
                          int* n_passes_out, const TiledPlan** plan_out) {
    uint64_t hash = 0xcbf29ce484222325ull ^ (uint64_t)T ^ ((uint64_t)h->n << 56);
    hash = hash_words(xs, T * 8, hash);
    hash = hash_words(zs, T * 8, hash);
    hash = hash_words(cs, T * 8, hash);
    constexpr size_t kMaxCachedPlans = 8;  // a sharded energy walks through a few layouts, each with its own term list
    TiledPlan* found = nullptr;
    for (size_t i = 0; i < h->k2plans.size(); ++i)
        if (h->k2plans[i]->hash == hash && h->k2plans[i]->host->n == h->n && h->k2plans[i]->matches(T, xs, zs, cs)) {
            found = h->k2plans[i];
            h->k2plans.erase(h->k2plans.begin() + i);
            break;
        }
    if (!found) {
        HostPlan* hp = build_host_plan(h->n, T, xs, zs, cs);
        QC_REQUIRE(hp != nullptr, "qc_expectation: could not build the tiled plan");
        found = upload_plan(hp, hash);
        QC_REQUIRE(found != nullptr, "qc_expectation: could not upload the tiled plan");
        found->key_x.assign(xs, xs + T);
        found->key_z.assign(zs, zs + T);
        found->key_c.assign(cs, cs + T);
        h->h2d_bytes += hp->units.size() * sizeof(Unit16);
        if (h->k2plans.size() >= kMaxCachedPlans) {
            QC_CHECK(cudaStreamSynchronize(h->stream));  // the evicted plan may still be read by in-flight passes
            free_tiled_plan(h->k2plans.front());
            h->k2plans.erase(h->k2plans.begin());
        }
    }
    h->k2plans.push_back(found);  // most recently used last
    const TiledPlan* plan = found;
    const HostPlan* hp = plan->host;
    // "k2_real": the real-amplitude kernels, only when every imaginary part of the shard is EXACTLY zero (checked here,
    // one read of the shard per call); anything else runs the complex kernels
    bool real = false;
    if (h->k2_real) {
        double im_sum = 1.0;
        if (imag_abs_sum(h, &im_sum)) return 1;
        real = im_sum == 0.0;
    }
    h->last_k2_real = real ? 1 : 0;
    const size_t smem_complex = (size_t)kTileSize * sizeof(double2) + (size_t)kBlobCapUnits * 16;
    const size_t smem_real_max = (size_t)kTileSize * sizeof(double) + (size_t)kBlobCapUnits * 16;
    const int ctas = real ? 3 : 2;
    for (auto kernel : {expectation_tiled_kernel<true, false>, expectation_tiled_kernel<false, false>}) {
        QC_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_complex));
        QC_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    for (auto kernel : {expectation_tiled_kernel<true, true>, expectation_tiled_kernel<false, true>}) {
        QC_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_real_max));
        QC_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    const uint64_t n_tiles = 1ull << (h->n - kTileBits);
    const int nb = (int)std::min<uint64_t>(n_tiles, (uint64_t)h->sm_count * ctas);
    const size_t np = hp->passes.size();
    if (h->ensure_partials(np * nb) || h->ensure_out((size_t)out_slot + 1)) return 1;
    for (size_t p = 0; p < np; ++p) {
        const TPass& pass = hp->passes[p];
        // algorithmic bytes: one shard read per x-group served; the pass itself reads the shard ONCE, so
        // the actual HBM traffic of this class is launches * 16 * dim
        {
            LaunchScope ls(h, kK2Tiled, 16.0 * (double)h->dim * (double)pass.n_groups);
            h->h2d_bytes += sizeof(TPass);
            if (real) {
                // shared memory = the tile + THIS pass's blob: passes whose blob stays below ~43 KB run three CTAs per SM
                const size_t smem = (size_t)kTileSize * sizeof(double) + (((size_t)pass.blob_units * 16 + 1023) & ~(size_t)1023);
                auto kernel = pass.n_generic ? expectation_tiled_kernel<true, true> : expectation_tiled_kernel<false, true>;
                // a persistent grid must not exceed what is resident at once (a pass with a full 44 KB blob fits two
                // CTAs per SM, the others three); unused partial slots of the pass are zeroed
                int occ = 2;
                QC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kTileThreads, smem));
                const int nbp = (int)std::min<uint64_t>(n_tiles, (uint64_t)h->sm_count * (uint64_t)std::max(1, std::min(occ, ctas)));
                if (nbp < nb) QC_CHECK(cudaMemsetAsync(h->d_partials + p * nb + nbp, 0, (size_t)(nb - nbp) * sizeof(double), h->stream));
                kernel<<<nbp, kTileThreads, smem, h->stream>>>(h->psi, h->n, pass, plan->d_units, h->d_partials + p * nb);
            } else {
                auto kernel = pass.n_generic ? expectation_tiled_kernel<true, false> : expectation_tiled_kernel<false, false>;
                kernel<<<nb, kTileThreads, smem_complex, h->stream>>>(h->psi, h->n, pass, plan->d_units, h->d_partials + p * nb);
            }
        }
        QC_CHECK(cudaGetLastError());
    }
    if (np) {
        if (fold_partials(h, h->d_partials, 1, (int)(np * nb), (int)(np * nb), h->d_out + out_slot)) return 1;
    } else {
        QC_CHECK(cudaMemsetAsync(h->d_out + out_slot, 0, sizeof(double), h->stream));
    }
    if (n_passes_out) *n_passes_out = (int)np;
    if (plan_out) *plan_out = plan;
    return 0;
}

void tiled_plan_leftovers(const TiledPlan* plan, const uint64_t** x, const uint64_t** z, const double** c, size_t* count) {
    *x = plan->host->left_x.data();
    *z = plan->host->left_z.data();
    *c = plan->host->left_c.data();
    *count = plan->host->left_x.size();
}

}  // namespace qc

// Planner statistics + self-check; host only (usable without a GPU).  stats_out: 16 x uint64
// {x-masks, tiled x-masks, passes, sub-cubes, DOT records, LIN entries, leftover terms, record units,
//  sub-cubes holding 0,1,...,>=7 x-masks}.  check_out (may be NULL): see check_host_plan.
extern "C" int qc_k2_plan_stats(int n_local_qubits, size_t T, const uint64_t* xmask, const uint64_t* zmask,
                                const double* coeff, uint64_t* stats_out, double* check_out) {
    using namespace qc;
    QC_REQUIRE(xmask && zmask && coeff && stats_out, "qc_k2_plan_stats: null argument");
    QC_REQUIRE(n_local_qubits >= kTileBits && n_local_qubits <= 40, "qc_k2_plan_stats: needs 12..40 local qubits");
    for (size_t k = 0; k < T; ++k)
        QC_REQUIRE((xmask[k] >> n_local_qubits) == 0 && (zmask[k] >> n_local_qubits) == 0,
                   "qc_k2_plan_stats: mask has bits outside the shard");
    HostPlan* p = build_host_plan(n_local_qubits, T, xmask, zmask, coeff);
    QC_REQUIRE(p != nullptr, "qc_k2_plan_stats: planner failed");
    const PlanStats& s = p->stats;
    const uint64_t v[8] = {s.n_groups, s.n_tiled_groups, s.n_passes, s.n_rblocks, s.n_dot, s.n_lin, s.n_leftover_terms,
                           (uint64_t)p->units.size()};
    for (int i = 0; i < 8; ++i) stats_out[i] = v[i], stats_out[8 + i] = s.rblock_hist[i];
    if (check_out) *check_out = check_host_plan(p, T, xmask, zmask, coeff, 20000, 12345);
    delete p;
    return 0;
}
