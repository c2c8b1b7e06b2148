// K1: fused Pauli rotations  psi <- prod_k exp(-i phi_k P_k / 2) psi  for a run of commuting strings that share
// one x-mask, in ONE in-place pass over the shard (pairs i <-> i^x).
//
// Reference behaviour replaced: qibochem `_expi_pauli` (src/qibochem/ansatz/_ansatz.py:55-92) expands every
// string into basis changes + CNOT ladder + RZ + inverse, and qibo's backend executes each of those 16-40 gates
// as a full-state pass.  `ucc_circuit` (src/qibochem/ansatz/ansatz.py:340-349) emits 2 (single) or 8 (double)
// such strings per excitation; the 8 strings of a double share x and commute, so they fuse to one pass here.
//
// Math.  On the pair {i, j=i^x} (i has the highest x bit clear) every string of the run acts as
// [[0, al_k],[conj(al_k), 0]] with al_k = (-i)^{nY_k} (-1)^{popc(i & z_k)}.  Commuting same-x strings have
// al_k = eps_k(i) al_0, eps = +-1, so the product of their exponentials is ONE rotation of angle
// Phi(i) = sum_k eps_k(i) phi_k about string 0.  eps_k(i) depends on i only through the bits of
// dz_k = z_k ^ z_0; Phi is tabulated on the host over the union support of the dz_k (<= kMaxSupport bits) and
// the kernel looks up (cos, sin)(Phi/2) per pair.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <functional>

#include "common.cuh"

namespace qc {

constexpr int kMaxSupport = 10;  // table of <= 1024 (cos, sin) entries per fused run
constexpr int kMaxRun = 256;     // strings per fused run

struct RotDesc {
    uint64_t x;         // pairing mask (0 => diagonal run)
    uint64_t z0;        // z-mask of the run's first string
    uint64_t sup_mask;  // union support of z_k ^ z_0, pair bit removed
    int hb;             // highest set bit of x (-1 if x == 0)
    double war, wai;    // a' = c a + f * (wa * b),   wa = (-i)^{nY0 + 1}
    double wbr, wbi;    // b' = c b + f * (wb * a),   wb = (-i)^{1 - nY0};   f = sin(Phi/2) * (-1)^{popc(i & z0)}
    uint32_t table_off; // first table entry (double2 units)
};

__device__ __forceinline__ double2 cmul(double wr, double wi, double2 v) {
    return make_double2(wr * v.x - wi * v.y, wr * v.y + wi * v.x);
}

// ---- general case: highest x bit >= 5, one thread per amplitude pair, U pairs in flight ----------------------
template <int U>
__global__ void __launch_bounds__(kThreads) rot_pairs_kernel(double2* __restrict__ psi, uint64_t npairs,
                                                             const RotDesc d, const double2* __restrict__ table) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    const double2* tab = table + d.table_off;
    for (uint64_t p0 = (uint64_t)blockIdx.x * kThreads + threadIdx.x; p0 < npairs; p0 += stride * U) {
        uint64_t idx[U];
        double2 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t p = p0 + (uint64_t)u * stride;
            idx[u] = insert_zero_bit(p, d.hb);
            if (p < npairs) {
                a[u] = psi[idx[u]];
                b[u] = psi[idx[u] ^ d.x];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t p = p0 + (uint64_t)u * stride;
            if (p < npairs) {
                const double2 cs = __ldg(&tab[d.sup_mask ? gather_bits(idx[u], d.sup_mask) : 0u]);
                const double f = flip_sign(cs.y, idx[u], d.z0);
                const double2 wb = cmul(d.war, d.wai, b[u]);
                const double2 wa = cmul(d.wbr, d.wbi, a[u]);
                psi[idx[u]] = make_double2(cs.x * a[u].x + f * wb.x, cs.x * a[u].y + f * wb.y);
                psi[idx[u] ^ d.x] = make_double2(cs.x * b[u].x + f * wa.x, cs.x * b[u].y + f * wa.y);
            }
        }
    }
}

// ---- x confined to the low 5 index bits: one thread per amplitude, partner through a warp shuffle ------------
__global__ void __launch_bounds__(kThreads) rot_lowx_kernel(double2* __restrict__ psi, uint64_t dim, const RotDesc d,
                                                            const double2* __restrict__ table) {
    const double2* tab = table + d.table_off;
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    const int xl = (int)d.x;  // < 32
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: every lane takes part in the shuffle even when the shard has < 32 amplitudes
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i - lane < dim; i += stride) {
        const bool valid = i < dim;
        const double2 a = valid ? psi[i] : make_double2(0.0, 0.0);
        double2 b;
        b.x = __shfl_xor_sync(0xffffffffu, a.x, xl);
        b.y = __shfl_xor_sync(0xffffffffu, a.y, xl);
        if (valid) {
            const uint64_t ci = ((i >> d.hb) & 1ull) ? (i ^ d.x) : i;
            const double2 cs = __ldg(&tab[d.sup_mask ? gather_bits(ci, d.sup_mask) : 0u]);
            const double f = flip_sign(cs.y, i, d.z0);
            const double2 wb = cmul(d.war, d.wai, b);
            psi[i] = make_double2(cs.x * a.x + f * wb.x, cs.x * a.y + f * wb.y);
        }
    }
}

// ---- x == 0: diagonal run, psi_i *= exp(-i s0(i) Phi(i) / 2) -------------------------------------------------
template <int U>
__global__ void __launch_bounds__(kThreads) rot_diag_kernel(double2* __restrict__ psi, uint64_t dim, const RotDesc d,
                                                            const double2* __restrict__ table) {
    const double2* tab = table + d.table_off;
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t i0 = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i0 < dim; i0 += stride * U) {
        double2 a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            if (i < dim) a[u] = psi[i];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            if (i < dim) {
                const double2 cs = __ldg(&tab[d.sup_mask ? gather_bits(i, d.sup_mask) : 0u]);
                const double ms = -flip_sign(cs.y, i, d.z0);  // multiply by (c - i s0 s)
                psi[i] = make_double2(cs.x * a[u].x - ms * a[u].y, cs.x * a[u].y + ms * a[u].x);
            }
        }
    }
}

// ---- small registers (n <= kSmallMaxQubits): the WHOLE program in one launch, state in shared memory ---------------
// A 12-qubit LiH-sized UCCSD program is ~90 fused runs over a 64 KB state: one launch per run makes it launch-bound
// (~6 us per run).  Here ONE CTA of 1024 threads keeps the state in shared memory and applies every run of the call in
// order (one __syncthreads per run); reference behaviour replaced: the 15 108 gate passes of qibo's gate loop for the
// 12-qubit circuit_ansatz.
constexpr int kSmallMaxQubits = 13;   // 2^13 x 16 B = 128 KB of shared memory
constexpr int kSmallThreads = 1024;

__device__ __forceinline__ uint32_t pext32_small(uint32_t v, uint32_t mask) {
    uint32_t out = 0, k = 0;
    while (mask) {
        const int b = __ffs((int)mask) - 1;
        out |= ((v >> b) & 1u) << k;
        ++k;
        mask &= mask - 1u;
    }
    return out;
}
__device__ __forceinline__ double sign32(double v, int par) {
    return __longlong_as_double(__double_as_longlong(v) ^ ((long long)(par & 1) << 63));
}

// STAGED = true: the run descriptors and their (cos, sin) tables sit in shared memory behind the state (they fit for the
// LiH / BeH2-sized programs); from global memory the two dependent loads per run (descriptor, then table entry) cost
// ~1.5 us per run with nothing to hide them behind -- one CTA, one run at a time.
template <bool STAGED>
__global__ void __launch_bounds__(kSmallThreads, 1)
    rot_small_kernel(double2* __restrict__ psi, int n, int n_runs, const RotDesc* __restrict__ descs,
                     const double2* __restrict__ table, uint32_t n_table, long long basis_index) {
    extern __shared__ __align__(16) unsigned char small_raw[];
    double2* s = reinterpret_cast<double2*>(small_raw);
    const uint32_t dim = 1u << n, npairs = dim >> 1;
    // basis_index >= 0: start from that basis state (K0 fused into the launch); -1: from the amplitudes in HBM;
    // -2: from the all-zero vector (a shard that does not hold the basis state)
    if (basis_index == -1) {
        for (uint32_t i = threadIdx.x; i < dim; i += kSmallThreads) s[i] = psi[i];
    } else {
        for (uint32_t i = threadIdx.x; i < dim; i += kSmallThreads)
            s[i] = make_double2((long long)i == basis_index ? 1.0 : 0.0, 0.0);
    }
    const double2* tab_base = table;
    const RotDesc* dsc = descs;
    if (STAGED) {
        double2* stab = s + dim;
        uint64_t* sdesc = reinterpret_cast<uint64_t*>(stab + n_table);
        for (uint32_t i = threadIdx.x; i < n_table; i += kSmallThreads) stab[i] = table[i];
        const uint64_t* gdesc = reinterpret_cast<const uint64_t*>(descs);
        const uint32_t n_units = (uint32_t)n_runs * (uint32_t)(sizeof(RotDesc) / 8);
        for (uint32_t i = threadIdx.x; i < n_units; i += kSmallThreads) sdesc[i] = gdesc[i];
        tab_base = stab;
        dsc = reinterpret_cast<const RotDesc*>(sdesc);
    }
    __syncthreads();
    for (int r = 0; r < n_runs; ++r) {
        const RotDesc d = dsc[r];  // warp-uniform loads
        const double2* tab = tab_base + d.table_off;
        // a register of <= 13 qubits: all index arithmetic in 32 bits
        const uint32_t x32 = (uint32_t)d.x, z32 = (uint32_t)d.z0, sup32 = (uint32_t)d.sup_mask;
        if (x32 == 0u) {
            for (uint32_t i = threadIdx.x; i < dim; i += kSmallThreads) {
                const double2 cs = tab[sup32 ? pext32_small(i, sup32) : 0u];
                const double ms = -sign32(cs.y, __popc(i & z32));
                const double2 a = s[i];
                s[i] = make_double2(cs.x * a.x - ms * a.y, cs.x * a.y + ms * a.x);
            }
        } else {
            const uint32_t lowmask = (1u << d.hb) - 1u;
            for (uint32_t p = threadIdx.x; p < npairs; p += kSmallThreads) {
                const uint32_t i = ((p & ~lowmask) << 1) | (p & lowmask), j = i ^ x32;
                const double2 a = s[i], b = s[j];
                const double2 cs = tab[sup32 ? pext32_small(i, sup32) : 0u];
                const double f = sign32(cs.y, __popc(i & z32));
                const double2 wb = cmul(d.war, d.wai, b);
                const double2 wa = cmul(d.wbr, d.wbi, a);
                s[i] = make_double2(cs.x * a.x + f * wb.x, cs.x * a.y + f * wb.y);
                s[j] = make_double2(cs.x * b.x + f * wa.x, cs.x * b.y + f * wa.y);
            }
        }
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < dim; i += kSmallThreads) psi[i] = s[i];
}

// ---- K1T: a BATCH of consecutive fused runs applied in one in-place HBM pass --------------------------------------
// When the x-masks of consecutive runs fit, together with index bits 0..2 (128-byte runs in HBM), into a set S of 12
// index bits, a CTA stages the 4096 amplitudes that differ only in the S bits (64 KB) in shared memory, applies every
// run of the batch there (pairs t <-> t ^ xl inside the tile; signs and table indices split into a per-tile part
// taken from the tile origin and a per-pair part from the tile-local index) and writes the tile back: one read and
// one write of the shard per BATCH instead of per run.  UCCSD streams batch ~4.6 excitations (37 strings) per pass.
constexpr int kRotTileBits = 12;
constexpr int kRotTileSize = 1 << kRotTileBits;
constexpr int kRotTileThreads = 256;
constexpr int kMaxBatchRuns = 16;

struct TileRun {
    uint64_t z0_outer;    // z0 outside S
    uint64_t sup_outer;   // support bits outside S (index positions)
    uint32_t xl;          // x in tile-local positions (0 => diagonal run)
    uint32_t z0_l;        // z0 inside S, tile-local
    uint32_t sup_l;       // support bits inside S, tile-local (pair bit removed)
    uint32_t table_off;   // first table entry (double2 units); index = pext(origin, sup_outer) << n_sup_l | pext(t, sup_l)
    int8_t hb_l;          // highest set bit of xl (-1 if xl == 0)
    int8_t n_sup_l;
    int8_t ma, mb;        // a' = c a + f (-i)^ma b,   b' = c b + f (-i)^mb a
    // per slot group k (warp-uniform part of a thread's slots): tile slot bits | pext(bits, sup_l) << 12 | parity(bits & z0_l) << 31
    uint32_t kpack[kRotTileSize / kRotTileThreads];
};
struct TileBatch {
    uint64_t s_mask;
    uint8_t spos[kRotTileBits];
    int n_runs;
    TileRun run[kMaxBatchRuns];
};

__device__ __forceinline__ uint32_t pext32(uint32_t v, uint32_t mask) {
    uint32_t out = 0;
    int k = 0;
    while (mask) {
        const int b = __ffs((int)mask) - 1;
        out |= ((v >> b) & 1u) << k;
        ++k;
        mask &= mask - 1;
    }
    return out;
}
// (-i)^m * v
__device__ __forceinline__ double2 mul_pow_minus_i(int m, double2 v) {
    switch (m & 3) {
        case 0: return v;
        case 1: return make_double2(v.y, -v.x);
        case 2: return make_double2(-v.x, -v.y);
        default: return make_double2(-v.y, v.x);
    }
}

constexpr int kTileRunTable = 16;  // (cos, sin) entries per run staged in shared memory (support <= 4 bits)

__global__ void __launch_bounds__(kRotTileThreads, 3)
    rot_tiled_kernel(double2* __restrict__ psi, int n, const TileBatch batch, const double2* __restrict__ table) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    double2* stab = tile + kRotTileSize;                                                      // [n_runs][kTileRunTable]
    uint8_t* spre = reinterpret_cast<uint8_t*>(stab + kMaxBatchRuns * kTileRunTable);         // [n_runs][threads]
    for (int e = threadIdx.x; e < batch.n_runs * kTileRunTable; e += kRotTileThreads) {
        const TileRun& d = batch.run[e / kTileRunTable];
        const int k = e % kTileRunTable;
        if (k < (1 << (d.n_sup_l + __popcll(d.sup_outer)))) stab[e] = __ldg(&table[d.table_off + k]);
    }
    // The part of a slot's table index and sign that comes from the thread id does not depend on the tile: computed once.
    for (int r = 0; r < batch.n_runs; ++r) {
        const TileRun& d = batch.run[r];
        const uint32_t t0 = d.xl ? (uint32_t)insert_zero_bit(threadIdx.x, d.hb_l) : threadIdx.x;
        spre[r * kRotTileThreads + threadIdx.x] =
            (uint8_t)((d.sup_l ? pext32(t0, d.sup_l) : 0u) | ((__popc(t0 & d.z0_l) & 1u) << 7));
    }
    const uint64_t n_tiles = 1ull << (n - kRotTileBits);
    // amplitude offset of tile slot t = k * 256 + tid: low 8 tile bits from tid, high 4 from k
    uint64_t off_lo = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) off_lo |= (uint64_t)((threadIdx.x >> j) & 1u) << batch.spos[j];

    for (uint64_t tau = blockIdx.x; tau < n_tiles; tau += gridDim.x) {
        uint64_t base = tau;
        {
            uint64_t m = batch.s_mask;
            while (m) {
                const int b = __ffsll((long long)m) - 1;
                base = insert_zero_bit(base, b);
                m &= m - 1;
            }
        }
        double2* src = psi + base + off_lo;
        __syncthreads();  // the previous tile has been written back (first time: tables staged)
#pragma unroll
        for (int k = 0; k < kRotTileSize / kRotTileThreads; ++k) {
            uint64_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k & (1 << j)) o |= 1ull << batch.spos[8 + j];
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(&tile[k * kRotTileThreads + threadIdx.x]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(src + o));
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();

        for (int r = 0; r < batch.n_runs; ++r) {
            const TileRun& d = batch.run[r];
            const double2* tab = stab + r * kTileRunTable;
            uint32_t outer_idx = 0;
            if (d.sup_outer) outer_idx = gather_bits(base, d.sup_outer) << d.n_sup_l;
            const uint32_t pre = spre[r * kRotTileThreads + threadIdx.x];
            const uint32_t idx0 = outer_idx | (pre & 0x7fu);
            const unsigned par0 = (unsigned)__popcll(base & d.z0_outer) + (pre >> 7);
            if (d.xl == 0) {
                const uint32_t t0 = threadIdx.x;
#pragma unroll
                for (int k = 0; k < kRotTileSize / kRotTileThreads; ++k) {
                    const uint32_t kp = d.kpack[k];
                    const uint32_t t = t0 | (kp & 0xfffu);
                    const double2 a = tile[t];
                    const double2 cs = tab[idx0 | ((kp >> 12) & 0x3ffu)];
                    const double ms = ((par0 + (kp >> 31)) & 1u) ? cs.y : -cs.y;  // multiply by (c - i s0 s)
                    tile[t] = make_double2(cs.x * a.x - ms * a.y, cs.x * a.y + ms * a.x);
                }
            } else {
                const uint32_t t0 = (uint32_t)insert_zero_bit(threadIdx.x, d.hb_l);
#pragma unroll
                for (int k = 0; k < kRotTileSize / 2 / kRotTileThreads; ++k) {
                    const uint32_t kp = d.kpack[k];
                    const uint32_t t = t0 | (kp & 0xfffu), u = t ^ d.xl;
                    const double2 a = tile[t], b = tile[u];
                    const double2 cs = tab[idx0 | ((kp >> 12) & 0x3ffu)];
                    const double f = ((par0 + (kp >> 31)) & 1u) ? -cs.y : cs.y;
                    const double2 ub = mul_pow_minus_i(d.ma, b), ua = mul_pow_minus_i(d.mb, a);
                    tile[t] = make_double2(cs.x * a.x + f * ub.x, cs.x * a.y + f * ub.y);
                    tile[u] = make_double2(cs.x * b.x + f * ua.x, cs.x * b.y + f * ua.y);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < kRotTileSize / kRotTileThreads; ++k) {
            uint64_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k & (1 << j)) o |= 1ull << batch.spos[8 + j];
            src[o] = tile[k * kRotTileThreads + threadIdx.x];
        }
    }
}

// ---- K1O: a PHASE of consecutive fused runs applied from registers in one in-place HBM pass ------------------------
// The x-masks of a phase are linearly independent over GF(2) (at most 5 distinct ones, each possibly used by several
// runs, plus diagonal runs).  They generate a group of 2^d index translations; a thread loads the 32 amplitudes of
// 2^(5-d) orbits of that group (element e = (orbit << d) | E lives at  c ^ XOR_{i in E} x_i,  c = the orbit's
// representative: the thread's global number deposited into the index bits that are not pivots of the x-masks),
// applies EVERY run of the phase to them -- run i pairs element E with E ^ (1 << basis(i)), a compile-time butterfly
// -- and stores them back: one read and one write of the shard per phase, no shared-memory traffic.  Loads and stores
// stay coalesced because consecutive lanes own consecutive representatives and an x-mask only permutes a 32-lane
// block.  UCCSD streams give 5 excitations (40 strings) per pass.
constexpr int kOrbitBits = 5;
constexpr int kOrbitSize = 1 << kOrbitBits;
constexpr int kOrbitThreads = 128;
constexpr int kMaxPhaseRuns = 16;
constexpr int kPhaseTable = 16;  // (cos, sin) entries per run staged in shared memory (support <= 4 bits)

struct PhaseRun {
    uint64_t z0;         // z-mask of the run's first string
    uint64_t sup_mask;   // union support of z_k ^ z_0, pair bit removed (table index = its bits of the canonical element)
    uint32_t table_off;  // first (cos, sin) entry (double2 units)
    uint32_t kx;         // key of the pairing mask x itself | swap flag (see sel)
    int8_t hb;           // highest set bit of x (-1 if x == 0: diagonal run)
    int8_t basis;        // which basis vector x is (-1 for a diagonal run)
    int8_t ma;           // a' = c a + f (-i)^ma b,  b' = c b + f (-i)^(2-ma) a   for the pair (a = hb clear, b)
    int8_t pad[5];
    // A pair's table index, sign parity and orientation are XOR-linear in the element's index c ^ elem[e] -- also
    // through the "canonical element = the one with bit hb clear" rule, because substituting e ^ x XORs a constant
    // (kx).  key = table index (bits 0..3) | sign parity (bit 4) | swap flag (bit 5);  key(thread, e) = ckey ^ sel[e]
    // with ckey from the thread's representative.  Butterfly runs use sel[j], j = 0..15 = the j-th element with the
    // run's basis bit clear; diagonal runs use all 32.
    uint8_t sel[32];
};
struct PhaseDesc {
    uint64_t elem[kOrbitSize];  // index offset of register element e relative to the thread's first representative
    uint8_t pivot[kOrbitBits];  // pivot bit of each basis vector, ascending
    int n_pivots;
    int n_runs;
    PhaseRun run[kMaxPhaseRuns];
};

constexpr int kKeyTable = 64;  // per run: (cos, signed sin) for every key

// Run on basis vector I: the j-th element e with bit I clear pairs with e | 1 << I.  ODD = ma is odd (w = -+i), else
// w = +-1.  tb = the run's key table, sl = its 16 selector words (byte offsets into tb), ck = the thread's key << 4.
// All sign handling lives in the table, so a pair costs one LOP, one LDS.128 and 8 FP64 instructions.
template <int I, bool ODD>
__device__ __forceinline__ void orbit_butterfly(double2 (&a)[kOrbitSize], const char* __restrict__ tb,
                                                const uint4* __restrict__ sl, uint32_t ck) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 s4 = sl[q];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * q + jj;
            const int e = ((j >> I) << (I + 1)) | (j & ((1 << I) - 1));
            const int e1 = e | (1 << I);
            const uint32_t sj = jj == 0 ? s4.x : jj == 1 ? s4.y : jj == 2 ? s4.z : s4.w;
            const double2 cs = *reinterpret_cast<const double2*>(tb + (sj ^ ck));
            const double c = cs.x, g = cs.y;
            const double2 u0 = a[e], u1 = a[e1];
            if (!ODD) {
                // a' = c a + g b, b' = c b - g a   (g carries w, the sign parity and the orientation)
                a[e] = make_double2(fma(g, u1.x, c * u0.x), fma(g, u1.y, c * u0.y));
                a[e1] = make_double2(fma(-g, u0.x, c * u1.x), fma(-g, u0.y, c * u1.y));
            } else {
                // a' = c a + i g b, b' = c b + i g a   (the same for both orientations)
                a[e] = make_double2(fma(-g, u1.y, c * u0.x), fma(g, u1.x, c * u0.y));
                a[e1] = make_double2(fma(-g, u0.y, c * u1.x), fma(g, u0.x, c * u1.y));
            }
        }
    }
}

__global__ void __launch_bounds__(kOrbitThreads, 2)
    rot_orbit_kernel(double2* __restrict__ psi, uint64_t n_threads_total, const PhaseDesc ph, const double2* __restrict__ table) {
    __shared__ __align__(16) double2 stab[kMaxPhaseRuns * kKeyTable];
    __shared__ __align__(16) uint32_t ssel[kMaxPhaseRuns * kOrbitSize];
    for (int e = threadIdx.x; e < ph.n_runs * kKeyTable; e += kOrbitThreads) {
        const PhaseRun& d = ph.run[e / kKeyTable];
        const uint32_t key = (uint32_t)e % kKeyTable, idx = key & 15u, par = (key >> 4) & 1u, sw = key >> 5;
        double2 out = make_double2(0.0, 0.0);
        if (idx < (1u << __popcll(d.sup_mask))) {
            const double2 cs = __ldg(&table[d.table_off + idx]);
            double g;
            if (d.basis < 0) {
                g = par ? cs.y : -cs.y;  // multiply by (c - i s0 s): a' = (c v.x - g v.y, c v.y + g v.x)
            } else {
                // w = (-i)^ma: +1, -i, -1, +i
                const double wsign = (d.ma == 0 || d.ma == 3) ? 1.0 : -1.0;
                g = (par ? -cs.y : cs.y) * wsign;
                if (!(d.ma & 1) && sw) g = -g;  // roles of a and b swap with the orientation
            }
            out = make_double2(cs.x, g);
        }
        stab[e] = out;
    }
    for (int e = threadIdx.x; e < ph.n_runs * kOrbitSize; e += kOrbitThreads)
        ssel[e] = (uint32_t)ph.run[e / kOrbitSize].sel[e % kOrbitSize] << 4;
    __syncthreads();
    for (uint64_t tg = (uint64_t)blockIdx.x * kOrbitThreads + threadIdx.x; tg < n_threads_total;
         tg += (uint64_t)gridDim.x * kOrbitThreads) {
        uint64_t c0 = tg;  // first representative: the thread number deposited into the non-pivot index bits
        for (int j = 0; j < ph.n_pivots; ++j) c0 = insert_zero_bit(c0, ph.pivot[j]);
        double2 a[kOrbitSize];
#pragma unroll
        for (int e = 0; e < kOrbitSize; ++e) a[e] = psi[c0 ^ ph.elem[e]];
        for (int r = 0; r < ph.n_runs; ++r) {
            const PhaseRun& d = ph.run[r];
            uint32_t ck = (d.sup_mask ? gather_bits(c0, d.sup_mask) : 0u) | (((uint32_t)__popcll(c0 & d.z0) & 1u) << 4);
            if (d.hb >= 0 && ((c0 >> d.hb) & 1ull)) ck ^= d.kx;
            ck <<= 4;
            const char* tb = reinterpret_cast<const char*>(stab + r * kKeyTable);
            const uint4* sl = reinterpret_cast<const uint4*>(ssel + r * kOrbitSize);
            if (d.basis < 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 s4 = sl[q];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int e = 4 * q + jj;
                        const uint32_t sj = jj == 0 ? s4.x : jj == 1 ? s4.y : jj == 2 ? s4.z : s4.w;
                        const double2 cs = *reinterpret_cast<const double2*>(tb + (sj ^ ck));
                        const double2 v = a[e];
                        a[e] = make_double2(fma(-cs.y, v.y, cs.x * v.x), fma(cs.y, v.x, cs.x * v.y));
                    }
                }
                continue;
            }
#define QC_BFLY(I)                                      \
    case I:                                             \
        if (d.ma & 1)                                   \
            orbit_butterfly<I, true>(a, tb, sl, ck);    \
        else                                            \
            orbit_butterfly<I, false>(a, tb, sl, ck);   \
        break;
            switch (d.basis) {
                QC_BFLY(0) QC_BFLY(1) QC_BFLY(2) QC_BFLY(3)
                default:
                    if (d.ma & 1)
                        orbit_butterfly<4, true>(a, tb, sl, ck);
                    else
                        orbit_butterfly<4, false>(a, tb, sl, ck);
                    break;
            }
#undef QC_BFLY
        }
#pragma unroll
        for (int e = 0; e < kOrbitSize; ++e) psi[c0 ^ ph.elem[e]] = a[e];
    }
}

static inline void pow_minus_i(int m, double& re, double& im) {
    switch (((m % 4) + 4) % 4) {
        case 0: re = 1.0; im = 0.0; break;
        case 1: re = 0.0; im = -1.0; break;
        case 2: re = -1.0; im = 0.0; break;
        default: re = 0.0; im = 1.0; break;
    }
}

struct Run {
    size_t first, count;
    uint64_t x, z0, sup;
    int hb;
};

// Host-side planner: split the ordered string list into maximal fusable runs.
static void plan_runs(size_t R, const uint64_t* xs, const uint64_t* zs, std::vector<Run>& runs) {
    size_t r = 0;
    while (r < R) {
        Run g;
        g.first = r;
        g.count = 1;
        g.x = xs[r];
        g.z0 = zs[r];
        g.sup = 0;
        g.hb = g.x ? 63 - __builtin_clzll(g.x) : -1;
        const uint64_t hbbit = g.x ? (1ull << g.hb) : 0ull;
        size_t k = r + 1;
        while (k < R && g.count < (size_t)kMaxRun && xs[k] == g.x) {
            const uint64_t dz = zs[k] ^ g.z0;
            if (__builtin_popcountll(g.x & dz) & 1) break;  // anticommutes with the run
            const uint64_t nsup = g.sup | (dz & ~hbbit);
            if (__builtin_popcountll(nsup) > kMaxSupport) break;
            g.sup = nsup;
            ++g.count;
            ++k;
        }
        runs.push_back(g);
        r = k;
    }
}



// Fused runs of a string list (first string, count) for callers that walk a program run by run (gradient.cu).
void plan_rotation_runs(size_t R, const uint64_t* xs, const uint64_t* zs, std::vector<std::pair<size_t, size_t>>& out) {
    std::vector<Run> runs;
    plan_runs(R, xs, zs, runs);
    out.clear();
    for (const Run& g : runs) out.push_back({g.first, g.count});
}

// Applies the rotations to `target` (a shard-sized array of this handle: psi itself or the scratch array).
int apply_rotations_to(qc_state* h, double2* target, size_t R, const uint64_t* xmask, const uint64_t* zmask,
                       const double* angle, int k1_mode, int* n_passes_out, long long basis_index) {
    std::vector<Run> runs;
    plan_runs(R, xmask, zmask, runs);

    if (h->n <= kSmallMaxQubits && k1_mode == 0 && (runs.size() >= 2 || basis_index != -1)) {
        // ---- small register: every run of the call in ONE launch, state in shared memory
        size_t total_entries = 0;
        for (const Run& g : runs) total_entries += (size_t)1 << __builtin_popcountll(g.sup);
        const size_t bytes_desc = (runs.size() * sizeof(RotDesc) + 255) & ~(size_t)255;
        QC_CHECK(cudaStreamSynchronize(h->stream));
        if (h->arena.ensure(bytes_desc + total_entries * sizeof(double2))) return 1;
        RotDesc* hdesc = (RotDesc*)h->arena.host;
        double2* htab = (double2*)((char*)h->arena.host + bytes_desc);
        size_t off = 0;
        for (size_t gi = 0; gi < runs.size(); ++gi) {
            const Run& g = runs[gi];
            int sup_bits[kMaxSupport], nsup = 0;
            for (uint64_t m = g.sup; m; m &= m - 1) sup_bits[nsup++] = __builtin_ctzll(m);
            const int ny0 = __builtin_popcountll(g.x & g.z0);
            for (uint32_t pat = 0; pat < (1u << nsup); ++pat) {
                uint64_t bits = 0;
                for (int k = 0; k < nsup; ++k)
                    if ((pat >> k) & 1u) bits |= 1ull << sup_bits[k];
                double phi = 0.0;
                for (size_t k = 0; k < g.count; ++k) {
                    const size_t r = g.first + k;
                    const uint64_t dz = zmask[r] ^ g.z0;
                    const int nyk = __builtin_popcountll(g.x & zmask[r]);
                    double sg = ((((nyk - ny0) % 4) + 4) % 4 == 0) ? 1.0 : -1.0;
                    if (__builtin_popcountll(bits & dz) & 1) sg = -sg;
                    phi += sg * angle[r];
                }
                htab[off + pat] = make_double2(cos(0.5 * phi), sin(0.5 * phi));
            }
            RotDesc& d = hdesc[gi];
            memset(&d, 0, sizeof(d));
            d.x = g.x;
            d.z0 = g.z0;
            d.sup_mask = g.sup;
            d.hb = g.hb;
            pow_minus_i(ny0 + 1, d.war, d.wai);
            pow_minus_i(1 - ny0, d.wbr, d.wbi);
            d.table_off = (uint32_t)off;
            off += (size_t)1 << nsup;
        }
        QC_CHECK(cudaMemcpyAsync(h->arena.dev, h->arena.host, bytes_desc + total_entries * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
        h->h2d_bytes += bytes_desc + total_entries * sizeof(double2);
        static_assert(sizeof(RotDesc) % 8 == 0, "RotDesc is copied in 8-byte units");
        const size_t smem_state = (size_t)h->dim * sizeof(double2);
        const size_t smem_staged = smem_state + total_entries * sizeof(double2) + runs.size() * sizeof(RotDesc);
        const bool staged = smem_staged <= 220 * 1024;
        const size_t smem = staged ? smem_staged : smem_state;
        auto kernel = staged ? rot_small_kernel<true> : rot_small_kernel<false>;
        QC_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        {
            LaunchScope ls(h, kK1Rotation, 32.0 * (double)h->dim * (double)runs.size());
            kernel<<<1, kSmallThreads, smem, h->stream>>>(target, h->n, (int)runs.size(), (const RotDesc*)h->arena.dev,
                                                          (const double2*)((const char*)h->arena.dev + bytes_desc), (uint32_t)total_entries,
                                                          basis_index);
        }
        QC_CHECK(cudaGetLastError());
        if (n_passes_out) *n_passes_out = 1;
        return 0;
    }

    // ---- batches.  k1_mode 0 (auto) / 3: register-blocked PHASES (K1O): consecutive runs whose distinct x-masks are
    // linearly independent (<= 5 of them); 2: shared-memory TILES (K1T): consecutive runs whose x-masks fit, with index
    // bits 0..2, into one 12-bit tile; 1: one sweep per fused run.
    struct Batch {
        size_t first, count;
        uint64_t S;  // tile bits of a K1T batch (0 otherwise)
        int phase;   // index into `phases` for a K1O batch (-1 otherwise)
    };
    struct Phase {
        std::vector<uint64_t> basis;  // distinct x-masks in order of first use
        std::vector<uint64_t> red;    // the same span in reduced row echelon form (pivot = highest set bit)
    };
    std::vector<Batch> batches;
    std::vector<Phase> phases;
    const bool orbit_ok = h->n >= 10 && (k1_mode == 0 || k1_mode == 3);
    if (orbit_ok) {
        auto small = [&](const Run& g) { return (1 << __builtin_popcountll(g.sup)) <= kPhaseTable; };
        auto reduce = [](const std::vector<uint64_t>& red, uint64_t v) {
            for (uint64_t r : red)
                if ((v >> (63 - __builtin_clzll(r))) & 1ull) v ^= r;
            return v;
        };
        size_t i = 0;
        while (i < runs.size()) {
            if (!small(runs[i])) {
                batches.push_back({i, 1, 0ull, -1});
                ++i;
                continue;
            }
            Phase ph;
            size_t j = i;
            while (j < runs.size() && j - i < (size_t)kMaxPhaseRuns && small(runs[j])) {
                const uint64_t x = runs[j].x;
                if (x != 0 && std::find(ph.basis.begin(), ph.basis.end(), x) == ph.basis.end()) {
                    const uint64_t v = reduce(ph.red, x);
                    if (v == 0 || ph.basis.size() == (size_t)kOrbitBits) break;  // dependent on the span / orbit is full
                    const int pv = 63 - __builtin_clzll(v);
                    for (uint64_t& r : ph.red)
                        if ((r >> pv) & 1ull) r ^= v;  // keep the echelon form fully reduced: pivot columns are unit columns
                    ph.red.push_back(v);
                    std::sort(ph.red.begin(), ph.red.end(), std::greater<uint64_t>());
                    ph.basis.push_back(x);
                }
                ++j;
            }
            if (j - i >= 2 || k1_mode == 3) {
                batches.push_back({i, j - i, 0ull, (int)phases.size()});
                phases.push_back(ph);
            } else {
                batches.push_back({i, 1, 0ull, -1});
                j = i + 1;
            }
            i = j;
        }
    } else {
        const bool tiles_ok = h->n >= kRotTileBits && k1_mode == 2;
        const uint64_t low = 7ull;
        size_t i = 0;
        while (i < runs.size()) {
            uint64_t U = runs[i].x;
            size_t j = i + 1;
            auto small = [&](const Run& g) { return (1 << __builtin_popcountll(g.sup)) <= kTileRunTable; };
            if (tiles_ok && small(runs[i]) && __builtin_popcountll(U | low) <= kRotTileBits) {
                while (j < runs.size() && j - i < (size_t)kMaxBatchRuns && small(runs[j]) &&
                       __builtin_popcountll(U | runs[j].x | low) <= kRotTileBits) {
                    U |= runs[j].x;
                    ++j;
                }
                uint64_t S = U | low;
                for (int b = 0; __builtin_popcountll(S) < kRotTileBits; ++b) S |= 1ull << b;
                batches.push_back({i, j - i, S, -1});
                i = j;
                continue;
            }
            batches.push_back({i, 1, 0ull, -1});
            i += 1;
        }
    }

    // ---- tables: (cos, sin)(Phi/2) over the support patterns of every run.  Index bit k of a table <-> the k-th
    // entry of the run's ordered support-bit list: ascending index bits for the sweep kernels; tile bits first, then
    // the bits outside the tile, for a run inside a tiled batch.
    size_t total_entries = 0;
    for (const Run& g : runs) total_entries += (size_t)1 << __builtin_popcountll(g.sup);
    // The arena may still be in use by a previous call's in-flight kernels only if the stream has not drained;
    // the copy below is stream-ordered, and the host buffer is rewritten only after the previous copy finished.
    QC_CHECK(cudaStreamSynchronize(h->stream));
    if (h->arena.ensure(total_entries * sizeof(double2))) return 1;
    double2* htab = (double2*)h->arena.host;
    std::vector<uint32_t> table_off(runs.size());
    {
        size_t off = 0;
        for (const Batch& bt : batches)
            for (size_t gi = bt.first; gi < bt.first + bt.count; ++gi) {
                const Run& g = runs[gi];
                int sup_bits[kMaxSupport], nsup = 0;
                for (int pass = 0; pass < 2; ++pass) {  // tile bits first (every bit when S == 0: ascending order)
                    uint64_t m = pass == 0 ? (bt.S ? g.sup & bt.S : g.sup) : (bt.S ? g.sup & ~bt.S : 0ull);
                    while (m) {
                        sup_bits[nsup++] = __builtin_ctzll(m);
                        m &= m - 1;
                    }
                }
                const int ny0 = __builtin_popcountll(g.x & g.z0);
                for (uint32_t pat = 0; pat < (1u << nsup); ++pat) {
                    uint64_t bits = 0;
                    for (int k = 0; k < nsup; ++k)
                        if ((pat >> k) & 1u) bits |= 1ull << sup_bits[k];
                    double phi = 0.0;
                    for (size_t k = 0; k < g.count; ++k) {
                        const size_t r = g.first + k;
                        const uint64_t dz = zmask[r] ^ g.z0;
                        const int nyk = __builtin_popcountll(g.x & zmask[r]);
                        double sg = ((((nyk - ny0) % 4) + 4) % 4 == 0) ? 1.0 : -1.0;  // (-i)^{nYk-nY0}, an even power
                        if (__builtin_popcountll(bits & dz) & 1) sg = -sg;
                        phi += sg * angle[r];
                    }
                    htab[off + pat] = make_double2(cos(0.5 * phi), sin(0.5 * phi));
                }
                table_off[gi] = (uint32_t)off;
                off += (size_t)1 << nsup;
            }
    }
    QC_CHECK(cudaMemcpyAsync(h->arena.dev, h->arena.host, total_entries * sizeof(double2), cudaMemcpyHostToDevice,
                             h->stream));
    h->h2d_bytes += total_entries * sizeof(double2);
    const double2* dtab = (const double2*)h->arena.dev;

    const size_t tile_smem = (size_t)(kRotTileSize + kMaxBatchRuns * kTileRunTable) * sizeof(double2) + (size_t)kMaxBatchRuns * kRotTileThreads;
    for (const Batch& bt : batches) {
        LaunchScope ls(h, kK1Rotation, 32.0 * (double)h->dim);  // every amplitude read + written once per pass
        if (bt.phase >= 0) {
            const Phase& phs = phases[bt.phase];
            PhaseDesc pd;
            memset(&pd, 0, sizeof(pd));
            const int d = (int)phs.basis.size();
            std::vector<int> piv;
            for (uint64_t r : phs.red) piv.push_back(63 - __builtin_clzll(r));
            std::sort(piv.begin(), piv.end());
            pd.n_pivots = d;
            for (int j = 0; j < d; ++j) pd.pivot[j] = (uint8_t)piv[j];
            auto deposit = [&](uint64_t v) {  // insert zeros at the pivot positions (ascending)
                for (int pv : piv) v = ((v >> pv) << (pv + 1)) | (v & ((1ull << pv) - 1ull));
                return v;
            };
            for (int e = 0; e < kOrbitSize; ++e) {
                const uint64_t o = (uint64_t)e >> d;
                uint64_t off = deposit(o << (h->n - kOrbitBits));
                for (int b = 0; b < d; ++b)
                    if ((e >> b) & 1) off ^= phs.basis[b];
                pd.elem[e] = off;
            }
            pd.n_runs = (int)bt.count;
            for (size_t k = 0; k < bt.count; ++k) {
                const Run& g = runs[bt.first + k];
                PhaseRun& pr = pd.run[k];
                const int ny0 = __builtin_popcountll(g.x & g.z0);
                pr.z0 = g.z0;
                pr.sup_mask = g.sup;
                pr.table_off = table_off[bt.first + k];
                pr.hb = (int8_t)g.hb;
                pr.basis = -1;
                for (int b = 0; b < d; ++b)
                    if (phs.basis[b] == g.x) pr.basis = (int8_t)b;
                pr.ma = (int8_t)((((ny0 + 1) % 4) + 4) % 4);
                auto pext = [&](uint64_t v) {
                    uint32_t out = 0;
                    int pos = 0;
                    for (uint64_t m = g.sup; m; m &= m - 1, ++pos)
                        if ((v >> __builtin_ctzll(m)) & 1ull) out |= 1u << pos;
                    return out;
                };
                // key of an index v: table-index bits | sign parity << 4, XOR kx when bit hb of v is set (the
                // canonical element is then v ^ x; bit 5 of kx records the swap)
                pr.kx = g.hb >= 0 ? (pext(g.x) | ((uint32_t)(__builtin_popcountll(g.x & g.z0) & 1) << 4) | 32u) : 0u;
                auto key_of = [&](uint64_t v) {
                    uint32_t key = pext(v) | ((uint32_t)(__builtin_popcountll(v & g.z0) & 1) << 4);
                    if (g.hb >= 0 && ((v >> g.hb) & 1ull)) key ^= pr.kx;
                    return key;
                };
                if (pr.basis < 0) {
                    for (int e = 0; e < kOrbitSize; ++e) pr.sel[e] = (uint8_t)key_of(pd.elem[e]);
                } else {
                    const int I = pr.basis;
                    for (int j = 0; j < kOrbitSize / 2; ++j) {
                        const int e = ((j >> I) << (I + 1)) | (j & ((1 << I) - 1));
                        pr.sel[j] = (uint8_t)key_of(pd.elem[e]);
                    }
                }
            }
            const uint64_t n_threads = h->dim >> kOrbitBits;
            const int nb = (int)std::min<uint64_t>((n_threads + kOrbitThreads - 1) / kOrbitThreads, (uint64_t)h->sm_count * 16);
            h->h2d_bytes += sizeof(PhaseDesc);  // kernel parameters
            rot_orbit_kernel<<<nb, kOrbitThreads, 0, h->stream>>>(target, n_threads, pd, dtab);
            continue;
        }
        if (bt.S) {
            TileBatch tb;
            memset(&tb, 0, sizeof(tb));
            tb.s_mask = bt.S;
            int spos[kRotTileBits];
            {
                uint64_t m = bt.S;
                for (int j = 0; j < kRotTileBits; ++j) {
                    spos[j] = __builtin_ctzll(m);
                    tb.spos[j] = (uint8_t)spos[j];
                    m &= m - 1;
                }
            }
            auto local = [&](uint64_t mask) {
                uint32_t out = 0;
                for (int j = 0; j < kRotTileBits; ++j)
                    if ((mask >> spos[j]) & 1ull) out |= 1u << j;
                return out;
            };
            tb.n_runs = (int)bt.count;
            for (size_t k = 0; k < bt.count; ++k) {
                const Run& g = runs[bt.first + k];
                TileRun& tr = tb.run[k];
                const int ny0 = __builtin_popcountll(g.x & g.z0);
                tr.z0_outer = g.z0 & ~bt.S;
                tr.sup_outer = g.sup & ~bt.S;
                tr.xl = local(g.x);
                tr.z0_l = local(g.z0 & bt.S);
                tr.sup_l = local(g.sup & bt.S);
                tr.table_off = table_off[bt.first + k];
                tr.hb_l = tr.xl ? (int8_t)(31 - __builtin_clz(tr.xl)) : (int8_t)-1;
                tr.n_sup_l = (int8_t)__builtin_popcount(tr.sup_l);
                tr.ma = (int8_t)((((ny0 + 1) % 4) + 4) % 4);
                tr.mb = (int8_t)((((1 - ny0) % 4) + 4) % 4);
                const int n_k = tr.xl ? kRotTileSize / 2 / kRotTileThreads : kRotTileSize / kRotTileThreads;
                for (int kk = 0; kk < n_k; ++kk) {
                    uint32_t tk = (uint32_t)kk * kRotTileThreads;
                    if (tr.xl) tk = ((tk >> tr.hb_l) << (tr.hb_l + 1)) | (tk & ((1u << tr.hb_l) - 1u));  // insert a 0 at hb_l
                    uint32_t ext = 0;
                    int pos = 0;
                    for (uint32_t m = tr.sup_l; m; m &= m - 1, ++pos)
                        if ((tk >> __builtin_ctz(m)) & 1u) ext |= 1u << pos;
                    tr.kpack[kk] = tk | (ext << 12) | ((uint32_t)(__builtin_popcount(tk & tr.z0_l) & 1) << 31);
                }
            }
            QC_CHECK(cudaFuncSetAttribute(rot_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
            const uint64_t n_tiles = 1ull << (h->n - kRotTileBits);
            const int nb = (int)std::min<uint64_t>(n_tiles, (uint64_t)h->sm_count * 3);
            h->h2d_bytes += sizeof(TileBatch);  // kernel parameters
            rot_tiled_kernel<<<nb, kRotTileThreads, tile_smem, h->stream>>>(target, h->n, tb, dtab);
            continue;
        }
        const Run& g = runs[bt.first];
        RotDesc d;
        const int ny0 = __builtin_popcountll(g.x & g.z0);
        d.x = g.x;
        d.z0 = g.z0;
        d.sup_mask = g.sup;
        d.hb = g.hb;
        pow_minus_i(ny0 + 1, d.war, d.wai);
        pow_minus_i(1 - ny0, d.wbr, d.wbi);
        d.table_off = table_off[bt.first];
        if (d.x == 0) {
            const int nb = grid_for(h->dim, 4, h->sm_count);
            rot_diag_kernel<4><<<nb, kThreads, 0, h->stream>>>(target, h->dim, d, dtab);
        } else if (d.hb < 5) {
            const int nb = grid_for(h->dim, 4, h->sm_count);
            rot_lowx_kernel<<<nb, kThreads, 0, h->stream>>>(target, h->dim, d, dtab);
        } else {
            const uint64_t npairs = h->dim >> 1;
            const int nb = grid_for(npairs, 4, h->sm_count);
            rot_pairs_kernel<4><<<nb, kThreads, 0, h->stream>>>(target, npairs, d, dtab);
        }
    }
    QC_CHECK(cudaGetLastError());
    if (n_passes_out) *n_passes_out = (int)batches.size();
    return 0;
}

}  // namespace qc

extern "C" int qc_apply_pauli_rotations(qc_state* h, size_t R, const uint64_t* xmask, const uint64_t* zmask,
                                        const double* angle, int* n_passes_out) {
    using namespace qc;
    QC_REQUIRE(h, "qc_apply_pauli_rotations: null handle");
    if (n_passes_out) *n_passes_out = 0;
    if (R == 0) return 0;
    QC_REQUIRE(xmask && zmask && angle, "qc_apply_pauli_rotations: null array");
    for (size_t r = 0; r < R; ++r)
        QC_REQUIRE(xmask[r] < h->dim && zmask[r] < h->dim, "qc_apply_pauli_rotations: mask has bits outside the shard");
    QC_CHECK(cudaSetDevice(h->device));
    return apply_rotations_to(h, h->psi, R, xmask, zmask, angle, h->k1_mode, n_passes_out, -1);
}

extern "C" int qc_prepare_state(qc_state* h, uint64_t basis_index, int has_one, size_t R, const uint64_t* xmask,
                                const uint64_t* zmask, const double* angle, int* n_passes_out) {
    using namespace qc;
    QC_REQUIRE(h, "qc_prepare_state: null handle");
    if (n_passes_out) *n_passes_out = 0;
    QC_REQUIRE(!has_one || basis_index < h->dim, "qc_prepare_state: index out of range");
    QC_REQUIRE(R == 0 || (xmask && zmask && angle), "qc_prepare_state: null array");
    for (size_t r = 0; r < R; ++r)
        QC_REQUIRE(xmask[r] < h->dim && zmask[r] < h->dim, "qc_prepare_state: mask has bits outside the shard");
    QC_CHECK(cudaSetDevice(h->device));
    if (h->n <= kSmallMaxQubits && h->k1_mode == 0 && R > 0) {
        // small register: basis state + the whole program in ONE launch
        h->prof.bytes[kK0Init] += 16.0 * (double)h->dim;
        return apply_rotations_to(h, h->psi, R, xmask, zmask, angle, 0, n_passes_out, has_one ? (long long)basis_index : -2);
    }
    if (qc_set_basis_state(h, basis_index, has_one)) return 1;
    if (R == 0) return 0;
    return apply_rotations_to(h, h->psi, R, xmask, zmask, angle, h->k1_mode, n_passes_out, -1);
}
