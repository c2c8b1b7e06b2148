// K2T: tiled expectation.  One HBM read of the shard serves EVERY x-mask whose support lies inside the tile's bit set.
//
// The sweep kernel (expectation.cu) reads the whole shard once per distinct x-mask; a molecular Hamiltonian of 10^5
// terms has ~2*10^4 x-masks, so the sweeps are 98 % of an energy evaluation.  Here a PASS fixes a set S of 12 index
// bits (the 3 lowest + 9 chosen by a greedy cover of the x-masks); a CTA stages one tile (the 4096 amplitudes that
// differ only in the S bits, 64 KB) in shared memory and evaluates every x-group with x inside S from it.  Inside a
// pass the groups are bucketed into R-BLOCKS: 5 tile bits R with x inside R.  Each of the 128 threads gathers the
// 32 amplitudes of its R-sub-cube into registers ONCE per R-block and evaluates all groups of the block from
// registers (the pair products q = conj(a_p) a_{p^x} need no further shared-memory traffic), which cuts the
// shared-memory reads per x-mask by the number of groups per block.
//
// Weights.  For a group (x; terms z_k, d_k) and the canonical element p of a pair inside the sub-cube,
// (-1)^{popc(i & z_k)} factors into  sign(tile origin & z_k outside S) * sign(thread's rest bits & z_k) *
// sign(p & z_k inside R).  Terms that agree outside R form a SUBGROUP whose last factor is tabulated on the host:
// tab[pair] = 2 sum_k d_k (-1)^{popc(p & zR_k)} (no factor 2 on the diagonal).  A quad group of a JW Hamiltonian
// (XXYY, YYXX, XYYX, YXXY) is one subgroup; a hopping group (X..X / Y..Y decorated with one Z_r) is ~29.
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>
#include <unordered_map>

#include "common.cuh"

namespace qc {

constexpr int kTileBits = 12;
constexpr int kTileSize = 1 << kTileBits;
constexpr int kLowBits = 3;   // the tile always contains index bits 0..2: 128-byte runs in HBM
constexpr int kRBits = 5;     // register sub-cube: 32 amplitudes per thread
constexpr int kTileThreads = kTileSize >> kRBits;  // 128
constexpr int kMaxPairs = 16;

struct TPass {
    uint64_t s_mask;      // the 12 index bits of the tile
    uint32_t first_rblock, n_rblocks;
    uint32_t run_table;   // offset into run_off[]: 512 amplitude offsets of the 8-amplitude runs of a tile
    uint32_t pad;
};
struct TRBlock {
    uint16_t sbit[kRBits];  // swizzled tile offset of each R bit
    uint16_t pad[3];
    uint32_t rest_table;    // offset into rest_tab[]: 128 entries (unswizzled << 16 | swizzled) of the thread's origin
    uint32_t first_group, n_groups;
    uint32_t pad2;
};
struct TGroup {
    uint32_t xr;        // x inside R (5 bits); 0 => diagonal group
    uint32_t has_im;    // some subgroup uses Im q (terms with odd nY)
    uint32_t first_sub, n_sub;
};
struct TSub {
    uint64_t z_outer;   // z outside S (index-bit positions)
    uint32_t z_rest;    // z on the tile's rest bits (tile-local positions, R bits cleared)
    uint32_t use_im;    // 0: weights multiply Re q (or |a|^2), 1: Im q
    uint32_t table;     // offset into tables[] (16 doubles; diagonal: 32)
    uint32_t pad;
};

struct TiledPlan {
    uint64_t hash = 0;
    int n = 0;
    std::vector<TPass> passes;
    std::vector<uint32_t> pass_groups;  // x-groups served by each pass
    size_t n_groups = 0, n_rblocks = 0, n_subs = 0;
    // leftover terms (x does not fit a tile / a sub-cube): evaluated by the sweep kernel
    std::vector<uint64_t> left_x, left_z;
    std::vector<double> left_c;
    // device copies
    char* dev = nullptr;
    TPass* d_pass = nullptr;
    TRBlock* d_rblock = nullptr;
    TGroup* d_group = nullptr;
    TSub* d_sub = nullptr;
    double* d_tables = nullptr;
    uint32_t* d_rest = nullptr;
    uint64_t* d_run = nullptr;
};

void free_tiled_plan(TiledPlan* p) {
    if (!p) return;
    if (p->dev) cudaFree(p->dev);
    delete p;
}

__host__ __device__ inline uint32_t swizzle(uint32_t t) { return t ^ ((t >> 3) & 7u) ^ ((t >> 6) & 7u) ^ ((t >> 9) & 7u); }

// ---------------------------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int XR>
struct PairEval {
    // q[k] for the k-th canonical pair (p ascending over the patterns with the highest bit of XR clear)
    static __device__ __forceinline__ void re(const double2 (&a)[32], double (&q)[kMaxPairs]) {
        constexpr int HB = XR >= 16 ? 16 : XR >= 8 ? 8 : XR >= 4 ? 4 : XR >= 2 ? 2 : 1;
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            if (!(p & HB)) {
                const int k = ((p / (2 * HB)) * HB) | (p & (HB - 1));  // rank of p among the patterns with bit HB clear
                q[k] = a[p].x * a[p ^ XR].x + a[p].y * a[p ^ XR].y;
            }
        }
    }
    static __device__ __forceinline__ void im(const double2 (&a)[32], double (&q)[kMaxPairs]) {
        constexpr int HB = XR >= 16 ? 16 : XR >= 8 ? 8 : XR >= 4 ? 4 : XR >= 2 ? 2 : 1;
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            if (!(p & HB)) {
                const int k = ((p / (2 * HB)) * HB) | (p & (HB - 1));
                q[k] = a[p].x * a[p ^ XR].y - a[p].y * a[p ^ XR].x;
            }
        }
    }
};

template <bool IM>
__device__ __forceinline__ void pair_products(uint32_t xr, const double2 (&a)[32], double (&q)[kMaxPairs]) {
#define QC_CASE(X)                                                       \
    case X:                                                              \
        if (IM) PairEval<X>::im(a, q); else PairEval<X>::re(a, q);       \
        break;
    switch (xr) {
        QC_CASE(1) QC_CASE(2) QC_CASE(3) QC_CASE(4) QC_CASE(5) QC_CASE(6) QC_CASE(7) QC_CASE(8) QC_CASE(9) QC_CASE(10)
        QC_CASE(11) QC_CASE(12) QC_CASE(13) QC_CASE(14) QC_CASE(15) QC_CASE(16) QC_CASE(17) QC_CASE(18) QC_CASE(19)
        QC_CASE(20) QC_CASE(21) QC_CASE(22) QC_CASE(23) QC_CASE(24) QC_CASE(25) QC_CASE(26) QC_CASE(27) QC_CASE(28)
        QC_CASE(29) QC_CASE(30) QC_CASE(31)
        default: break;
    }
#undef QC_CASE
}

__device__ __forceinline__ double apply_sign(double v, unsigned par) {
    return __longlong_as_double(__double_as_longlong(v) ^ ((long long)(par & 1u) << 63));
}

__global__ void __launch_bounds__(kTileThreads, 2)
    expectation_tiled_kernel(const double2* __restrict__ psi, int n, const TPass pass, const TRBlock* __restrict__ rblocks,
                             const TGroup* __restrict__ groups, const TSub* __restrict__ subs,
                             const double* __restrict__ tables, const uint32_t* __restrict__ rest_tab,
                             const uint64_t* __restrict__ run_off, double* __restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    const uint64_t n_tiles = 1ull << (n - kTileBits);
    const uint64_t* runs = run_off + pass.run_table;
    double acc = 0.0;

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
        __syncthreads();  // previous tile fully consumed
#pragma unroll 4
        for (int k = 0; k < kTileSize / kTileThreads; ++k) {
            const uint32_t t = (uint32_t)k * kTileThreads + threadIdx.x;
            const uint64_t g = base + runs[t >> kLowBits] + (t & ((1u << kLowBits) - 1u));
            cp_async16(&tile[swizzle(t)], &psi[g]);
        }
        cp_async_wait_all();
        __syncthreads();

        for (uint32_t rb = 0; rb < pass.n_rblocks; ++rb) {
            const TRBlock blk = rblocks[pass.first_rblock + rb];
            const uint32_t packed = __ldg(&rest_tab[blk.rest_table + threadIdx.x]);
            const uint32_t org_sw = packed & 0xffffu, org = packed >> 16;
            double2 a[32];
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                uint32_t off = org_sw;
#pragma unroll
                for (int j = 0; j < kRBits; ++j)
                    if (p & (1 << j)) off ^= blk.sbit[j];
                a[p] = tile[off];
            }
            for (uint32_t gi = 0; gi < blk.n_groups; ++gi) {
                const TGroup grp = groups[blk.first_group + gi];
                // The pair products below depend only on a[], which is invariant in this loop: without this barrier the
                // compiler hoists the products of EVERY switch case out of the loop (LICM) and spills ~7 KB per thread.
#pragma unroll
                for (int p = 0; p < 32; ++p) asm volatile("" : "+d"(a[p].x), "+d"(a[p].y));
                double q[kMaxPairs];
                if (grp.xr == 0) {
                    // diagonal: 32 probabilities in two halves, subgroup tables hold 32 weights
                    for (int half = 0; half < 2; ++half) {
#pragma unroll
                        for (int k = 0; k < kMaxPairs; ++k) {
                            // static register indexing: select by the (uniform) half flag
                            const double2 lo = a[k], hi = a[k + 16];
                            const double2 v = half ? hi : lo;
                            q[k] = v.x * v.x + v.y * v.y;
                        }
                        for (uint32_t s = 0; s < grp.n_sub; ++s) {
                            const TSub sub = subs[grp.first_sub + s];
                            const unsigned par = (unsigned)__popcll(base & sub.z_outer) + (unsigned)__popc(org & sub.z_rest);
                            const double* tab = tables + sub.table + half * kMaxPairs;
                            double w = 0.0;
#pragma unroll
                            for (int k = 0; k < kMaxPairs; ++k) w += q[k] * __ldg(&tab[k]);
                            acc += apply_sign(w, par);
                        }
                    }
                    continue;
                }
                pair_products<false>(grp.xr, a, q);
                for (uint32_t s = 0; s < grp.n_sub; ++s) {
                    const TSub sub = subs[grp.first_sub + s];
                    if (sub.use_im) continue;
                    const unsigned par = (unsigned)__popcll(base & sub.z_outer) + (unsigned)__popc(org & sub.z_rest);
                    const double* tab = tables + sub.table;
                    double w = 0.0;
#pragma unroll
                    for (int k = 0; k < kMaxPairs; ++k) w += q[k] * __ldg(&tab[k]);
                    acc += apply_sign(w, par);
                }
                if (grp.has_im) {
                    pair_products<true>(grp.xr, a, q);
                    for (uint32_t s = 0; s < grp.n_sub; ++s) {
                        const TSub sub = subs[grp.first_sub + s];
                        if (!sub.use_im) continue;
                        const unsigned par = (unsigned)__popcll(base & sub.z_outer) + (unsigned)__popc(org & sub.z_rest);
                        const double* tab = tables + sub.table;
                        double w = 0.0;
#pragma unroll
                        for (int k = 0; k < kMaxPairs; ++k) w += q[k] * __ldg(&tab[k]);
                        acc += apply_sign(w, par);
                    }
                }
            }
        }
    }
    // CTA reduction (kTileThreads = 128 threads)
    __shared__ double wpart[kTileThreads / 32];
    acc = warp_sum(acc);
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

// ---------------------------------------------------------------------------------------------------------------
// host planner
// ---------------------------------------------------------------------------------------------------------------
static uint64_t fnv1a(const void* data, size_t bytes, uint64_t h) {
    const uint64_t* w = (const uint64_t*)data;
    const size_t nw = bytes / 8;
    for (size_t i = 0; i < nw; ++i) {
        h ^= w[i];
        h *= 0x100000001b3ull;
    }
    return h;
}

static inline uint32_t to_tile_local(uint64_t mask, const int* spos) {
    uint32_t out = 0;
    for (int j = 0; j < kTileBits; ++j)
        if ((mask >> spos[j]) & 1ull) out |= 1u << j;
    return out;
}

struct XG {  // terms of one x-mask
    uint64_t x;
    std::vector<uint32_t> terms;
};

static TiledPlan* build_plan(int n, size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs, uint64_t hash) {
    TiledPlan* plan = new TiledPlan();
    plan->hash = hash;
    plan->n = n;
    const uint64_t low_mask = (1ull << kLowBits) - 1ull;

    // ---- x-groups
    std::vector<XG> xg;
    {
        std::unordered_map<uint64_t, uint32_t> index;
        for (size_t k = 0; k < T; ++k) {
            auto it = index.find(xs[k]);
            if (it == index.end()) {
                index.emplace(xs[k], (uint32_t)xg.size());
                xg.push_back({xs[k], {}});
                xg.back().terms.push_back((uint32_t)k);
            } else {
                xg[it->second].terms.push_back((uint32_t)k);
            }
        }
    }
    // ---- which groups can be tiled at all
    std::vector<uint32_t> uncovered;
    int diag_group = -1;
    for (uint32_t g = 0; g < xg.size(); ++g) {
        const uint64_t x = xg[g].x;
        if (x == 0) {
            diag_group = (int)g;
        } else if (__builtin_popcountll(x) <= kRBits && __builtin_popcountll(x | low_mask) <= kTileBits) {
            uncovered.push_back(g);
        } else {
            for (uint32_t t : xg[g].terms) {
                plan->left_x.push_back(xs[t]);
                plan->left_z.push_back(zs[t]);
                plan->left_c.push_back(cs[t]);
            }
        }
    }

    std::vector<TRBlock> rblocks;
    std::vector<TGroup> groups;
    std::vector<TSub> subs;
    std::vector<double> tables;
    std::vector<uint32_t> rest_tab;
    std::vector<uint64_t> run_off;

    // ---- passes: greedy cover of the x-masks by 12-bit sets S that contain the low bits
    const int free_slots = std::min(kTileBits - kLowBits, n - kLowBits);
    bool diag_done = diag_group < 0;
    while (!uncovered.empty() || !diag_done) {
        uint64_t S = low_mask;
        for (int slot = 0; slot < free_slots; ++slot) {
            long long best_score = -1;
            int best_bit = -1;
            for (int b = kLowBits; b < n; ++b) {
                if ((S >> b) & 1ull) continue;
                const uint64_t cand = S | (1ull << b);
                long long score = 0;
                for (uint32_t g : uncovered) {
                    const int missing = __builtin_popcountll(xg[g].x & ~cand);
                    // a mask still needs `missing` more bits; it can only be completed if enough slots remain
                    if (missing <= free_slots - slot - 1) score += 4096ll >> (3 * missing);
                }
                if (score > best_score) {
                    best_score = score;
                    best_bit = b;
                }
            }
            S |= 1ull << best_bit;
        }
        int spos[kTileBits];
        {
            uint64_t m = S;
            for (int j = 0; j < kTileBits; ++j) {
                spos[j] = __builtin_ctzll(m);
                m &= m - 1;
            }
        }
        // groups of this pass
        std::vector<uint32_t> mine, rest_unc;
        for (uint32_t g : uncovered) ((xg[g].x & ~S) == 0 ? mine : rest_unc).push_back(g);
        if (mine.empty() && diag_done) {
            // cannot make progress (should not happen): hand the rest to the sweep kernel
            for (uint32_t g : uncovered)
                for (uint32_t t : xg[g].terms) {
                    plan->left_x.push_back(xs[t]);
                    plan->left_z.push_back(zs[t]);
                    plan->left_c.push_back(cs[t]);
                }
            uncovered.clear();
            break;
        }
        uncovered.swap(rest_unc);
        if (!diag_done) {
            mine.push_back((uint32_t)diag_group);
            diag_done = true;
        }

        TPass pass;
        pass.s_mask = S;
        pass.first_rblock = (uint32_t)rblocks.size();
        pass.run_table = (uint32_t)run_off.size();
        pass.pad = 0;
        for (uint32_t u = 0; u < (1u << (kTileBits - kLowBits)); ++u) {
            uint64_t off = 0;
            for (int j = kLowBits; j < kTileBits; ++j)
                if ((u >> (j - kLowBits)) & 1u) off |= 1ull << spos[j];
            run_off.push_back(off);
        }

        // ---- R-blocks: bucket the pass's groups into 5-bit sub-cubes (largest masks first)
        std::vector<uint32_t> xl(mine.size());
        for (size_t i = 0; i < mine.size(); ++i) xl[i] = to_tile_local(xg[mine[i]].x, spos);
        std::vector<size_t> ord(mine.size());
        std::iota(ord.begin(), ord.end(), 0);
        std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return __builtin_popcount(xl[a]) > __builtin_popcount(xl[b]); });
        std::vector<char> assigned(mine.size(), 0);
        for (size_t oi = 0; oi < ord.size(); ++oi) {
            const size_t seed = ord[oi];
            if (assigned[seed]) continue;
            uint32_t R = xl[seed];
            while (__builtin_popcount(R) < kRBits) {
                int best_b = -1, best_cnt = -1;
                for (int b = 0; b < kTileBits; ++b) {
                    if ((R >> b) & 1u) continue;
                    const uint32_t cand = R | (1u << b);
                    int cnt = 0;
                    for (size_t i = 0; i < mine.size(); ++i)
                        if (!assigned[i] && (xl[i] & ~cand) == 0) ++cnt;
                    if (cnt > best_cnt) {
                        best_cnt = cnt;
                        best_b = b;
                    }
                }
                R |= 1u << best_b;
            }
            int rpos[kRBits];
            {
                uint32_t m = R;
                for (int j = 0; j < kRBits; ++j) {
                    rpos[j] = __builtin_ctz(m);
                    m &= m - 1;
                }
            }
            uint64_t r_global = 0;
            for (int j = 0; j < kRBits; ++j) r_global |= 1ull << spos[rpos[j]];

            TRBlock blk;
            memset(&blk, 0, sizeof(blk));
            for (int j = 0; j < kRBits; ++j) blk.sbit[j] = (uint16_t)swizzle(1u << rpos[j]);
            // thread bit -> rest position; the three lowest thread bits get positions with distinct residues mod 3
            // (distinct swizzle columns => conflict-free 128-bit shared loads inside a quarter warp)
            std::vector<int> rest_pos;
            for (int b = 0; b < kTileBits; ++b)
                if (!((R >> b) & 1u)) rest_pos.push_back(b);
            std::vector<int> perm;
            {
                std::vector<char> used(rest_pos.size(), 0);
                for (int res = 0; res < 3; ++res)
                    for (size_t i = 0; i < rest_pos.size(); ++i)
                        if (!used[i] && rest_pos[i] % 3 == res) {
                            perm.push_back(rest_pos[i]);
                            used[i] = 1;
                            break;
                        }
                for (size_t i = 0; i < rest_pos.size(); ++i)
                    if (!used[i]) perm.push_back(rest_pos[i]);
            }
            blk.rest_table = (uint32_t)rest_tab.size();
            for (uint32_t r = 0; r < (uint32_t)kTileThreads; ++r) {
                uint32_t org = 0;
                for (size_t j = 0; j < perm.size(); ++j)
                    if ((r >> j) & 1u) org |= 1u << perm[j];
                rest_tab.push_back((org << 16) | swizzle(org));
            }
            blk.first_group = (uint32_t)groups.size();

            for (size_t i = 0; i < mine.size(); ++i) {
                if (assigned[i] || (xl[i] & ~R) != 0) continue;
                assigned[i] = 1;
                const XG& g = xg[mine[i]];
                TGroup tg;
                tg.xr = 0;
                for (int j = 0; j < kRBits; ++j)
                    if ((xl[i] >> rpos[j]) & 1u) tg.xr |= 1u << j;
                tg.has_im = 0;
                tg.first_sub = (uint32_t)subs.size();
                const int npairs_tab = tg.xr ? kMaxPairs : 32;
                const int hb = tg.xr ? (31 - __builtin_clz(tg.xr)) : -1;
                // subgroups keyed by (z outside R, parity type)
                std::map<std::pair<uint64_t, int>, uint32_t> key_to_sub;
                for (uint32_t t : g.terms) {
                    const uint64_t z = zs[t];
                    const int ny = __builtin_popcountll(g.x & z);
                    const int odd = ny & 1;
                    const double d = cs[t] * (((ny >> 1) & 1) ? -1.0 : 1.0);
                    const uint64_t zkey = z & ~r_global;
                    auto key = std::make_pair(zkey, odd);
                    auto it = key_to_sub.find(key);
                    uint32_t si;
                    if (it == key_to_sub.end()) {
                        si = (uint32_t)subs.size();
                        key_to_sub.emplace(key, si);
                        TSub sub;
                        sub.z_outer = zkey & ~S;
                        sub.z_rest = to_tile_local(zkey & S, spos);
                        sub.use_im = (uint32_t)odd;
                        sub.table = (uint32_t)tables.size();
                        sub.pad = 0;
                        subs.push_back(sub);
                        tables.resize(tables.size() + npairs_tab, 0.0);
                        if (odd) tg.has_im = 1;
                    } else {
                        si = it->second;
                    }
                    uint32_t zr = 0;  // z inside R as a 5-bit pattern
                    for (int j = 0; j < kRBits; ++j)
                        if ((z >> spos[rpos[j]]) & 1ull) zr |= 1u << j;
                    double* tab = tables.data() + subs[si].table;
                    if (tg.xr == 0) {
                        for (int p = 0; p < 32; ++p) tab[p] += (__builtin_popcount(p & zr) & 1) ? -d : d;
                    } else {
                        int kk = 0;
                        for (int p = 0; p < 32; ++p) {
                            if ((p >> hb) & 1) continue;
                            tab[kk++] += 2.0 * ((__builtin_popcount(p & zr) & 1) ? -d : d);
                        }
                    }
                }
                tg.n_sub = (uint32_t)subs.size() - tg.first_sub;
                groups.push_back(tg);
            }
            blk.n_groups = (uint32_t)groups.size() - blk.first_group;
            rblocks.push_back(blk);
        }
        pass.n_rblocks = (uint32_t)rblocks.size() - pass.first_rblock;
        plan->pass_groups.push_back((uint32_t)mine.size());
        plan->passes.push_back(pass);
    }
    plan->n_groups = groups.size();
    plan->n_rblocks = rblocks.size();
    plan->n_subs = subs.size();

    // ---- one device allocation for everything
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_rb = 0;
    const size_t o_gr = align(o_rb + rblocks.size() * sizeof(TRBlock));
    const size_t o_sb = align(o_gr + groups.size() * sizeof(TGroup));
    const size_t o_tb = align(o_sb + subs.size() * sizeof(TSub));
    const size_t o_rt = align(o_tb + tables.size() * sizeof(double));
    const size_t o_ru = align(o_rt + rest_tab.size() * sizeof(uint32_t));
    const size_t total = align(o_ru + run_off.size() * sizeof(uint64_t)) + 256;
    if (cudaMalloc(&plan->dev, total) != cudaSuccess) {
        delete plan;
        return nullptr;
    }
    std::vector<char> staging(total, 0);
    memcpy(staging.data() + o_rb, rblocks.data(), rblocks.size() * sizeof(TRBlock));
    memcpy(staging.data() + o_gr, groups.data(), groups.size() * sizeof(TGroup));
    memcpy(staging.data() + o_sb, subs.data(), subs.size() * sizeof(TSub));
    memcpy(staging.data() + o_tb, tables.data(), tables.size() * sizeof(double));
    memcpy(staging.data() + o_rt, rest_tab.data(), rest_tab.size() * sizeof(uint32_t));
    memcpy(staging.data() + o_ru, run_off.data(), run_off.size() * sizeof(uint64_t));
    if (cudaMemcpy(plan->dev, staging.data(), total, cudaMemcpyHostToDevice) != cudaSuccess) {
        free_tiled_plan(plan);
        return nullptr;
    }
    plan->d_rblock = (TRBlock*)(plan->dev + o_rb);
    plan->d_group = (TGroup*)(plan->dev + o_gr);
    plan->d_sub = (TSub*)(plan->dev + o_sb);
    plan->d_tables = (double*)(plan->dev + o_tb);
    plan->d_rest = (uint32_t*)(plan->dev + o_rt);
    plan->d_run = (uint64_t*)(plan->dev + o_ru);
    return plan;
}

// Tiled evaluation of sum_k c_k Re<psi|P_k|psi>.  The device-side sum of all passes lands in h->d_out[slot].
// Leftover terms (not tileable) are returned through the plan for the sweep kernel.
int run_tiled_expectation(qc_state* h, size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs, int out_slot,
                          int* n_passes_out, const TiledPlan** plan_out) {
    uint64_t hash = 0xcbf29ce484222325ull ^ (uint64_t)T ^ ((uint64_t)h->n << 56);
    hash = fnv1a(xs, T * 8, hash);
    hash = fnv1a(zs, T * 8, hash);
    hash = fnv1a(cs, T * 8, hash);
    if (!h->k2plan || h->k2plan->hash != hash || h->k2plan->n != h->n) {
        QC_CHECK(cudaStreamSynchronize(h->stream));
        free_tiled_plan(h->k2plan);
        h->k2plan = build_plan(h->n, T, xs, zs, cs, hash);
        QC_REQUIRE(h->k2plan != nullptr, "qc_expectation: could not build / upload the tiled plan");
    }
    const TiledPlan* plan = h->k2plan;
    static bool attr_set = false;
    const size_t smem = (size_t)kTileSize * sizeof(double2);
    if (!attr_set) {
        QC_CHECK(cudaFuncSetAttribute(expectation_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const uint64_t n_tiles = 1ull << (h->n - kTileBits);
    const int nb = (int)std::min<uint64_t>(n_tiles, (uint64_t)h->sm_count * 2);
    const size_t np = plan->passes.size();
    if (h->ensure_partials(np * nb) || h->ensure_out((size_t)out_slot + 1)) return 1;
    for (size_t p = 0; p < np; ++p) {
        const TPass& pass = plan->passes[p];
        // algorithmic bytes: one shard read per x-group served; the pass itself reads the shard ONCE, so
        // the actual HBM traffic of this class is launches * 16 * dim
        {
            LaunchScope ls(h, kK2Tiled, 16.0 * (double)h->dim * (double)plan->pass_groups[p]);
            expectation_tiled_kernel<<<nb, kTileThreads, smem, h->stream>>>(h->psi, h->n, pass, plan->d_rblock, plan->d_group,
                                                                          plan->d_sub, plan->d_tables, plan->d_rest,
                                                                          plan->d_run, h->d_partials + p * nb);
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
    *x = plan->left_x.data();
    *z = plan->left_z.data();
    *c = plan->left_c.data();
    *count = plan->left_x.size();
}

void tiled_plan_stats(const TiledPlan* plan, size_t* passes, size_t* rblocks, size_t* groups, size_t* subs) {
    *passes = plan->passes.size();
    *rblocks = plan->n_rblocks;
    *groups = plan->n_groups;
    *subs = plan->n_subs;
}

}  // namespace qc
