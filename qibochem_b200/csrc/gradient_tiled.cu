// K6T: lambda (+)= H psi through shared-memory tiles (first stage of the adjoint gradient, gradient.cu).
//
// Reference behaviour replaced: qibo.derivative.parameter_shift (doc/source/tutorials/adaptive.rst:46-61, :118-141 of
// the reference) evaluates the energy twice PER parameter; the adjoint method needs lambda = H psi once.  The sweep
// version (hpsi_kernel) reads the shard once per <= 16 x-masks; here a PASS fixes 12 index bits S (the tile sets of
// the K2T planner, k2_plan.cpp) and one HBM read of psi plus one read-modify-write of lambda serves EVERY x-mask
// inside S (~80 per pass for a molecular Hamiltonian):
//
//   lambda_i += sum_{x in pass} (Wre_x(i) + i Wim_x(i)) psi_{i ^ x},   W_x(i) = sum_k d_k (-1)^{popc(i & z_k)}
//
// A CTA of 256 threads stages the 4096 amplitudes of a tile (64 KB, cp.async); thread t owns the 16 tile elements
// e_j = t | j << 8 and keeps their lambda in registers for the whole tile.  The sign of a term splits into
// (tile origin, thread bits) -- one POPC per term, thread and tile -- and the 4 element ("j") bits.  On the element bits
// OUTSIDE x the terms of an x-mask are grouped into SUB-SEGMENTS that share one 16-bit Walsh word; on the element bits
// INSIDE x (the Y sites of XXYY / YYXX / ... differ there) every term carries its 2^M signs, M = the number of such
// bits, and the 2^M partial sums are indexed at compile time (16-way switch on x's element-bit part, which also
// makes the partner offsets compile-time constants).  The partner amplitudes psi_{e ^ x} come from shared memory
// (conflict-free: XOR with x permutes a warp's 32 consecutive slots), so the kernel is bound by those LDS.128
// (64 KB per x-mask and tile), not by HBM; all per-pass metadata is a blob in shared memory.
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>

#include "common.cuh"
#include "k2_plan.h"

namespace qc {

constexpr int kHpThreads = 256;
constexpr int kHpElems = kTileSize / kHpThreads;  // 16 elements per thread: tile-local bits 8..11

// Per pass, everything the inner loops read besides the amplitudes is one BLOB of 16-byte units staged in shared memory
// next to the tile (like K2T): [HpGroup x n_groups][HpSeg x n_segs][HpTerm x n_terms].  From global memory the three
// dependent loads per x-mask (group -> segment -> terms) cost an L2 round trip each (the L1 is carved down to ~28 KB by the
// tile) and the kernel ran latency-bound at a tenth of its shared-memory bound.
constexpr uint32_t kHpBlobCapUnits = 2816;  // 44 KB: tile 64 KB + blob, two CTAs per SM

struct HpPass {            // kernel parameter
    uint64_t s_mask;
    uint8_t spos[kTileBits];  // index-bit position of tile-local bit b (0..7 thread bits, 8..11 element bits)
    uint32_t blob_off;     // first unit of the pass's blob in the plan's unit array
    uint32_t blob_units;
    uint32_t n_groups, n_segs;
};
struct HpGroup {           // 16 bytes
    uint32_t x_loc;        // pairing mask in tile-local positions
    uint32_t first_seg;    // unit index inside the blob
    uint16_t n_seg_re, n_seg_im;  // segments of Re-type (even nY) terms first
    uint32_t pad;
};
struct HpSeg {             // 16 bytes: a sub-segment (Re-type terms) or a segment (Im-type terms)
    uint32_t first_term;   // unit index inside the blob
    uint32_t n_terms;
    uint32_t walsh;        // bit j = parity(j & z_j), z_j = the common z on the element bits (Re-type: outside x only)
    uint32_t pad;
};
struct HpTerm {            // 16 bytes
    uint64_t z;            // bits 0..39: z outside the tile (index-bit positions, n <= 40); bits 40..55 (Re-type terms): the
                           // 2^M signs of the term's z on the element bits inside x; bits 56..63: z on the 8 thread bits
    double d;              // signed coefficient (conventions of hpsi_kernel)
};
static_assert(sizeof(HpGroup) == 16 && sizeof(HpSeg) == 16 && sizeof(HpTerm) == 16, "blob units");

struct HpPlan {
    std::vector<uint64_t> key_x, key_z;
    std::vector<double> key_c;
    int n = 0;
    std::vector<HpPass> passes;
    uint4* d_units = nullptr;
    std::vector<uint64_t> left_x, left_z;
    std::vector<double> left_c;
    size_t n_groups = 0;
    ~HpPlan() {
        if (d_units) cudaFree(d_units);
    }
};

void free_hp_plan(HpPlan* p) { delete p; }

__device__ __forceinline__ void hp_cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

__device__ __forceinline__ double hp_sign(double v, unsigned par) {
    return __longlong_as_double(__double_as_longlong(v) ^ ((long long)(par & 1u) << 63));
}

// All Re-type (even nY) terms of one x-mask.  XJ = the x-mask's bits on the 4 element bits, a template parameter, so that
//  (a) the partner of element j sits at the COMPILE-TIME offset (j ^ XJ) * 256 slots from the thread's base: the 16
//      loads need no address arithmetic;
//  (b) the terms of an x-mask differ in z on the x-mask's OWN bits (the Y sites of XXYY / YYXX / ...), and a random quad
//      has 1.3 of its four bits among the element bits, which used to split it into 2.7 "segments" of 16 signed
//      additions each.  Instead: a SUB-SEGMENT collects the terms that agree on the element bits OUTSIDE x (one Walsh
//      word for the 16 elements), each term carries the 2^M signs of its z on the M = popcount(XJ) element bits INSIDE x,
//      V[u] = sum_k (+-)c_k over u = 0 .. 2^M - 1 is formed once, and element j takes V[pext(j, XJ)] -- a compile-time
//      index -- times its Walsh sign.
__host__ __device__ constexpr int hp_pext4(int j, int mask) {
    int out = 0, pos = 0;
    for (int b = 0; b < 4; ++b)
        if ((mask >> b) & 1) {
            out |= ((j >> b) & 1) << pos;
            ++pos;
        }
    return out;
}

template <int XJ>
__device__ __forceinline__ void hp_group_re(double2 (&acc)[kHpElems], const double2* __restrict__ base,
                                            const uint4* __restrict__ blob, uint32_t first_sub, uint32_t n_sub, uint32_t sw_lo,
                                            uint32_t sw_hi) {
    constexpr int M = ((XJ >> 0) & 1) + ((XJ >> 1) & 1) + ((XJ >> 2) & 1) + ((XJ >> 3) & 1);
    constexpr int NV = 1 << M;
    auto sub_values = [&](const uint4& sraw, double (&V)[NV]) {
#pragma unroll
        for (int u = 0; u < NV; ++u) V[u] = 0.0;
        const uint4* tp = blob + sraw.x;
        for (uint32_t k = 0; k < sraw.y; ++k) {
            const uint4 q = tp[k];
            const double c = hp_sign(__hiloint2double((int)q.w, (int)q.z), __popc((sw_lo & q.x) ^ (sw_hi & q.y)));
            const uint32_t wy = q.y >> 8;  // bits 8..23 of the high word: the term's 2^M signs (bit u)
#pragma unroll
            for (int u = 0; u < NV; ++u) V[u] += NV == 1 ? c : hp_sign(c, wy >> u);
        }
    };
    // one apply per sub-segment (a quad has one; a hopping group a handful: its extra partner loads are cheaper than
    // a 16-entry weight array next to V and the accumulators in 128 registers)
    for (uint32_t s = 0; s < n_sub; ++s) {
        const uint4 sraw = blob[first_sub + s];
        double V[NV];
        sub_values(sraw, V);
        const uint32_t walsh = sraw.z;
        if (walsh == 0u) {
#pragma unroll
            for (int j = 0; j < kHpElems; ++j) {
                const double2 v = base[(j ^ XJ) * kHpThreads];
                const double wj = V[hp_pext4(j, XJ)];
                acc[j].x = fma(wj, v.x, acc[j].x);
                acc[j].y = fma(wj, v.y, acc[j].y);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kHpElems; ++j) {
                const double2 v = base[(j ^ XJ) * kHpThreads];
                const double wj = hp_sign(V[hp_pext4(j, XJ)], walsh >> j);
                acc[j].x = fma(wj, v.x, acc[j].x);
                acc[j].y = fma(wj, v.y, acc[j].y);
            }
        }
    }
}

__global__ void __launch_bounds__(kHpThreads, 2)
    hpsi_tiled_kernel(const double2* __restrict__ psi, double2* __restrict__ lam, int n, const HpPass pass,
                      const uint4* __restrict__ units, int accumulate) {
    extern __shared__ __align__(16) unsigned char hp_smem[];
    double2* tile = reinterpret_cast<double2*>(hp_smem);
    const uint4* blob = reinterpret_cast<const uint4*>(hp_smem + (size_t)kTileSize * sizeof(double2));
    {
        uint4* dst = reinterpret_cast<uint4*>(hp_smem + (size_t)kTileSize * sizeof(double2));
        const uint4* src = units + pass.blob_off;
        for (uint32_t u = threadIdx.x; u < pass.blob_units; u += kHpThreads) dst[u] = __ldg(&src[u]);
    }
    const uint32_t t = threadIdx.x;
    const uint64_t n_tiles = 1ull << (n - kTileBits);
    uint64_t off_thr = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) off_thr |= (uint64_t)((t >> b) & 1u) << pass.spos[b];
    // offset of element j (recomputed where needed: 16 live 64-bit offsets would not fit next to the accumulators)
    auto off_el = [&](int j) {
        uint64_t o = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (j & (1 << b)) o |= 1ull << pass.spos[8 + b];
        return o;
    };
    const uint32_t t_hi = t << 24;  // thread bits aligned with bits 56..63 of a term's z word
    for (uint64_t tau = blockIdx.x; tau < n_tiles; tau += gridDim.x) {
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
        const double2* src = psi + base + off_thr;
#pragma unroll
        for (int j = 0; j < kHpElems; ++j) hp_cp_async16(&tile[(uint32_t)j * kHpThreads + t], src + off_el(j));
        double2 acc[kHpElems];
        if (accumulate) {
            const double2* lsrc = lam + base + off_thr;
#pragma unroll
            for (int j = 0; j < kHpElems; ++j) acc[j] = lsrc[off_el(j)];
        } else {
#pragma unroll
            for (int j = 0; j < kHpElems; ++j) acc[j] = make_double2(0.0, 0.0);
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();

        // sign word of (tile origin, thread): parity of (word & term.z) = parity of (origin & z_outer) ^ (thread & z_thr)
        const uint32_t sw_lo = (uint32_t)base, sw_hi = ((uint32_t)(base >> 32) & 0x00ffffffu) | t_hi;
        uint4 gnext = blob[0];
        for (uint32_t g = 0; g < pass.n_groups; ++g) {
            const uint4 graw = gnext;
            if (g + 1 < pass.n_groups) gnext = blob[g + 1];  // the next header is in flight while this group runs
            const uint32_t x_loc = graw.x;
            const double2* partner = tile + (t ^ (x_loc & 0xffu));
            const uint32_t xj = x_loc >> 8;
            // ---- Re-type terms (even nY): everything in one templated call selected by the element-bit part of x
            {
                const uint32_t n_sub = graw.z & 0xffffu;
                if (n_sub) {
                    const double2* pb = tile + (t ^ (x_loc & 0xffu));
#define QC_HP_CASE(X) case X: hp_group_re<X>(acc, pb, blob, graw.y, n_sub, sw_lo, sw_hi); break;
                    switch (xj) {
                        QC_HP_CASE(0) QC_HP_CASE(1) QC_HP_CASE(2) QC_HP_CASE(3) QC_HP_CASE(4) QC_HP_CASE(5) QC_HP_CASE(6)
                        QC_HP_CASE(7) QC_HP_CASE(8) QC_HP_CASE(9) QC_HP_CASE(10) QC_HP_CASE(11) QC_HP_CASE(12)
                        QC_HP_CASE(13) QC_HP_CASE(14) QC_HP_CASE(15)
                    }
#undef QC_HP_CASE
                }
            }
            // ---- Im-type terms (odd nY; none in a molecular Hamiltonian): segments keyed by the whole z on the element bits
            const uint32_t ns_im = graw.z >> 16;
            if (ns_im) {
                const uint32_t s0 = graw.y + (graw.z & 0xffffu);
                double w[kHpElems];
#pragma unroll
                for (int j = 0; j < kHpElems; ++j) w[j] = 0.0;
                for (uint32_t s = 0; s < ns_im; ++s) {
                    const uint4 sraw = blob[s0 + s];  // first_term, n_terms, walsh
                    const uint4* tp = blob + sraw.x;
                    double bsum = 0.0;
                    for (uint32_t k = 0; k < sraw.y; ++k) {
                        const uint4 q = tp[k];
                        bsum += hp_sign(__hiloint2double((int)q.w, (int)q.z), __popc((sw_lo & q.x) ^ (sw_hi & q.y)));
                    }
#pragma unroll
                    for (int j = 0; j < kHpElems; ++j) w[j] += hp_sign(bsum, sraw.z >> j);
                }
#pragma unroll
                for (int j = 0; j < kHpElems; ++j) {
                    const double2 v = partner[(uint32_t)(j ^ xj) * kHpThreads];
                    acc[j].x = fma(-w[j], v.y, acc[j].x);
                    acc[j].y = fma(w[j], v.x, acc[j].y);
                }
            }
        }
        double2* dst = lam + base + off_thr;
#pragma unroll
        for (int j = 0; j < kHpElems; ++j) dst[off_el(j)] = acc[j];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
// h == nullptr: host-only planning (statistics; no device upload)
static HpPlan* build_hp_plan(qc_state* h, int n, size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc,
                             uint64_t* stats = nullptr) {
    HostPlan* kp = build_host_plan(n, T, hx, hz, hc);  // only its tile sets are used
    if (!kp) return nullptr;
    std::vector<uint64_t> sets;
    for (const TPass& p : kp->passes)
        if (sets.empty() || sets.back() != p.s_mask) sets.push_back(p.s_mask);
    delete kp;
    HpPlan* plan = new HpPlan();
    plan->n = n;
    // x-groups (stable order), assigned to the first tile set that contains them
    std::vector<uint32_t> order(T);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hx[a] < hx[b]; });
    struct XG {
        uint64_t x;
        size_t first, count;
        int pass;
    };
    std::vector<XG> xgs;
    for (size_t k = 0; k < T;) {
        size_t e = k;
        while (e < T && hx[order[e]] == hx[order[k]]) ++e;
        xgs.push_back({hx[order[k]], k, e - k, -1});
        k = e;
    }
    std::vector<std::vector<size_t>> members(sets.size());
    for (size_t g = 0; g < xgs.size(); ++g) {
        for (size_t p = 0; p < sets.size(); ++p)
            if ((xgs[g].x & ~sets[p]) == 0) {
                xgs[g].pass = (int)p;
                members[p].push_back(g);
                break;
            }
        if (xgs[g].pass < 0)
            for (size_t k = xgs[g].first; k < xgs[g].first + xgs[g].count; ++k) {
                plan->left_x.push_back(hx[order[k]]);
                plan->left_z.push_back(hz[order[k]]);
                plan->left_c.push_back(hc[order[k]]);
            }
    }
    std::vector<Unit16> units;  // all blobs, back to back
    for (size_t p = 0; p < sets.size(); ++p) {
        if (members[p].empty()) continue;
        const uint64_t S = sets[p];
        // element bits: the 4 tile bits (never index bit 0) that the z-masks of the pass's terms touch least
        std::vector<int> sbits;
        for (int b = 0; b < 64; ++b)
            if ((S >> b) & 1ull) sbits.push_back(b);
        // score of a tile bit: how many terms of the pass carry z there (a bit no z touches keeps every group at one
        // segment with an empty Walsh word: the kernel's cheapest path)
        std::vector<uint64_t> vary(sbits.size(), 0);
        for (size_t g : members[p])
            for (size_t k = xgs[g].first; k < xgs[g].first + xgs[g].count; ++k)
                for (size_t i = 0; i < sbits.size(); ++i)
                    if ((hz[order[k]] >> sbits[i]) & 1ull) vary[i] += 1;
        std::vector<size_t> idx(sbits.size());
        std::iota(idx.begin(), idx.end(), (size_t)0);
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
            const bool a0 = sbits[a] == 0, b0 = sbits[b] == 0;  // index bit 0 stays a thread bit (32-byte sectors)
            if (a0 != b0) return b0;
            return vary[a] < vary[b];
        });
        HpPass base_pass;
        memset(&base_pass, 0, sizeof(base_pass));
        base_pass.s_mask = S;
        std::vector<int> el(idx.begin(), idx.begin() + 4), th(idx.begin() + 4, idx.end());
        std::sort(el.begin(), el.end());
        std::sort(th.begin(), th.end());
        for (int b = 0; b < 8; ++b) base_pass.spos[b] = (uint8_t)sbits[th[b]];
        for (int b = 0; b < 4; ++b) base_pass.spos[8 + b] = (uint8_t)sbits[el[b]];
        auto to_local = [&](uint64_t m) {
            uint32_t out = 0;
            for (int b = 0; b < kTileBits; ++b)
                if ((m >> base_pass.spos[b]) & 1ull) out |= 1u << b;
            return out;
        };
        // groups of the pass with their segments and terms; a blob that would exceed the cap starts another pass over S
        struct GroupOut {
            HpGroup hdr;
            std::vector<HpSeg> segs;
            std::vector<HpTerm> terms;  // seg.first_term relative to this vector until the blob is laid out
        };
        std::vector<GroupOut> cur;
        uint32_t cur_units = 0;
        auto flush = [&]() {
            if (cur.empty()) return;
            HpPass pass = base_pass;
            pass.blob_off = (uint32_t)units.size();
            pass.n_groups = (uint32_t)cur.size();
            uint32_t n_segs = 0;
            for (const GroupOut& go : cur) n_segs += (uint32_t)go.segs.size();
            pass.n_segs = n_segs;
            uint32_t seg_at = pass.n_groups, term_at = pass.n_groups + n_segs;
            std::vector<Unit16> blob(term_at);
            std::vector<Unit16> term_units;
            uint32_t gi = 0;
            for (GroupOut& go : cur) {
                go.hdr.first_seg = seg_at;
                for (HpSeg sg : go.segs) {
                    sg.first_term += term_at + (uint32_t)term_units.size();
                    memcpy(&blob[seg_at++], &sg, 16);
                }
                for (const HpTerm& tm : go.terms) {
                    Unit16 u;
                    memcpy(&u, &tm, 16);
                    term_units.push_back(u);
                }
                memcpy(&blob[gi++], &go.hdr, 16);
            }
            blob.insert(blob.end(), term_units.begin(), term_units.end());
            pass.blob_units = (uint32_t)blob.size();
            units.insert(units.end(), blob.begin(), blob.end());
            plan->passes.push_back(pass);
            plan->n_groups += cur.size();
            cur.clear();
            cur_units = 0;
        };
        for (size_t g : members[p]) {
            GroupOut go;
            memset(&go.hdr, 0, sizeof(go.hdr));
            go.hdr.x_loc = to_local(xgs[g].x);
            for (int type = 0; type < 2; ++type) {
                // segments keyed by the z pattern on the element bits
                std::map<uint32_t, std::vector<HpTerm>> by_zj;
                for (size_t k = xgs[g].first; k < xgs[g].first + xgs[g].count; ++k) {
                    const uint32_t tix = order[k];
                    const int ny = __builtin_popcountll(hx[tix] & hz[tix]);
                    if ((ny & 1) != type) continue;
                    HpTerm tm;
                    const uint32_t zl = to_local(hz[tix]);
                    tm.z = (hz[tix] & ~S) | ((uint64_t)(zl & 0xffu) << 56);
                    tm.d = type ? -hc[tix] * ((((ny - 1) >> 1) & 1) ? -1.0 : 1.0) : hc[tix] * (((ny >> 1) & 1) ? -1.0 : 1.0);
                    uint32_t key = zl >> 8;
                    if (type == 0) {
                        // Re-type: sub-segment = z on the element bits OUTSIDE x; the part inside x becomes 2^M sign bits
                        const uint32_t xj = go.hdr.x_loc >> 8, inside = key & xj;
                        key &= ~xj;
                        uint32_t yk = 0;
                        int pos = 0, m = 0;
                        for (int b = 0; b < 4; ++b)
                            if ((xj >> b) & 1u) {
                                yk |= ((inside >> b) & 1u) << pos;
                                ++pos, ++m;
                            }
                        uint64_t wy = 0;
                        for (uint32_t u = 0; u < (1u << m); ++u)
                            if (__builtin_popcount(u & yk) & 1) wy |= 1ull << u;
                        tm.z |= wy << 40;  // bits 40..55: unused by z_outer (n <= 40 qubits) and by the thread bits (56..63)
                    }
                    by_zj[key].push_back(tm);
                }
                for (auto& kv : by_zj) {
                    HpSeg sg;
                    memset(&sg, 0, sizeof(sg));
                    sg.first_term = (uint32_t)go.terms.size();
                    sg.n_terms = (uint32_t)kv.second.size();
                    for (uint32_t j = 0; j < 16; ++j)
                        if (__builtin_popcount(j & kv.first) & 1) sg.walsh |= 1u << j;
                    go.terms.insert(go.terms.end(), kv.second.begin(), kv.second.end());
                    go.segs.push_back(sg);
                    (type ? go.hdr.n_seg_im : go.hdr.n_seg_re) += 1;
                }
            }
            const uint32_t need = 1u + (uint32_t)go.segs.size() + (uint32_t)go.terms.size();
            if (need > kHpBlobCapUnits) {  // a single x-mask with thousands of terms: left to the sweep kernel
                for (size_t k = xgs[g].first; k < xgs[g].first + xgs[g].count; ++k) {
                    plan->left_x.push_back(hx[order[k]]);
                    plan->left_z.push_back(hz[order[k]]);
                    plan->left_c.push_back(hc[order[k]]);
                }
                continue;
            }
            if (cur_units + need > kHpBlobCapUnits) flush();
            cur_units += need;
            cur.push_back(std::move(go));
        }
        flush();
    }
    if (stats) {
        // {passes, x-masks, segments, terms, largest number of x-masks in a pass, leftover terms, x-masks with ONE
        //  segment, of those with an empty Walsh word}
        uint64_t n_seg = 0, n_term = 0, max_g = 0, single = 0, plain = 0;
        for (const HpPass& ps : plan->passes) {
            n_seg += ps.n_segs;
            n_term += ps.blob_units - ps.n_groups - ps.n_segs;
            max_g = std::max<uint64_t>(max_g, ps.n_groups);
            for (uint32_t g = 0; g < ps.n_groups; ++g) {
                HpGroup gr;
                memcpy(&gr, &units[ps.blob_off + g], 16);
                if (gr.n_seg_re + gr.n_seg_im == 1) {
                    HpSeg sg;
                    memcpy(&sg, &units[ps.blob_off + gr.first_seg], 16);
                    single += 1;
                    plain += sg.walsh == 0;
                }
            }
        }
        const uint64_t v[8] = {plan->passes.size(), plan->n_groups, n_seg, n_term, max_g, plan->left_x.size(), single, plain};
        for (int i = 0; i < 8; ++i) stats[i] = v[i];
    }
    if (h) {
        const size_t bytes = std::max<size_t>(units.size(), 1) * sizeof(Unit16);
        if (cudaMalloc((void**)&plan->d_units, bytes) != cudaSuccess ||
            (!units.empty() && cudaMemcpy(plan->d_units, units.data(), units.size() * sizeof(Unit16), cudaMemcpyHostToDevice) != cudaSuccess)) {
            delete plan;
            return nullptr;
        }
        h->h2d_bytes += units.size() * sizeof(Unit16);
    }
    plan->key_x.assign(hx, hx + T);
    plan->key_z.assign(hz, hz + T);
    plan->key_c.assign(hc, hc + T);
    return plan;
}

int hpsi_sweeps(qc_state* h, size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc, bool accumulate);

// lambda (+)= H psi through tiles when the shard is large enough; *handled = false leaves everything to the sweeps.
int hpsi_tiled(qc_state* h, size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc, bool accumulate,
               bool* handled) {
    *handled = false;
    if (!(h->n >= 12 && (h->k2_mode == 2 || (h->k2_mode == 0 && h->n >= 16 && T >= 512)))) return 0;
    HpPlan* plan = h->hp_plan;
    if (!(plan && plan->n == h->n && plan->key_x.size() == T && memcmp(plan->key_x.data(), hx, T * 8) == 0 &&
          memcmp(plan->key_z.data(), hz, T * 8) == 0 && memcmp(plan->key_c.data(), hc, T * 8) == 0)) {
        QC_CHECK(cudaStreamSynchronize(h->stream));  // the old plan may still be read by in-flight passes
        delete h->hp_plan;
        h->hp_plan = nullptr;
        plan = build_hp_plan(h, h->n, T, hx, hz, hc);
        QC_REQUIRE(plan != nullptr, "qc_hpsi: could not build the tiled plan");
        h->hp_plan = plan;
    }
    const size_t smem = (size_t)kTileSize * sizeof(double2) + (size_t)kHpBlobCapUnits * 16;
    QC_CHECK(cudaFuncSetAttribute(hpsi_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QC_CHECK(cudaFuncSetAttribute(hpsi_tiled_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const uint64_t n_tiles = 1ull << (h->n - kTileBits);
    const int nb = (int)std::min<uint64_t>(n_tiles, (uint64_t)h->sm_count * 2);
    bool acc = accumulate;
    if (plan->passes.empty() && !accumulate) QC_CHECK(cudaMemsetAsync(h->scratch_psi, 0, h->dim * sizeof(double2), h->stream));
    for (const HpPass& pass : plan->passes) {
        {
            // algorithmic bytes: psi read + lambda read-modify-write per x-mask served (what the sweep version moves)
            LaunchScope ls(h, kGradient, 48.0 * (double)h->dim * (double)pass.n_groups);
            h->h2d_bytes += sizeof(HpPass);
            hpsi_tiled_kernel<<<nb, kHpThreads, smem, h->stream>>>(h->psi, h->scratch_psi, h->n, pass, plan->d_units, acc ? 1 : 0);
        }
        QC_CHECK(cudaGetLastError());
        acc = true;
    }
    if (!plan->left_x.empty() &&
        hpsi_sweeps(h, plan->left_x.size(), plan->left_x.data(), plan->left_z.data(), plan->left_c.data(), true))
        return 1;
    *handled = true;
    return 0;
}

}  // namespace qc

// Host-only (no GPU needed): plans the tiled H|psi> of a term list and reports its shape.  stats_out: 8 x uint64
// {passes, x-masks served by tiles, segments, terms, largest number of x-masks in one pass, leftover terms (sweep
//  kernel), x-masks with a single segment, of those with an empty Walsh word}.
extern "C" int qc_hpsi_plan_stats(int n_local_qubits, size_t T, const uint64_t* xmask, const uint64_t* zmask,
                                  const double* coeff, uint64_t* stats_out) {
    using namespace qc;
    QC_REQUIRE(xmask && zmask && coeff && stats_out, "qc_hpsi_plan_stats: null argument");
    QC_REQUIRE(n_local_qubits >= kTileBits && n_local_qubits <= 40, "qc_hpsi_plan_stats: needs 12..40 local qubits");
    for (size_t k = 0; k < T; ++k)
        QC_REQUIRE((xmask[k] >> n_local_qubits) == 0 && (zmask[k] >> n_local_qubits) == 0,
                   "qc_hpsi_plan_stats: mask has bits outside the shard");
    HpPlan* p = build_hp_plan(nullptr, n_local_qubits, T, xmask, zmask, coeff, stats_out);
    QC_REQUIRE(p != nullptr, "qc_hpsi_plan_stats: planner failed");
    delete p;
    return 0;
}
