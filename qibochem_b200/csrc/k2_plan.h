// Plan of the tiled expectation (K2T): which x-masks are served by which HBM pass / register sub-cube, and the
// record stream the kernel walks.  Pure host data + the few layout helpers the kernel shares with the planner.
// The planner (k2_plan.cpp) needs no GPU; expectation_tiled.cu uploads a HostPlan and runs it.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#if defined(__CUDACC__)
#define QC_HD __host__ __device__
#else
#define QC_HD
#endif

namespace qc {

constexpr int kTileBits = 12;                       // amplitudes of one tile: 4096 (64 KB of shared memory)
constexpr int kTileSize = 1 << kTileBits;
constexpr int kRBits = 5;                           // register sub-cube: 32 amplitudes per thread
constexpr int kTileThreads = kTileSize >> kRBits;   // 128
constexpr int kPairsPerCube = 16;                   // pairs {p, p^x} of a sub-cube
constexpr int kSlots = 4;                           // pattern sums A[0..3] shared by the LIN entries of a group

// tile-local amplitude index -> shared-memory slot (16-byte units); XORs every 3-bit field into the lowest one so
// that the 8 lanes of a quarter warp hit 8 distinct 16-byte bank groups for any choice of sub-cube bits
QC_HD inline uint32_t k2_swizzle(uint32_t t) { return t ^ ((t >> 3) & 7u) ^ ((t >> 6) & 7u) ^ ((t >> 9) & 7u); }

// Per-pass "blob": [TRBlock headers][record stream], staged ONCE per CTA into shared memory next to the tile, so
// that walking the records costs broadcast LDS instead of dependent global loads.
constexpr uint32_t kBlobCapUnits = 2816;            // 44 KB of 16-byte units per pass (tile 64 KB + blob: 2 CTAs/SM)

struct TPass {             // kernel parameter (constant bank)
    uint64_t s_mask;       // the kTileBits index bits of the tile (always contains bit 0)
    uint8_t spos[kTileBits];  // index-bit position of every tile-local bit, ascending
    uint32_t blob_off;     // first unit of this pass's blob in the plan's unit array
    uint32_t blob_units;   // <= kBlobCapUnits
    uint32_t n_rblocks;
    uint32_t n_groups;     // x-masks served by this pass
    uint32_t n_generic;    // records outside the straight-line quad path (0: the lean kernel variant runs the pass)
    uint32_t pad;
};

// A sub-cube visit.  The 32 amplitudes a thread holds are  c ^ XOR_{i in E} gen_i  (E = 0..31) for five GENERATORS
// gen_i (tile-local masks, linearly independent) and the thread's representative c:
//   * parity cube: gen_i = {q_i, q_5} for six tile bits q_0..q_5 -- the span is every EVEN-weight subset of the six
//     bits, so one gather serves up to 15 four-bit and 15 two-bit x-masks (molecular Hamiltonians only have even ones);
//   * unit cube:   gen_i = {r_i} for five tile bits (x-masks of odd weight).
// An x-mask in the span has the 5-bit coordinate vector xr (which generators it is the XOR of); the pairs of a record
// are the elements E and E ^ xr.
struct TRBlock {           // 32 bytes = 2 units, followed by 6 units of origin tables (see kRBlockUnits)
    uint16_t sbit[kRBits]; // swizzled tile slot offset of each generator (the swizzle is XOR-linear)
    uint16_t quad_mask;    // bit s: the visit starts with a plain DOT/ACC/Re record for xr = kQuadKappa[s] ("simple
                           // quad"), in ascending s -- the kernel runs these from straight-line code without dispatch
    uint32_t pad;          // bit 0: the first quad record carries two tables
    uint32_t rec_off;      // first record (quad records first), in units from the start of the blob
    uint32_t n_recs;       // generic records following the popcount(quad_mask) quad records
    uint32_t n_groups;
    uint32_t hop_mask;     // bit s: a HOP record for xr = kHopKappa[s] follows the quad records, in ascending s
};
static_assert(sizeof(TRBlock) == 32, "TRBlock layout");
// coordinate vectors with a straight-line path in the kernel: the weight-4 and weight-3 ones (in a parity cube these
// are exactly the 15 four-bit x-masks: weight 3 means the sixth bit takes part)
constexpr int kQuadSlots = 15;
constexpr uint32_t kQuadKappa[kQuadSlots] = {30, 29, 27, 23, 15, 7, 11, 13, 14, 19, 21, 22, 25, 26, 28};
// HOP record: the x-masks of weight 2 of a molecular Hamiltonian ("hopping" strings X_p Z..Z X_q, Y_p Z..Z Y_q, bare
// and decorated with one Z_r) need one ACC table plus two pattern sums A0, A1 that many LIN entries combine.  All
// three tables weight the SAME 16 products, so one record carries them and the kernel forms the products once, from
// straight-line code selected by hop_mask (xr of weight 1 or 2 in a parity cube):
//   header (mode kModeHop, count = LIN2 entries; kMetaHopFlip0/1: A_j changes sign with the representative's sixth
//   bit rho) + 8 units ACC table + 2 x 8 units pattern tables [+ 8 units: the ACC table for rho = 1, kMetaTwoTables]
//   + count entries of 32 bytes {z_outer, z_rest, pad, c0, c1}:
//     acc += (-1)^{popc(i & z_hdr)} w_acc  +  sum_e (-1)^{popc(i & z_e)} (c0 A0 + c1 A1)
constexpr int kHopSlots = 15;
constexpr uint32_t kHopKappa[kHopSlots] = {1, 2, 4, 8, 16, 3, 5, 6, 9, 10, 12, 17, 18, 20, 24};
constexpr uint32_t kHopHeadUnits = 1 + 3 * 8;
constexpr uint32_t kHopEntryUnits = 2;
constexpr uint32_t kMetaHopFlip0 = 1u << 14, kMetaHopFlip1 = 1u << 15;
// A sub-cube visit in the blob: TRBlock (2 units) + origin tables (6 units): uint32 lo[16], hi[8] with
// (origin << 16 | swizzled origin) of thread t = lo[t & 15] ^ hi[t >> 4]  (the swizzle is XOR-linear and the two
// halves deposit disjoint bits), so a thread finds its representative with two LDS instead of ~40 ALU ops.
constexpr uint32_t kRBlockUnits = 8;

// ---- record stream (16-byte units) ----------------------------------------------------------------------------
// DOT / DIAG record: header + 16 doubles.   w = sum_k tab[k] * prod_k  with
//     DOT   : prod_k = Re or Im of conj(a_p) a_{p^xr}, p = k-th sub-cube index with the highest bit of xr clear
//     DIAG_x: prod_k = |a_{k + 16 x}|^2
//   mode ACC: acc += (-1)^{popc(i & z)} w   (z = header masks, i = thread's index outside the sub-cube)
//   mode SET / ADD: A[slot] = / += w
// LIN record: header (count in meta) + count entries of 48 bytes {z_outer, z_rest, pad, c[4]}:
//     acc += (-1)^{popc(i & z)} (c0 A0 + c1 A1 + c2 A2 + c3 A3)
struct TRecHdr {
    uint64_t z_outer;      // z outside the tile (index-bit positions)
    uint32_t z_rest;       // z on the tile bits outside the sub-cube (tile-local positions)
    uint32_t meta;
};
enum : uint32_t { kRecDot = 0, kRecDiagLo = 1, kRecDiagHi = 2, kRecLin = 3 };
enum : uint32_t { kModeAcc = 0, kModeSet = 1, kModeAdd = 2, kModeHop = 3 };
QC_HD inline uint32_t rec_meta(uint32_t xr, uint32_t im, uint32_t kind, uint32_t mode, uint32_t slot, uint32_t count) {
    return (xr & 31u) | ((im & 1u) << 5) | ((kind & 3u) << 6) | ((mode & 3u) << 8) | ((slot & 3u) << 10) | (count << 16);
}
QC_HD inline uint32_t meta_xr(uint32_t m) { return m & 31u; }
QC_HD inline uint32_t meta_im(uint32_t m) { return (m >> 5) & 1u; }
QC_HD inline uint32_t meta_kind(uint32_t m) { return (m >> 6) & 3u; }
QC_HD inline uint32_t meta_mode(uint32_t m) { return (m >> 8) & 3u; }
QC_HD inline uint32_t meta_slot(uint32_t m) { return (m >> 10) & 3u; }
QC_HD inline uint32_t meta_count(uint32_t m) { return m >> 16; }
// A DOT / DIAG record of a parity cube whose terms differ in z on the cube's sixth bit carries TWO weight tables: the
// representative's sixth bit rho (thread-id bit 6, warp-uniform) selects tab[rho]  (w(E, rho) = A(E) + (-1)^rho B(E)).
constexpr uint32_t kMetaTwoTables = 1u << 12;
// Set on a simple-quad record when the NEXT quad record of the visit carries two tables (TRBlock::pad bit 0 says it
// for the first one): the kernel then knows a record's table address before it has loaded the record's own header.
constexpr uint32_t kMetaNextTwoTables = 1u << 13;
constexpr uint32_t kDotUnits = 1 + 8;   // header + 16 doubles
constexpr uint32_t kDot2Units = 1 + 16; // header + 2 x 16 doubles
QC_HD inline uint32_t dot_units(uint32_t meta) { return (meta & kMetaTwoTables) ? kDot2Units : kDotUnits; }
constexpr uint32_t kLinEntryUnits = 3;  // 48 bytes
QC_HD inline uint32_t hop_units(uint32_t meta) {
    return kHopHeadUnits + ((meta & kMetaTwoTables) ? 8u : 0u) + kHopEntryUnits * meta_count(meta);
}

struct Unit16 {
    uint64_t lo, hi;
};

struct PlanStats {
    uint64_t n_groups = 0;       // distinct x-masks in the term list
    uint64_t n_tiled_groups = 0; // served by tiles
    uint64_t n_passes = 0, n_rblocks = 0, n_dot = 0, n_lin = 0, n_leftover_terms = 0;
    uint64_t rblock_hist[8] = {0};  // R-blocks by number of groups (>= 7 in the last bin)
};

struct HostPlan {
    int n = 0;
    std::vector<TPass> passes;
    std::vector<Unit16> units;   // all blobs, back to back
    // terms whose x-mask does not fit a tile / a sub-cube: evaluated by the sweep kernel
    std::vector<uint64_t> left_x, left_z;
    std::vector<double> left_c;
    PlanStats stats;
};

// Builds the plan for sum_k c_k Re<psi|P_k|psi> on an n-bit shard (n >= kTileBits).  No GPU needed.
HostPlan* build_host_plan(int n, size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs);

// Planner self-check (tests): for `samples` random (term-group, amplitude index) picks, the weight
// W_x(i) = sum_{k in group x} d_k (-1)^{popc(i & z_k)} reconstructed from the plan's records vs computed from the
// term list.  Returns the largest absolute difference (0 when the plan is exact), or -1 if a group is served by
// no record / more than one sub-cube.
double check_host_plan(const HostPlan* plan, size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs,
                       size_t samples, uint64_t seed);

}  // namespace qc
