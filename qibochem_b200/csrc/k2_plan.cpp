// Host-side planner of the tiled expectation (K2T).  See k2_plan.h for the data it produces and
// expectation_tiled.cu for the kernel that consumes it.
//
// Cost model behind the heuristics (B200, per SM): a sub-cube gather moves the whole 64 KB tile through shared
// memory once (512 clk at 128 B/clk); an x-mask evaluated from registers costs 48 DFMA per thread (~96 clk of the
// FP64 pipe per tile).  So the planner (1) packs as many x-masks as it can into every 5-bit sub-cube R, and (2) lets
// the x-masks that would sit alone in a sub-cube wait for a later pass where they may find company.
#include "k2_plan.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>
#include <unordered_map>

namespace qc {
namespace {

inline int popc64(uint64_t v) { return __builtin_popcountll(v); }
inline int popc32(uint32_t v) { return __builtin_popcount(v); }

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

struct XG {  // terms of one x-mask
    uint64_t x;
    std::vector<uint32_t> terms;
    float weight;  // rough evaluation cost in units of one 4-term group
};

struct Term {  // one term seen from inside a sub-cube
    uint64_t z_not_r;
    uint32_t pk;  // sign pattern over the cube's elements (pair groups: highest x coordinate cleared); bit 5: z on the
                  // parity cube's sixth bit (the sign then also depends on the representative's sixth bit rho)
    int odd;
    double d;
};

uint32_t compress_bits(uint64_t mask, const int* pos, int count) {
    uint32_t out = 0;
    for (int j = 0; j < count; ++j)
        if ((mask >> pos[j]) & 1ull) out |= 1u << j;
    return out;
}

void push_units(std::vector<Unit16>& recs, const void* data, size_t bytes) {
    const size_t units = bytes / 16;
    const size_t at = recs.size();
    recs.resize(at + units);
    memcpy(recs.data() + at, data, bytes);
}

// sign pattern of z-bits `pk` over the 16 canonical pair indices (x != 0) or over one half of the cube (diagonal)
void pattern_table(uint32_t xr, uint32_t pk, int half, double scale, double* tab /*16*/) {
    if (xr == 0) {
        for (int k = 0; k < 16; ++k) tab[k] = (popc32((uint32_t)(k + 16 * half) & pk) & 1) ? -scale : scale;
        return;
    }
    const int hb = 31 - __builtin_clz(xr);
    int kk = 0;
    for (int p = 0; p < 32; ++p) {
        if ((p >> hb) & 1) continue;
        tab[kk++] = (popc32((uint32_t)p & pk) & 1) ? -scale : scale;
    }
}

struct DotRec {
    TRecHdr h;
    double tab[16];
};
struct LinEntry {
    uint64_t z_outer;
    uint32_t z_rest;
    uint32_t pad;
    double c[kSlots];
};
struct HopEntry {
    uint64_t z_outer;
    uint32_t z_rest;
    uint32_t pad;
    double c[2];
};
static_assert(sizeof(HopEntry) == 16 * kHopEntryUnits, "record layout");
static_assert(sizeof(DotRec) == 16 * kDotUnits, "record layout");
static_assert(sizeof(LinEntry) == 16 * kLinEntryUnits, "record layout");

// Records of one x-group inside one sub-cube, as ATOMS that may be placed in different passes (a pass's blob is
// capped): `setup` (the SET/ADD records forming the pattern sums A[slot]) must precede, inside the same sub-cube
// visit, every atom with needs_setup.
struct Atom {
    std::vector<Unit16> units;
    uint32_t n_recs = 0;
    bool needs_setup = false;
    int quad_slot = -1;  // j when this is the single record of a "simple quad" (xr = 31 ^ (1 << j))
    int hop_slot = -1;   // s when this is a HOP record (xr = kHopKappa[s])
};
struct GroupRecs {
    std::vector<Unit16> setup;
    uint32_t setup_recs = 0;
    std::vector<Atom> atoms;
};
constexpr size_t kLinChunk = 200;  // LIN entries per record

// rglob: index-bit positions of the cube's five COORDINATE bits (bit i of an element index E is read there);
// gglob: the five generators as index-bit masks; r_global: the union of the coordinate bits.
// rho_pos: index-bit position of a parity cube's sixth bit (-1 for a unit cube).
GroupRecs emit_group(const XG& g, const uint64_t* zs, const double* cs, uint64_t S, const int* spos, uint64_t r_global,
                     const int* rglob, const uint64_t* gglob, int rho_pos, PlanStats& st) {
    const uint64_t rho_bit = rho_pos >= 0 ? (1ull << rho_pos) : 0ull;
    GroupRecs out;
    const uint32_t xr = compress_bits(g.x, rglob, kRBits);  // coordinate vector of x (x is in the generators' span)
    auto zeta = [&](uint64_t z) {  // sign of element E under z: (-1)^{popc(E & zeta)}  (beyond the representative's part)
        uint32_t out_bits = 0;
        for (int j = 0; j < kRBits; ++j) out_bits |= (uint32_t)(popc64(gglob[j] & z) & 1) << j;
        return out_bits;
    };
    const uint32_t hbbit = xr ? (1u << (31 - __builtin_clz(xr))) : 0u;
    const double pair_scale = xr ? 2.0 : 1.0;
    const int halves = xr ? 1 : 2;

    // ---- subgroups: terms that agree outside the sub-cube (and in parity type)
    std::map<std::pair<uint64_t, int>, std::vector<Term>> subs;
    for (uint32_t t : g.terms) {
        const uint64_t z = zs[t];
        const int ny = popc64(g.x & z);
        Term tm;
        tm.z_not_r = z & ~(r_global | rho_bit);
        tm.pk = (zeta(z) & ~hbbit) | ((z & rho_bit) ? 32u : 0u);
        tm.odd = ny & 1;
        tm.d = cs[t] * (((ny >> 1) & 1) ? -1.0 : 1.0);
        subs[{tm.z_not_r, tm.odd}].push_back(tm);
    }
    // ---- shared sign patterns: worth a slot when >= 3 subgroups use them
    std::map<std::pair<uint32_t, int>, int> freq;  // (pk, odd) -> number of subgroups containing it
    for (auto& kv : subs) {
        std::vector<uint32_t> seen;
        for (const Term& tm : kv.second)
            if (std::find(seen.begin(), seen.end(), tm.pk) == seen.end()) seen.push_back(tm.pk);
        for (uint32_t pk : seen) freq[{pk, kv.first.second}] += 1;
    }
    std::vector<std::pair<int, std::pair<uint32_t, int>>> by_freq;
    for (auto& kv : freq)
        if (kv.second >= 3) by_freq.push_back({kv.second, kv.first});
    std::stable_sort(by_freq.begin(), by_freq.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    if (by_freq.size() > (size_t)kSlots) by_freq.resize(kSlots);
    auto slot_of = [&](uint32_t pk, int odd) -> int {
        for (size_t j = 0; j < by_freq.size(); ++j)
            if (by_freq[j].second.first == pk && by_freq[j].second.second == odd) return (int)j;
        return -1;
    };
    std::vector<const std::vector<Term>*> lin_subs, tab_subs;
    std::vector<std::pair<uint64_t, int>> lin_keys, tab_keys;
    bool slot_used[kSlots] = {false, false, false, false};
    for (auto& kv : subs) {
        bool linear = !by_freq.empty();
        for (const Term& tm : kv.second)
            if (slot_of(tm.pk, tm.odd) < 0) linear = false;
        if (linear) {
            lin_subs.push_back(&kv.second), lin_keys.push_back(kv.first);
        } else {
            tab_subs.push_back(&kv.second), tab_keys.push_back(kv.first);
        }
    }
    if (lin_subs.size() < 3) {  // not worth the pattern sums
        for (size_t i = 0; i < lin_subs.size(); ++i) tab_subs.push_back(lin_subs[i]), tab_keys.push_back(lin_keys[i]);
        lin_subs.clear(), lin_keys.clear();
    }
    for (const auto* sub : lin_subs)
        for (const Term& tm : *sub) slot_used[slot_of(tm.pk, tm.odd)] = true;

    auto header = [&](uint64_t z_not_r, uint32_t meta) {
        TRecHdr h;
        h.z_outer = z_not_r & ~S;
        h.z_rest = compress_bits(z_not_r & S, spos, kTileBits);
        h.meta = meta;
        return h;
    };
    // DOT / DIAG record: one table, or two (rho = 0, 1) when tab1 is given
    auto push_dot = [&](std::vector<Unit16>& units, TRecHdr h, const double* tab0, const double* tab1) {
        if (tab1) h.meta |= kMetaTwoTables;
        push_units(units, &h, sizeof(h));
        push_units(units, tab0, 16 * sizeof(double));
        if (tab1) push_units(units, tab1, 16 * sizeof(double));
    };
    // ---- HOP record: pattern sums on at most two slots + at most one table of its own, all real, xr of weight <= 2
    {
        int hop_slot = -1;
        for (int sl = 0; sl < kHopSlots; ++sl)
            if (kHopKappa[sl] == xr) hop_slot = sl;
        int used[kSlots], n_used = 0;
        bool ok = hop_slot >= 0 && rho_pos >= 0 && !lin_subs.empty() && tab_subs.size() <= 1 && lin_subs.size() <= 1200;
        for (size_t j = 0; ok && j < by_freq.size(); ++j)
            if (slot_used[j]) {
                if (n_used == 2 || by_freq[j].second.second != 0) ok = false;
                else used[n_used++] = (int)j;
            }
        bool any0 = false, any1 = false;
        if (ok && tab_subs.size() == 1) {
            for (const Term& tm : *tab_subs[0]) ((tm.pk & 32u) ? any1 : any0) = true;
            if (tab_keys[0].second != 0) ok = false;
        }
        if (ok) {
            Atom atom;
            atom.hop_slot = hop_slot;
            atom.n_recs = 1;
            uint32_t meta = rec_meta(xr, 0, kRecDot, kModeHop, 0, (uint32_t)lin_subs.size());
            double tabs[3][16], tab_rho1[16];
            for (int t = 0; t < 3; ++t)
                for (int k = 0; k < 16; ++k) tabs[t][k] = 0.0;
            for (int k = 0; k < 16; ++k) tab_rho1[k] = 0.0;
            uint64_t key = 0;
            const bool two = any0 && any1;  // mixed: the representative's sixth bit selects the ACC table
            if (tab_subs.size() == 1) {
                key = tab_keys[0].first | ((any1 && !any0) ? rho_bit : 0ull);
                double pat[16];
                for (const Term& tm : *tab_subs[0]) {
                    pattern_table(xr, tm.pk & 31u, 0, pair_scale * tm.d, pat);
                    const double s1 = ((tm.pk & 32u) && any0) ? -1.0 : 1.0;
                    for (int k = 0; k < 16; ++k) tabs[0][k] += pat[k], tab_rho1[k] += s1 * pat[k];
                }
            }
            if (two) meta |= kMetaTwoTables;
            for (int j = 0; j < n_used; ++j) {
                const uint32_t pk = by_freq[used[j]].second.first;
                pattern_table(xr, pk & 31u, 0, pair_scale, tabs[1 + j]);
                if (pk & 32u) meta |= j == 0 ? kMetaHopFlip0 : kMetaHopFlip1;
            }
            const TRecHdr h = header(key, meta);
            push_units(atom.units, &h, sizeof(h));
            push_units(atom.units, tabs, sizeof(tabs));
            if (two) push_units(atom.units, tab_rho1, sizeof(tab_rho1));
            for (size_t i = 0; i < lin_subs.size(); ++i) {
                HopEntry e;
                memset(&e, 0, sizeof(e));
                const TRecHdr eh = header(lin_keys[i].first, 0);
                e.z_outer = eh.z_outer;
                e.z_rest = eh.z_rest;
                for (const Term& tm : *lin_subs[i]) {
                    const int sl = slot_of(tm.pk, tm.odd);
                    e.c[sl == used[0] ? 0 : 1] += tm.d;
                }
                push_units(atom.units, &e, sizeof(e));
                ++st.n_lin;
            }
            out.atoms.push_back(std::move(atom));
            return out;
        }
    }
    // ---- pattern sums A[slot]
    if (!lin_subs.empty()) {
        for (size_t j = 0; j < by_freq.size(); ++j) {
            if (!slot_used[j]) continue;
            const uint32_t pk = by_freq[j].second.first;
            const int odd = by_freq[j].second.second;
            for (int half = 0; half < halves; ++half) {
                const uint32_t kind = xr ? kRecDot : (half ? kRecDiagHi : kRecDiagLo);
                const TRecHdr h = header(0, rec_meta(xr, (uint32_t)odd, kind, half ? kModeAdd : kModeSet, (uint32_t)j, 0));
                double t0[16], t1[16];
                pattern_table(xr, pk & 31u, half, pair_scale, t0);
                for (int k = 0; k < 16; ++k) t1[k] = -t0[k];  // a pattern with z on the sixth bit flips with rho
                push_dot(out.setup, h, t0, (pk & 32u) ? t1 : nullptr);
                ++out.setup_recs;
            }
        }
        for (size_t done = 0; done < lin_subs.size(); done += kLinChunk) {
            const size_t cnt = std::min(lin_subs.size() - done, kLinChunk);
            Atom atom;
            atom.needs_setup = true;
            atom.n_recs = 1;
            TRecHdr h = header(0, rec_meta(xr, 0, kRecLin, kModeAcc, 0, (uint32_t)cnt));
            push_units(atom.units, &h, sizeof(h));
            for (size_t i = done; i < done + cnt; ++i) {
                LinEntry e;
                memset(&e, 0, sizeof(e));
                const TRecHdr eh = header(lin_keys[i].first, 0);
                e.z_outer = eh.z_outer;
                e.z_rest = eh.z_rest;
                for (const Term& tm : *lin_subs[i]) e.c[slot_of(tm.pk, tm.odd)] += tm.d;
                push_units(atom.units, &e, sizeof(e));
                ++st.n_lin;
            }
            out.atoms.push_back(std::move(atom));
        }
    }
    // ---- subgroups with their own weight table
    for (size_t i = 0; i < tab_subs.size(); ++i) {
        Atom atom;
        bool any0 = false, any1 = false;
        for (const Term& tm : *tab_subs[i]) ((tm.pk & 32u) ? any1 : any0) = true;
        for (int half = 0; half < halves; ++half) {
            const uint32_t kind = xr ? kRecDot : (half ? kRecDiagHi : kRecDiagLo);
            // every term has z on the sixth bit: one table, the representative's sixth bit joins the thread's sign
            const uint64_t key = tab_keys[i].first | ((any1 && !any0) ? rho_bit : 0ull);
            const TRecHdr h = header(key, rec_meta(xr, (uint32_t)tab_keys[i].second, kind, kModeAcc, 0, 0));
            double t0[16], t1[16], pat[16];
            for (int k = 0; k < 16; ++k) t0[k] = t1[k] = 0.0;
            for (const Term& tm : *tab_subs[i]) {
                pattern_table(xr, tm.pk & 31u, half, pair_scale * tm.d, pat);
                const double s1 = ((tm.pk & 32u) && any0) ? -1.0 : 1.0;  // mixed record: rho = 1 flips these terms
                for (int k = 0; k < 16; ++k) t0[k] += pat[k], t1[k] += s1 * pat[k];
            }
            push_dot(atom.units, h, t0, (any0 && any1) ? t1 : nullptr);
            ++atom.n_recs;
        }
        out.atoms.push_back(std::move(atom));
    }
    if (out.atoms.size() == 1 && out.atoms[0].n_recs == 1 && !out.atoms[0].needs_setup && tab_keys.size() == 1 &&
        tab_keys[0].second == 0)
        for (int sl = 0; sl < kQuadSlots; ++sl)
            if (kQuadKappa[sl] == xr) out.atoms[0].quad_slot = sl;
    return out;
}

// Sub-cube candidates inside a tile (tile-local positions).  `conflict_free`: the positions the thread id is mapped
// to (everything outside the coordinate bits) cover all three residues mod 3 -- only then can the lanes of a quarter
// warp be mapped to distinct shared-memory bank groups (the others still work, with 2-way conflicts on the gather, and
// are used as a last resort).
struct Cube {
    uint32_t span_mask;            // tile bits the cube's x-masks may touch (6 for a parity cube, 5 for a unit cube)
    uint32_t gens[kRBits];         // generators, tile-local masks
    int coord[kRBits];             // coordinate bit of each generator (its lowest bit), ascending
    int rest[kTileBits - kRBits];  // tile positions the 7 thread-id bits are deposited at
    bool conflict_free;
};
struct CubeCands {
    std::vector<Cube> parity, unit;
};
CubeCands make_cube_candidates() {
    CubeCands cc;
    for (uint32_t m = 0; m < (1u << kTileBits); ++m) {
        const int w = popc32(m);
        if (w != kRBits && w != kRBits + 1) continue;
        Cube c;
        c.span_mask = m;
        int pos[kRBits + 1], np = 0;
        for (int b = 0; b < kTileBits; ++b)
            if ((m >> b) & 1u) pos[np++] = b;
        for (int j = 0; j < kRBits; ++j) {
            c.coord[j] = pos[j];
            c.gens[j] = (1u << pos[j]) | (w == kRBits + 1 ? (1u << pos[kRBits]) : 0u);
        }
        int nr = 0;
        for (int b = 0; b < kTileBits; ++b)
            if (!((m >> b) & 1u)) c.rest[nr++] = b;
        if (w == kRBits + 1) c.rest[nr++] = pos[kRBits];  // the sixth bit: set by the thread id, flipped by odd elements
        bool res[3] = {false, false, false};
        for (int j = 0; j < nr; ++j)
            if (!(w == kRBits + 1 && c.rest[j] == pos[kRBits])) res[c.rest[j] % 3] = true;  // lanes never map to the sixth bit
        c.conflict_free = res[0] && res[1] && res[2];
        (w == kRBits ? cc.unit : cc.parity).push_back(c);
    }
    return cc;
}
const CubeCands& cube_candidates() {
    static const CubeCands cc = make_cube_candidates();  // thread-safe one-time initialisation
    return cc;
}

}  // namespace

HostPlan* build_host_plan(int n, size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs) {
    HostPlan* plan = new HostPlan();
    plan->n = n;
    PlanStats& st = plan->stats;
    const int tau0 = env_int("QC_K2_TAU", 9);            // first passes only build sub-cubes of >= tau x-masks
    const int min_gain = env_int("QC_K2_MIN_GAIN", 128);  // ... until a pass would serve fewer x-masks than this

    // ---- x-groups
    std::vector<XG> xg;
    {
        std::unordered_map<uint64_t, uint32_t> index;
        index.reserve(T / 2 + 16);
        for (size_t k = 0; k < T; ++k) {
            auto it = index.find(xs[k]);
            if (it == index.end()) {
                index.emplace(xs[k], (uint32_t)xg.size());
                xg.push_back({xs[k], {}, 1.0f});
                xg.back().terms.push_back((uint32_t)k);
            } else {
                xg[it->second].terms.push_back((uint32_t)k);
            }
        }
    }
    st.n_groups = xg.size();
    std::vector<uint32_t> uncovered;
    int diag_group = -1;
    for (uint32_t g = 0; g < xg.size(); ++g) {
        const uint64_t x = xg[g].x;
        xg[g].weight = xg[g].terms.size() <= 8 ? 1.0f : 4.0f;
        if (x == 0) {
            diag_group = (int)g;
        } else if (popc64(x) <= kRBits && popc64(x | 1ull) <= kTileBits) {
            uncovered.push_back(g);
        } else {
            for (uint32_t t : xg[g].terms) {
                plan->left_x.push_back(xs[t]);
                plan->left_z.push_back(zs[t]);
                plan->left_c.push_back(cs[t]);
            }
        }
    }
    st.n_leftover_terms = plan->left_x.size();
    st.n_tiled_groups = uncovered.size() + (diag_group >= 0 ? 1 : 0);

    const CubeCands& cube_cands = cube_candidates();
    const int free_slots = kTileBits - 1;
    bool diag_pending = diag_group >= 0;
    int tau = std::max(1, tau0);

    while (!uncovered.empty() || diag_pending) {
        // ---- tile bits S: greedy, one bit at a time, scored by how close it brings the uncovered x-masks to fitting
        uint64_t S = 1ull;
        for (int slot = 0; slot < free_slots; ++slot) {
            const int remaining_after = free_slots - slot - 1;
            double delta[64] = {0.0};
            for (uint32_t g : uncovered) {
                const uint64_t need = xg[g].x & ~S;
                const int miss = popc64(need);
                if (miss == 0 || miss - 1 > remaining_after) continue;
                const double f0 = miss <= remaining_after ? (double)(4096 >> (3 * std::min(miss, 4))) : 0.0;
                const double f1 = (double)(4096 >> (3 * std::min(miss - 1, 4)));
                const double dlt = (f1 - f0) * xg[g].weight;
                uint64_t m = need;
                while (m) {
                    delta[__builtin_ctzll(m)] += dlt;
                    m &= m - 1;
                }
            }
            int best_bit = -1;
            double best = 0.0;
            for (int b = 1; b < n; ++b) {
                if ((S >> b) & 1ull) continue;
                if (best_bit < 0 || delta[b] > best) best = delta[b], best_bit = b;
            }
            S |= 1ull << best_bit;
        }
        // local search: swap one tile bit for one outside bit while that covers more of the uncovered weight
        if (env_int("QC_K2_LOCAL_SEARCH", 1)) {
            auto covered = [&](uint64_t mask) {
                double w = 0.0;
                for (uint32_t g : uncovered)
                    if ((xg[g].x & ~mask) == 0) w += xg[g].weight;
                return w;
            };
            double cur = covered(S);
            for (int iter = 0; iter < 4; ++iter) {
                bool improved = false;
                for (int out = 1; out < n; ++out) {
                    if (!((S >> out) & 1ull)) continue;
                    // only x-masks that miss exactly one bit can be gained by one swap: count gains per incoming bit
                    const uint64_t S1 = S & ~(1ull << out);
                    double lost = 0.0, gain[64] = {0.0};
                    for (uint32_t g : uncovered) {
                        const uint64_t miss = xg[g].x & ~S1;
                        if (miss == 0) continue;  // covered with or without `out`
                        if (miss == (1ull << out)) lost += xg[g].weight;
                        else if ((miss & (miss - 1)) == 0) gain[__builtin_ctzll(miss)] += xg[g].weight;
                    }
                    int best_in = -1;
                    double best_gain = lost;
                    for (int in = 1; in < n; ++in)
                        if (!((S >> in) & 1ull) && gain[in] > best_gain) best_gain = gain[in], best_in = in;
                    if (best_in >= 0) {
                        S = S1 | (1ull << best_in);
                        cur += best_gain - lost;
                        improved = true;
                    }
                }
                if (!improved) break;
            }
            (void)cur;
        }
        int spos[kTileBits];
        {
            uint64_t m = S;
            for (int j = 0; j < kTileBits; ++j) {
                spos[j] = __builtin_ctzll(m);
                m &= m - 1;
            }
        }
        std::vector<uint32_t> mine, others;
        for (uint32_t g : uncovered) ((xg[g].x & ~S) == 0 ? mine : others).push_back(g);

        // ---- sub-cubes: greedy cover of this pass's x-masks -- parity cubes (six tile bits) for the even-weight
        // x-masks, unit cubes (five tile bits) for the odd-weight ones
        std::vector<uint32_t> xl(mine.size());
        for (size_t i = 0; i < mine.size(); ++i) xl[i] = compress_bits(xg[mine[i]].x, spos, kTileBits);
        std::vector<char> assigned(mine.size(), 0);
        std::vector<std::pair<const Cube*, std::vector<uint32_t>>> cubes;  // (cube, group ids)
        float total_w = 0.0f;
        for (int round = 0; round < 2; ++round) {
            const std::vector<Cube>& cands = round == 0 ? cube_cands.parity : cube_cands.unit;
            auto eligible = [&](size_t i) { return !assigned[i] && ((popc32(xl[i]) & 1) == round); };
            std::vector<float> cnt(cands.size(), 0.0f);
            for (size_t i = 0; i < mine.size(); ++i)
                if (eligible(i))
                    for (size_t c = 0; c < cands.size(); ++c)
                        if ((xl[i] & ~cands[c].span_mask) == 0) cnt[c] += xg[mine[i]].weight;
            for (;;) {
                size_t best_c = 0;
                float best = -1.0f;
                for (size_t c = 0; c < cands.size(); ++c) {
                    const float v = cands[c].conflict_free ? cnt[c] : 0.5f * cnt[c];
                    if (v > best) best = v, best_c = c;
                }
                if (best <= 0.0f) break;
                best = cnt[best_c];
                // lower the bar when this pass would otherwise be too thin
                const int bar = round == 0 ? tau : 1;
                if (best < (float)bar) {
                    if (total_w >= (float)min_gain || tau == 1) break;
                    --tau;
                    continue;
                }
                const Cube& cube = cands[best_c];
                std::vector<uint32_t> members;
                for (size_t i = 0; i < mine.size(); ++i) {
                    if (!eligible(i) || (xl[i] & ~cube.span_mask) != 0) continue;
                    members.push_back(mine[i]);
                    total_w += xg[mine[i]].weight;
                    for (size_t c = 0; c < cands.size(); ++c)
                        if ((xl[i] & ~cands[c].span_mask) == 0) cnt[c] -= xg[mine[i]].weight;
                    assigned[i] = 1;
                }
                cnt[best_c] = 0.0f;
                cubes.push_back({&cube, members});
            }
        }
        for (size_t i = 0; i < mine.size(); ++i)
            if (!assigned[i]) others.push_back(mine[i]);
        if (cubes.empty() && !diag_pending) {
            // cannot happen (tau reaches 1 and every tileable x-mask fits some candidate); guard against looping
            for (uint32_t g : others)
                for (uint32_t t : xg[g].terms) {
                    plan->left_x.push_back(xs[t]);
                    plan->left_z.push_back(zs[t]);
                    plan->left_c.push_back(cs[t]);
                }
            st.n_leftover_terms = plan->left_x.size();
            others.clear();
            uncovered.clear();
            break;
        }
        uncovered.swap(others);
        if (diag_pending) {
            if (cubes.empty()) cubes.push_back({&cube_cands.parity[0], {}});
            cubes[0].second.push_back((uint32_t)diag_group);
            diag_pending = false;
        }

        // ---- emit the pass (split into several passes over the same S when its blob would exceed the cap)
        struct RB {
            TRBlock hdr;
            int perm[kTileBits - kRBits];  // thread-id bit j -> tile position
            std::vector<Unit16> quads[kQuadSlots];  // simple-quad records by slot
            std::vector<Unit16> hops[kHopSlots];    // HOP records by slot
            std::vector<Unit16> recs;           // generic records
        };
        std::vector<RB> cur;       // sub-cube visits of the pass being assembled
        uint32_t cur_units = 0;    // headers + records
        uint32_t cur_groups = 0;
        auto flush_pass = [&]() {
            if (cur.empty()) return;
            TPass pass;
            memset(&pass, 0, sizeof(pass));
            pass.s_mask = S;
            for (int j = 0; j < kTileBits; ++j) pass.spos[j] = (uint8_t)spos[j];
            pass.blob_off = (uint32_t)plan->units.size();
            pass.n_rblocks = (uint32_t)cur.size();
            pass.n_groups = cur_groups;
            uint32_t rec_off = (uint32_t)cur.size() * kRBlockUnits;
            for (RB& rb : cur) {
                pass.n_generic += rb.hdr.n_recs + (uint32_t)popc32(rb.hdr.hop_mask);
                rb.hdr.rec_off = rec_off;
                rec_off += (uint32_t)rb.recs.size();
                for (int j = 0; j < kQuadSlots; ++j) rec_off += (uint32_t)rb.quads[j].size();
                for (int j = 0; j < kHopSlots; ++j) rec_off += (uint32_t)rb.hops[j].size();
                {   // every quad record announces whether the NEXT one carries two tables (the first: the visit header)
                    Unit16* prev = nullptr;
                    rb.hdr.pad = 0;
                    for (int j = 0; j < kQuadSlots; ++j) {
                        if (rb.quads[j].empty()) continue;
                        TRecHdr h;
                        memcpy(&h, &rb.quads[j][0], sizeof(h));
                        const bool two = (h.meta & kMetaTwoTables) != 0;
                        if (prev) {
                            TRecHdr ph;
                            memcpy(&ph, prev, sizeof(ph));
                            ph.meta = two ? (ph.meta | kMetaNextTwoTables) : (ph.meta & ~kMetaNextTwoTables);
                            memcpy(prev, &ph, sizeof(ph));
                        } else if (two) {
                            rb.hdr.pad = 1;
                        }
                        prev = &rb.quads[j][0];
                    }
                }
                push_units(plan->units, &rb.hdr, sizeof(TRBlock));
                uint32_t origin[24];  // lo[16] from thread bits 0..3, hi[8] from thread bits 4..6
                for (uint32_t t = 0; t < 24; ++t) {
                    const uint32_t bits = t < 16 ? t : (t - 16) << 4;
                    uint32_t org = 0;
                    for (int j = 0; j < kTileBits - kRBits; ++j)
                        if ((bits >> j) & 1u) org |= 1u << rb.perm[j];
                    origin[t] = (org << 16) | k2_swizzle(org);
                }
                push_units(plan->units, origin, sizeof(origin));
            }
            for (RB& rb : cur) {
                for (int j = 0; j < kQuadSlots; ++j)
                    plan->units.insert(plan->units.end(), rb.quads[j].begin(), rb.quads[j].end());
                for (int j = 0; j < kHopSlots; ++j)
                    plan->units.insert(plan->units.end(), rb.hops[j].begin(), rb.hops[j].end());
                plan->units.insert(plan->units.end(), rb.recs.begin(), rb.recs.end());
            }
            pass.blob_units = rec_off;
            plan->passes.push_back(pass);
            st.n_rblocks += cur.size();
            cur.clear();
            cur_units = 0, cur_groups = 0;
        };
        for (auto& cube : cubes) {
            const Cube& cb = *cube.first;
            int rglob[kRBits];
            uint64_t gglob[kRBits], r_global = 0;
            for (int j = 0; j < kRBits; ++j) {
                rglob[j] = spos[cb.coord[j]];
                r_global |= 1ull << rglob[j];
                gglob[j] = 0;
                for (int b = 0; b < kTileBits; ++b)
                    if ((cb.gens[j] >> b) & 1u) gglob[j] |= 1ull << spos[b];
            }
            const bool parity_cube = popc32(cb.span_mask) == kRBits + 1;
            const int rho_local = parity_cube ? 31 - __builtin_clz(cb.span_mask) : -1;  // the sixth (highest) cube bit
            const int rho_pos = parity_cube ? spos[rho_local] : -1;
            TRBlock blk;
            memset(&blk, 0, sizeof(blk));
            for (int j = 0; j < kRBits; ++j) blk.sbit[j] = (uint16_t)k2_swizzle(cb.gens[j]);
            // thread bit -> tile position; the three lowest thread bits get positions with distinct residues mod 3
            // (distinct swizzle columns => conflict-free 128-bit shared loads inside a quarter warp)
            int perm_arr[kTileBits - kRBits];
            {
                // a parity cube's sixth bit goes to thread-id bit 6 (warp-uniform: it selects the table of a two-table record)
                std::vector<int> rest_pos, perm;
                for (int j = 0; j < kTileBits - kRBits; ++j)
                    if (cb.rest[j] != rho_local) rest_pos.push_back(cb.rest[j]);
                std::sort(rest_pos.begin(), rest_pos.end());
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
                if (parity_cube) perm.push_back(rho_local);
                for (size_t j = 0; j < perm.size(); ++j) perm_arr[j] = perm[j];
            }
            // groups sorted by their in-cube mask: consecutive records then take the same branch of the kernel
            std::vector<uint32_t> members = cube.second;
            std::stable_sort(members.begin(), members.end(), [&](uint32_t a, uint32_t b) {
                return compress_bits(xg[a].x, rglob, kRBits) < compress_bits(xg[b].x, rglob, kRBits);
            });
            bool open = false;  // a visit of this sub-cube is open in `cur`
            uint32_t visit_groups = 0;
            auto open_visit = [&]() {
                cur.push_back(RB());
                cur.back().hdr = blk;
                memcpy(cur.back().perm, perm_arr, sizeof(perm_arr));
                cur_units += kRBlockUnits;
                open = true;
                visit_groups = 0;
            };
            auto close_visit = [&]() {
                if (open) st.rblock_hist[std::min<uint32_t>(visit_groups, 7)] += 1;
                open = false;
            };
            for (uint32_t g : members) {
                GroupRecs gr = emit_group(xg[g], zs, cs, S, spos, r_global, rglob, gglob, rho_pos, st);
                bool setup_here = false, counted = false;
                for (Atom& atom : gr.atoms) {
                    const bool want_setup = atom.needs_setup && !setup_here;
                    const uint32_t need = (uint32_t)atom.units.size() + (want_setup ? (uint32_t)gr.setup.size() : 0u) +
                                          (open ? 0u : kRBlockUnits);
                    if (cur_units + need > kBlobCapUnits) {
                        close_visit();
                        flush_pass();
                        setup_here = false;
                    }
                    if (!open) open_visit();
                    RB& rb = cur.back();
                    if (atom.needs_setup && !setup_here) {
                        rb.recs.insert(rb.recs.end(), gr.setup.begin(), gr.setup.end());
                        rb.hdr.n_recs += gr.setup_recs;
                        cur_units += (uint32_t)gr.setup.size();
                        st.n_dot += gr.setup_recs;
                        setup_here = true;
                    }
                    if (atom.quad_slot >= 0 && !((rb.hdr.quad_mask >> atom.quad_slot) & 1u)) {
                        rb.quads[atom.quad_slot] = atom.units;
                        rb.hdr.quad_mask |= (uint16_t)(1u << atom.quad_slot);
                    } else if (atom.hop_slot >= 0) {  // one x-mask per coordinate vector and visit: the slot is free
                        rb.hops[atom.hop_slot] = atom.units;
                        rb.hdr.hop_mask |= 1u << atom.hop_slot;
                    } else {
                        rb.recs.insert(rb.recs.end(), atom.units.begin(), atom.units.end());
                        rb.hdr.n_recs += atom.n_recs;
                    }
                    cur_units += (uint32_t)atom.units.size();
                    if (!atom.needs_setup) st.n_dot += atom.n_recs;
                    if (!counted) {
                        counted = true;
                        rb.hdr.n_groups += 1, ++visit_groups, ++cur_groups;
                    }
                }
            }
            close_visit();
        }
        flush_pass();
    }
    st.n_passes = plan->passes.size();
    return plan;
}

double check_host_plan(const HostPlan* plan, size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs,
                       size_t samples, uint64_t seed) {
    // inverse of the (XOR-linear, bijective) swizzle
    std::vector<uint16_t> unswz(kTileSize);
    for (uint32_t t = 0; t < (uint32_t)kTileSize; ++t) unswz[k2_swizzle(t)] = (uint16_t)t;
    struct CubeGeom {
        uint64_t gen[kRBits];   // generators as index-bit masks
        int piv[kRBits];        // pivot positions (index bits): the tile bits no representative ever sets
        uint32_t inv[kRBits];   // rows of the inverse of the generators' pivot-bit matrix: element bit j = parity(inv[j] & v)
        int spos[kTileBits];
        // coordinates (which generators) of the span component of an index, read off its pivot bits
        uint32_t elem_of(uint64_t idx) const {
            uint32_t v = 0, e = 0;
            for (int k = 0; k < kRBits; ++k) v |= (uint32_t)((idx >> piv[k]) & 1ull) << k;
            for (int j = 0; j < kRBits; ++j) e |= (uint32_t)(popc32(inv[j] & v) & 1) << j;
            return e;
        }
    };
    // hdr_unit: the visit's TRBlock (its origin tables follow): the representatives are XOR-combinations of seven tile
    // bits; the other five are the pivots, on which the five generators must form an invertible matrix
    auto geometry = [&](const TPass& pass, const TRBlock& blk, uint32_t hdr_unit, CubeGeom& cg) {
        for (int j = 0; j < kTileBits; ++j) cg.spos[j] = pass.spos[j];
        uint32_t origin[24];
        memcpy(origin, &plan->units[hdr_unit + 2], sizeof(origin));
        uint32_t free_local = 0;
        for (uint32_t t = 0; t < 24; ++t) free_local |= origin[t] >> 16;
        if (free_local >= (uint32_t)kTileSize || popc32(free_local) != kTileBits - kRBits) return false;
        int np = 0;
        for (int b = 0; b < kTileBits; ++b)
            if (!((free_local >> b) & 1u)) cg.piv[np++] = pass.spos[b];
        for (int j = 0; j < kRBits; ++j) {
            if (blk.sbit[j] >= kTileSize) return false;
            const uint32_t local = unswz[blk.sbit[j]];
            if (local == 0) return false;
            cg.gen[j] = 0;
            for (int b = 0; b < kTileBits; ++b)
                if ((local >> b) & 1u) cg.gen[j] |= 1ull << pass.spos[b];
        }
        // N[k] = bits (over generators j) of pivot k:  v = N E;  invert by Gauss-Jordan on [N | I]
        uint32_t rows[kRBits];
        for (int k = 0; k < kRBits; ++k) {
            uint32_t r = 0;
            for (int j = 0; j < kRBits; ++j) r |= (uint32_t)((cg.gen[j] >> cg.piv[k]) & 1ull) << j;
            rows[k] = r | (1u << (kRBits + k));
        }
        for (int col = 0; col < kRBits; ++col) {
            int pr = -1;
            for (int k = col; k < kRBits; ++k)
                if ((rows[k] >> col) & 1u) {
                    pr = k;
                    break;
                }
            if (pr < 0) return false;  // generators not independent of the representatives
            std::swap(rows[col], rows[pr]);
            for (int k = 0; k < kRBits; ++k)
                if (k != col && ((rows[k] >> col) & 1u)) rows[k] ^= rows[col];
        }
        for (int j = 0; j < kRBits; ++j) cg.inv[j] = rows[j] >> kRBits;
        return true;
    };
    // a parity cube's sixth bit: the one bit all five two-bit generators share (-1 for any other cube)
    auto parity_rho_pos = [](const uint64_t* gen) {
        uint64_t common = ~0ull;
        for (int j = 0; j < kRBits; ++j) {
            if (popc64(gen[j]) != 2) return -1;
            common &= gen[j];
        }
        return popc64(common) == 1 ? (int)__builtin_ctzll(common) : -1;
    };
    // index: x-mask -> record segments (pass, sub-cube visit, first unit, unit count), in execution order
    struct Where {
        uint32_t pass, rblock_unit, first, units;
    };
    std::unordered_map<uint64_t, std::vector<Where>> where;
    for (uint32_t p = 0; p < plan->passes.size(); ++p) {
        const TPass& pass = plan->passes[p];
        if (pass.blob_units > kBlobCapUnits) return -1.0;
        for (uint32_t rb = 0; rb < pass.n_rblocks; ++rb) {
            TRBlock blk;
            const uint32_t hdr_unit = pass.blob_off + rb * kRBlockUnits;
            memcpy(&blk, &plan->units[hdr_unit], sizeof(blk));
            CubeGeom cg;
            if (!geometry(pass, blk, hdr_unit, cg)) return -1.0;
            // the 128 thread representatives x 32 elements must tile the 4096 slots exactly once
            uint32_t origin[24];
            memcpy(origin, &plan->units[hdr_unit + 2], sizeof(origin));
            std::vector<char> seen(kTileSize, 0);
            uint32_t lgen[kRBits];
            for (int j = 0; j < kRBits; ++j) lgen[j] = unswz[blk.sbit[j]];
            for (uint32_t t = 0; t < (uint32_t)kTileThreads; ++t) {
                const uint32_t packed = origin[t & 15u] ^ origin[16u + (t >> 4)];
                const uint32_t org = packed >> 16;
                if (k2_swizzle(org) != (packed & 0xffffu)) return -1.0;
                for (uint32_t e = 0; e < 32; ++e) {
                    uint32_t slot = org;
                    for (int j = 0; j < kRBits; ++j)
                        if ((e >> j) & 1u) slot ^= lgen[j];
                    if (seen[slot]) return -1.0;
                    seen[slot] = 1;
                }
            }
            uint32_t u = pass.blob_off + blk.rec_off;
            const uint32_t n_quads = (uint32_t)popc32(blk.quad_mask);
            {   // the quad records must be plain DOT/ACC/Re records of the announced coordinate vectors, in slot order
                uint32_t uq = u;
                for (int sl = 0; sl < kQuadSlots; ++sl) {
                    if (!((blk.quad_mask >> sl) & 1u)) continue;
                    TRecHdr h;
                    memcpy(&h, &plan->units[uq], sizeof(h));
                    if (meta_kind(h.meta) != kRecDot || meta_mode(h.meta) != kModeAcc || meta_im(h.meta) != 0 ||
                        meta_xr(h.meta) != kQuadKappa[sl])
                        return -1.0;
                    uq += dot_units(h.meta);
                }
            }
            const uint32_t n_hops = (uint32_t)popc32(blk.hop_mask);
            if (blk.hop_mask >> kHopSlots) return -1.0;
            {   // the HOP records follow the quads, in slot order, with the announced coordinate vectors
                uint32_t uh = u;
                for (uint32_t r = 0; r < n_quads; ++r) {
                    TRecHdr h;
                    memcpy(&h, &plan->units[uh], sizeof(h));
                    uh += dot_units(h.meta);
                }
                for (int sl = 0; sl < kHopSlots; ++sl) {
                    if (!((blk.hop_mask >> sl) & 1u)) continue;
                    TRecHdr h;
                    memcpy(&h, &plan->units[uh], sizeof(h));
                    if (meta_kind(h.meta) != kRecDot || meta_mode(h.meta) != kModeHop || meta_im(h.meta) != 0 ||
                        meta_xr(h.meta) != kHopKappa[sl])
                        return -1.0;
                    uh += hop_units(h.meta);
                }
            }
            for (uint32_t r = 0; r < n_quads + n_hops + blk.n_recs; ++r) {
                TRecHdr h;
                memcpy(&h, &plan->units[u], sizeof(h));
                const bool is_hop = meta_kind(h.meta) == kRecDot && meta_mode(h.meta) == kModeHop;
                if (is_hop != (r >= n_quads && r < n_quads + n_hops)) return -1.0;
                const uint32_t units = is_hop ? hop_units(h.meta)
                                              : meta_kind(h.meta) == kRecLin ? 1 + kLinEntryUnits * meta_count(h.meta)
                                                                             : dot_units(h.meta);
                uint64_t x = 0;
                for (int j = 0; j < kRBits; ++j)
                    if ((meta_xr(h.meta) >> j) & 1u) x ^= cg.gen[j];
                auto& v = where[x];
                if (!v.empty() && v.back().rblock_unit == hdr_unit && v.back().first + v.back().units == u)
                    v.back().units += units;
                else
                    v.push_back({p, hdr_unit, u, units});
                u += units;
            }
            if (u > pass.blob_off + pass.blob_units) return -1.0;
        }
    }
    // term groups
    std::unordered_map<uint64_t, std::vector<uint32_t>> groups;
    for (size_t k = 0; k < T; ++k) groups[xs[k]].push_back((uint32_t)k);
    std::vector<uint64_t> keys;
    for (auto& kv : groups) keys.push_back(kv.first);
    std::sort(keys.begin(), keys.end());
    std::unordered_map<uint64_t, char> is_left;
    for (uint64_t x : plan->left_x) is_left[x] = 1;
    for (uint64_t x : keys)
        if (!is_left.count(x) && !where.count(x)) return -1.0;  // an x-mask nobody serves

    uint64_t rng = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    auto next = [&]() {
        rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17;
        return rng;
    };
    double worst = 0.0;
    for (size_t s = 0; s < samples; ++s) {
        const uint64_t x = keys[next() % keys.size()];
        if (is_left.count(x)) {
            if (where.count(x)) return -1.0;
            continue;
        }
        const uint64_t i_raw = next() & ((1ull << plan->n) - 1ull);
        double got[2] = {0.0, 0.0}, ref[2] = {0.0, 0.0};
        uint64_t i_ref = i_raw;
        bool first_seg = true;
        for (const Where& w : where[x]) {
            const TPass& pass = plan->passes[w.pass];
            TRBlock blk;
            memcpy(&blk, &plan->units[w.rblock_unit], sizeof(blk));
            CubeGeom cg;
            geometry(pass, blk, w.rblock_unit, cg);
            // coordinate vector of x (x is in the span, checked when `where` was built)
            const uint32_t xr = cg.elem_of(x);
            uint64_t i = i_raw;
            auto elem_of = [&](uint64_t idx) { return cg.elem_of(idx); };
            int k = -1;  // index of the element among the products of a record
            if (xr) {
                const int hb = 31 - __builtin_clz(xr);
                if ((elem_of(i) >> hb) & 1u) i ^= x;  // canonical element of the pair in THIS sub-cube
                const uint32_t p = elem_of(i);
                int kk = 0;
                for (int q = 0; q < 32; ++q) {
                    if ((q >> hb) & 1) continue;
                    if ((uint32_t)q == p) k = kk;
                    ++kk;
                }
            }
            // every segment must weight the SAME pair: W is invariant under i -> i^x up to the sign (-1)^{nY}, which
            // only matters for odd-nY (Im) terms; compare per segment against the direct weight at its own i
            if (first_seg) i_ref = i;
            const uint32_t p = elem_of(i);
            uint64_t c = i;  // the thread's representative: element 0 of the orbit of i
            for (int j = 0; j < kRBits; ++j)
                if ((p >> j) & 1u) c ^= cg.gen[j];
            const uint32_t org_local = compress_bits(c, cg.spos, kTileBits);
            const double flip_im = (i == i_ref) ? 1.0 : -1.0;  // Im q changes sign under i <-> i^x
            double A[kSlots] = {0, 0, 0, 0};
            int a_type[kSlots] = {0, 0, 0, 0};
            uint32_t u = w.first;
            while (u < w.first + w.units) {
                TRecHdr h;
                memcpy(&h, &plan->units[u], sizeof(h));
                const uint32_t kind = meta_kind(h.meta);
                if (kind == kRecLin) {
                    for (uint32_t e = 0; e < meta_count(h.meta); ++e) {
                        LinEntry le;
                        memcpy(&le, &plan->units[u + 1 + kLinEntryUnits * e], sizeof(le));
                        const int par = (popc64(c & le.z_outer) + popc32(org_local & le.z_rest)) & 1;
                        for (int j = 0; j < kSlots; ++j)
                            got[a_type[j]] += (par ? -1.0 : 1.0) * le.c[j] * A[j] * (a_type[j] ? flip_im : 1.0);
                    }
                    u += 1 + kLinEntryUnits * meta_count(h.meta);
                    continue;
                }
                if (kind == kRecDot && meta_mode(h.meta) == kModeHop) {
                    if (k < 0) return -1.0;
                    const int rho_pos = parity_rho_pos(cg.gen);
                    if (rho_pos < 0) return -1.0;
                    const uint32_t rho = (uint32_t)((c >> rho_pos) & 1ull);
                    double tabs[3][16];
                    memcpy(tabs, &plan->units[u + 1], sizeof(tabs));
                    const uint32_t two8 = (h.meta & kMetaTwoTables) ? 8u : 0u;
                    if (two8 && rho) memcpy(tabs[0], &plan->units[u + kHopHeadUnits], sizeof(tabs[0]));
                    const int par = (popc64(c & h.z_outer) + popc32(org_local & h.z_rest)) & 1;
                    got[0] += par ? -tabs[0][k] : tabs[0][k];
                    const double a0 = tabs[1][k] * (((h.meta & kMetaHopFlip0) && rho) ? -1.0 : 1.0);
                    const double a1 = tabs[2][k] * (((h.meta & kMetaHopFlip1) && rho) ? -1.0 : 1.0);
                    for (uint32_t e = 0; e < meta_count(h.meta); ++e) {
                        HopEntry he;
                        memcpy(&he, &plan->units[u + kHopHeadUnits + two8 + kHopEntryUnits * e], sizeof(he));
                        const int pe = (popc64(c & he.z_outer) + popc32(org_local & he.z_rest)) & 1;
                        got[0] += (pe ? -1.0 : 1.0) * (he.c[0] * a0 + he.c[1] * a1);
                    }
                    u += hop_units(h.meta);
                    continue;
                }
                DotRec r;
                memcpy(&r.h, &plan->units[u], sizeof(r.h));
                {
                    // thread-id bit 6 of a parity cube = the representative's sixth bit (highest bit of the generators)
                    const int rho_pos = parity_rho_pos(cg.gen);
                    const uint32_t rho = (rho_pos >= 0 && (h.meta & kMetaTwoTables)) ? (uint32_t)((c >> rho_pos) & 1ull) : 0u;
                    memcpy(r.tab, &plan->units[u + 1 + 8 * rho], sizeof(r.tab));
                }
                u += dot_units(h.meta);
                double wgt;
                if (kind == kRecDot) {
                    wgt = r.tab[k];
                } else {
                    const int half = kind == kRecDiagHi;
                    wgt = ((int)(p >> 4) == half) ? r.tab[p & 15] : 0.0;
                }
                const uint32_t mode = meta_mode(h.meta);
                if (mode == kModeAcc) {
                    const int par = (popc64(c & h.z_outer) + popc32(org_local & h.z_rest)) & 1;
                    got[meta_im(h.meta)] += (par ? -wgt : wgt) * (meta_im(h.meta) ? flip_im : 1.0);
                } else {
                    const uint32_t sl = meta_slot(h.meta);
                    A[sl] = (mode == kModeSet) ? wgt : A[sl] + wgt;
                    a_type[sl] = (int)meta_im(h.meta);
                }
            }
            first_seg = false;
        }
        for (uint32_t t : groups[x]) {
            const int ny = popc64(x & zs[t]);
            const double d = cs[t] * (((ny >> 1) & 1) ? -1.0 : 1.0) * (x ? 2.0 : 1.0);
            ref[ny & 1] += (popc64(i_ref & zs[t]) & 1) ? -d : d;
        }
        worst = std::max(worst, std::max(std::abs(got[0] - ref[0]), std::abs(got[1] - ref[1])));
    }
    return worst;
}

}  // namespace qc
