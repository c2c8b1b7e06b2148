// Host-side grouping of Pauli terms into mutually (qubit-wise) commuting sets: greedy colouring of the
// non-commutation graph with networkx's default rule ("largest_first": nodes by decreasing degree, ties in input
// order; every node takes the smallest colour none of its neighbours has).
//
// Reference behaviour replaced: src/qibochem/measurement/util.py:22-79 parses the strings of every pair of terms in
// Python (O(T^2) string operations) and calls networkx.  Here a pair test is three bit operations on the (x, z)
// masks; degrees are counted in parallel (OpenMP), and the colouring scans colour classes with early exit.
#include <stdint.h>
#include <stddef.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

namespace qc {
void set_error(const std::string& msg);
}

namespace {
inline bool conflict(uint64_t x1, uint64_t z1, uint64_t x2, uint64_t z2, bool qubitwise) {
    if (qubitwise) {
        const uint64_t common = (x1 | z1) & (x2 | z2);
        return (((x1 ^ x2) | (z1 ^ z2)) & common) != 0;
    }
    return __builtin_popcountll((x1 & z2) ^ (z1 & x2)) & 1;
}
}  // namespace

// colour_out[k] <- colour (group index) of term k; n_colours_out <- number of groups.  Masks in any consistent bit
// convention (only equality of positions matters).
extern "C" int qc_group_commuting(size_t T, const uint64_t* xmask, const uint64_t* zmask, int qubitwise,
                                  int32_t* colour_out, int32_t* n_colours_out) {
    if ((T && (!xmask || !zmask || !colour_out)) || !n_colours_out) {
        qc::set_error("qc_group_commuting: null argument");
        return 2;
    }
    *n_colours_out = 0;
    if (T == 0) return 0;
    const bool qw = qubitwise != 0;
    std::vector<int64_t> degree(T, 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)T; ++i) {
        int64_t d = 0;
        const uint64_t xi = xmask[i], zi = zmask[i];
        for (size_t j = 0; j < T; ++j) d += conflict(xi, zi, xmask[j], zmask[j], qw);
        degree[i] = d;
    }
    std::vector<uint32_t> order(T);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return degree[a] > degree[b]; });
    std::vector<std::vector<uint32_t>> classes;
    for (uint32_t u : order) {
        const uint64_t xu = xmask[u], zu = zmask[u];
        size_t c = 0;
        for (; c < classes.size(); ++c) {
            bool clash = false;
            for (uint32_t v : classes[c])
                if (conflict(xu, zu, xmask[v], zmask[v], qw)) {
                    clash = true;
                    break;
                }
            if (!clash) break;
        }
        if (c == classes.size()) classes.emplace_back();
        classes[c].push_back(u);
        colour_out[u] = (int32_t)c;
    }
    *n_colours_out = (int32_t)classes.size();
    return 0;
}
