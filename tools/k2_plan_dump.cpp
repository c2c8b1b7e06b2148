// Per-pass shape of a tiled-expectation plan, for fitting / applying the cost model of DESIGN.md (4.1):
//   time(pass) = 2.69 ms + 0.446 ms x visits + 0.139 ms x quads + 0.93 ms x hops   (30 qubits, B200, final round-1 kernel)
// Input: a binary term list  int32 n | uint64 T | uint64 x[T] | uint64 z[T] | double c[T]  (see tools/k2_cost_model.py).
// Output: one line per pass: index visits quads hops generic_records x_masks blob_units low_bits_in_tile; totals on stderr.
// Build (from tools/): g++ -O2 -std=c++17 -o k2_plan_dump k2_plan_dump.cpp ../qibochem_b200/csrc/k2_plan.cpp
#include "../qibochem_b200/csrc/k2_plan.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace qc;
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    size_t T; int n;
    if (fread(&n, sizeof(n), 1, f) != 1 || fread(&T, sizeof(T), 1, f) != 1) return 1;
    std::vector<uint64_t> xs(T), zs(T); std::vector<double> cs(T);
    if (fread(xs.data(), 8, T, f) != T || fread(zs.data(), 8, T, f) != T || fread(cs.data(), 8, T, f) != T) return 1;
    fclose(f);
    HostPlan* p = build_host_plan(n, T, xs.data(), zs.data(), cs.data());
    size_t tot_v = 0, tot_q = 0;
    for (size_t i = 0; i < p->passes.size(); ++i) {
        const TPass& ps = p->passes[i];
        int quads = 0, hops = 0, lin = 0, gen = 0;
        for (uint32_t rb = 0; rb < ps.n_rblocks; ++rb) {
            TRBlock b;
            memcpy(&b, &p->units[ps.blob_off + rb * kRBlockUnits], sizeof(b));
            quads += __builtin_popcount(b.quad_mask);
            hops += __builtin_popcount(b.hop_mask);
            gen += b.n_recs;
        }
        int lowbits = __builtin_popcountll(ps.s_mask & 7ull);
        printf("%zu %u %d %d %d %u %u %d\n", i, ps.n_rblocks, quads, hops, gen, ps.n_groups, ps.blob_units, lowbits);
        tot_v += ps.n_rblocks; tot_q += quads;
    }
    fprintf(stderr, "passes %zu visits %zu quads %zu\n", p->passes.size(), tot_v, tot_q);
    return 0;
}
