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

#include <algorithm>

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

    std::vector<Run> runs;
    plan_runs(R, xmask, zmask, runs);

    // ---- tables: (cos, sin)(Phi/2) over the support patterns of every run
    size_t total_entries = 0;
    for (const Run& g : runs) total_entries += (size_t)1 << __builtin_popcountll(g.sup);
    // The arena may still be in use by a previous call's in-flight kernels only if the stream has not drained;
    // the copy below is stream-ordered, and the host buffer is rewritten only after the previous copy finished.
    QC_CHECK(cudaStreamSynchronize(h->stream));
    if (h->arena.ensure(total_entries * sizeof(double2))) return 1;
    double2* htab = (double2*)h->arena.host;
    std::vector<RotDesc> descs(runs.size());
    size_t off = 0;
    for (size_t gi = 0; gi < runs.size(); ++gi) {
        const Run& g = runs[gi];
        const int nsup = __builtin_popcountll(g.sup);
        int sup_bits[kMaxSupport];
        {
            uint64_t m = g.sup;
            for (int k = 0; k < nsup; ++k) {
                sup_bits[k] = __builtin_ctzll(m);
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
        RotDesc& d = descs[gi];
        d.x = g.x;
        d.z0 = g.z0;
        d.sup_mask = g.sup;
        d.hb = g.hb;
        pow_minus_i(ny0 + 1, d.war, d.wai);
        pow_minus_i(1 - ny0, d.wbr, d.wbi);
        d.table_off = (uint32_t)off;
        off += (size_t)1 << nsup;
    }
    QC_CHECK(cudaMemcpyAsync(h->arena.dev, h->arena.host, total_entries * sizeof(double2), cudaMemcpyHostToDevice,
                             h->stream));
    const double2* dtab = (const double2*)h->arena.dev;

    for (const RotDesc& d : descs) {
        LaunchScope ls(h, kK1Rotation, 32.0 * (double)h->dim);  // every amplitude read + written once per pass
        if (d.x == 0) {
            const int nb = grid_for(h->dim, 4, h->sm_count);
            rot_diag_kernel<4><<<nb, kThreads, 0, h->stream>>>(h->psi, h->dim, d, dtab);
        } else if (d.hb < 5) {
            const int nb = grid_for(h->dim, 4, h->sm_count);
            rot_lowx_kernel<<<nb, kThreads, 0, h->stream>>>(h->psi, h->dim, d, dtab);
        } else {
            const uint64_t npairs = h->dim >> 1;
            const int nb = grid_for(npairs, 4, h->sm_count);
            rot_pairs_kernel<4><<<nb, kThreads, 0, h->stream>>>(h->psi, npairs, d, dtab);
        }
    }
    QC_CHECK(cudaGetLastError());
    if (n_passes_out) *n_passes_out = (int)descs.size();
    return 0;
}
