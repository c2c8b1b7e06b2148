// K2: exact expectation  E = sum_k c_k Re<psi|P_k|psi>, terms grouped by x-mask: one read-only sweep of the shard
// per distinct x-mask evaluates every z-mask of that group.
//
// Reference behaviour replaced: qibo's SymbolicHamiltonian.expectation(circuit) (called at e.g.
// tests/test_ansatz.py:131, examples/ucc_example1.py:45 of the reference) builds H|psi> term by term -- a state
// copy plus one full-state pass per non-identity factor of every term -- and then takes <psi|H psi>.
//
// Math.  <psi|P|psi> = sum_i conj(psi_i) g s(i) psi_{i^x},  g = (-i)^{nY}, s(i) = (-1)^{popc(i&z)}.
// With q_i = conj(psi_i) psi_{i^x} and the pair symmetry q_{i^x} = conj(q_i), s(i^x) = s(i) (-1)^{nY}:
//   nY even:  pair contribution = 2 (-1)^{nY/2}      s(i) Re q_i
//   nY odd :  pair contribution = 2 (-1)^{(nY-1)/2}  s(i) Im q_i
// so each term is a real weight d_k = c_k (-1)^{floor(nY/2)} applied to Re q or Im q.
#include <string.h>

#include <algorithm>
#include <numeric>

#include "common.cuh"

namespace qc {

constexpr int kMaxGroupTerms = 1024;  // terms of one sweep staged in shared memory (16 KB)
constexpr int kMaxGroupsPerLaunch = 2048;

struct XGroup {
    uint64_t x;
    int hb;           // highest bit of x, -1 for the diagonal group
    uint32_t first;   // first term of the group in the device term arrays
    uint32_t n_even;  // terms using Re q (come first)
    uint32_t n_odd;   // terms using Im q
};

struct TermDev {
    uint64_t z;
    double d;
};

__device__ __forceinline__ double signed_weight(double d, uint64_t i, uint64_t z) {
    const uint64_t v = i & z;
    const unsigned par = __popc((unsigned)v ^ (unsigned)(v >> 32)) & 1u;
    return __longlong_as_double(__double_as_longlong(d) ^ ((long long)par << 63));
}

template <int U>
__global__ void __launch_bounds__(kThreads) expectation_kernel(const double2* __restrict__ psi, uint64_t dim,
                                                               const XGroup* __restrict__ groups,
                                                               const TermDev* __restrict__ terms,
                                                               double* __restrict__ partials) {
    __shared__ TermDev sterm[kMaxGroupTerms];
    const XGroup g = groups[blockIdx.y];
    const uint32_t m = g.n_even + g.n_odd;
    for (uint32_t t = threadIdx.x; t < m; t += kThreads) sterm[t] = terms[g.first + t];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    double acc = 0.0;

    if (g.hb < 0) {
        // diagonal group: sum_i |psi_i|^2 sum_k d_k s_k(i)
        for (uint64_t i0 = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i0 < dim; i0 += stride * U) {
            double pr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t i = i0 + (uint64_t)u * stride;
                pr[u] = 0.0;
                if (i < dim) {
                    const double2 a = psi[i];
                    pr[u] = a.x * a.x + a.y * a.y;
                }
            }
            double w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) w[u] = 0.0;
            for (uint32_t t = 0; t < m; ++t) {
                const TermDev td = sterm[t];
#pragma unroll
                for (int u = 0; u < U; ++u) w[u] += signed_weight(td.d, i0 + (uint64_t)u * stride, td.z);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += pr[u] * w[u];
        }
    } else if (g.hb >= 5) {
        // pairs (i, i^x), i has bit hb clear; the factor 2 of the pair symmetry is applied at the end
        const uint64_t npairs = dim >> 1;
        for (uint64_t p0 = (uint64_t)blockIdx.x * kThreads + threadIdx.x; p0 < npairs; p0 += stride * U) {
            uint64_t idx[U];
            double qre[U], qim[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t p = p0 + (uint64_t)u * stride;
                idx[u] = insert_zero_bit(p, g.hb);
                qre[u] = qim[u] = 0.0;
                if (p < npairs) {
                    const double2 a = psi[idx[u]];
                    const double2 b = psi[idx[u] ^ g.x];
                    qre[u] = a.x * b.x + a.y * b.y;
                    qim[u] = a.x * b.y - a.y * b.x;
                }
            }
            double wre[U], wim[U];
#pragma unroll
            for (int u = 0; u < U; ++u) wre[u] = wim[u] = 0.0;
            for (uint32_t t = 0; t < g.n_even; ++t) {
                const TermDev td = sterm[t];
#pragma unroll
                for (int u = 0; u < U; ++u) wre[u] += signed_weight(td.d, idx[u], td.z);
            }
            for (uint32_t t = g.n_even; t < m; ++t) {
                const TermDev td = sterm[t];
#pragma unroll
                for (int u = 0; u < U; ++u) wim[u] += signed_weight(td.d, idx[u], td.z);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += 2.0 * (qre[u] * wre[u] + qim[u] * wim[u]);
        }
    } else {
        // x inside the low 5 bits: one thread per amplitude, partner by shuffle, no pair factor
        const int xl = (int)g.x;
        const int lane = threadIdx.x & 31;
        for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i - lane < dim; i += stride) {
            const bool valid = i < dim;
            const double2 a = valid ? psi[i] : make_double2(0.0, 0.0);
            double2 b;
            b.x = __shfl_xor_sync(0xffffffffu, a.x, xl);
            b.y = __shfl_xor_sync(0xffffffffu, a.y, xl);
            const double qre = a.x * b.x + a.y * b.y;
            const double qim = a.x * b.y - a.y * b.x;
            double wre = 0.0, wim = 0.0;
            for (uint32_t t = 0; t < g.n_even; ++t) wre += signed_weight(sterm[t].d, i, sterm[t].z);
            for (uint32_t t = g.n_even; t < m; ++t) wim += signed_weight(sterm[t].d, i, sterm[t].z);
            if (valid) acc += qre * wre + qim * wim;
        }
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// Grouped device program of a term list, kept on the handle between calls: an optimiser loop evaluates the SAME
// Hamiltonian thousands of times, and for a small register the per-call sort + upload cost more than the kernels.
struct SweepPlan {
    std::vector<uint64_t> key_x, key_z;
    std::vector<double> key_c;
    bool per_term = false;
    XGroup* d_groups = nullptr;
    TermDev* d_terms = nullptr;
    size_t ng = 0;
    ~SweepPlan() {
        if (d_groups) cudaFree(d_groups);
        if (d_terms) cudaFree(d_terms);
    }
    bool matches(size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs, bool pt) const {
        return per_term == pt && key_x.size() == T && memcmp(key_x.data(), xs, T * 8) == 0 &&
               memcmp(key_z.data(), zs, T * 8) == 0 && (pt || memcmp(key_c.data(), cs, T * 8) == 0);
    }
};
void free_sweep_plan(SweepPlan* p) { delete p; }

// Sweep path: one launch row per x-group.  Total -> h->d_out[out_slot]; per-group values -> h->d_out[out_slot+1 ...].
static int run_sweeps(qc_state* h, size_t T, const uint64_t* xmask, const uint64_t* zmask, const double* coeff,
                      bool per_term, size_t out_slot, size_t* ng_out) {
    SweepPlan* sp = h->sweep_plan;
    if (!(sp && sp->matches(T, xmask, zmask, coeff, per_term))) {
        // ---- group the terms by x-mask (stable), even-nY terms first inside a group
        std::vector<uint32_t> order(T);
        std::iota(order.begin(), order.end(), 0u);
        if (!per_term) {
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                if (xmask[a] != xmask[b]) return xmask[a] < xmask[b];
                const int oa = __builtin_popcountll(xmask[a] & zmask[a]) & 1, ob = __builtin_popcountll(xmask[b] & zmask[b]) & 1;
                return oa < ob;
            });
        }
        std::vector<XGroup> groups;
        std::vector<TermDev> tdev(T);
        for (size_t k = 0; k < T; ++k) {
            const uint32_t t = order[k];
            const int ny = __builtin_popcountll(xmask[t] & zmask[t]);
            const double sg = ((ny >> 1) & 1) ? -1.0 : 1.0;
            tdev[k].z = zmask[t];
            tdev[k].d = (per_term ? 1.0 : coeff[t]) * sg;
            const bool new_group = per_term || groups.empty() || groups.back().x != xmask[t] ||
                                   groups.back().n_even + groups.back().n_odd >= (uint32_t)kMaxGroupTerms;
            if (new_group) {
                XGroup g;
                g.x = xmask[t];
                g.hb = g.x ? 63 - __builtin_clzll(g.x) : -1;
                g.first = (uint32_t)k;
                g.n_even = g.n_odd = 0;
                groups.push_back(g);
            }
            if (ny & 1) {
                ++groups.back().n_odd;
            } else {
                // a group split at kMaxGroupTerms keeps "even first" because the sort put all evens before odds
                QC_REQUIRE(groups.back().n_odd == 0, "internal: term order inside an x-group");
                ++groups.back().n_even;
            }
        }
        // ---- ship the program (kept on the device until a different term list arrives)
        QC_CHECK(cudaStreamSynchronize(h->stream));  // the old program may still be read by in-flight sweeps
        delete h->sweep_plan;
        h->sweep_plan = nullptr;
        sp = new SweepPlan();
        sp->per_term = per_term;
        sp->ng = groups.size();
        if (cudaMalloc(&sp->d_groups, groups.size() * sizeof(XGroup)) != cudaSuccess ||
            cudaMalloc(&sp->d_terms, T * sizeof(TermDev)) != cudaSuccess ||
            cudaMemcpy(sp->d_groups, groups.data(), groups.size() * sizeof(XGroup), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(sp->d_terms, tdev.data(), T * sizeof(TermDev), cudaMemcpyHostToDevice) != cudaSuccess) {
            delete sp;
            set_error("qc_expectation: could not upload the sweep program");
            return 1;
        }
        h->h2d_bytes += groups.size() * sizeof(XGroup) + T * sizeof(TermDev);
        sp->key_x.assign(xmask, xmask + T);
        sp->key_z.assign(zmask, zmask + T);
        if (!per_term) sp->key_c.assign(coeff, coeff + T);
        h->sweep_plan = sp;
    }
    const size_t ng = sp->ng;
    const XGroup* dgroups = sp->d_groups;
    const TermDev* dterms = sp->d_terms;

    const int nbx = grid_for(h->dim >> 1, 4, h->sm_count);
    const size_t chunk = std::min(ng, (size_t)kMaxGroupsPerLaunch);
    if (h->ensure_partials(chunk * nbx) || h->ensure_out(out_slot + ng + 1)) return 1;
    for (size_t g0 = 0; g0 < ng; g0 += chunk) {
        const size_t gn = std::min(chunk, ng - g0);
        dim3 grid(nbx, (unsigned)gn);
        {
            LaunchScope ls(h, kK2Expectation, 16.0 * (double)h->dim * (double)gn);  // one read sweep per x-group
            expectation_kernel<4><<<grid, kThreads, 0, h->stream>>>(h->psi, h->dim, dgroups + g0, dterms, h->d_partials);
        }
        QC_CHECK(cudaGetLastError());
        if (fold_partials(h, h->d_partials, (int)gn, nbx, nbx, h->d_out + out_slot + 1 + g0)) return 1;
    }
    if (!per_term)
        if (fold_partials(h, h->d_out + out_slot + 1, 1, (int)ng, (int)ng, h->d_out + out_slot)) return 1;
    *ng_out = ng;
    return 0;
}

int run_tiled_expectation(qc_state* h, size_t T, const uint64_t* xs, const uint64_t* zs, const double* cs, int out_slot,
                          int* n_passes_out, const TiledPlan** plan_out);
void tiled_plan_leftovers(const TiledPlan* plan, const uint64_t** x, const uint64_t** z, const double** c, size_t* count);

}  // namespace qc

extern "C" int qc_expectation(qc_state* h, size_t T, const uint64_t* xmask, const uint64_t* zmask, const double* coeff,
                              double* energy_out, double* per_term, int* n_sweeps_out) {
    using namespace qc;
    QC_REQUIRE(h && energy_out, "qc_expectation: null argument");
    *energy_out = 0.0;
    if (n_sweeps_out) *n_sweeps_out = 0;
    if (T == 0) return 0;
    QC_REQUIRE(xmask && zmask && coeff, "qc_expectation: null array");
    QC_REQUIRE(T < (1ull << 31), "qc_expectation: too many terms");
    for (size_t k = 0; k < T; ++k)
        QC_REQUIRE(xmask[k] < h->dim && zmask[k] < h->dim, "qc_expectation: mask has bits outside the shard");
    QC_CHECK(cudaSetDevice(h->device));

    const bool tiled = !per_term && h->n >= 12 && (h->k2_mode == 2 || (h->k2_mode == 0 && h->n >= 16 && T >= 512));
    if (tiled) {
        int passes = 0;
        const TiledPlan* plan = nullptr;
        if (run_tiled_expectation(h, T, xmask, zmask, coeff, 0, &passes, &plan)) return 1;
        const uint64_t *lx, *lz;
        const double* lc;
        size_t nleft = 0, ng = 0;
        tiled_plan_leftovers(plan, &lx, &lz, &lc, &nleft);
        if (nleft && run_sweeps(h, nleft, lx, lz, lc, false, 1, &ng)) return 1;
        QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        h->d2h_bytes += 2 * sizeof(double);
        QC_CHECK(cudaStreamSynchronize(h->stream));
        *energy_out = h->h_out[0] + (nleft ? h->h_out[1] : 0.0);
        if (n_sweeps_out) *n_sweeps_out = passes + (int)ng;
        return 0;
    }

    size_t ng = 0;
    if (run_sweeps(h, T, xmask, zmask, coeff, per_term != nullptr, 0, &ng)) return 1;
    if (per_term) {
        QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_out, (ng + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        h->d2h_bytes += (ng + 1) * sizeof(double);
        QC_CHECK(cudaStreamSynchronize(h->stream));
        double e = 0.0;
        for (size_t k = 0; k < T; ++k) {
            per_term[k] = h->h_out[1 + k];
            e += coeff[k] * per_term[k];
        }
        *energy_out = e;
    } else {
        QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        h->d2h_bytes += sizeof(double);
        QC_CHECK(cudaStreamSynchronize(h->stream));
        *energy_out = h->h_out[0];
    }
    if (n_sweeps_out) *n_sweeps_out = (int)ng;
    return 0;
}
