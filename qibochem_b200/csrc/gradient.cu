// Adjoint-mode gradient of  E(phi) = <psi0| U(phi)^dagger H U(phi) |psi0>,  U = prod_k exp(-i phi_k P_k / 2),
// for ALL rotation angles at the cost of a few passes per fused run.
//
// Reference behaviour replaced: the reference differentiates with
// qibo.derivative.parameter_shift (doc/source/tutorials/adaptive.rst:46-61, :118-141), i.e. two full energy
// evaluations PER parameter.  Here, with psi = U psi0 resident in the handle:
//     lambda = H psi                          (sweeps over the x-groups of H into the scratch array)
//     E      = Re <psi|lambda>
//     for every fused run, last to first:      dE/dphi_j = Im <lambda| P_j |psi>   for the strings j of the run
//                                              psi <- U_run^dagger psi,  lambda <- U_run^dagger lambda
// The strings of a fused run commute, so the same (psi, lambda) pair serves all of them.  On return psi = psi0.
#include <string.h>

#include <algorithm>
#include <numeric>

#include "common.cuh"

namespace qc {

void plan_rotation_runs(size_t R, const uint64_t* xs, const uint64_t* zs, std::vector<std::pair<size_t, size_t>>& out);
int apply_rotations_to(qc_state* h, double2* target, size_t R, const uint64_t* xmask, const uint64_t* zmask,
                       const double* angle, int k1_mode, int* n_passes_out, long long basis_index = -1);

constexpr int kHGroupsPerLaunch = 16;
constexpr int kHTermsPerLaunch = 1024;
constexpr int kGradStrings = 8;  // strings of one run evaluated per sweep

struct HGroup {
    uint64_t x;
    uint32_t first, n_re, n_im, pad;
};
struct HTerm {
    uint64_t z;
    double d;
};

__device__ __forceinline__ double signed_d(double d, uint64_t i, uint64_t z) {
    const uint64_t v = i & z;
    const unsigned par = __popc((unsigned)v ^ (unsigned)(v >> 32)) & 1u;
    return __longlong_as_double(__double_as_longlong(d) ^ ((long long)par << 63));
}

// lambda_i (+)= sum_g (Wre_g(i) + i Wim_g(i)) psi_{i ^ x_g}
__global__ void __launch_bounds__(kThreads) hpsi_kernel(const double2* __restrict__ psi, double2* __restrict__ lam,
                                                        uint64_t dim, const HGroup* __restrict__ groups, int n_groups,
                                                        const HTerm* __restrict__ terms, int n_terms, int accumulate) {
    __shared__ HTerm sterm[kHTermsPerLaunch];
    __shared__ HGroup sgroup[kHGroupsPerLaunch];
    for (int t = threadIdx.x; t < n_terms; t += kThreads) sterm[t] = terms[t];
    if (threadIdx.x < n_groups) sgroup[threadIdx.x] = groups[threadIdx.x];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < dim; i += stride) {
        double ax = 0.0, ay = 0.0;
        for (int g = 0; g < n_groups; ++g) {
            const HGroup gr = sgroup[g];
            const double2 v = psi[i ^ gr.x];
            double wr = 0.0, wi = 0.0;
            for (uint32_t t = gr.first; t < gr.first + gr.n_re; ++t) wr += signed_d(sterm[t].d, i, sterm[t].z);
            for (uint32_t t = gr.first + gr.n_re; t < gr.first + gr.n_re + gr.n_im; ++t) wi += signed_d(sterm[t].d, i, sterm[t].z);
            ax += wr * v.x - wi * v.y;
            ay += wr * v.y + wi * v.x;
        }
        if (accumulate) {
            const double2 o = lam[i];
            ax += o.x, ay += o.y;
        }
        lam[i] = make_double2(ax, ay);
    }
}

// partials[0][b] + i partials[1][b] = sum_i conj(lam_i) psi_i over CTA b's amplitudes
__global__ void __launch_bounds__(kThreads) inner_kernel(const double2* __restrict__ lam, const double2* __restrict__ psi,
                                                         uint64_t dim, double* __restrict__ partials) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    double re = 0.0, im = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < dim; i += stride) {
        const double2 l = lam[i], p = psi[i];
        re += l.x * p.x + l.y * p.y;
        im += l.x * p.y - l.y * p.x;
    }
    const double tr = block_sum(re);
    const double ti = block_sum(im);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = tr;
        partials[gridDim.x + blockIdx.x] = ti;
    }
}

struct GradRun {
    uint64_t x;
    uint64_t z[kGradStrings];
    int count;
};

// C_j = sum_i (-1)^{popc(i & z_j)} conj(lam_i) psi_{i ^ x}   ->  partials[(2j + c) * gridDim.x + b]
__global__ void __launch_bounds__(kThreads) grad_inner_kernel(const double2* __restrict__ lam, const double2* __restrict__ psi,
                                                              uint64_t dim, const GradRun run, double* __restrict__ partials) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    double cr[kGradStrings], ci[kGradStrings];
#pragma unroll
    for (int j = 0; j < kGradStrings; ++j) cr[j] = ci[j] = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < dim; i += stride) {
        const double2 l = lam[i], p = psi[i ^ run.x];
        const double tx = l.x * p.x + l.y * p.y, ty = l.x * p.y - l.y * p.x;
#pragma unroll
        for (int j = 0; j < kGradStrings; ++j) {
            if (j < run.count) {
                cr[j] += signed_d(tx, i, run.z[j]);
                ci[j] += signed_d(ty, i, run.z[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kGradStrings; ++j) {
        const double tr = block_sum(cr[j]);
        const double ti = block_sum(ci[j]);
        if (threadIdx.x == 0) {
            partials[(size_t)(2 * j) * gridDim.x + blockIdx.x] = tr;
            partials[(size_t)(2 * j + 1) * gridDim.x + blockIdx.x] = ti;
        }
    }
}

}  // namespace qc

namespace qc {

// lambda (+)= sum_k c_k P_k psi over the given terms (all x-masks local to the shard); lambda = h->scratch_psi.
int hpsi_sweeps(qc_state* h, size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc, bool accumulate) {
    double2* lam = h->scratch_psi;
    // x-groups, Re-type (even nY) terms first
    std::vector<uint32_t> order(T);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (hx[a] != hx[b]) return hx[a] < hx[b];
        return (__builtin_popcountll(hx[a] & hz[a]) & 1) < (__builtin_popcountll(hx[b] & hz[b]) & 1);
    });
    std::vector<HGroup> groups;
    std::vector<HTerm> terms(T);
    {
        // launches: <= kHGroupsPerLaunch groups and <= kHTermsPerLaunch terms; a larger group is split (additive)
        size_t k = 0;
        while (k < T) {
            HGroup g;
            g.x = hx[order[k]];
            g.first = (uint32_t)k;
            g.n_re = g.n_im = 0, g.pad = 0;
            while (k < T && hx[order[k]] == g.x && g.n_re + g.n_im < (uint32_t)kHTermsPerLaunch) {
                const uint32_t t = order[k];
                const int ny = __builtin_popcountll(hx[t] & hz[t]);
                const bool odd = ny & 1;
                if (!odd && g.n_im) break;  // keep Re-type terms first inside a (sub)group
                terms[k].z = hz[t];
                terms[k].d = odd ? -hc[t] * ((((ny - 1) >> 1) & 1) ? -1.0 : 1.0) : hc[t] * (((ny >> 1) & 1) ? -1.0 : 1.0);
                odd ? ++g.n_im : ++g.n_re;
                ++k;
            }
            groups.push_back(g);
        }
    }
    HGroup* d_groups = nullptr;
    HTerm* d_terms = nullptr;
    QC_CHECK(cudaMalloc(&d_groups, groups.size() * sizeof(HGroup)));
    if (cudaMalloc(&d_terms, terms.size() * sizeof(HTerm)) != cudaSuccess) {
        cudaFree(d_groups);
        set_error("hpsi: cudaMalloc failed");
        return 1;
    }
    auto cleanup = [&]() {
        cudaStreamSynchronize(h->stream);
        cudaFree(d_groups);
        cudaFree(d_terms);
    };
    const int nb = grid_for(h->dim, 2, h->sm_count);
    size_t g0 = 0;
    bool first_launch = !accumulate;
    std::vector<HGroup> launch_groups;
    while (g0 < groups.size()) {
        size_t g1 = g0;
        uint32_t nt = 0;
        launch_groups.clear();
        while (g1 < groups.size() && g1 - g0 < (size_t)kHGroupsPerLaunch && nt + groups[g1].n_re + groups[g1].n_im <= (uint32_t)kHTermsPerLaunch) {
            HGroup g = groups[g1];
            g.first -= groups[g0].first;  // relative to the launch's first term
            launch_groups.push_back(g);
            nt += groups[g1].n_re + groups[g1].n_im;
            ++g1;
        }
        if (cudaMemcpyAsync(d_groups + g0, launch_groups.data(), launch_groups.size() * sizeof(HGroup), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
            cudaMemcpyAsync(d_terms + groups[g0].first, terms.data() + groups[g0].first, nt * sizeof(HTerm), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
            cleanup();
            set_error("hpsi: upload failed");
            return 1;
        }
        h->h2d_bytes += launch_groups.size() * sizeof(HGroup) + nt * sizeof(HTerm);
        // pageable host memory: the async copies above are staged synchronously, so the vectors may be reused
        {
            LaunchScope ls(h, kGradient, 16.0 * (double)h->dim * (double)(launch_groups.size() + 2));
            hpsi_kernel<<<nb, kThreads, 0, h->stream>>>(h->psi, lam, h->dim, d_groups + g0, (int)launch_groups.size(),
                                                        d_terms + groups[g0].first, (int)nt, first_launch ? 0 : 1);
        }
        first_launch = false;
        g0 = g1;
    }
    const cudaError_t e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) {
        set_error(std::string("hpsi launch failed: ") + cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

int hpsi_tiled(qc_state* h, size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc, bool accumulate,
               bool* handled);  // gradient_tiled.cu

static int ensure_lambda(qc_state* h) {
    if (!h->scratch_psi) QC_CHECK(cudaMalloc(&h->scratch_psi, h->dim * sizeof(double2)));
    return 0;
}

}  // namespace qc

extern "C" int qc_hpsi(qc_state* h, size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc, int accumulate) {
    using namespace qc;
    QC_REQUIRE(h, "qc_hpsi: null handle");
    QC_REQUIRE(T == 0 || (hx && hz && hc), "qc_hpsi: null term array");
    QC_REQUIRE(T < (1ull << 31), "qc_hpsi: too many terms");
    for (size_t k = 0; k < T; ++k)
        QC_REQUIRE(hx[k] < h->dim && hz[k] < h->dim, "qc_hpsi: term mask has bits outside the shard");
    QC_CHECK(cudaSetDevice(h->device));
    if (ensure_lambda(h)) return 1;
    if (T == 0) {
        if (!accumulate) QC_CHECK(cudaMemsetAsync(h->scratch_psi, 0, h->dim * sizeof(double2), h->stream));
        return 0;
    }
    bool handled = false;
    if (hpsi_tiled(h, T, hx, hz, hc, accumulate != 0, &handled)) return 1;
    if (handled) return 0;
    return hpsi_sweeps(h, T, hx, hz, hc, accumulate != 0);
}

extern "C" int qc_lambda_inner(qc_state* h, double* re_out, double* im_out) {
    using namespace qc;
    QC_REQUIRE(h && re_out && im_out, "qc_lambda_inner: null argument");
    QC_REQUIRE(h->scratch_psi, "qc_lambda_inner: no lambda (call qc_hpsi first)");
    QC_CHECK(cudaSetDevice(h->device));
    const int nb = grid_for(h->dim, 2, h->sm_count);
    if (h->ensure_partials((size_t)nb * 2) || h->ensure_out(2)) return 1;
    {
        LaunchScope ls(h, kGradient, 32.0 * (double)h->dim);
        inner_kernel<<<nb, kThreads, 0, h->stream>>>(h->scratch_psi, h->psi, h->dim, h->d_partials);
    }
    QC_CHECK(cudaGetLastError());
    if (fold_partials(h, h->d_partials, 2, nb, nb, h->d_out)) return 1;
    QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->d2h_bytes += 2 * sizeof(double);
    QC_CHECK(cudaStreamSynchronize(h->stream));
    *re_out = h->h_out[0];
    *im_out = h->h_out[1];
    return 0;
}

extern "C" int qc_gradient_backward(qc_state* h, size_t R, const uint64_t* rx, const uint64_t* rz, const double* angle,
                                    double* grad_out) {
    using namespace qc;
    QC_REQUIRE(h, "qc_gradient_backward: null handle");
    if (R == 0) return 0;
    QC_REQUIRE(rx && rz && angle && grad_out, "qc_gradient_backward: null rotation array");
    QC_REQUIRE(h->scratch_psi, "qc_gradient_backward: no lambda (call qc_hpsi first)");
    for (size_t r = 0; r < R; ++r)
        QC_REQUIRE(rx[r] < h->dim && rz[r] < h->dim, "qc_gradient_backward: rotation mask has bits outside the shard");
    QC_CHECK(cudaSetDevice(h->device));
    double2* lam = h->scratch_psi;
    const int nb = grid_for(h->dim, 2, h->sm_count);
    if (h->ensure_partials((size_t)nb * 2 * kGradStrings) || h->ensure_out(2 * R + 2 * kGradStrings)) return 1;
    std::vector<std::pair<size_t, size_t>> runs;
    plan_rotation_runs(R, rx, rz, runs);
    std::vector<double> neg(R);
    for (size_t r = 0; r < R; ++r) neg[r] = -angle[r];
    for (size_t ri = runs.size(); ri-- > 0;) {
        const size_t first = runs[ri].first, count = runs[ri].second;
        for (size_t c0 = 0; c0 < count; c0 += kGradStrings) {
            GradRun gr;
            memset(&gr, 0, sizeof(gr));
            gr.x = rx[first];
            gr.count = (int)std::min<size_t>(kGradStrings, count - c0);
            for (int j = 0; j < gr.count; ++j) gr.z[j] = rz[first + c0 + j];
            {
                LaunchScope ls(h, kGradient, 32.0 * (double)h->dim);
                grad_inner_kernel<<<nb, kThreads, 0, h->stream>>>(lam, h->psi, h->dim, gr, h->d_partials);
            }
            QC_CHECK(cudaGetLastError());
            if (fold_partials(h, h->d_partials, 2 * gr.count, nb, nb, h->d_out + 2 * (first + c0))) return 1;
        }
        if (apply_rotations_to(h, h->psi, count, rx + first, rz + first, neg.data() + first, 1, nullptr) ||
            apply_rotations_to(h, lam, count, rx + first, rz + first, neg.data() + first, 1, nullptr))
            return 1;
    }
    h->d2h_bytes += 2 * R * sizeof(double);
    QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_out, 2 * R * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QC_CHECK(cudaStreamSynchronize(h->stream));
    for (size_t r = 0; r < R; ++r) {
        // dE/dphi_r = Im[(-i)^{nY} C_r]
        const double cr = h->h_out[2 * r], ci = h->h_out[2 * r + 1];
        switch (__builtin_popcountll(rx[r] & rz[r]) & 3) {
            case 0: grad_out[r] = ci; break;
            case 1: grad_out[r] = -cr; break;
            case 2: grad_out[r] = -ci; break;
            default: grad_out[r] = cr; break;
        }
    }
    return 0;
}

extern "C" int qc_energy_gradient(qc_state* h, size_t R, const uint64_t* rx, const uint64_t* rz, const double* angle,
                                  size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc,
                                  double* energy_out, double* grad_out) {
    using namespace qc;
    QC_REQUIRE(h && energy_out, "qc_energy_gradient: null argument");
    QC_REQUIRE(R == 0 || (rx && rz && angle && grad_out), "qc_energy_gradient: null rotation array");
    QC_REQUIRE(T == 0 || (hx && hz && hc), "qc_energy_gradient: null term array");
    *energy_out = 0.0;
    for (size_t r = 0; r < R; ++r) grad_out[r] = 0.0;
    if (T == 0) return 0;
    double im = 0.0;
    if (qc_hpsi(h, T, hx, hz, hc, 0)) return 1;
    if (qc_lambda_inner(h, energy_out, &im)) return 1;
    return qc_gradient_backward(h, R, rx, rz, angle, grad_out);
}
