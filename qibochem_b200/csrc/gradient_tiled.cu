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
// (tile origin, thread bits) -- one POPC per term, thread and tile -- and the 4 "j" bits, whose 16 parities come from a
// Walsh word; terms of an x-mask that agree on the j bits form a SEGMENT and are summed before they touch the 16
// elements.  The partner amplitudes psi_{e ^ x} come from shared memory (conflict-free: XOR with x permutes a warp's
// 32 consecutive slots), so the kernel is bound by those LDS.128 (64 KB per x-mask and tile), not by HBM.
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>

#include "common.cuh"
#include "k2_plan.h"

namespace qc {

constexpr int kHpThreads = 256;
constexpr int kHpElems = kTileSize / kHpThreads;  // 16 elements per thread: tile-local bits 8..11

struct HpPass {            // kernel parameter
    uint64_t s_mask;
    uint8_t spos[kTileBits];  // index-bit position of tile-local bit b (0..7 thread bits, 8..11 element bits)
    uint32_t first_group, n_groups;
};
struct HpGroup {           // 16 bytes
    uint32_t x_loc;        // pairing mask in tile-local positions
    uint32_t first_seg;
    uint16_t n_seg_re, n_seg_im;  // segments of Re-type (even nY) terms first
    uint32_t pad;
};
struct HpSeg {             // 16 bytes
    uint32_t first_term, n_terms;
    uint32_t walsh;        // bit j = parity(j & z_j) of the segment's common z on the element bits
    uint32_t pad;
};
struct HpTerm {            // 32 bytes
    uint64_t z_outer;      // z outside the tile (index-bit positions)
    double d;              // signed coefficient (conventions of hpsi_kernel)
    uint32_t z_thr;        // z on the 8 thread bits (tile-local positions 0..7)
    uint32_t pad[3];
};

struct HpPlan {
    std::vector<uint64_t> key_x, key_z;
    std::vector<double> key_c;
    int n = 0;
    std::vector<HpPass> passes;
    HpGroup* d_groups = nullptr;
    HpSeg* d_segs = nullptr;
    HpTerm* d_terms = nullptr;
    std::vector<uint64_t> left_x, left_z;
    std::vector<double> left_c;
    size_t n_groups = 0;
    ~HpPlan() {
        if (d_groups) cudaFree(d_groups);
        if (d_segs) cudaFree(d_segs);
        if (d_terms) cudaFree(d_terms);
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

__global__ void __launch_bounds__(kHpThreads, 2)
    hpsi_tiled_kernel(const double2* __restrict__ psi, double2* __restrict__ lam, int n, const HpPass pass,
                      const HpGroup* __restrict__ groups, const HpSeg* __restrict__ segs, const HpTerm* __restrict__ terms,
                      int accumulate) {
    extern __shared__ __align__(16) unsigned char hp_smem[];
    double2* tile = reinterpret_cast<double2*>(hp_smem);
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
        __syncthreads();  // previous tile fully consumed
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

        const uint32_t base_lo = (uint32_t)base, base_hi = (uint32_t)(base >> 32);
        for (uint32_t g = 0; g < pass.n_groups; ++g) {
            HpGroup gr;
            {
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&groups[pass.first_group + g]));
                gr.x_loc = raw.x, gr.first_seg = raw.y, gr.n_seg_re = (uint16_t)(raw.z & 0xffffu), gr.n_seg_im = (uint16_t)(raw.z >> 16);
            }
            const double2* partner = tile + (t ^ (gr.x_loc & 0xffu));
            const uint32_t xj = gr.x_loc >> 8;
#pragma unroll 1
            for (int type = 0; type < 2; ++type) {
                const uint32_t s0 = gr.first_seg + (type ? gr.n_seg_re : 0u);
                const uint32_t ns = type ? gr.n_seg_im : gr.n_seg_re;
                if (ns == 0) continue;
                double w[kHpElems];
#pragma unroll
                for (int j = 0; j < kHpElems; ++j) w[j] = 0.0;
                for (uint32_t s = 0; s < ns; ++s) {
                    HpSeg sg;
                    {
                        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(&segs[s0 + s]));
                        sg.first_term = raw.x, sg.n_terms = raw.y, sg.walsh = raw.z;
                    }
                    double b0 = 0.0, b1 = 0.0;
                    uint32_t k = 0;
                    for (; k + 1 < sg.n_terms; k += 2) {
                        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(&terms[sg.first_term + k]));
                        const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(&terms[sg.first_term + k]) + 1);
                        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(&terms[sg.first_term + k + 1]));
                        const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(&terms[sg.first_term + k + 1]) + 1);
                        const unsigned p0 = __popc((base_lo & q0.x) ^ (base_hi & q0.y) ^ (t & r0.x));
                        const unsigned p1 = __popc((base_lo & q1.x) ^ (base_hi & q1.y) ^ (t & r1.x));
                        b0 += hp_sign(__hiloint2double((int)q0.w, (int)q0.z), p0);
                        b1 += hp_sign(__hiloint2double((int)q1.w, (int)q1.z), p1);
                    }
                    if (k < sg.n_terms) {
                        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(&terms[sg.first_term + k]));
                        const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(&terms[sg.first_term + k]) + 1);
                        const unsigned p0 = __popc((base_lo & q0.x) ^ (base_hi & q0.y) ^ (t & r0.x));
                        b0 += hp_sign(__hiloint2double((int)q0.w, (int)q0.z), p0);
                    }
                    const double bsum = b0 + b1;
#pragma unroll
                    for (int j = 0; j < kHpElems; ++j) w[j] += hp_sign(bsum, sg.walsh >> j);
                }
                if (type == 0) {
#pragma unroll
                    for (int j = 0; j < kHpElems; ++j) {
                        const double2 v = partner[(uint32_t)(j ^ xj) * kHpThreads];
                        acc[j].x = fma(w[j], v.x, acc[j].x);
                        acc[j].y = fma(w[j], v.y, acc[j].y);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kHpElems; ++j) {
                        const double2 v = partner[(uint32_t)(j ^ xj) * kHpThreads];
                        acc[j].x = fma(-w[j], v.y, acc[j].x);
                        acc[j].y = fma(w[j], v.x, acc[j].y);
                    }
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
static HpPlan* build_hp_plan(qc_state* h, size_t T, const uint64_t* hx, const uint64_t* hz, const double* hc) {
    const int n = h->n;
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
    std::vector<HpGroup> groups;
    std::vector<HpSeg> segs;
    std::vector<HpTerm> terms;
    for (size_t p = 0; p < sets.size(); ++p) {
        if (members[p].empty()) continue;
        const uint64_t S = sets[p];
        // element bits: the 4 tile bits (never index bit 0) on which the z-masks of the pass's groups vary least
        std::vector<int> sbits;
        for (int b = 0; b < 64; ++b)
            if ((S >> b) & 1ull) sbits.push_back(b);
        std::vector<uint64_t> vary(sbits.size(), 0);
        for (size_t g : members[p]) {
            uint64_t d = 0;
            const uint64_t z0 = hz[order[xgs[g].first]];
            for (size_t k = xgs[g].first; k < xgs[g].first + xgs[g].count; ++k) d |= hz[order[k]] ^ z0;
            for (size_t i = 0; i < sbits.size(); ++i)
                if ((d >> sbits[i]) & 1ull) vary[i] += 1;
        }
        std::vector<size_t> idx(sbits.size());
        std::iota(idx.begin(), idx.end(), (size_t)0);
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
            const bool a0 = sbits[a] == 0, b0 = sbits[b] == 0;  // index bit 0 stays a thread bit (32-byte sectors)
            if (a0 != b0) return b0;
            return vary[a] < vary[b];
        });
        HpPass pass;
        memset(&pass, 0, sizeof(pass));
        pass.s_mask = S;
        std::vector<int> el(idx.begin(), idx.begin() + 4), th(idx.begin() + 4, idx.end());
        std::sort(el.begin(), el.end());
        std::sort(th.begin(), th.end());
        for (int b = 0; b < 8; ++b) pass.spos[b] = (uint8_t)sbits[th[b]];
        for (int b = 0; b < 4; ++b) pass.spos[8 + b] = (uint8_t)sbits[el[b]];
        auto to_local = [&](uint64_t m) {
            uint32_t out = 0;
            for (int b = 0; b < kTileBits; ++b)
                if ((m >> pass.spos[b]) & 1ull) out |= 1u << b;
            return out;
        };
        pass.first_group = (uint32_t)groups.size();
        for (size_t g : members[p]) {
            HpGroup gr;
            memset(&gr, 0, sizeof(gr));
            gr.x_loc = to_local(xgs[g].x);
            gr.first_seg = (uint32_t)segs.size();
            for (int type = 0; type < 2; ++type) {
                // segments keyed by the z pattern on the element bits
                std::map<uint32_t, std::vector<HpTerm>> by_zj;
                for (size_t k = xgs[g].first; k < xgs[g].first + xgs[g].count; ++k) {
                    const uint32_t tix = order[k];
                    const int ny = __builtin_popcountll(hx[tix] & hz[tix]);
                    if ((ny & 1) != type) continue;
                    HpTerm tm;
                    memset(&tm, 0, sizeof(tm));
                    const uint32_t zl = to_local(hz[tix]);
                    tm.z_outer = hz[tix] & ~S;
                    tm.z_thr = zl & 0xffu;
                    tm.d = type ? -hc[tix] * ((((ny - 1) >> 1) & 1) ? -1.0 : 1.0) : hc[tix] * (((ny >> 1) & 1) ? -1.0 : 1.0);
                    by_zj[zl >> 8].push_back(tm);
                }
                for (auto& kv : by_zj) {
                    HpSeg sg;
                    memset(&sg, 0, sizeof(sg));
                    sg.first_term = (uint32_t)terms.size();
                    sg.n_terms = (uint32_t)kv.second.size();
                    for (uint32_t j = 0; j < 16; ++j)
                        if (__builtin_popcount(j & kv.first) & 1) sg.walsh |= 1u << j;
                    terms.insert(terms.end(), kv.second.begin(), kv.second.end());
                    segs.push_back(sg);
                    (type ? gr.n_seg_im : gr.n_seg_re) += 1;
                }
            }
            groups.push_back(gr);
        }
        pass.n_groups = (uint32_t)groups.size() - pass.first_group;
        plan->passes.push_back(pass);
    }
    plan->n_groups = groups.size();
    auto up = [&](void** dptr, const void* src, size_t bytes) {
        if (cudaMalloc(dptr, std::max<size_t>(bytes, 16)) != cudaSuccess) return false;
        return bytes == 0 || cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!up((void**)&plan->d_groups, groups.data(), groups.size() * sizeof(HpGroup)) ||
        !up((void**)&plan->d_segs, segs.data(), segs.size() * sizeof(HpSeg)) ||
        !up((void**)&plan->d_terms, terms.data(), terms.size() * sizeof(HpTerm))) {
        delete plan;
        return nullptr;
    }
    h->h2d_bytes += groups.size() * sizeof(HpGroup) + segs.size() * sizeof(HpSeg) + terms.size() * sizeof(HpTerm);
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
        plan = build_hp_plan(h, T, hx, hz, hc);
        QC_REQUIRE(plan != nullptr, "qc_hpsi: could not build the tiled plan");
        h->hp_plan = plan;
    }
    const size_t smem = (size_t)kTileSize * sizeof(double2);
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
            hpsi_tiled_kernel<<<nb, kHpThreads, smem, h->stream>>>(h->psi, h->scratch_psi, h->n, pass, plan->d_groups, plan->d_segs,
                                                                   plan->d_terms, acc ? 1 : 0);
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
