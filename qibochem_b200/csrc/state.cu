// State life cycle, K0 (basis state), host<->device copies, norm.
#include <string.h>

#include "common.cuh"

namespace qc {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

int Arena::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    size_t ncap = cap ? cap : (1u << 16);
    while (ncap < bytes) ncap *= 2;
    release();
    QC_CHECK(cudaMallocHost(&host, ncap));
    QC_CHECK(cudaMalloc(&dev, ncap));
    cap = ncap;
    return 0;
}
void Arena::release() {
    if (host) cudaFreeHost(host);
    if (dev) cudaFree(dev);
    host = dev = nullptr;
    cap = 0;
}

cudaEvent_t Profiler::get_event() {
    if (!pool.empty()) {
        cudaEvent_t e = pool.back();
        pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void Profiler::collect() {
    for (const Rec& r : pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms[r.cls] += t;
        pool.push_back(r.a);
        pool.push_back(r.b);
    }
    pending.clear();
}
void Profiler::release() {
    for (const Rec& r : pending) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
    pending.clear();
    pool.clear();
}

__global__ void set_one_kernel(double2* psi, uint64_t index) { psi[index] = make_double2(1.0, 0.0); }

__global__ void __launch_bounds__(kThreads) norm2_kernel(const double2* __restrict__ psi, uint64_t dim,
                                                         double* __restrict__ partials) {
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < dim; i += (uint64_t)gridDim.x * kThreads) {
        const double2 a = psi[i];
        acc += a.x * a.x + a.y * a.y;
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// out[g] (+)= sum_b partials[g * stride + b], one CTA per g, fixed summation order.
__global__ void __launch_bounds__(kThreads) fold_partials_kernel(const double* __restrict__ partials, int nb,
                                                                 int stride, double* __restrict__ out) {
    const double* p = partials + (size_t)blockIdx.x * stride;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nb; b += kThreads) acc += p[b];
    const double t = block_sum(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}

int fold_partials(qc_state* h, const double* partials, int ngroups, int nb, int stride, double* out) {
    LaunchScope ls(h, kReduce, 0.0);
    fold_partials_kernel<<<ngroups, kThreads, 0, h->stream>>>(partials, nb, stride, out);
    QC_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace qc

int qc_state::ensure_partials(size_t doubles) {
    if (doubles <= partials_cap) return 0;
    if (d_partials) cudaFree(d_partials);
    d_partials = nullptr;
    partials_cap = 0;
    QC_CHECK(cudaMalloc(&d_partials, doubles * sizeof(double)));
    partials_cap = doubles;
    return 0;
}
int qc_state::ensure_out(size_t doubles) {
    if (doubles <= out_cap) return 0;
    size_t ncap = 1024;
    while (ncap < doubles) ncap *= 2;
    double *nd = nullptr, *nh = nullptr;
    QC_CHECK(cudaMalloc(&nd, ncap * sizeof(double)));
    QC_CHECK(cudaMallocHost(&nh, ncap * sizeof(double)));
    if (d_out) {
        // growth keeps what earlier launches of the same call already left in the buffer (the tiled sum in slot 0
        // while the leftover sweeps ask for more slots)
        QC_CHECK(cudaStreamSynchronize(stream));
        QC_CHECK(cudaMemcpy(nd, d_out, out_cap * sizeof(double), cudaMemcpyDeviceToDevice));
        memcpy(nh, h_out, out_cap * sizeof(double));
        cudaFree(d_out);
        cudaFreeHost(h_out);
    }
    d_out = nd;
    h_out = nh;
    out_cap = ncap;
    return 0;
}

extern "C" {

const char* qc_last_error(void) { return qc::g_last_error.c_str(); }
int qc_version(void) { return 100; }

int qc_device_count(int* count) {
    QC_REQUIRE(count != nullptr, "qc_device_count: null pointer");
    *count = 0;
    QC_CHECK(cudaGetDeviceCount(count));
    QC_REQUIRE(*count > 0, "qc_device_count: no CUDA device visible (there is no CPU fallback)");
    return 0;
}

int qc_create(int device, int n_local_qubits, void* stream, qc_state** out) {
    QC_REQUIRE(out != nullptr, "qc_create: null out pointer");
    QC_REQUIRE(n_local_qubits >= 1 && n_local_qubits <= 40, "qc_create: n_local_qubits must be in [1, 40]");
    int count = 0;
    if (qc_device_count(&count)) return 1;
    QC_REQUIRE(device >= 0 && device < count, "qc_create: bad device index");
    QC_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    QC_CHECK(cudaGetDeviceProperties(&prop, device));
    QC_REQUIRE(prop.major >= 10, "qc_create: this library is built for sm_100a (B200) only");
    qc_state* h = new qc_state();
    h->device = device;
    h->n = n_local_qubits;
    h->dim = 1ull << n_local_qubits;
    h->sm_count = prop.multiProcessorCount;
    if (stream) {
        h->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete h;
            qc::set_error(std::string("cudaStreamCreate failed: ") + cudaGetErrorString(e));
            return 1;
        }
        h->owns_stream = true;
    }
    cudaError_t e = cudaMalloc(&h->psi, h->dim * sizeof(double2));
    if (e != cudaSuccess) {
        qc::set_error(std::string("qc_create: cudaMalloc of the state failed: ") + cudaGetErrorString(e));
        if (h->owns_stream) cudaStreamDestroy(h->stream);
        delete h;
        return 1;
    }
    *out = h;
    return 0;
}

int qc_destroy(qc_state* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->psi) cudaFree(h->psi);
    if (h->scratch_psi) cudaFree(h->scratch_psi);
    if (h->d_partials) cudaFree(h->d_partials);
    if (h->d_out) cudaFree(h->d_out);
    if (h->h_out) cudaFreeHost(h->h_out);
    h->arena.release();
    h->prof.release();
    for (qc::TiledPlan* p : h->k2plans) qc::free_tiled_plan(p);
    if (h->hp_plan) qc::free_hp_plan(h->hp_plan);
    if (h->sweep_plan) qc::free_sweep_plan(h->sweep_plan);
    if (h->owns_stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int qc_num_qubits(const qc_state* h) { return h ? h->n : -1; }

int qc_state_device_ptr(qc_state* h, void** ptr) {
    QC_REQUIRE(h && ptr, "qc_state_device_ptr: null argument");
    *ptr = h->psi;
    return 0;
}

int qc_synchronize(qc_state* h) {
    QC_REQUIRE(h, "qc_synchronize: null handle");
    QC_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

int qc_set_basis_state(qc_state* h, uint64_t index, int has_one) {
    QC_REQUIRE(h, "qc_set_basis_state: null handle");
    QC_REQUIRE(!has_one || index < h->dim, "qc_set_basis_state: index out of range");
    QC_CHECK(cudaSetDevice(h->device));
    QC_CHECK(cudaMemsetAsync(h->psi, 0, h->dim * sizeof(double2), h->stream));
    h->prof.bytes[qc::kK0Init] += 16.0 * (double)h->dim;  // the memset writes the shard once
    if (has_one) {
        qc::LaunchScope ls(h, qc::kK0Init, 0.0);
        qc::set_one_kernel<<<1, 1, 0, h->stream>>>(h->psi, index);
        QC_CHECK(cudaGetLastError());
    }
    return 0;
}

int qc_norm2(qc_state* h, double* out) {
    QC_REQUIRE(h && out, "qc_norm2: null argument");
    QC_CHECK(cudaSetDevice(h->device));
    const int nb = qc::grid_for(h->dim, 4, h->sm_count);
    if (h->ensure_partials(nb) || h->ensure_out(1)) return 1;
    {
        qc::LaunchScope ls(h, qc::kReduce, 16.0 * (double)h->dim);
        qc::norm2_kernel<<<nb, qc::kThreads, 0, h->stream>>>(h->psi, h->dim, h->d_partials);
    }
    QC_CHECK(cudaGetLastError());
    if (qc::fold_partials(h, h->d_partials, 1, nb, nb, h->d_out)) return 1;
    QC_CHECK(cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QC_CHECK(cudaStreamSynchronize(h->stream));
    *out = h->h_out[0];
    return 0;
}

int qc_set_option(qc_state* h, const char* key, int64_t value) {
    QC_REQUIRE(h && key, "qc_set_option: null argument");
    if (strcmp(key, "k2_mode") == 0) {
        QC_REQUIRE(value >= 0 && value <= 2, "qc_set_option: k2_mode must be 0 (auto), 1 (sweeps) or 2 (tiles)");
        h->k2_mode = (int)value;
        return 0;
    }
    if (strcmp(key, "k1_mode") == 0) {
        QC_REQUIRE(value >= 0 && value <= 3, "qc_set_option: k1_mode must be 0 (auto), 1 (one pass per run), 2 (tiles) or 3 (register phases)");
        h->k1_mode = (int)value;
        return 0;
    }
    if (strcmp(key, "k2_real") == 0) {
        QC_REQUIRE(value == 0 || value == 1, "qc_set_option: k2_real must be 0 or 1");
        h->k2_real = (int)value;
        return 0;
    }
    if (strcmp(key, "k2_real_taken") == 0) {  // query: did the last tiled expectation run the real-amplitude kernels?
        return h->last_k2_real ? 0 : 3;
    }
    qc::set_error(std::string("qc_set_option: unknown key ") + key);
    return 2;
}

int qc_transfer_bytes(qc_state* h, uint64_t* h2d_bytes, uint64_t* d2h_bytes, int reset) {
    QC_REQUIRE(h, "qc_transfer_bytes: null handle");
    if (h2d_bytes) *h2d_bytes = h->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = h->d2h_bytes;
    if (reset) h->h2d_bytes = h->d2h_bytes = 0;
    return 0;
}

int qc_profile_enable(qc_state* h, int enabled) {
    QC_REQUIRE(h, "qc_profile_enable: null handle");
    h->prof.enabled = enabled != 0;
    return 0;
}

int qc_profile_reset(qc_state* h) {
    QC_REQUIRE(h, "qc_profile_reset: null handle");
    QC_CHECK(cudaSetDevice(h->device));
    QC_CHECK(cudaStreamSynchronize(h->stream));
    h->prof.collect();
    for (int c = 0; c < qc::kNumClasses; ++c) {
        h->prof.ms[c] = 0.0;
        h->prof.bytes[c] = 0.0;
        h->prof.launches[c] = 0;
    }
    return 0;
}

int qc_profile_get(qc_state* h, int kernel_class, double* ms_total, uint64_t* launches, double* algorithmic_bytes) {
    QC_REQUIRE(h, "qc_profile_get: null handle");
    QC_REQUIRE(kernel_class >= 0 && kernel_class < qc::kNumClasses, "qc_profile_get: bad kernel class");
    QC_CHECK(cudaSetDevice(h->device));
    QC_CHECK(cudaStreamSynchronize(h->stream));
    h->prof.collect();
    if (ms_total) *ms_total = h->prof.ms[kernel_class];
    if (launches) *launches = h->prof.launches[kernel_class];
    if (algorithmic_bytes) *algorithmic_bytes = h->prof.bytes[kernel_class];
    return 0;
}

int qc_copy_state_to_host(qc_state* h, double* out, uint64_t offset, uint64_t count) {
    QC_REQUIRE(h && out, "qc_copy_state_to_host: null argument");
    QC_REQUIRE(offset + count <= h->dim, "qc_copy_state_to_host: range out of bounds");
    QC_CHECK(cudaSetDevice(h->device));
    QC_CHECK(cudaMemcpyAsync(out, h->psi + offset, count * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    QC_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

int qc_copy_state_from_host(qc_state* h, const double* in, uint64_t offset, uint64_t count) {
    QC_REQUIRE(h && in, "qc_copy_state_from_host: null argument");
    QC_REQUIRE(offset + count <= h->dim, "qc_copy_state_from_host: range out of bounds");
    QC_CHECK(cudaSetDevice(h->device));
    QC_CHECK(cudaMemcpyAsync(h->psi + offset, in, count * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    h->h2d_bytes += count * sizeof(double2);
    QC_CHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
