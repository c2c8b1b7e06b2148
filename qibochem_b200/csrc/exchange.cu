// K4: exchange of a "global" (rank-selecting) index bit with a local index bit between two shards, in place, over
// peer memory (NVLink P2P through CUDA IPC, or two shards of one process).
//
// The reference is single-process (no counterpart); this is the one exchange step of the sharded path: a Pauli
// string with X/Y on a global bit needs the partner shard's amplitudes, so the host layer first swaps that global
// bit with a local bit that upcoming strings do not touch, after which every kernel above runs unchanged.
//
// Two shards A (global bit = 0) and B (global bit = 1).  Swapping global bit g with local bit l exchanges
//     A[i | 1<<l]  <->  B[i]          for every i with bit l clear  (2^(n-1) element pairs)
// and leaves A[bit l = 0] and B[bit l = 1] in place.  Each element pair is moved by exactly one thread of exactly
// one rank (rank side s handles the pairs p with p's top bit == s... see `part`), reading one local and one remote
// amplitude and writing both, so the exchange needs no staging buffer and each NVLink direction carries the
// minimum of B_local/2 bytes per rank (half as read responses, half as writes).
#include <string.h>

#include "common.cuh"

namespace qc {

template <int U>
__global__ void __launch_bounds__(kThreads) swap_bit_kernel(double2* __restrict__ a_shard, double2* __restrict__ b_shard,
                                                            uint64_t p_begin, uint64_t p_end, int bit) {
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    const uint64_t bm = 1ull << bit;
    for (uint64_t p0 = p_begin + (uint64_t)blockIdx.x * kThreads + threadIdx.x; p0 < p_end; p0 += stride * U) {
        uint64_t idx[U];
        double2 va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t p = p0 + (uint64_t)u * stride;
            idx[u] = insert_zero_bit(p, bit);
            if (p < p_end) {
                va[u] = a_shard[idx[u] | bm];
                vb[u] = b_shard[idx[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t p = p0 + (uint64_t)u * stride;
            if (p < p_end) {
                a_shard[idx[u] | bm] = vb[u];
                b_shard[idx[u]] = va[u];
            }
        }
    }
}

}  // namespace qc

extern "C" {

int qc_ipc_export(qc_state* h, unsigned char handle[64]) {
    QC_REQUIRE(h && handle, "qc_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    QC_CHECK(cudaSetDevice(h->device));
    cudaIpcMemHandle_t hd;
    QC_CHECK(cudaIpcGetMemHandle(&hd, h->psi));
    memcpy(handle, &hd, 64);
    return 0;
}

int qc_ipc_export_scratch(qc_state* h, unsigned char handle[64]) {
    QC_REQUIRE(h && handle, "qc_ipc_export_scratch: null argument");
    QC_CHECK(cudaSetDevice(h->device));
    if (!h->scratch_psi) QC_CHECK(cudaMalloc(&h->scratch_psi, h->dim * sizeof(double2)));
    cudaIpcMemHandle_t hd;
    QC_CHECK(cudaIpcGetMemHandle(&hd, h->scratch_psi));
    memcpy(handle, &hd, 64);
    return 0;
}

int qc_scratch_device_ptr(qc_state* h, void** ptr) {
    QC_REQUIRE(h && ptr, "qc_scratch_device_ptr: null argument");
    QC_CHECK(cudaSetDevice(h->device));
    if (!h->scratch_psi) QC_CHECK(cudaMalloc(&h->scratch_psi, h->dim * sizeof(double2)));
    *ptr = h->scratch_psi;
    return 0;
}

int qc_ipc_open(qc_state* h, const unsigned char handle[64], void** peer_ptr) {
    QC_REQUIRE(h && handle && peer_ptr, "qc_ipc_open: null argument");
    QC_CHECK(cudaSetDevice(h->device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    QC_CHECK(cudaIpcOpenMemHandle(peer_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int qc_ipc_close(qc_state* h, void* peer_ptr) {
    QC_REQUIRE(h, "qc_ipc_close: null handle");
    if (!peer_ptr) return 0;
    QC_CHECK(cudaSetDevice(h->device));
    QC_CHECK(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}

static int swap_array_with_peer(qc_state* h, double2* mine, void* peer_ptr, int local_bit, int my_side, int part);

int qc_swap_with_peer(qc_state* h, void* peer_ptr, int local_bit, int my_side, int part) {
    QC_REQUIRE(h && peer_ptr, "qc_swap_with_peer: null argument");
    return swap_array_with_peer(h, h->psi, peer_ptr, local_bit, my_side, part);
}

int qc_swap_scratch_with_peer(qc_state* h, void* peer_scratch_ptr, int local_bit, int my_side, int part) {
    QC_REQUIRE(h && peer_scratch_ptr, "qc_swap_scratch_with_peer: null argument");
    QC_REQUIRE(h->scratch_psi, "qc_swap_scratch_with_peer: this shard has no scratch array yet");
    return swap_array_with_peer(h, h->scratch_psi, peer_scratch_ptr, local_bit, my_side, part);
}

static int swap_array_with_peer(qc_state* h, double2* mine, void* peer_ptr, int local_bit, int my_side, int part) {
    using namespace qc;
    QC_REQUIRE(local_bit >= 0 && local_bit < h->n, "qc_swap_with_peer: local_bit out of range");
    QC_REQUIRE((my_side == 0 || my_side == 1) && (part == 0 || part == 1), "qc_swap_with_peer: side/part must be 0 or 1");
    QC_CHECK(cudaSetDevice(h->device));
    const uint64_t npairs = h->dim >> 1;
    const uint64_t half = npairs >> 1;  // n >= 2 shards exchange two halves; a 1-qubit shard is moved by part 0
    const uint64_t p_begin = part == 0 ? 0 : (half ? half : npairs);
    const uint64_t p_end = part == 0 ? (half ? half : npairs) : npairs;
    if (p_begin >= p_end) return 0;
    double2* a = my_side == 0 ? mine : (double2*)peer_ptr;
    double2* b = my_side == 0 ? (double2*)peer_ptr : mine;
    const int nb = grid_for(p_end - p_begin, 4, h->sm_count);
    LaunchScope ls(h, kK4Exchange, 32.0 * (double)(p_end - p_begin));
    swap_bit_kernel<4><<<nb, kThreads, 0, h->stream>>>(a, b, p_begin, p_end, local_bit);
    QC_CHECK(cudaGetLastError());
    return 0;
}

int qc_swap_local_pair(qc_state* lo, qc_state* hi, int local_bit) {
    using namespace qc;
    QC_REQUIRE(lo && hi, "qc_swap_local_pair: null handle");
    QC_REQUIRE(lo->n == hi->n, "qc_swap_local_pair: shards differ in size");
    QC_REQUIRE(local_bit >= 0 && local_bit < lo->n, "qc_swap_local_pair: local_bit out of range");
    QC_CHECK(cudaSetDevice(lo->device));
    QC_CHECK(cudaStreamSynchronize(hi->stream));
    const uint64_t npairs = lo->dim >> 1;
    const int nb = grid_for(npairs, 4, lo->sm_count);
    LaunchScope ls(lo, kK4Exchange, 32.0 * (double)npairs);
    swap_bit_kernel<4><<<nb, kThreads, 0, lo->stream>>>(lo->psi, hi->psi, 0, npairs, local_bit);
    QC_CHECK(cudaGetLastError());
    QC_CHECK(cudaStreamSynchronize(lo->stream));
    return 0;
}

}  // extern "C"
