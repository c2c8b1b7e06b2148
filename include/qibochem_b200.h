/*
 * qibochem_b200 -- C ABI of the B200-native UCC-energy hot path.
 *
 * This is the drop-in boundary for ONE path of qiboteam/qibochem (reference citations are relative to the
 * reference checkout): HF basis state -> UCC Pauli rotations exp(-i phi P / 2) on a complex128 state vector ->
 * Hamiltonian expectation sum_k c_k <psi|P_k|psi>, exact and sampled (QWC-grouped).  The reference reaches this
 * arithmetic through qibo's object protocol (Circuit.__call__, SymbolicHamiltonian.expectation /
 * .expectation_from_samples), not through an FFI; each entry point names the reference call it replaces.
 *
 * Conventions
 *   - All masks are in INDEX-BIT positions of the local shard: qubit q of an n-qubit register is bit (n-1-q)
 *     (qibo puts qubit 0 on the most significant index bit).  x = X|Y sites, z = Z|Y sites of a Pauli string.
 *   - (P psi)_i = (-i)^{popcount(x&z)} * (-1)^{popcount(i&z)} * psi_{i^x}.
 *   - A rotation with angle phi applies  psi <- cos(phi/2) psi - i sin(phi/2) P psi   (== qibochem's
 *     _expi_pauli(n, P, theta) with phi = -2 theta, src/qibochem/ansatz/_ansatz.py:55-92).
 *   - Every function returns 0 on success, non-zero on error; qc_last_error() returns the message of the last
 *     failure on the calling thread.  There is no CPU fallback: without a CUDA device every call fails.
 *   - Host arrays are borrowed for the duration of the call.  A handle is not thread-safe.
 *   - A handle owns ONE shard of 2^n_local amplitudes on one GPU.  Multi-GPU runs use one handle (one process)
 *     per GPU; the host layer moves index bits between "global" (rank) and local positions with
 *     qc_swap_with_peer and folds the rank-dependent signs of global Z factors into angles / coefficients.
 */
#ifndef QIBOCHEM_B200_H
#define QIBOCHEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qc_state qc_state;

/* ---- library / device ------------------------------------------------------------------------------------ */
const char* qc_last_error(void);
int qc_version(void);
/* Number of visible CUDA devices (0 and an error if none). */
int qc_device_count(int* count);

/* ---- state life cycle  (replaces: allocation of the state inside qibo Circuit.__call__) ------------------- */
/* stream: a cudaStream_t created by the caller in the same CUDA context (e.g. torch's current stream), or NULL
 * for a stream owned by the handle.  Allocates 16 * 2^n_local bytes of HBM. */
int qc_create(int device, int n_local_qubits, void* stream, qc_state** out);
int qc_destroy(qc_state* h);
int qc_num_qubits(const qc_state* h);
/* Raw device pointer of the shard (complex128[2^n_local]); for the host layer's NCCL / IPC plumbing only. */
int qc_state_device_ptr(qc_state* h, void** ptr);
int qc_synchronize(qc_state* h);

/* ---- K0: basis state  (replaces hf_circuit's X-gate passes, src/qibochem/ansatz/ansatz.py:276-282) --------- */
/* psi <- 0; if has_one, psi[index] <- 1.  (A sharded caller passes has_one=1 only on the owning rank.) */
int qc_set_basis_state(qc_state* h, uint64_t index, int has_one);

/* ---- K1: Pauli rotations  (replaces the 16-40 gate passes per _expi_pauli, _ansatz.py:80-92, executed by
 *      qibo's backend gate loop reached from Circuit.__call__) ---------------------------------------------- */
/* Applies R rotations IN ORDER.  Consecutive strings that share an x-mask and commute are fused into one pass
 * (the 8 strings of a JW double excitation always do), and consecutive runs whose x-masks are linearly independent
 * (up to 5 distinct masks, e.g. 5 double excitations = 40 strings) are applied in ONE in-place pass.  n_passes_out (may be NULL) receives the number of full-state
 * passes actually launched. */
int qc_apply_pauli_rotations(qc_state* h, size_t R, const uint64_t* xmask, const uint64_t* zmask,
                             const double* angle, int* n_passes_out);

/* K0 + K1 in one call (what Circuit.__call__ does for hf_circuit + ucc_circuit, src/qibochem/ansatz/ansatz.py:253-350):
 * identical to qc_set_basis_state followed by qc_apply_pauli_rotations; on a small register (<= 13 local qubits) the
 * basis state and the WHOLE program run in one launch with the state in shared memory. */
int qc_prepare_state(qc_state* h, uint64_t basis_index, int has_one, size_t R, const uint64_t* xmask,
                     const uint64_t* zmask, const double* angle, int* n_passes_out);

/* ---- K2: exact expectation  (replaces SymbolicHamiltonian.expectation(circuit), called e.g. at
 *      tests/test_ansatz.py:131, examples/ucc_example1.py:45) ------------------------------------------------ */
/* energy_out <- sum_k coeff[k] * Re<psi|P_k|psi> over this shard (no constant term; the host adds it and
 * all-reduces across shards).  per_term (may be NULL) receives the T individual values Re<psi|P_k|psi>.
 * Terms are grouped by x-mask; every group costs one read-only sweep of the shard. n_sweeps_out may be NULL. */
int qc_expectation(qc_state* h, size_t T, const uint64_t* xmask, const uint64_t* zmask, const double* coeff,
                   double* energy_out, double* per_term, int* n_sweeps_out);
/* Options: "k2_mode" = 0 auto (default), 1 one HBM sweep per x-mask, 2 shared-memory tiles serving many x-masks per
 * HBM read (needs >= 12 local qubits; falls back to sweeps for x-masks that do not fit a tile).
 * "k1_mode" = 0 auto (default: consecutive fused runs with linearly independent x-masks -- up to 5 distinct ones --
 * are applied from registers in one in-place HBM pass), 1 one HBM pass per fused run, 2 shared-memory tiles
 * (consecutive runs whose x-masks fit a 12-bit tile), 3 register phases even for a lone run.
 * "k2_real" = 1: the tiled expectation checks (one read of the shard) whether every imaginary part is exactly zero --
 * true for a Hartree-Fock determinant evolved by UCC rotations -- and then runs real-amplitude kernels (one
 * multiplication per pair product, odd-nY terms vanish, three CTAs per SM); default 0: always the complex128 kernels.
 * ("k2_real_taken", any value) returns 0 if the last tiled expectation took the real path, 3 if not. */
int qc_set_option(qc_state* h, const char* key, int64_t value);
/* Host-only (no GPU needed): plans the tiled expectation of a term list and reports its shape.  stats_out: 16 x
 * uint64 {x-masks, tiled x-masks, HBM passes, register sub-cubes, DOT records, LIN entries, leftover terms (sweep
 * kernel), record units, then the number of sub-cubes holding 0,1,..,>=7 x-masks}.  check_out (may be NULL) receives
 * the planner self-check: max |W_plan - W_direct| over random (x-mask, index) samples of the per-pair weights
 * W_x(i) = sum_k d_k (-1)^{popcount(i & z_k)}; negative if some x-mask is not served exactly once. */
int qc_k2_plan_stats(int n_local_qubits, size_t T, const uint64_t* xmask, const uint64_t* zmask, const double* coeff,
                     uint64_t* stats_out, double* check_out);
/* ---- adjoint-mode gradient (replaces qibo.derivative.parameter_shift, used at doc/source/tutorials/adaptive.rst:46-61,
 *      :118-141 of the reference: two full energy evaluations PER parameter) ------------------------------------- */
/* Precondition: the handle holds psi = U psi0 where U is the given rotation program (qc_apply_pauli_rotations with
 * the same arrays).  energy_out <- sum_k coeff[k] Re<psi|P_k|psi>;  grad_out[r] <- dE/d angle[r] for all R rotations,
 * from lambda = H psi and one backward walk over the fused runs (dE/dphi_r = Im<lambda_r|P_r|psi_r>).  Postcondition:
 * the handle holds psi0 again (the program has been un-applied).  Uses one scratch array of the shard's size. */
int qc_energy_gradient(qc_state* h, size_t R, const uint64_t* xmask, const uint64_t* zmask, const double* angle,
                       size_t T, const uint64_t* hxmask, const uint64_t* hzmask, const double* coeff,
                       double* energy_out, double* grad_out);
/* The three stages of qc_energy_gradient as separate calls, so that a SHARDED run (one process per GPU) can re-layout
 * psi and lambda between them (x-masks on a global bit need qc_swap_with_peer + qc_swap_scratch_with_peer first):
 *   qc_hpsi:              lambda (+)= sum_k coeff[k] P_k psi   (lambda = the handle's scratch array; accumulate != 0 adds)
 *   qc_lambda_inner:      <lambda|psi> over this shard (re = the energy's partial sum)
 *   qc_gradient_backward: for the given rotations, last to first: grad_out[r] <- this shard's part of dE/d angle[r]
 *                         = Im<lambda|P_r|psi>, then psi and lambda <- U_r^dagger psi, U_r^dagger lambda. */
int qc_hpsi(qc_state* h, size_t T, const uint64_t* hxmask, const uint64_t* hzmask, const double* coeff, int accumulate);
int qc_lambda_inner(qc_state* h, double* re_out, double* im_out);
int qc_gradient_backward(qc_state* h, size_t R, const uint64_t* xmask, const uint64_t* zmask, const double* angle,
                         double* grad_out);
/* Host-only (no GPU needed): shape of the tiled H|psi> plan of a term list.  stats_out: 8 x uint64 {passes, x-masks
 * served by tiles, segments, terms, largest number of x-masks in one pass, leftover terms (sweep kernel), x-masks
 * with a single segment, of those with an empty sign word}. */
int qc_hpsi_plan_stats(int n_local_qubits, size_t T, const uint64_t* hxmask, const uint64_t* hzmask, const double* coeff,
                       uint64_t* stats_out);
/* <psi|psi> over this shard. */
int qc_norm2(qc_state* h, double* out);

/* ---- K5: sampling in a product Pauli basis (replaces Circuit.__call__(nshots=) with gates.M(q, basis=...),
 *      src/qibochem/measurement/result.py:106-114) ---------------------------------------------------------- */
/* Draws n_shots indices i with probability |<i|B|psi>|^2 where B rotates the qubits in x_basis_mask with H and
 * those in y_basis_mask with H S^dagger (psi itself is left untouched: the rotation runs on a scratch copy).
 * samples_out: DEVICE or HOST buffer of n_shots uint64 (flag samples_on_device).  Philox4x32-10, counter = shot
 * index + shot_offset, so shards / repeated calls can draw disjoint streams.  total_prob_out may be NULL. */
int qc_sample(qc_state* h, uint64_t x_basis_mask, uint64_t y_basis_mask, size_t n_shots, uint64_t seed,
              uint64_t shot_offset, uint64_t* samples_out, int samples_on_device, double* total_prob_out);

/* ---- K3: integer shot statistics (replaces the Python loops of qibo's expectation_from_samples reached from
 *      src/qibochem/measurement/result.py:38-54; bit-exact) -------------------------------------------------- */
/* S_out[j] = sum_s weight_s * (-1)^{popcount(samples[s] & zmask[j])}, weight_s = counts[s] or 1 if counts==NULL.
 * samples/counts: device or host buffers according to on_device. */
int qc_parity_sums(qc_state* h, const uint64_t* samples, const int64_t* counts, size_t n, int on_device,
                   const uint64_t* zmask, size_t J, int64_t* S_out);
/* Histogram of samples over `keep_mask` bits (result.frequencies()): keys_out/counts_out are HOST buffers of
 * capacity `cap`; keys are the compacted bit patterns (pext(sample, keep_mask)), ascending.  n_unique_out
 * receives the number of distinct keys (error if > cap). */
int qc_frequencies(qc_state* h, const uint64_t* samples, size_t n, int on_device, uint64_t keep_mask,
                   uint64_t* keys_out, int64_t* counts_out, size_t cap, size_t* n_unique_out);

/* Squared deviations of a term group's value over the DISTINCT outcomes of a measurement (replaces the Python loop of
 * sample_statistics, src/qibochem/measurement/result.py:150-155):  v_u = sum_j coeff[j] (-1)^{popcount(keys[u] & zmask[j])};
 * sumsq_unweighted_out <- sum_u (v_u - mean)^2  (what the reference divides by n_shots - 1: every distinct outcome
 * counted once), sumsq_weighted_out <- sum_u counts[u] (v_u - mean)^2  (the usual sample variance numerator).
 * keys / counts / zmask / coeff are host arrays. */
int qc_outcome_variance(qc_state* h, const uint64_t* keys, const int64_t* counts, size_t n_unique,
                        const uint64_t* zmask, const double* coeff, size_t J, double mean, double* sumsq_unweighted_out,
                        double* sumsq_weighted_out);

/* ---- host-side term grouping (replaces the O(T^2) string parsing + networkx colouring of
 *      src/qibochem/measurement/util.py:22-79 reached from _measurement_basis_rotations("qwc")) -------------------- */
/* Greedy colouring ("largest_first", networkx's default) of the non-commutation graph of T Pauli terms given as
 * masks; qubitwise != 0: qubit-wise commutation, else general commutation.  colour_out[k] <- group of term k. */
int qc_group_commuting(size_t T, const uint64_t* xmask, const uint64_t* zmask, int qubitwise, int32_t* colour_out,
                       int32_t* n_colours_out);

/* ---- live profiler / launch counter (bench.py: roofline of the dominant kernel, gpu_launches) ----------------- */
/* Kernel classes: 0 K0 init, 1 K1 rotation pass, 2 K2 expectation sweeps, 3 reductions, 4 K5 sampling, 5 K3 shot
 * statistics, 6 K4 exchange, 7 K2 tiled expectation passes, 8 gradient sweeps.  Launch counts and algorithmic bytes are always accumulated; when enabled, every
 * launch is additionally bracketed by CUDA events on the handle's stream and its device time accumulated. */
/* Host->device / device->host bytes the compute entry points (programs, tables, plans, kernel parameters, results,
 * qc_copy_state_from_host) have moved since the last reset: bench.py's e2e.h2d_bytes_per_step / d2h_bytes_per_step. */
int qc_transfer_bytes(qc_state* h, uint64_t* h2d_bytes, uint64_t* d2h_bytes, int reset);
int qc_profile_enable(qc_state* h, int enabled);
int qc_profile_reset(qc_state* h);
int qc_profile_get(qc_state* h, int kernel_class, double* ms_total, uint64_t* launches, double* algorithmic_bytes);

/* ---- host <-> device copies of amplitudes (tests, checkpoints; n <= ~30) ----------------------------------- */
/* interleaved (re, im) doubles */
int qc_copy_state_to_host(qc_state* h, double* out, uint64_t offset, uint64_t count);
int qc_copy_state_from_host(qc_state* h, const double* in, uint64_t offset, uint64_t count);

/* ---- K4: global <-> local index-bit exchange between two shards over NVLink peer memory -------------------- */
/* 64-byte CUDA IPC handle of this shard, to be passed to the partner process. */
int qc_ipc_export(qc_state* h, unsigned char handle[64]);
/* The same for the scratch array (lambda of the adjoint gradient), allocated on first use; and its device pointer. */
int qc_ipc_export_scratch(qc_state* h, unsigned char handle[64]);
int qc_scratch_device_ptr(qc_state* h, void** ptr);
int qc_ipc_open(qc_state* h, const unsigned char handle[64], void** peer_ptr);
int qc_ipc_close(qc_state* h, void* peer_ptr);
/* Swaps the half of this shard where `local_bit` == 1 - my_side with the partner's half where `local_bit` ==
 * my_side ... see DESIGN.md "K4".  Each rank of the pair moves one half of the exchanged elements: `part` in
 * {0,1} selects which. Caller brackets the call with a cross-rank barrier on both sides. */
int qc_swap_with_peer(qc_state* h, void* peer_ptr, int local_bit, int my_side, int part);
/* The same exchange on the scratch arrays of the two shards (lambda follows psi through every re-layout). */
int qc_swap_scratch_with_peer(qc_state* h, void* peer_scratch_ptr, int local_bit, int my_side, int part);
/* Same exchange inside ONE process between two handles' shards (single-GPU emulation of two ranks for tests, or
 * a single process driving peer-enabled GPUs). */
int qc_swap_local_pair(qc_state* lo, qc_state* hi, int local_bit);

#ifdef __cplusplus
}
#endif
#endif
