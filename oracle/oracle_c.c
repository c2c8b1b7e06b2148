/*
 * CPU ORACLE (C restatement) -- TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing under qibochem_b200/ links this.
 *
 * Restates, gate by gate, what the reference path costs on host cores:
 *   - `_expi_pauli` gate programs (src/qibochem/ansatz/_ansatz.py:55-92): H, S, SDG, CNOT, RZ (+ X for hf_circuit,
 *     src/qibochem/ansatz/ansatz.py:276-282), each gate ONE full pass over the 2^n complex128 amplitudes, applied in
 *     place with OpenMP over all host threads (the "qibojit-equivalent" engine; qibo's default numpy backend does
 *     the same arithmetic out of place and single-threaded [recalled]);
 *   - SymbolicHamiltonian.expectation [recalled]: per term, copy the state, apply each factor's 2x2 matrix (one
 *     pass per factor), accumulate coeff * <psi|tmp>.
 * Conventions: qubit q <-> index bit n-1-q;  RZ(phi) = diag(e^{-i phi/2}, e^{+i phi/2}).
 * Parity: pinned through the numpy oracle (tests/test_oracle_c.py compares both on random inputs) which in turn is
 * pinned to the reference's goldens (tests/test_oracle_goldens.py).  qibo itself is not installed: "parity
 * unpinned" beyond those goldens.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

enum { G_X = 0, G_H = 1, G_S = 2, G_SDG = 3, G_CNOT = 4, G_RZ = 5, G_Y = 6, G_Z = 7 };

int oc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline sets its thread count explicitly instead of
 * inheriting that (0 = all logical cores). */
void oc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}

static inline uint64_t insert0(uint64_t p, int b) { return ((p >> b) << (b + 1)) | (p & ((1ull << b) - 1)); }

/* one 1-qubit gate pass, in place */
static void gate1(cplx* s, int n, int q, const cplx m[4]) {
    const int b = n - 1 - q;
    const uint64_t half = 1ull << (n - 1), bm = 1ull << b;
#pragma omp parallel for schedule(static)
    for (uint64_t p = 0; p < half; ++p) {
        const uint64_t i = insert0(p, b), j = i | bm;
        const cplx a = s[i], c = s[j];
        s[i] = m[0] * a + m[1] * c;
        s[j] = m[2] * a + m[3] * c;
    }
}

static void cnot(cplx* s, int n, int control, int target) {
    const int bc = n - 1 - control, bt = n - 1 - target;
    const int lo = bc < bt ? bc : bt, hi = bc < bt ? bt : bc;
    const uint64_t quarter = 1ull << (n - 2);
#pragma omp parallel for schedule(static)
    for (uint64_t p = 0; p < quarter; ++p) {
        const uint64_t base = insert0(insert0(p, lo), hi) | (1ull << bc);
        const uint64_t j = base | (1ull << bt);
        const cplx t = s[base];
        s[base] = s[j];
        s[j] = t;
    }
}

/* kinds[g], q0[g] (target / control), q1[g] (CNOT target), params[g] (RZ angle) */
void oc_run_gates(double* state, int n, int64_t ngates, const int32_t* kinds, const int32_t* q0, const int32_t* q1,
                  const double* params) {
    cplx* s = (cplx*)state;
    const double r = 0.70710678118654752440;
    for (int64_t g = 0; g < ngates; ++g) {
        cplx m[4];
        switch (kinds[g]) {
            case G_X: m[0] = 0; m[1] = 1; m[2] = 1; m[3] = 0; gate1(s, n, q0[g], m); break;
            case G_Y: m[0] = 0; m[1] = -I; m[2] = I; m[3] = 0; gate1(s, n, q0[g], m); break;
            case G_Z: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = -1; gate1(s, n, q0[g], m); break;
            case G_H: m[0] = r; m[1] = r; m[2] = r; m[3] = -r; gate1(s, n, q0[g], m); break;
            case G_S: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = I; gate1(s, n, q0[g], m); break;
            case G_SDG: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = -I; gate1(s, n, q0[g], m); break;
            case G_RZ:
                m[0] = cexp(-0.5 * I * params[g]); m[1] = 0; m[2] = 0; m[3] = cexp(0.5 * I * params[g]);
                gate1(s, n, q0[g], m);
                break;
            case G_CNOT: cnot(s, n, q0[g], q1[g]); break;
            default: break;
        }
    }
}

/* qibo-style expectation: sum_k coeff_k Re<psi| (prod factors) |psi>, factors applied as gate passes on a copy.
 * term k owns factors [offsets[k], offsets[k+1]): qubits[] and paulis[] (0=X, 1=Y, 2=Z). tmp: scratch of 2^n. */
double oc_expectation_terms(const double* state, double* tmp, int n, int64_t nterms, const double* coeffs,
                            const int64_t* offsets, const int32_t* qubits, const int32_t* paulis) {
    const cplx* s = (const cplx*)state;
    cplx* t = (cplx*)tmp;
    const uint64_t dim = 1ull << n;
    double total = 0.0;
    for (int64_t k = 0; k < nterms; ++k) {
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < dim; ++i) t[i] = s[i];
        for (int64_t f = offsets[k]; f < offsets[k + 1]; ++f) {
            cplx m[4];
            if (paulis[f] == 0) { m[0] = 0; m[1] = 1; m[2] = 1; m[3] = 0; }
            else if (paulis[f] == 1) { m[0] = 0; m[1] = -I; m[2] = I; m[3] = 0; }
            else { m[0] = 1; m[1] = 0; m[2] = 0; m[3] = -1; }
            gate1(t, n, qubits[f], m);
        }
        double re = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : re)
        for (uint64_t i = 0; i < dim; ++i) re += creal(conj(s[i]) * t[i]);
        total += coeffs[k] * re;
    }
    return total;
}

/* mask-level restatements (contract of K1 / K2), used to cross-check the gate path at larger n */
void oc_pauli_rotation(double* state, int n, uint64_t x, uint64_t z, double phi) {
    cplx* s = (cplx*)state;
    const uint64_t dim = 1ull << n;
    const double c = cos(0.5 * phi), sn = sin(0.5 * phi);
    const int ny = __builtin_popcountll(x & z);
    cplx g = 1;
    for (int k = 0; k < (ny & 3); ++k) g *= -I;
    if (x == 0) {
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < dim; ++i) {
            const double sg = (__builtin_popcountll(i & z) & 1) ? -1.0 : 1.0;
            s[i] = (c - I * sn * sg) * s[i];
        }
        return;
    }
    const int hb = 63 - __builtin_clzll(x);
#pragma omp parallel for schedule(static)
    for (uint64_t p = 0; p < dim / 2; ++p) {
        const uint64_t i = insert0(p, hb), j = i ^ x;
        const double si = (__builtin_popcountll(i & z) & 1) ? -1.0 : 1.0;
        const double sj = (__builtin_popcountll(j & z) & 1) ? -1.0 : 1.0;
        const cplx a = s[i], b = s[j];
        s[i] = c * a - I * sn * g * si * b;
        s[j] = c * b - I * sn * g * sj * a;
    }
}

double oc_expectation_masks(const double* state, int n, int64_t nterms, const uint64_t* xs, const uint64_t* zs,
                            const double* coeffs) {
    const cplx* s = (const cplx*)state;
    const uint64_t dim = 1ull << n;
    double total = 0.0;
    for (int64_t k = 0; k < nterms; ++k) {
        const uint64_t x = xs[k], z = zs[k];
        const int ny = __builtin_popcountll(x & z);
        cplx g = 1;
        for (int q = 0; q < (ny & 3); ++q) g *= -I;
        double re = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : re)
        for (uint64_t i = 0; i < dim; ++i) {
            const double sg = (__builtin_popcountll(i & z) & 1) ? -1.0 : 1.0;
            re += creal(conj(s[i]) * g * sg * s[i ^ x]);
        }
        total += coeffs[k] * re;
    }
    return total;
}

/* per-term values Re<psi|P_k|psi> (the contract of qc_expectation's per_term output) */
void oc_expectation_masks_per_term(const double* state, int n, int64_t nterms, const uint64_t* xs, const uint64_t* zs,
                                   double* out) {
    const cplx* s = (const cplx*)state;
    const uint64_t dim = 1ull << n;
    for (int64_t k = 0; k < nterms; ++k) {
        const uint64_t x = xs[k], z = zs[k];
        const int ny = __builtin_popcountll(x & z);
        cplx g = 1;
        for (int q = 0; q < (ny & 3); ++q) g *= -I;
        double re = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : re)
        for (uint64_t i = 0; i < dim; ++i) {
            const double sg = (__builtin_popcountll(i & z) & 1) ? -1.0 : 1.0;
            re += creal(conj(s[i]) * g * sg * s[i ^ x]);
        }
        out[k] = re;
    }
}
