"""
Expectation values of a Hamiltonian for a circuit: exact (state vector, K2) and from samples (K5 + K3).

Reference: ``src/qibochem/measurement/result.py`` -- ``_constant_term`` :24-35,
``_pauli_term_measurement_expectation`` :38-54, ``expectation_from_samples`` :57-118 -- plus the qibo methods they
call.  Differences in mechanism (not in result): the reference re-simulates the whole circuit once per measurement
group (:106-113) and reduces Counter dictionaries in Python; here the state is prepared once, each group is a
basis-change + sampling pass on the device, and the per-term statistics are exact integer parity sums.
"""

from __future__ import annotations

from collections import Counter

import numpy as np

from qibochem_b200.measurement.optimization import _measurement_basis_rotations
from qibochem_b200.measurement.shot_allocation import allocate_shots, allocate_shots_by_variance
from qibochem_b200.pauli import PauliSum
from qibochem_b200.runtime import get_runtime


def expectation(circuit, hamiltonian) -> float:
    """Exact ``<psi|H|psi>`` with ``psi = circuit|0...0>`` (the function spelling of ``hamiltonian.expectation(circuit)``)."""
    return hamiltonian.expectation(circuit)


def _constant_term(hamiltonian) -> float:
    """Identity coefficient of the Hamiltonian (result.py:24-35)."""
    return hamiltonian.constant


def _frequencies_to_arrays(frequencies):
    """Counter {"0110": 5, ...} -> (keys as integers read MSB-first, counts, total shots)."""
    keys = np.array([int(k, 2) if isinstance(k, str) else int(k) for k in frequencies.keys()], dtype=np.uint64)
    counts = np.array(list(frequencies.values()), dtype=np.int64)
    return keys, counts, int(counts.sum())


def _z_terms_from_frequencies(keys: np.ndarray, counts: np.ndarray, masks: np.ndarray) -> np.ndarray:
    """Exact integer sums  S_j = sum_keys count * (-1)^{popcount(key & mask_j)}  on the device (K3)."""
    if keys.size == 0 or masks.size == 0:
        return np.zeros(masks.size, dtype=np.int64)
    st = get_runtime().context_state()  # K3 only needs a device context
    return st.parity_sums(keys, masks, counts=counts)


def _pauli_term_measurement_expectation(expression: PauliSum, frequencies: Counter, qubit_map) -> float:
    """Expectation of an expression of Pauli terms from frequencies measured in the terms' own bases: X/Y are read
    as Z (result.py:46); character k of a key is the outcome of qubit ``qubit_map[k]``
    (goldens: ``tests/test_expectation_samples.py:22-26`` of the reference)."""
    qubit_map = list(qubit_map)
    keys, counts, total = _frequencies_to_arrays(frequencies)
    if total == 0:
        return 0.0
    nchar = len(next(iter(frequencies)))
    coeffs, masks = [], []
    const = 0.0
    for (x, z), c in expression.terms.items():
        support = x | z
        if support == 0:
            const += complex(c).real
            continue
        m = 0
        q = 0
        while support >> q:
            if (support >> q) & 1:
                m |= 1 << (nchar - 1 - qubit_map.index(q))
            q += 1
        coeffs.append(complex(c).real)
        masks.append(m)
    sums = _z_terms_from_frequencies(keys, counts, np.array(masks, dtype=np.uint64))
    return float(const + np.dot(np.array(coeffs), sums / total))


def expectation_from_samples(
    circuit,
    hamiltonian,
    n_shots: int = 1000,
    grouping: str | None = None,
    n_shots_per_pauli_term: bool = True,
    shot_allocation: list[int] | None = None,
    seed: int | None = None,
) -> float:
    """Expectation value of ``hamiltonian`` from sample measurements of ``circuit``.

    Args:
        circuit: ansatz circuit
        hamiltonian: ``SymbolicHamiltonian``
        n_shots: shots per group of terms (or the total budget if ``n_shots_per_pauli_term`` is False)
        grouping: ``None`` (every term measured separately) or ``"qwc"``
        n_shots_per_pauli_term: use ``n_shots`` for every group
        shot_allocation: shots per group when ``n_shots_per_pauli_term`` is False
        seed: Philox key of the device sampler (extra to the reference's signature; default: fresh entropy)
    """
    from qibochem_b200.circuit import _next_seed

    grouped_terms, device_plan = _sampling_plan(hamiltonian, grouping, circuit.nqubits)
    if not n_shots_per_pauli_term:
        if shot_allocation is None:
            shot_allocation = allocate_shots(grouped_terms, n_shots)
        if len(shot_allocation) != len(grouped_terms):
            raise ValueError(f"{len(shot_allocation) = } doesn't match the number of grouped terms ({len(grouped_terms)})!")

    total = _constant_term(hamiltonian)
    if not grouped_terms:
        return total
    state = circuit.run_on_device()  # psi is prepared ONCE; the reference re-runs the circuit per group
    base_seed = _next_seed() if seed is None else int(seed)
    for i, (xb, yb, coeffs, masks) in enumerate(device_plan):
        nshots = n_shots if n_shots_per_pauli_term else int(shot_allocation[i])
        if not nshots:
            continue
        samples = state.sample(xb, yb, nshots, (base_seed + 0x9E3779B97F4A7C15 * (i + 1)) % (1 << 64), on_device=True)
        sums = state.parity_sums(samples, masks)
        total += float(np.dot(coeffs, sums / nshots))
    return total


def _sampling_plan(hamiltonian, grouping, n: int):
    """(grouped terms, per group: X-basis mask, Y-basis mask, coefficients, support masks in index-bit positions).
    Grouping and mask construction depend only on the Hamiltonian, and an optimiser loop samples the same Hamiltonian
    thousands of times: both are cached on the Hamiltonian object (per grouping and register size)."""
    cache = hamiltonian.__dict__.setdefault("_sampling_plans", {})
    key = (grouping, int(n))
    if key not in cache:
        grouped_terms = _measurement_basis_rotations(hamiltonian, grouping=grouping)
        plan = []
        for expression, measurement_gates, _rotation_gates in grouped_terms:
            xb = sum(1 << (n - 1 - g.qubits[0]) for g in measurement_gates if g.basis[0] == "X")
            yb = sum(1 << (n - 1 - g.qubits[0]) for g in measurement_gates if g.basis[0] == "Y")
            coeffs, masks = [], []
            for (x, z), c in expression.terms.items():
                support = x | z
                m, q = 0, 0
                while support >> q:
                    if (support >> q) & 1:
                        m |= 1 << (n - 1 - q)
                    q += 1
                coeffs.append(complex(c).real)
                masks.append(m)
            plan.append((xb, yb, np.array(coeffs, dtype=np.float64), np.array(masks, dtype=np.uint64)))
        cache[key] = (grouped_terms, plan)
    return cache[key]


def _group_masks(expression: PauliSum, n: int):
    """(coefficients, support masks in index-bit positions, constant) of a group expression."""
    coeffs, masks, const = [], [], 0.0
    for (x, z), c in expression.terms.items():
        support = x | z
        if support == 0:
            const += complex(c).real
            continue
        m, q = 0, 0
        while support >> q:
            if (support >> q) & 1:
                m |= 1 << (n - 1 - q)
            q += 1
        coeffs.append(complex(c).real)
        masks.append(m)
    return np.array(coeffs, dtype=np.float64), np.array(masks, dtype=np.uint64), const


def sample_statistics(circuit, grouped_terms, n_shots: int = 1000, seed: int | None = None):
    """Sample mean and sample variance of every term group from ``n_shots`` shots per group
    (reference: ``src/qibochem/measurement/result.py:121-158``).

    The reference's variance sums ``(value(outcome) - mean)^2`` over the DISTINCT measured outcomes -- every outcome
    once, whatever its count -- and divides by ``n_shots - 1`` (result.py:150-155; its goldens,
    ``tests/test_expectation_samples.py:139-155``, pin exactly that).  The same number is returned here; the state is
    prepared once, and the per-outcome values and squared deviations are reduced on the device (K3).
    """
    from qibochem_b200.circuit import _next_seed

    n = circuit.nqubits
    state = circuit.run_on_device()
    base_seed = _next_seed() if seed is None else int(seed)
    means, variances = [], []
    for i, (expression, measurement_gates, _rotation_gates) in enumerate(grouped_terms):
        xb = sum(1 << (n - 1 - g.qubits[0]) for g in measurement_gates if g.basis[0] == "X")
        yb = sum(1 << (n - 1 - g.qubits[0]) for g in measurement_gates if g.basis[0] == "Y")
        keep = sum(1 << (n - 1 - g.qubits[0]) for g in measurement_gates)
        samples = state.sample(xb, yb, n_shots, (base_seed + 0x9E3779B97F4A7C15 * (i + 1)) % (1 << 64), on_device=True)
        coeffs, masks, const = _group_masks(expression, n)
        sums = state.parity_sums(samples, masks)  # exact integers
        mean = float(const + np.dot(coeffs, sums / n_shots))
        # distinct outcomes over the measured qubits; keys come back compacted, so expand them to index-bit positions
        keys, counts = state.frequencies(samples, keep)
        kept_bits = [b for b in range(n) if (keep >> b) & 1]
        full = np.zeros(keys.size, dtype=np.uint64)
        for j, b in enumerate(kept_bits):
            full |= ((keys >> np.uint64(j)) & np.uint64(1)) << np.uint64(b)
        unweighted, _weighted = state.outcome_variance(full, counts, masks, coeffs, mean - const)
        means.append(mean)
        variances.append(unweighted / (n_shots - 1))
    return means, variances


def v_expectation(circuit, hamiltonian, n_shots: int, n_trial_shots: int, grouping: str | None = None, method: str = "vmsa") -> float:
    """Expectation value with variance-based shot allocation (VMSA / VPSR; reference ``result.py:161-220``):
    ``n_trial_shots`` per group give the sample variances, the remaining budget is split by
    ``allocate_shots_by_variance`` and both estimates are combined weighted by their shot numbers."""
    from qibochem_b200.hamiltonian import SymbolicHamiltonian

    assert method in ("vmsa", "vpsr"), f"Unknown shot assignment method ({method}) called"
    grouped_terms = _measurement_basis_rotations(hamiltonian, grouping=grouping)
    assert (
        n_trial_shots * len(grouped_terms) <= n_shots
    ), f"n(Trial shots = {n_trial_shots}) * n(Term groups = {len(grouped_terms)}) > n(Total shots = {n_shots})"
    sample_means, sample_variances = sample_statistics(circuit, grouped_terms, n_shots=n_trial_shots)
    remaining = allocate_shots_by_variance(n_shots, n_trial_shots, sample_variances, method=method)
    total = 0.0
    for (expression, _, _), mean0, extra in zip(grouped_terms, sample_means, remaining):
        new_mean = 0.0
        if extra:
            sub = SymbolicHamiltonian(expression, nqubits=hamiltonian.nqubits)
            new_mean = expectation_from_samples(circuit, sub, n_shots=extra, grouping=grouping)
        total += (n_trial_shots * mean0 + extra * new_mean) / (n_trial_shots + extra)
    return total + _constant_term(hamiltonian)
