"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the arithmetic that the qibochem hot path runs inside qibo's numpy backend.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  Nothing under ``qibochem_b200/`` imports it; the product path fails loudly without CUDA.

Parity status
-------------
qibo 0.3.2 (``/root/reference/poetry.lock:2761``), openfermion 1.7.1 (``poetry.lock:1907``) and pyscf are NOT
installed in this image, so the third-party half of the path is a *restatement* of the published algorithm
("[recalled]" below), pinned against every golden the reference holds for the path (``tests/test_oracle_goldens.py``):
``doc/source/tutorials/hamiltonian.rst:59`` + ``ansatz.rst:73,80`` (H2 energies), ``measurement.rst:155-208`` and
``:290-325`` (RX/RZ/CNOT circuits), ``tests/test_expectation_samples.py:22-26`` (integer frequency goldens),
``tests/test_ansatz.py:170-186`` (UCC strings/signs), ``doc/source/tutorials/adaptive.rst:21-30`` (gate census).
``hamiltonian.expectation(circuit)`` comparisons that the reference computes at test time with qibo are
"parity unpinned" beyond those goldens.

Conventions: qubit 0 is the MOST significant index bit, ``RZ(phi)=diag(e^{-i phi/2}, e^{+i phi/2})``.
"""

from __future__ import annotations

from collections import Counter

import numpy as np

# --------------------------------------------------------------------------------------------------------------
# Gate level: what qibo's NumpyBackend.apply_gate does for the gate set `_expi_pauli` emits   [recalled]
# --------------------------------------------------------------------------------------------------------------

_SQ2 = 1.0 / np.sqrt(2.0)
_FIXED = {
    "I": np.eye(2, dtype=complex),
    "X": np.array([[0, 1], [1, 0]], dtype=complex),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=complex),
    "Z": np.array([[1, 0], [0, -1]], dtype=complex),
    "H": np.array([[_SQ2, _SQ2], [_SQ2, -_SQ2]], dtype=complex),
    "S": np.array([[1, 0], [0, 1j]], dtype=complex),
    "SDG": np.array([[1, 0], [0, -1j]], dtype=complex),
    "CNOT": np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex),
    "CZ": np.diag([1, 1, 1, -1]).astype(complex),
    "SWAP": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex),
}


def gate_matrix(name: str, param=None) -> np.ndarray:
    """Matrix of a qibo gate.  RX/RY/RZ follow qibo's definitions [recalled], validated in tests against
    ``measurement.rst:325`` (-0.0171209) and ``measurement.rst:200-205`` (0.02847 / -0.19465)."""
    if name in _FIXED:
        return _FIXED[name]
    c, s = np.cos(param / 2.0), np.sin(param / 2.0)
    if name == "RZ":
        return np.array([[np.exp(-0.5j * param), 0], [0, np.exp(0.5j * param)]], dtype=complex)
    if name == "RX":
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)
    if name == "RY":
        return np.array([[c, -s], [s, c]], dtype=complex)
    raise KeyError(name)


def apply_gate(state: np.ndarray, n: int, matrix: np.ndarray, qubits) -> np.ndarray:
    """Out-of-place application of a k-qubit matrix on `qubits` (qubit 0 = leading tensor axis), the way the numpy
    backend does it with reshape + einsum [recalled].  For CNOT, qubits = (control, target)."""
    k = len(qubits)
    psi = state.reshape((2,) * n)
    m = matrix.reshape((2,) * (2 * k))
    out = np.tensordot(m, psi, axes=(list(range(k, 2 * k)), list(qubits)))
    out = np.moveaxis(out, list(range(k)), list(qubits))
    return np.ascontiguousarray(out).reshape(-1)


def expi_pauli_gates(pauli_string: str, theta) -> list[tuple]:
    """Gate program for exp(i*theta*P) exactly as built by ``src/qibochem/ansatz/_ansatz.py:55-92``:
    sort ops by qubit (:69); Y -> S^dagger then H, X -> H (:73-78); CNOT ladder from the last op to the first (:85);
    RZ(first qubit, -2 theta) (:87); un-ladder (:89); reversed daggered basis changes (:91).
    Returns a list of (name, qubits, param)."""
    ops = sorted(((int(o[1:]), o[0]) for o in pauli_string.split()), key=lambda t: t[0])
    basis = []
    for q, p in ops:
        if p == "Y":
            basis.append(("SDG", (q,), None))
        if p not in ("I", "Z"):
            basis.append(("H", (q,), None))
    gates = list(basis)
    m = len(ops)
    gates += [("CNOT", (ops[i][0], ops[i - 1][0]), None) for i in range(m - 1, 0, -1)]
    gates.append(("RZ", (ops[0][0],), -2.0 * theta))
    gates += [("CNOT", (ops[i + 1][0], ops[i][0]), None) for i in range(m - 1)]
    dag = {"SDG": "S", "S": "SDG", "H": "H"}
    gates += [(dag[g[0]], g[1], None) for g in reversed(basis)]
    return gates


def run_gates(n: int, gates, state: np.ndarray | None = None) -> np.ndarray:
    """``Circuit.__call__`` on the numpy backend: start from |0..0> and apply the queue in order [recalled]."""
    if state is None:
        state = np.zeros(2**n, dtype=complex)
        state[0] = 1.0
    for name, qubits, param in gates:
        p = param.real if isinstance(param, complex) else param
        state = apply_gate(state, n, gate_matrix(name, p), qubits)
    return state


def hf_gates(n: int, nelectrons: int) -> list[tuple]:
    """``hf_circuit`` (JW): X on qubits 0..ne-1 (``src/qibochem/ansatz/ansatz.py:274-282``)."""
    return [("X", (q,), None) for q in range(nelectrons)]


# --------------------------------------------------------------------------------------------------------------
# Excitation -> Pauli strings (what openfermion.jordan_wigner(T - T^dagger) returns; ansatz.py:311-349)  [recalled]
# --------------------------------------------------------------------------------------------------------------

_PAULI_MUL = {
    ("I", "I"): (1, "I"), ("I", "X"): (1, "X"), ("I", "Y"): (1, "Y"), ("I", "Z"): (1, "Z"),
    ("X", "I"): (1, "X"), ("X", "X"): (1, "I"), ("X", "Y"): (1j, "Z"), ("X", "Z"): (-1j, "Y"),
    ("Y", "I"): (1, "Y"), ("Y", "X"): (-1j, "Z"), ("Y", "Y"): (1, "I"), ("Y", "Z"): (1j, "X"),
    ("Z", "I"): (1, "Z"), ("Z", "X"): (1j, "Y"), ("Z", "Y"): (-1j, "X"), ("Z", "Z"): (1, "I"),
}  # fmt: skip


def _qmul(a: dict, b: dict) -> dict:
    """Product of two Pauli sums stored as {((q, 'X'), ...): coeff}; insertion order is left-major like
    openfermion's SymbolicOperator.__imul__ [recalled]."""
    out: dict = {}
    for ta, ca in a.items():
        for tb, cb in b.items():
            coeff = ca * cb
            acc = dict(ta)
            for q, p in tb:
                ph, r = _PAULI_MUL[(acc.get(q, "I"), p)]
                coeff *= ph
                acc[q] = r
            key = tuple(sorted((q, p) for q, p in acc.items() if p != "I"))
            out[key] = out.get(key, 0.0) + coeff
    return out


def jw_ladder(p: int, dagger: bool) -> dict:
    """a_p^(dagger) = Z_0..Z_{p-1} (X_p -/+ iY_p)/2."""
    zs = tuple((q, "Z") for q in range(p))
    return {zs + ((p, "X"),): 0.5, zs + ((p, "Y"),): (-0.5j if dagger else 0.5j)}


def ucc_pauli_terms(excitation) -> list[tuple[str, complex]]:
    """JW image of T - T^dagger for one excitation, in openfermion's dict order [recalled]; pinned by the order and
    signs in ``tests/test_ansatz.py:170-186`` and the drawing at ``doc/source/tutorials/ansatz.rst:127-140``.
    Returns [("X0 X1 Y2 X3", coeff), ...] with purely imaginary coefficients."""
    orbs = sorted(excitation, reverse=True)
    half = len(orbs) // 2
    t_ops = [(o, True) for o in orbs[:half]] + [(o, False) for o in orbs[half:]]
    tdag_ops = [(o, not d) for o, d in reversed(t_ops)]
    total: dict = {}
    for ops, sign in ((t_ops, 1.0), (tdag_ops, -1.0)):
        term = {(): sign}
        for o, d in ops:
            term = _qmul(term, jw_ladder(o, d))
        for key, c in term.items():
            if key in total:
                total[key] += c
                if abs(total[key]) < 1e-8:  # openfermion drops terms that cancel (EQ_TOLERANCE) [recalled]
                    del total[key]
            else:
                total[key] = c
    return [(" ".join(f"{p}{q}" for q, p in key), c) for key, c in total.items()]


def ucc_gates(excitation, theta: float, trotter_steps: int = 1) -> list[tuple]:
    """``ucc_circuit`` gate list (``ansatz.py:340-349``)."""
    gates = []
    for _ in range(trotter_steps):
        for string, coeff in ucc_pauli_terms(excitation):
            gates += expi_pauli_gates(string, (-1.0j * coeff * theta / trotter_steps).real)
    return gates


# --------------------------------------------------------------------------------------------------------------
# Mask level: the same unitary / expectation in one pass (the contract of kernels K1 / K2)
# --------------------------------------------------------------------------------------------------------------


def pauli_masks(n: int, pauli_string: str) -> tuple[int, int]:
    """(xmask, zmask) in index-bit positions: qubit q <-> bit n-1-q.  x = X|Y sites, z = Z|Y sites."""
    x = z = 0
    for op in pauli_string.split():
        p, q = op[0], int(op[1:])
        b = 1 << (n - 1 - q)
        if p in ("X", "Y"):
            x |= b
        if p in ("Z", "Y"):
            z |= b
    return x, z


def _parity_sign(idx: np.ndarray, mask: int) -> np.ndarray:
    v = idx & np.uint64(mask)
    par = np.zeros(idx.shape, dtype=np.uint64)
    while True:
        par ^= v & np.uint64(1)
        v = v >> np.uint64(1)
        if not v.any():
            break
    return 1.0 - 2.0 * par.astype(np.float64)


def apply_pauli(state: np.ndarray, x: int, z: int) -> np.ndarray:
    """(P psi)_i = (-i)^{nY} (-1)^{popcount(i & z)} psi_{i ^ x}   (SURVEY.md 3.6)."""
    idx = np.arange(state.size, dtype=np.uint64)
    ny = bin(x & z).count("1")
    return ((-1j) ** ny) * _parity_sign(idx, z) * state[idx ^ np.uint64(x)]


def apply_pauli_rotation(state: np.ndarray, x: int, z: int, phi: float) -> np.ndarray:
    """exp(-i phi P / 2) psi = cos(phi/2) psi - i sin(phi/2) P psi."""
    return np.cos(phi / 2.0) * state - 1j * np.sin(phi / 2.0) * apply_pauli(state, x, z)


def expectation_masks(state: np.ndarray, xs, zs, coeffs, constant: float = 0.0) -> float:
    """const + sum_k c_k Re<psi|P_k|psi>."""
    total = constant * np.vdot(state, state).real
    for x, z, c in zip(xs, zs, coeffs):
        total += c * np.vdot(state, apply_pauli(state, int(x), int(z))).real
    return float(total)


def expectation_terms_gatepath(state: np.ndarray, n: int, terms, constant: float = 0.0) -> float:
    """qibo's ``SymbolicHamiltonian.expectation`` on a state [recalled]: H|psi> is accumulated term by term (copy
    the state, apply each factor's 2x2 matrix with the backend, scale by the coefficient) and the result is
    Re<psi|H psi>.  `terms` = [(coeff, [(qubit, 'X'|'Y'|'Z'), ...]), ...]."""
    hpsi = constant * state
    for coeff, factors in terms:
        tmp = state.copy()
        for q, p in factors:
            tmp = apply_gate(tmp, n, _FIXED[p], (q,))
        hpsi = hpsi + coeff * tmp
    return float(np.vdot(state, hpsi).real)


# --------------------------------------------------------------------------------------------------------------
# Sampled path (result.py:57-118 + qibo's sampling / expectation_from_samples)  [recalled]
# --------------------------------------------------------------------------------------------------------------


def measurement_basis_gates(qubit_basis: dict[int, str]) -> list[tuple]:
    """Rotation that ``gates.M(q, basis=...)`` inserts before a Z measurement [recalled]: X -> H, Y -> S^dagger, H."""
    gates = []
    for q in sorted(qubit_basis):
        if qubit_basis[q] == "Y":
            gates.append(("SDG", (q,), None))
        if qubit_basis[q] in ("X", "Y"):
            gates.append(("H", (q,), None))
    return gates


def probabilities(state: np.ndarray, n: int, measured) -> np.ndarray:
    """|psi|^2 summed over the unmeasured axes; remaining axes in the order given by `measured` [recalled]."""
    p = (np.abs(state) ** 2).reshape((2,) * n)
    rest = tuple(q for q in range(n) if q not in measured)
    p = p.sum(axis=rest) if rest else p
    kept = [q for q in range(n) if q in measured]
    p = np.transpose(p, [kept.index(q) for q in measured])
    return p.reshape(-1)


def sample_frequencies(state: np.ndarray, n: int, measured, nshots: int, rng) -> Counter:
    """Shots drawn from the marginal; ``frequencies(binary=True)``: keys are bit strings over the measured qubits
    in gate order [recalled]."""
    p = probabilities(state, n, list(measured))
    p = p / p.sum()
    shots = rng.choice(p.size, size=nshots, p=p)
    m = len(measured)
    vals, counts = np.unique(shots, return_counts=True)
    return Counter({format(int(v), f"0{m}b"): int(c) for v, c in zip(vals, counts)})


def z_expectation_from_frequencies(freq, qubit_map, term_qubits, coeff: float = 1.0) -> float:
    """qibo's ``SymbolicHamiltonian.expectation_from_samples(freq, qubit_map)`` for ONE Z-only term [recalled]:
    ``counts = values/sum(values)``; for each key the sign is (-1)^{# of '1' at positions qubit_map.index(q)}.
    Pinned by ``tests/test_expectation_samples.py:22-26``."""
    keys = list(freq.keys())
    counts = np.array(list(freq.values()), dtype=float) / sum(freq.values())
    qm = list(qubit_map)
    total = 0.0
    for key, cnt in zip(keys, counts):
        ones = sum(1 for q in term_qubits if key[qm.index(q)] == "1")
        total += (-1.0) ** ones * cnt
    return coeff * total


def parity_sums(samples: np.ndarray, masks) -> np.ndarray:
    """Integer contract of kernel K3: S_j = sum_shots (-1)^{popcount(s & m_j)} (int64, exact)."""
    samples = np.asarray(samples, dtype=np.uint64)
    out = np.zeros(len(masks), dtype=np.int64)
    for j, m in enumerate(masks):
        sgn = _parity_sign(samples, int(m))
        out[j] = int(sgn.sum())
    return out


def frequencies_from_samples(samples: np.ndarray):
    """Sorted unique keys + counts (contract of `qc_frequencies`)."""
    vals, counts = np.unique(np.asarray(samples, dtype=np.uint64), return_counts=True)
    return vals, counts.astype(np.int64)


# --------------------------------------------------------------------------------------------------------------
# Philox4x32-10 (counter-based RNG used by the device sampler; restated so tests can regenerate the uniforms)
# --------------------------------------------------------------------------------------------------------------


def philox4x32(counter: np.ndarray, key: tuple[int, int]) -> np.ndarray:
    """Philox4x32-10 of Salmon et al. (SC'11).  counter: uint32[N,4]; key: two uint32.  Returns uint32[N,4]."""
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    w0, w1 = 0x9E3779B9, 0xBB67AE85
    c = counter.astype(np.uint64)
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = m0 * c[:, 0]
        p1 = m1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[:, 1] ^ np.uint64(k0), lo1, hi0 ^ c[:, 3] ^ np.uint64(k1), lo0], axis=1)
        k0 = (k0 + w0) & 0xFFFFFFFF
        k1 = (k1 + w1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def philox_uniform53(nshots: int, seed: int, stream: int = 0, shot_offset: int = 0) -> np.ndarray:
    """One double in [0,1) per shot: counter = (shot_lo, shot_hi, stream, 0), key = (seed_lo, seed_hi);
    u = ((r0 << 21) ^ (r1 >> 11)) * 2^-53 using the first two output words (same recipe as the device sampler)."""
    idx = np.arange(nshots, dtype=np.uint64) + np.uint64(shot_offset)
    ctr = np.stack(
        [idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.full(nshots, stream, np.uint64), np.zeros(nshots, np.uint64)],
        axis=1,
    )
    r = philox4x32(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.uint64)
    bits = ((r[:, 0] << np.uint64(21)) ^ (r[:, 1] >> np.uint64(11))) & np.uint64((1 << 53) - 1)
    return bits.astype(np.float64) * (2.0**-53)
