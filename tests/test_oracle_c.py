"""The C restatement of the oracle agrees with the numpy oracle (which is pinned to the reference's goldens)."""

import numpy as np

from oracle import oracle as O
from oracle import oracle_c as OC
from tests import fixtures as F


def _rand(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return v / np.linalg.norm(v)


def test_gate_programs_agree():
    n = 7
    gates = O.hf_gates(n, 3) + O.ucc_gates([0, 1, 4, 6], 0.3) + O.ucc_gates([2, 5], -0.4) + [("Y", (3,), None), ("Z", (0,), None)]
    a = O.run_gates(n, gates)
    b = np.zeros(2**n, dtype=complex)
    b[0] = 1
    OC.run_gates(b, n, gates)
    assert np.abs(a - b).max() < 1e-14


def test_expectation_paths_agree():
    n = 6
    psi = _rand(n, 1)
    rng = np.random.default_rng(2)
    terms = []
    for _ in range(25):
        qs = sorted(rng.choice(n, size=rng.integers(1, 5), replace=False).tolist())
        terms.append((float(rng.normal()), [(q, "XYZ"[rng.integers(0, 3)]) for q in qs]))
    ref = O.expectation_terms_gatepath(psi, n, terms)
    assert abs(OC.expectation_terms(psi.copy(), n, terms) - ref) < 1e-13
    xs, zs = zip(*[O.pauli_masks(n, " ".join(f"{p}{q}" for q, p in f)) for _, f in terms])
    assert abs(OC.expectation_masks(psi, n, xs, zs, [c for c, _ in terms]) - ref) < 1e-13
    assert abs(O.expectation_masks(psi, xs, zs, [c for c, _ in terms]) - ref) < 1e-13


def test_mask_rotation_agrees():
    n = 8
    psi = _rand(n, 3)
    rng = np.random.default_rng(4)
    ref = psi.copy()
    got = psi.copy()
    for _ in range(10):
        x, z, phi = int(rng.integers(0, 2**n)), int(rng.integers(0, 2**n)), float(rng.uniform(-3, 3))
        ref = O.apply_pauli_rotation(ref, x, z, phi)
        OC.pauli_rotation(got, n, x, z, phi)
    assert np.abs(ref - got).max() < 1e-13


def test_h2_known_answers_through_c_oracle():
    psi = np.zeros(16, dtype=complex)
    psi[0] = 1
    OC.run_gates(psi, 4, O.hf_gates(4, 2))
    e = F.H2_JW_TERMS[()] + OC.expectation_terms(psi, 4, F.term_list(F.H2_JW_TERMS))
    assert abs(e - F.H2_HF_ENERGY) < 1e-13
    assert OC.num_threads() >= 1
