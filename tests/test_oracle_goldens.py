"""Pins the CPU oracle against every golden the reference holds for the hot path. CPU only."""

import numpy as np
import pytest

from oracle import oracle as O
from tests import fixtures as F


def test_h2_hf_and_vacuum_energy_gate_path():
    tl = F.term_list(F.H2_JW_TERMS)
    const = F.H2_JW_TERMS[()]
    hf = O.run_gates(4, O.hf_gates(4, 2))
    assert np.nonzero(hf)[0].tolist() == [0b1100]  # qubit 0 = most significant bit
    assert O.expectation_terms_gatepath(hf, 4, tl, const) == pytest.approx(F.H2_HF_ENERGY, abs=1e-13)
    vac = O.run_gates(4, [])
    assert O.expectation_terms_gatepath(vac, 4, tl, const) == pytest.approx(F.H2_VACUUM_ENERGY, abs=1e-13)


def test_mask_expectation_equals_gate_path():
    rng = np.random.default_rng(0)
    psi = rng.normal(size=16) + 1j * rng.normal(size=16)
    psi /= np.linalg.norm(psi)
    const, xs, zs, cs = F.masks_from_terms(4, F.H2_JW_TERMS)
    a = O.expectation_masks(psi, xs, zs, cs, const)
    b = O.expectation_terms_gatepath(psi, 4, F.term_list(F.H2_JW_TERMS), const)
    assert a == pytest.approx(b, abs=1e-13)


def test_doc_circuit_expectations():
    """measurement.rst:325 (-0.0171209) and :200-205 (XIZ 0.02847, IYZ -0.19465)."""
    st = O.run_gates(2, F.doc_circuit_gates(2))
    assert O.expectation_terms_gatepath(st, 2, F.BK_H2_TERMS, F.BK_H2_CONST) == pytest.approx(-0.0171209, abs=5e-8)
    st = O.run_gates(3, F.doc_circuit_gates(3))
    assert O.expectation_terms_gatepath(st, 3, [(1.0, [(0, "X"), (2, "Z")])]) == pytest.approx(0.02847, abs=5e-6)
    assert O.expectation_terms_gatepath(st, 3, [(1.0, [(1, "Y"), (2, "Z")])]) == pytest.approx(-0.19465, abs=5e-6)


@pytest.mark.parametrize("string", ["Z1", "Z0 Z1", "X0 X1", "Y0 Y1", "X0 Z1 Y2 X3", "Y0 Y1 Y2 X3", "Z0 X3", "Y2"])
def test_expi_pauli_gate_program_equals_mask_rotation(string):
    """tests/test_ansatz.py:137-164 analogue: the gate program of _expi_pauli(theta) == exp(-i phi P/2), phi=-2 theta."""
    n = 4
    rng = np.random.default_rng(len(string))
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)
    theta = 0.37
    a = O.run_gates(n, O.expi_pauli_gates(string, theta), psi)
    x, z = O.pauli_masks(n, string)
    b = O.apply_pauli_rotation(psi, x, z, -2 * theta)
    assert np.abs(a - b).max() < 1e-14
    # and equals the matrix exponential definition exp(i theta P)
    P = np.array([[1.0]], dtype=complex)
    ops = {int(o[1:]): o[0] for o in string.split()}
    for q in range(n):
        P = np.kron(P, O.gate_matrix(ops.get(q, "I")))
    c = np.cos(theta) * psi + 1j * np.sin(theta) * (P @ psi)
    assert np.abs(a - c).max() < 1e-14


def test_ucc_strings_order_and_signs():
    """tests/test_ansatz.py:170-186, examples/ucc_example1.py:17, doc/source/tutorials/ansatz.rst:127-140."""
    doubles = O.ucc_pauli_terms([0, 1, 2, 3])
    assert [s for s, _ in doubles] == [
        "X0 X1 Y2 X3", "Y0 Y1 Y2 X3", "Y0 X1 X2 X3", "X0 Y1 X2 X3",
        "Y0 X1 Y2 Y3", "X0 Y1 Y2 Y3", "X0 X1 X2 Y3", "Y0 Y1 X2 Y3",
    ]  # fmt: skip
    rz_mult = [(2j * c).real for _, c in doubles]  # RZ angle = -2 * (-1j * coeff * theta)
    assert rz_mult == pytest.approx([-0.25, 0.25, 0.25, 0.25, -0.25, -0.25, -0.25, 0.25])
    singles = O.ucc_pauli_terms([0, 2])
    assert [s for s, _ in singles] == ["Y0 Z1 X2", "X0 Z1 Y2"]
    assert [(-1j * c).real for _, c in singles] == pytest.approx([0.5, -0.5])  # tests/test_ansatz.py:170
    assert [(2j * c).real for _, c in singles] == pytest.approx([-1.0, 1.0])  # examples/ucc_example1.py:17


def test_uccd_reaches_fci_of_fixture():
    const, xs, zs, cs = F.masks_from_terms(4, F.H2_JW_TERMS)
    best = min(
        O.expectation_masks(O.run_gates(4, O.hf_gates(4, 2) + O.ucc_gates([0, 1, 2, 3], th)), xs, zs, cs, const)
        for th in np.linspace(0.113, 0.116, 61)
    )
    assert best == pytest.approx(F.H2_FCI_ENERGY, abs=1e-8)


def test_gate_census_matches_adaptive_rst():
    """doc/source/tutorials/adaptive.rst:21-30: 8-qubit LiH UCCSD = 3300 gates, rz 160, cx 1312, h 1216, sdg/s 304, x 4.
    The excitation list is the one printed at examples/ucc_example1.py:64 scaled to (n=8, ne=4); here we count the
    gates of every JW single/double with the same occupied/virtual split."""
    import json
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "excitations.json")
    exc = json.load(open(path))["circuit_ansatz_order"]["8_4"]
    gates = O.hf_gates(8, 4)
    for e in exc:
        gates += O.ucc_gates(e, 0.1)
    names = [g[0] for g in gates]
    census = {k: names.count(k) for k in set(names)}
    assert len(gates) == 3300
    assert census == {"RZ": 160, "CNOT": 1312, "H": 1216, "SDG": 304, "S": 304, "X": 4}


@pytest.mark.parametrize(
    "term_qubits,freq,qubit_map,expected",
    [
        ([0], {"10": 5}, [0, 1], -1.0),
        ([2], {"010": 5}, [0, 2, 5], -1.0),
        ([4], {"110": 5}, [0, 2, 4], 1.0),
        ([0, 1], {"11": 5}, [0, 1], 1.0),
    ],
)
def test_frequency_goldens(term_qubits, freq, qubit_map, expected):
    """tests/test_expectation_samples.py:22-26 (X/Y are replaced by Z before the call, result.py:46)."""
    assert O.z_expectation_from_frequencies(freq, qubit_map, term_qubits) == expected


def test_frequency_golden_sum():
    f = {"11": 5}
    total = O.z_expectation_from_frequencies(f, [0, 1], [0, 1]) + O.z_expectation_from_frequencies(f, [0, 1], [0])
    assert total == 0.0  # tests/test_expectation_samples.py:26


def test_parity_sums_equal_frequency_expectation():
    rng = np.random.default_rng(3)
    samples = rng.integers(0, 8, size=1000, dtype=np.uint64)
    vals, counts = O.frequencies_from_samples(samples)
    freq = {format(int(v), "03b"): int(c) for v, c in zip(vals, counts)}
    # qubits (0,1,2) <-> string positions 0,1,2 <-> index bits 2,1,0
    for qs in ([0], [1, 2], [0, 1, 2]):
        mask = sum(1 << (2 - q) for q in qs)
        s = O.parity_sums(samples, [mask])[0]
        assert s / 1000 == pytest.approx(O.z_expectation_from_frequencies(freq, [0, 1, 2], qs), abs=1e-15)


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    out = O.philox4x32(np.array([[0, 0, 0, 0]], dtype=np.uint32), (0, 0))[0]
    assert [hex(int(v)) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = O.philox4x32(np.array([[0xFFFFFFFF] * 4], dtype=np.uint32), (0xFFFFFFFF, 0xFFFFFFFF))[0]
    assert [hex(int(v)) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    out = O.philox4x32(np.array([[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], dtype=np.uint32), (0xA4093822, 0x299F31D0))[0]
    assert [hex(int(v)) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_expi_pauli_programs_equal_reference_gate_lists():
    """Gate programs recorded from the reference's own `_expi_pauli` (tests/golden/make_golden.py)."""
    import json
    import os

    progs = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "expi_pauli_programs.json")))
    rename = {"S_dagger": "SDG", "H_dagger": "H", "S": "S", "H": "H", "CNOT": "CNOT", "RZ": "RZ"}
    for string, prog in progs.items():
        mine = O.expi_pauli_gates(string, 0.25)
        assert len(mine) == len(prog)
        for (name, qubits, param), (rname, rqubits, rparams) in zip(mine, prog):
            assert name == rename[rname] and list(qubits) == rqubits
            if rparams:
                assert param == pytest.approx(rparams[0])
