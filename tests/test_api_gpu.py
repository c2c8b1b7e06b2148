"""GPU tests of the qibochem-facing API; they mirror the reference's own tests of the hot path
(tests/test_ansatz.py, tests/test_expectation_samples.py) with the CPU oracle standing in for qibo."""

import numpy as np
import pytest

from oracle import oracle as O
from tests import fixtures as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _lib(native_lib):
    return native_lib


def _api():
    import qibochem_b200 as qb
    from qibochem_b200 import Circuit, SymbolicHamiltonian, gates
    from qibochem_b200.ansatz import circuit_ansatz, hf_circuit, ucc_circuit
    from qibochem_b200.ansatz.ansatz import _expi_pauli
    from qibochem_b200.measurement import expectation, expectation_from_samples
    from qibochem_b200.measurement.result import _pauli_term_measurement_expectation
    from qibochem_b200.pauli import X, Y, Z

    return locals()


def gate_list_of(circuit):
    """qibochem_b200 gate queue -> oracle gate list."""
    name = {"x": "X", "y": "Y", "z": "Z", "h": "H", "s": "S", "sdg": "SDG", "cx": "CNOT", "cz": "CZ", "rx": "RX", "ry": "RY", "rz": "RZ"}
    out = []
    for g in circuit.queue:
        if g.name == "pauli_rotation":
            out += O.expi_pauli_gates(g.pauli_string, -0.5 * complex(g.parameters[0]).real)
        elif g.name != "measure":
            out.append((name[g.name], g.qubits, g.parameters[0] if g.parameters else None))
    return out


def doc_circuit(A, n, f=0.1):
    g = A["gates"]
    c = A["Circuit"](n)
    c.add(g.RX(i, f) for i in range(n))
    c.add(g.RZ(i, f) for i in range(n))
    c.add(g.CNOT(i, i + 1) for i in range(n - 1))
    c.add(g.RX(i, 2 * f) for i in range(n))
    c.add(g.RZ(i, 2 * f) for i in range(n))
    return c


def test_hf_circuit_energy_jw_and_bk():
    """tests/test_ansatz.py:115-134 analogue on the doc fixture geometry (ansatz.rst:73)."""
    A = _api()
    mol = F.h2_molecule()
    for mapping in (None, "bk"):
        ham = mol.hamiltonian(ferm_qubit_map=mapping)
        circuit = A["hf_circuit"](mol.nso, mol.nelec, ferm_qubit_map=mapping)
        assert ham.expectation(circuit) == pytest.approx(F.H2_HF_ENERGY, abs=1e-12)
    assert A["expectation"](A["hf_circuit"](4, 2), mol.hamiltonian()) == pytest.approx(F.H2_HF_ENERGY, abs=1e-12)
    assert mol.hamiltonian().expectation(A["Circuit"](4)) == pytest.approx(F.H2_VACUUM_ENERGY, abs=1e-12)


@pytest.mark.parametrize("pauli_string", ["Z1", "Z0 Z1", "X0 X1", "Y0 Y1"])
def test_expi_pauli(pauli_string):
    """tests/test_ansatz.py:137-164: state of _expi_pauli(theta) == exp(i theta P)|00> (oracle gate program)."""
    A = _api()
    theta = 0.4321
    got = A["_expi_pauli"](2, pauli_string, theta)().state()
    ref = O.run_gates(2, O.expi_pauli_gates(pauli_string, theta))
    assert np.abs(got - ref).max() < 1e-12


@pytest.mark.parametrize("excitation,mapping", [([0, 2], None), ([0, 1, 2, 3], None), ([0, 2], "bk")])
def test_ucc_circuit(excitation, mapping):
    """tests/test_ansatz.py:167-217, but from the HF state (the reference's vacuum start is degenerate)."""
    A = _api()
    theta, n = 0.1, 4
    kw = {} if mapping is None else {"ferm_qubit_map": mapping}
    circuit = A["hf_circuit"](n, 2, **kw) + A["ucc_circuit"](n, excitation, theta=theta, **kw)
    got = circuit().state()
    ref = O.run_gates(n, gate_list_of(circuit))
    assert np.abs(got - ref).max() < 1e-12
    if mapping is None:
        ref2 = O.run_gates(n, O.hf_gates(n, 2) + O.ucc_gates(excitation, theta))
        assert np.abs(got - ref2).max() < 1e-12


def test_general_gate_circuit_state_and_expectation():
    """doc/source/tutorials/measurement.rst:155-208 and :290-325 through the public API."""
    A = _api()
    X, Y, Z, SH = A["X"], A["Y"], A["Z"], A["SymbolicHamiltonian"]
    c3 = doc_circuit(A, 3)
    ref = O.run_gates(3, gate_list_of(c3))
    assert np.abs(c3().state() - ref).max() < 1e-12
    assert SH(X(0) * Z(2), nqubits=3).expectation(c3) == pytest.approx(0.02847, abs=5e-6)
    assert SH(Y(1) * Z(2), nqubits=3).expectation(c3) == pytest.approx(-0.19465, abs=5e-6)
    bk = SH(-0.4584 + 0.3593 * Z(0) - 0.4826 * Z(1) + 0.5818 * Z(0) * Z(1) + 0.0896 * X(0) * X(1) + 0.0896 * Y(0) * Y(1))
    assert bk.expectation(doc_circuit(A, 2)) == pytest.approx(-0.0171209, abs=5e-8)
    assert bk.expectation(doc_circuit(A, 2)) == pytest.approx(
        O.expectation_terms_gatepath(O.run_gates(2, F.doc_circuit_gates(2)), 2, F.BK_H2_TERMS, F.BK_H2_CONST), abs=1e-12
    )


def test_vqe_energy_callable_reaches_fci():
    """README.md:25-48 / adaptive.rst:95-106: minimise E(params) over the 8 RZ angles of HF+UCCD with scipy."""
    from scipy.optimize import minimize

    A = _api()
    mol = F.h2_molecule()
    ham = mol.hamiltonian()
    circuit = A["hf_circuit"](4, 2) + A["ucc_circuit"](4, [0, 1, 2, 3])

    def energy(params):
        circuit.set_parameters(params)
        return ham.expectation(circuit)

    res = minimize(energy, np.zeros(8), method="BFGS")
    assert res.fun == pytest.approx(F.H2_FCI_ENERGY, abs=1e-8)


@pytest.mark.parametrize("n_spatial,nelec,seed", [(4, 4, 1), (6, 4, 2)])
def test_uccsd_energy_molecule_shaped(n_spatial, nelec, seed):
    """C2-like configuration (LiH-shaped: 12 qubits, 4 electrons, 92 excitations / 640 rotations): circuit_ansatz +
    exact expectation vs the oracle's gate-by-gate path (reference gate program + per-term expectation)."""
    A = _api()
    mol = F.synthetic_molecule(n_spatial, nelec, seed)
    n = mol.nso
    ham = mol.hamiltonian()
    rng = np.random.default_rng(seed)
    from qibochem_b200.ansatz import generate_excitations

    exc = []
    for order in (2, 1):
        exc += generate_excitations(order, range(nelec), range(nelec, n))
    thetas = rng.uniform(-0.1, 0.1, size=len(exc))
    circuit = A["circuit_ansatz"](mol, "ucc", excitations=exc, thetas=thetas)
    e = ham.expectation(circuit)
    gates_ref = O.hf_gates(n, nelec)
    for ex, th in zip(exc, thetas):
        gates_ref += O.ucc_gates(ex, th)
    psi = O.run_gates(n, gates_ref)
    assert np.abs(circuit().state() - psi).max() < 1e-12
    x, z, c = ham.index_masks()
    assert e == pytest.approx(O.expectation_masks(psi, x, z, c, ham.constant), abs=1e-10)
    if n <= 8:
        assert e == pytest.approx(O.expectation_terms_gatepath(psi, n, ham.terms, ham.constant), abs=1e-10)


@pytest.mark.parametrize(
    "term,frequencies,qubit_map,expected",
    [
        ("X0", {"10": 5}, [0, 1], -1.0),
        ("X2", {"010": 5}, [0, 2, 5], -1.0),
        ("Y4", {"110": 5}, [0, 2, 4], 1.0),
        ("X0 Y1", {"11": 5}, [0, 1], 1.0),
        ("X0 Y1 + X0", {"11": 5}, [0, 1], 0.0),
    ],
)
def test_pauli_term_measurement_expectation(term, frequencies, qubit_map, expected):
    """tests/test_expectation_samples.py:19-31 -- exact goldens."""
    A = _api()
    from qibochem_b200.pauli import PauliSum

    expr = PauliSum()
    for t in term.split(" + "):
        expr = expr + PauliSum.from_string(t)
    assert A["_pauli_term_measurement_expectation"](expr, frequencies, qubit_map) == expected


def test_z_hamiltonian_expectation_from_samples_method():
    A = _api()
    Z, SH = A["Z"], A["SymbolicHamiltonian"]
    ham = SH(0.5 * Z(0) * Z(2) - 2.0 * Z(1) + 0.25, nqubits=3)
    freq = {"000": 3, "101": 2, "011": 5}
    ref = 0.25 + 0.5 * O.z_expectation_from_frequencies(freq, [0, 1, 2], [0, 2]) - 2.0 * O.z_expectation_from_frequencies(freq, [0, 1, 2], [1])
    assert ham.expectation_from_samples(freq) == pytest.approx(ref, abs=1e-15)
    # qubit_map permutes which character belongs to which qubit
    ref2 = 0.25 + 0.5 * O.z_expectation_from_frequencies(freq, [2, 0, 1], [0, 2]) - 2.0 * O.z_expectation_from_frequencies(freq, [2, 0, 1], [1])
    assert ham.expectation_from_samples(freq, qubit_map=[2, 0, 1]) == pytest.approx(ref2, abs=1e-15)


def test_expectation_from_samples_deterministic_states():
    """tests/test_expectation_samples.py:34-47."""
    A = _api()
    g, SH, X, Z = A["gates"], A["SymbolicHamiltonian"], A["X"], A["Z"]
    for terms, to_add in ((Z(0), [g.X(0)]), (Z(0) * Z(1), [g.X(0)]), (X(0), [g.H(0)])):
        ham = SH(terms, nqubits=2)
        c = A["Circuit"](2)
        c.add(to_add)
        assert A["expectation_from_samples"](c, ham) == pytest.approx(ham.expectation(c))


def test_expectation_manual_and_invalid_shot_allocation():
    """tests/test_expectation_samples.py:57-82."""
    A = _api()
    g, SH, X, Z = A["gates"], A["SymbolicHamiltonian"], A["X"], A["Z"]
    ham = SH(X(0) + Z(0))
    for to_add, alloc, expected in (([g.H(0)], [10, 0], 1.0), ([g.X(0), g.Z(0)], [0, 10], -1.0)):
        c = A["Circuit"](1)
        c.add(to_add)
        assert A["expectation_from_samples"](c, ham, n_shots_per_pauli_term=False, shot_allocation=alloc) == pytest.approx(expected)
    with pytest.raises(ValueError):
        A["expectation_from_samples"](A["Circuit"](1), SH(Z(0) + X(0)), n_shots_per_pauli_term=False, shot_allocation=(1,))


def test_measurement_grouping_functionality():
    """tests/test_expectation_samples.py:84-112: sampled == exact within abs 0.08 at 10^4 shots, None and qwc."""
    A = _api()
    g, SH, X, Y, Z = A["gates"], A["SymbolicHamiltonian"], A["X"], A["Y"], A["Z"]
    hams = [
        SH(Z(2)),
        SH(0.2 * X(0) + Y(2) + 13.0),
        SH(Z(0) + X(0) * Y(1) + Z(0) * Y(2)),
        SH(Y(0) + Z(1) + X(0) * Z(2)),
        SH(0.1 * X(0) * X(1) * Y(2) + 0.2 * X(0) * Y(1) * Y(2) + 0.3 * Y(0) * X(1) * X(2) - 3.14 * Y(0) * Y(1) * X(2)),
    ]
    n = 3
    c = A["Circuit"](n)
    c.add(g.RX(i, 0.1 * i) for i in range(n))
    c.add(g.CNOT(i, i + 1) for i in range(n - 1))
    c.add(g.RZ(i, 0.2 * i) for i in range(n))
    for ham in hams:
        exact = ham.expectation(c)
        assert exact == pytest.approx(O.expectation_terms_gatepath(O.run_gates(n, gate_list_of(c)), n, ham.terms, ham.constant), abs=1e-12)
        for grouping in (None, "qwc"):
            assert A["expectation_from_samples"](c, ham, n_shots=10000, grouping=grouping) == pytest.approx(exact, abs=0.08)
        with pytest.raises(NotImplementedError):
            A["expectation_from_samples"](c, ham, grouping="gc")


def test_h2_hf_energy_from_samples():
    """tests/test_expectation_samples.py:115-136 (qwc instead of gc): abs 0.01 at 5*10^4 shots, both shot modes."""
    A = _api()
    ham = F.h2_molecule().hamiltonian()
    c = A["Circuit"](4)
    c.add(A["gates"].X(i) for i in range(2))
    for per_term in (True, False):
        e = A["expectation_from_samples"](c, ham, n_shots_per_pauli_term=per_term, n_shots=50000, grouping="qwc")
        assert e == pytest.approx(ham.expectation(c), abs=0.01)


def test_sampled_energy_matches_reference_arithmetic_on_identical_samples():
    """C3-like check (BeH2-shaped, 14 qubits, QWC groups): the same device samples pushed through the oracle's
    restatement of qibo's frequency arithmetic give the same group energies; integer sums are bit-exact."""
    A = _api()
    from qibochem_b200.measurement.optimization import _measurement_basis_rotations
    from qibochem_b200.runtime import get_runtime

    mol = F.synthetic_molecule(7, 6, 3)
    n = mol.nso
    ham = mol.hamiltonian()
    groups = _measurement_basis_rotations(ham, "qwc")
    assert 1 < len(groups) < ham.nterms
    circuit = A["hf_circuit"](n, 6) + A["ucc_circuit"](n, [0, 1, 6, 7], 0.3) + A["ucc_circuit"](n, [2, 9], -0.2)
    st = circuit.run_on_device()
    assert st is get_runtime()._states[n]
    nshots = 20000
    for gi, (expr, m_gates, _) in enumerate(groups[:6]):
        xb = sum(1 << (n - 1 - g.qubits[0]) for g in m_gates if g.basis[0] == "X")
        yb = sum(1 << (n - 1 - g.qubits[0]) for g in m_gates if g.basis[0] == "Y")
        samples = st.sample(xb, yb, nshots, 77 + gi)
        qubit_map = [g.qubits[0] for g in m_gates]
        # frequencies over the measured qubits, in gate order, as qibo would report them
        bits = np.stack([(samples >> np.uint64(n - 1 - q)) & np.uint64(1) for q in qubit_map], axis=1)
        keys = ["".join(map(str, row)) for row in bits.astype(int)]
        from collections import Counter

        freq = Counter(keys)
        ref = 0.0
        masks, coeffs = [], []
        for (x, z), c in expr.terms.items():
            qs = [q for q in range(n) if ((x | z) >> q) & 1]
            ref += O.z_expectation_from_frequencies(freq, qubit_map, qs, complex(c).real)
            masks.append(sum(1 << (n - 1 - q) for q in qs))
            coeffs.append(complex(c).real)
        sums = st.parity_sums(samples, masks)
        assert np.array_equal(sums, O.parity_sums(samples, masks))
        got = float(np.dot(coeffs, sums / nshots))
        assert got == pytest.approx(ref, abs=1e-12)
        assert A["_pauli_term_measurement_expectation"](expr, freq, qubit_map) == pytest.approx(ref, abs=1e-12)


def test_result_frequencies_protocol():
    """circuit(nshots=) -> frequencies(binary=True): character k <-> k-th measured qubit (result.py:113-116)."""
    A = _api()
    g = A["gates"]
    c = A["Circuit"](4)
    c.add([g.X(0), g.X(3)])
    c.add(g.M(3))
    c.add(g.M(1))
    c.add(g.M(0, basis="Z"))
    freq = c(nshots=100).frequencies(binary=True)
    assert dict(freq) == {"101": 100}
    c = A["Circuit"](2)
    c.add(g.H(0))
    c.add(g.M(0, basis=g.X))  # measured in the X basis: |+> always gives 0
    c.add(g.M(1))
    assert dict(c(nshots=64).frequencies()) == {"00": 64}
    c = A["Circuit"](1)
    c.add([g.H(0), g.S(0)])  # |+i>
    c.add(g.M(0, basis=g.Y))
    assert dict(c(nshots=50).frequencies()) == {"0": 50}


def test_adjoint_gradient_equals_parameter_shift_on_uccsd():
    """qibochem_b200.derivative: all partial derivatives from one adjoint pass == qibo-style parameter_shift per
    parameter (doc/source/tutorials/adaptive.rst:46-61 of the reference), on a molecule-shaped UCCSD circuit."""
    A = _api()
    from qibochem_b200.derivative import adjoint_gradient, parameter_shift

    mol = F.synthetic_molecule(4, 4, 1)
    ham = mol.hamiltonian()
    circuit = A["hf_circuit"](mol.nso, 4) + A["ucc_circuit"](mol.nso, [0, 1, 4, 5], 0.21) + A["ucc_circuit"](mol.nso, [1, 7], -0.4)
    circuit.add(A["gates"].RY(2, 0.3))
    circuit.add(A["gates"].CNOT(2, 5))
    circuit.add(A["gates"].RX(5, -0.7))
    e, grad = adjoint_gradient(circuit, ham)
    assert e == pytest.approx(ham.expectation(circuit), abs=1e-10)
    assert grad.shape == (len(circuit.get_parameters()),)
    for k in range(0, grad.size, 3):
        assert grad[k] == pytest.approx(parameter_shift(circuit, ham, k), abs=1e-9)
    with pytest.raises(ValueError):
        parameter_shift(circuit, ham, grad.size)


@pytest.mark.parametrize(
    "terms,grouping,expected_means,expected_variances",
    [
        ("X0", None, [1.0], [0.0]),
        ("X0 + Z0", None, [1.0, 0.0], [0.0, 0.0]),
        ("Z0 + X0*Z1", "qwc", [-1.0, 0.0], [0.0, 0.0]),
    ],
)
def test_sample_statistics(terms, grouping, expected_means, expected_variances):
    """tests/test_expectation_samples.py:139-155 of the reference (same circuit, shots and tolerances)."""
    A = _api()
    from qibochem_b200.measurement import sample_statistics
    from qibochem_b200.measurement.optimization import _measurement_basis_rotations

    X, Z = A["X"], A["Z"]
    form = {"X0": X(0), "X0 + Z0": X(0) + Z(0), "Z0 + X0*Z1": Z(0) + X(0) * Z(1)}[terms]
    ham = A["SymbolicHamiltonian"](form, nqubits=2)
    circuit = A["Circuit"](2)
    circuit.add(A["gates"].H(0))
    circuit.add(A["gates"].X(1))
    grouped = _measurement_basis_rotations(ham, grouping)
    means, variances = sample_statistics(circuit, grouped, n_shots=20000)
    order = np.argsort(means)  # group order is not part of the contract
    assert sorted(means) == pytest.approx(sorted(expected_means), abs=0.08)
    assert [variances[k] for k in order] == pytest.approx([0.0] * len(means), abs=0.1)


def test_sample_statistics_matches_reference_formula_on_identical_samples():
    """The device reduction equals a plain restatement of result.py:144-155 on the same samples: mean from the
    frequencies, sum over DISTINCT outcomes of (value - mean)^2, divided by n_shots - 1."""
    A = _api()
    from collections import Counter

    from qibochem_b200.measurement import sample_statistics
    from qibochem_b200.measurement.optimization import _measurement_basis_rotations
    from qibochem_b200.measurement.result import _group_masks

    X, Y, Z = A["X"], A["Y"], A["Z"]
    n = 5
    ham = A["SymbolicHamiltonian"](0.7 * Z(0) * Z(3) + 0.2 * X(1) * Z(2) - 1.1 * X(1) * Z(4) + 0.4 * Y(0) * Y(2) + 0.9 * Z(2), nqubits=n)
    g = A["gates"]
    circuit = A["Circuit"](n)
    circuit.add(g.RX(q, 0.3 + 0.2 * q) for q in range(n))
    circuit.add(g.CNOT(q, q + 1) for q in range(n - 1))
    circuit.add(g.RY(q, 0.5 - 0.1 * q) for q in range(n))
    grouped = _measurement_basis_rotations(ham, "qwc")
    nshots, seed = 5000, 1234
    means, variances = sample_statistics(circuit, grouped, n_shots=nshots, seed=seed)
    st = circuit.run_on_device()
    for i, (expr, m_gates, _) in enumerate(grouped):
        xb = sum(1 << (n - 1 - gt.qubits[0]) for gt in m_gates if gt.basis[0] == "X")
        yb = sum(1 << (n - 1 - gt.qubits[0]) for gt in m_gates if gt.basis[0] == "Y")
        samples = st.sample(xb, yb, nshots, (seed + 0x9E3779B97F4A7C15 * (i + 1)) % (1 << 64))
        coeffs, masks, const = _group_masks(expr, n)
        keep = sum(1 << (n - 1 - gt.qubits[0]) for gt in m_gates)
        freq = Counter((samples & np.uint64(keep)).tolist())
        value = lambda key: const + sum(c * (-1) ** bin(key & int(m)).count("1") for c, m in zip(coeffs, masks))  # noqa: E731
        ref_mean = sum(value(k) * c for k, c in freq.items()) / nshots
        ref_var = sum((value(k) - ref_mean) ** 2 for k in freq) / (nshots - 1)
        assert means[i] == pytest.approx(ref_mean, abs=1e-12)
        assert variances[i] == pytest.approx(ref_var, abs=1e-12)


@pytest.mark.parametrize("grouping", [None, "qwc"])
def test_v_expectation_vmsa(grouping):
    """tests/test_expectation_samples.py:158-185 of the reference (qwc instead of gc)."""
    A = _api()
    from qibochem_b200.measurement import v_expectation

    X, Y, Z = A["X"], A["Y"], A["Z"]
    g = A["gates"]
    for form in (0.2 * X(0) + Y(2) + 13.0, Y(0) + Z(1) + X(0) * Z(2)):
        ham = A["SymbolicHamiltonian"](form, nqubits=3)
        circuit = A["Circuit"](3)
        circuit.add(g.RX(i, 0.1 * i) for i in range(3))
        circuit.add(g.CNOT(i, i + 1) for i in range(2))
        circuit.add(g.RZ(i, 0.2 * i) for i in range(3))
        expected = ham.expectation(circuit)
        for method in ("vmsa", "vpsr"):
            got = v_expectation(circuit, ham, n_trial_shots=2000, n_shots=50000, grouping=grouping, method=method)
            assert got == pytest.approx(expected, abs=0.08)
    with pytest.raises(AssertionError):
        v_expectation(circuit, ham, n_trial_shots=2000, n_shots=10, grouping=grouping)


def _mcry_matrix(k, phi):
    """RY(phi) on the last of k+1 qubits, controlled on the first k being 1 (qubit 0 = most significant)."""
    dim = 2 ** (k + 1)
    m = np.eye(dim, dtype=complex)
    c, s = np.cos(phi / 2), np.sin(phi / 2)
    m[dim - 2 :, dim - 2 :] = [[c, -s], [s, c]]
    return m


@pytest.mark.parametrize("excitation", [[0, 2], [1, 3], [0, 1, 2, 3], [3, 0, 4, 1]])
def test_qeb_circuit_equals_reference_gate_sequence(excitation):
    """qeb_circuit == the gate list of the reference (ansatz.py:384-397: CNOT/X fan-in, multi-controlled RY(2 theta),
    fan-in reversed) executed gate by gate in the oracle, on a random state; one trainable parameter (the RY angle)."""
    A = _api()
    from qibochem_b200.ansatz import qeb_circuit

    n, theta = 5, 0.3173
    rng = np.random.default_rng(len(excitation) + excitation[0])
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)
    k = len(excitation) // 2
    i_arr, a_arr = excitation[:k], excitation[k:]
    fwd = [("CNOT", (i_arr[-1], i), None) for i in i_arr[-2::-1]]
    fwd += [("CNOT", (a_arr[-1], a), None) for a in a_arr[-2::-1]]
    fwd.append(("CNOT", (a_arr[-1], i_arr[-1]), None))
    fwd += [("X", (q,), None) for q in excitation if q not in (i_arr[-1], a_arr[-1])]
    ref = O.run_gates(n, fwd, state=psi.copy())
    ref = O.apply_gate(ref, n, _mcry_matrix(len(excitation) - 1, 2 * theta), list(excitation[:-1]) + [a_arr[-1]])
    ref = O.run_gates(n, fwd[::-1], state=ref)
    circuit = qeb_circuit(n, excitation, theta)
    assert len(circuit.get_parameters()) == 1 and circuit.get_parameters()[0][0] == pytest.approx(2 * theta)
    got = circuit(initial_state=psi).state()
    assert np.abs(got - ref).max() < 1e-12
    # the excitation moves |i..=1, a..=0> towards |i..=0, a..=1> and leaves the vacuum alone (tests/test_ansatz.py:208-212)
    vac = qeb_circuit(n, excitation, theta)().state()
    assert abs(vac[0] - 1.0) < 1e-12
    with pytest.raises(ValueError):
        qeb_circuit(n, [0, 1, 2])
    with pytest.raises(ValueError):
        qeb_circuit(n, [])


def test_givens_circuit_and_gates():
    """GIVENS / GeneralizedRBS as fused Pauli-rotation runs: the documented matrices [recalled from qibo], and the
    reference's own check (tests/test_ansatz.py:284-300): equal to the paper's CNOT/RY/H decomposition on |0000>."""
    A = _api()
    from qibochem_b200.ansatz import givens_circuit

    g, n, theta = A["gates"], 4, 0.27183
    c, s = np.cos(theta), np.sin(theta)
    for q0, q1 in ((0, 2), (3, 1)):
        circ = A["Circuit"](n)
        circ.add(g.GIVENS(q0, q1, theta))
        for b0, b1, expect in ((0, 1, {(0, 1): c, (1, 0): s}), (1, 0, {(0, 1): -s, (1, 0): c}), (1, 1, {(1, 1): 1.0}), (0, 0, {(0, 0): 1.0})):
            e0 = np.zeros(2**n, dtype=complex)
            e0[(b0 << (n - 1 - q0)) | (b1 << (n - 1 - q1))] = 1.0
            out = circ(initial_state=e0).state()
            want = np.zeros(2**n, dtype=complex)
            for (c0, c1), v in expect.items():
                want[(c0 << (n - 1 - q0)) | (c1 << (n - 1 - q1))] = v
            assert np.abs(out - want).max() < 1e-12
    # double excitation through GeneralizedRBS(in, out, -theta): rotation between |1100> and |0011> only
    circ = givens_circuit(n, [0, 1, 2, 3], theta)
    assert len(circ.get_parameters()) == 1
    e0 = np.zeros(2**n, dtype=complex)
    e0[0b1100] = 1.0
    out = circ(initial_state=e0).state()
    # sign pinned by the reference's own control circuit on non-vacuum states (tests/test_host_logic.py::
    # test_givens_sign_pinned_by_reference_decomposition_on_every_basis_state): |1100> -> cos|1100> - sin|0011>
    assert abs(out[0b1100] - c) < 1e-12 and abs(out[0b0011] + s) < 1e-12 and abs(np.linalg.norm(out) - 1) < 1e-12
    for excitation in ([0, 2], [0, 1, 2, 3]):
        vac = givens_circuit(n, excitation, theta)().state()
        assert abs(vac[0] - 1.0) < 1e-12 and np.abs(vac[1:]).max() < 1e-12
    with pytest.raises(ValueError):
        givens_circuit(n, [0, 1, 2])
    # the whole run fuses into ONE device pass per excitation
    from qibochem_b200.runtime import get_runtime

    circ.run_on_device()
    assert get_runtime().state(n).last_passes == 1


@pytest.mark.parametrize("builder", ["qeb", "givens"])
def test_vqe_loop_over_gates_that_lower_to_several_rotations(builder):
    """qeb / givens ansatz with several excitations: `set_parameters` AFTER the first execution (the cached program's
    fast path, a numpy vector as an optimiser passes it) must update every rotation of a gate with that gate's parameter
    -- it used to broadcast per rotation and fail; energies against the oracle's execution of the lowered program."""
    A = _api()
    from qibochem_b200.ansatz import givens_circuit, qeb_circuit

    n = 6
    build = qeb_circuit if builder == "qeb" else givens_circuit
    circ = A["hf_circuit"](n, 2)
    for exc, th in (([0, 1, 2, 3], 0.1), ([0, 4], 0.2), ([1, 2, 4, 5], 0.3)):
        circ += build(n, exc, th)
    rng = np.random.default_rng(3)
    strings = ("Z0", "X0 X1", "Y1 Y4", "Z2 Z5", "X0 Z1 X2", "Y0 Z1 Z2 Y3", "X2 X3 Y4 Y5")
    terms = {tuple((int(op[1:]), op[0]) for op in s.split()): float(rng.normal()) for s in strings}
    terms[()] = 0.37
    const, xs, zs, cs = F.masks_from_terms(n, terms)
    ham = A["SymbolicHamiltonian"].from_index_masks(n, xs, zs, cs, constant=const)

    def oracle_energy():
        prog = circ.program()
        psi = np.zeros(2**n, dtype=complex)
        psi[prog.basis_index] = 1.0
        for x, z, a in zip(prog.xmask.tolist(), prog.zmask.tolist(), prog.angle.tolist()):
            psi = O.apply_pauli_rotation(psi, x, z, a)
        return O.expectation_masks(psi, xs, zs, cs, const)

    e0 = ham.expectation(circ)
    assert abs(e0 - oracle_energy()) < 1e-10
    for trial in range(3):
        new = rng.uniform(-1, 1, size=len(circ.get_parameters()))
        circ.set_parameters(new)  # numpy vector: fast path on the cached program
        fresh = circ.copy()  # compiled from the gates' own parameters: must agree with the patched program
        assert np.allclose(fresh.program().angle, circ.program().angle, rtol=0, atol=0)
        e = ham.expectation(circ)
        assert abs(e - oracle_energy()) < 1e-10
        assert abs(e - e0) > 1e-6
    assert [p[0] for p in circ.get_parameters()] == pytest.approx(list(new))


def test_runtime_adopts_a_caller_owned_state():
    """`Runtime.adopt` / `forget`: the public API runs on a DeviceState the caller allocated (bench.py's e2e leg), no
    private attribute involved; adopting does not close the previous owner's state."""
    A = _api()
    from qibochem_b200.engine import DeviceState
    from qibochem_b200.runtime import get_runtime

    n = 5
    rt = get_runtime()
    circ = A["hf_circuit"](n, 2) + A["ucc_circuit"](n, [0, 1, 3, 4], 0.3)
    ham = A["SymbolicHamiltonian"](A["Z"](0) * A["Z"](3) + 0.5 * A["X"](1) * A["X"](4) + 0.25)
    e_ref = ham.expectation(circ)
    with DeviceState(n) as mine:
        rt.adopt(mine)
        assert rt.state(n) is mine
        assert abs(ham.expectation(circ) - e_ref) < 1e-13
        assert abs(mine.norm2() - 1.0) < 1e-13  # the circuit really ran on the adopted state
        rt.forget(mine)
        assert rt.state(n) is not mine
        assert abs(mine.norm2() - 1.0) < 1e-13  # still alive: forget() does not close it
    assert abs(ham.expectation(circ) - e_ref) < 1e-13
