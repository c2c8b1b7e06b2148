"""BASELINE config 3: BeH2-shaped (14 qubits, 6 electrons) UCCSD state, QWC-grouped sampled expectation, 10^6 shots per
group, through the public API (`expectation_from_samples`).  Integrals are synthetic (no PySCF): term structure and
sizes are those of the molecule, the numbers are not."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qibochem_b200.ansatz import circuit_ansatz  # noqa: E402
from qibochem_b200.measurement import expectation_from_samples  # noqa: E402
from qibochem_b200.measurement.optimization import _measurement_basis_rotations  # noqa: E402
from tests.fixtures import synthetic_molecule  # noqa: E402

shots = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mol = synthetic_molecule(7, 6, 3)
ham = mol.hamiltonian()
t0 = time.perf_counter()
circuit = circuit_ansatz(mol, "ucc", thetas=np.random.default_rng(4).uniform(-0.1, 0.1, size=204))
t_build = time.perf_counter() - t0
t0 = time.perf_counter()
groups = _measurement_basis_rotations(ham, "qwc")
t_group = time.perf_counter() - t0
exact = ham.expectation(circuit)
from qibochem_b200.runtime import get_runtime  # noqa: E402

st = get_runtime().state(mol.nso)
st.profile_enable(True)
for rep in range(2):
    st.profile_reset()
    t0 = time.perf_counter()
    e = expectation_from_samples(circuit, ham, n_shots=shots, grouping="qwc", seed=5 + rep)
    dt = time.perf_counter() - t0
print({k: (round(v["ms"], 1), v["launches"]) for k, v in st.profile().items() if v["launches"]})
print(f"BeH2-shaped: {mol.nso} qubits, {len(circuit.get_parameters())} rotations, {ham.nterms} terms, {len(groups)} QWC groups "
      f"(grouping {t_group:.2f}s, circuit build {t_build:.2f}s)")
print(f"exact {exact:.8f}  sampled {e:.8f}  |diff| {abs(e - exact):.2e}  with {shots} shots/group: {dt:.3f}s per sampled energy "
      f"= {len(groups) * shots / dt / 1e9:.2f} Gshots/s")
