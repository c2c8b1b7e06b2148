"""BASELINE configs 0-1 (the reference's own CPU-runnable cases) on the B200 path next to the CPU port of the reference
algorithm: H2 (4 qubits, HF + UCCD, doc fixture) and a LiH-shaped 12-qubit UCCSD energy (synthetic integrals: no PySCF).
Time per energy evaluation through the public API (`set_parameters` + `hamiltonian.expectation`)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402  (CPU baseline leg only)
from oracle import oracle_c as OC  # noqa: E402
from qibochem_b200.ansatz import circuit_ansatz  # noqa: E402
from tests.fixtures import h2_molecule, synthetic_molecule  # noqa: E402


def gate_list(circuit):
    out = []
    for g in circuit.queue:
        if g.name == "pauli_rotation":
            out += O.expi_pauli_gates(g.pauli_string, -0.5 * complex(g.parameters[0]).real)
        elif g.name == "x":
            out.append(("X", g.qubits, None))
    return out


for name, mol in (("H2 (4 qubits, doc fixture)", h2_molecule()), ("LiH-shaped (12 qubits)", synthetic_molecule(6, 4, 1))):
    n = mol.nso
    ham = mol.hamiltonian()
    n_exc = len(circuit_ansatz(mol, "ucc", thetas=None if mol.eps is not None else 0).get_parameters()) if False else None
    circuit = circuit_ansatz(mol, "ucc")
    params = np.array([p[0].real if isinstance(p[0], complex) else p[0] for p in circuit.get_parameters()], dtype=float)
    e = ham.expectation(circuit)  # warm-up: compiles the program, plans the Hamiltonian
    reps = 50
    t0 = time.perf_counter()
    for k in range(reps):
        circuit.set_parameters(params * (1.0 + 1e-3 * k))
        e = ham.expectation(circuit)
    t_gpu = (time.perf_counter() - t0) / reps
    # CPU port of the reference: gate-by-gate program + per-term expectation (all host threads)
    gates = gate_list(circuit)
    hx, hz, hc = ham.index_masks()
    state = np.zeros(1 << n, dtype=np.complex128)
    t0 = time.perf_counter()
    for _ in range(3):
        state[:] = 0
        state[0] = 1
        OC.run_gates(state, n, gates)
        e_cpu = ham.constant + OC.expectation_masks(state, n, hx, hz, hc)
    t_cpu = (time.perf_counter() - t0) / 3
    print(f"{name}: {len(params)} rotations ({len(gates)} reference gates), {hx.size} terms | B200 {1e3 * t_gpu:.2f} ms per energy "
          f"| CPU port ({OC.num_threads()} threads) {1e3 * t_cpu:.2f} ms | E {e:.10f} vs {e_cpu:.10f}")
