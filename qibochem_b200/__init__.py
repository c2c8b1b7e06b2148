"""
qibochem_b200 -- B200-native implementation of qibochem's UCC-energy hot path.

HF basis state -> fused Pauli rotations exp(-i phi P / 2) -> Hamiltonian expectation (exact and sampled), on a
complex128 state vector resident in HBM, behind qibochem's Python surface (``hf_circuit``, ``ucc_circuit``,
``circuit_ansatz``, ``expectation``, ``expectation_from_samples``).  All arithmetic on amplitudes runs in
hand-written CUDA kernels (sm_100a) reached through the C ABI of ``include/qibochem_b200.h``; there is no CPU path.
"""

__version__ = "0.1.0"

from qibochem_b200 import gates  # noqa: E402
from qibochem_b200.circuit import Circuit  # noqa: E402
from qibochem_b200.hamiltonian import SymbolicHamiltonian  # noqa: E402
from qibochem_b200 import pauli as symbols  # noqa: E402  (symbols.X / symbols.Y / symbols.Z / symbols.I)

__all__ = ["Circuit", "SymbolicHamiltonian", "gates", "symbols"]
