from qibochem_b200.ansatz.ansatz import circuit_ansatz, givens_circuit, hf_circuit, qeb_circuit, ucc_circuit
from qibochem_b200.ansatz.utils import generate_excitations, mp2_amplitude

__all__ = ["circuit_ansatz", "hf_circuit", "ucc_circuit", "qeb_circuit", "givens_circuit", "generate_excitations", "mp2_amplitude"]
