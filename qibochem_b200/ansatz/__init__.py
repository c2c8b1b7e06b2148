from qibochem_b200.ansatz.ansatz import circuit_ansatz, hf_circuit, ucc_circuit
from qibochem_b200.ansatz.utils import generate_excitations, mp2_amplitude

__all__ = ["circuit_ansatz", "hf_circuit", "ucc_circuit", "generate_excitations", "mp2_amplitude"]
