"""
Circuit builders of the UCC hot path with the reference's signatures
(``src/qibochem/ansatz/ansatz.py``: ``hf_circuit`` :253-282, ``ucc_circuit`` :285-350, ``circuit_ansatz`` :60-152).

Where the reference emits a gate queue of basis changes, CNOT ladders and one RZ per Pauli string
(``_expi_pauli``, ``_ansatz.py:55-92``), these builders emit one ``PauliRotation`` per string whose parameter is that
RZ angle, so ``circuit.get_parameters()`` / ``set_parameters()`` keep the reference's meaning and order while the
device applies each string (or each fused run of strings) in a single pass.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np

from qibochem_b200 import gates
from qibochem_b200.ansatz._fermion import bk_occupation_mask, excitation_generator
from qibochem_b200.ansatz.utils import generate_excitations, mp2_amplitude
from qibochem_b200.circuit import Circuit
from qibochem_b200.pauli import masks_to_string


def _check_map(ferm_qubit_map):
    if ferm_qubit_map is None:
        ferm_qubit_map = "jw"
    if ferm_qubit_map not in ("jw", "bk"):
        raise NotImplementedError("Fermon-to-qubit mapping must be either 'jw' or 'bk'")
    return ferm_qubit_map


def _expi_pauli(nqubits: int, pauli_string: str, theta, **kwargs) -> Circuit:
    """Circuit for exp(i*theta*pauli_string): ONE fused rotation instead of the reference's gate ladder
    (``_ansatz.py:55-92``).  The rotation's parameter is the reference's RZ angle, -2*theta (``_ansatz.py:87``)."""
    circuit = Circuit(nqubits, **kwargs)
    circuit.add(gates.PauliRotation(pauli_string, -2.0 * theta))
    return circuit


def hf_circuit(nqubits: int, nelectrons: int, ferm_qubit_map: str = "jw", **kwargs) -> Circuit:
    """Circuit preparing the Hartree-Fock state: X on the qubits that hold a 1 in the (mapped) occupation vector.

    Args:
        nqubits: number of qubits (spin-orbitals)
        nelectrons: number of electrons
        ferm_qubit_map: "jw" (default) or "bk"
        kwargs: accepted for signature compatibility with ``qibo.Circuit``
    """
    ferm_qubit_map = _check_map(ferm_qubit_map)
    if ferm_qubit_map == "jw":
        occupied = (1 << nelectrons) - 1
    else:
        occupied = bk_occupation_mask(range(nelectrons), nqubits)
    circuit = Circuit(nqubits, **kwargs)
    circuit.add(gates.X(q) for q in range(nqubits) if (occupied >> q) & 1)
    return circuit


def ucc_circuit(
    nqubits: int,
    excitation: Sequence[int],
    theta: float = 0.0,
    trotter_steps: int = 1,
    ferm_qubit_map: str = "jw",
    **kwargs,
) -> Circuit:
    r"""Unitary coupled-cluster circuit of a single excitation, :math:`e^{\theta (T - T^\dagger)}` Trotterised per
    Pauli string.

    Args:
        nqubits: number of qubits
        excitation: orbitals involved, e.g. ``[0, 1, 2, 3]`` excites electrons from (0, 1) to (2, 3)
        theta: UCC parameter
        trotter_steps: number of times the ansatz is applied with ``theta / trotter_steps``
        ferm_qubit_map: "jw" (default) or "bk"
    """
    n_orbitals = len(excitation)
    if not n_orbitals:
        raise ValueError("No excitations given")
    if n_orbitals % 2 != 0:
        raise ValueError(f"{excitation} must have an even number of items")
    if ferm_qubit_map not in ("jw", "bk"):
        raise NotImplementedError("Fermon-to-qubit mapping must be either 'jw' or 'bk'")
    generator = excitation_generator(excitation, ferm_qubit_map)
    if trotter_steps < 1:
        raise ValueError(f"{trotter_steps} must be > 0!")
    circuit = Circuit(nqubits, **kwargs)
    for _ in range(trotter_steps):
        for (x, z), coeff in generator.terms.items():
            # reference: _expi_pauli(n, P, -1j*coeff*theta/steps) whose RZ angle is -2 * that (complex typed, imag 0)
            circuit.add(gates.PauliRotation(masks_to_string(x, z), -2.0 * (-1.0j * coeff * theta / trotter_steps)))
    return circuit


def circuit_ansatz(
    molecule,
    ansatz: str,
    *,
    excitations: Sequence[Sequence[int]] | None = None,
    thetas: Sequence[float] | None = None,
    include_hf: bool = True,
    **kwargs,
) -> Circuit:
    """Parameterised ansatz circuit for a molecule.  ``"ucc"`` is the B200 hot path; ``"qeb"`` and ``"givens"`` are
    the same kind of program (fused runs of commuting Pauli rotations); the other ansatz families of the reference
    (br, symm, ham) are outside this package.

    Args:
        molecule: object with ``nso``, ``nelec``, ``n_active_orbs``, ``n_active_e``, ``eps``, ``tei``
        ansatz: ``"ucc"``, ``"qeb"`` or ``"givens"``
        excitations: fermionic excitations; default: all doubles, then all singles (reference order)
        thetas: one parameter per excitation; default: MP2 amplitudes
        include_hf: start from the Hartree-Fock state
    """
    if ansatz in ("br", "symm", "ham"):
        raise NotImplementedError(f'ansatz "{ansatz}" is outside the B200 hot path (only "ucc" is built)')
    if ansatz not in ("ucc", "qeb", "givens"):
        raise ValueError('Invalid argument for "ansatz"')
    builder = {"ucc": ucc_circuit, "qeb": qeb_circuit, "givens": givens_circuit}[ansatz]
    nqubits = molecule.nso if molecule.n_active_orbs is None else molecule.n_active_orbs
    nelec = molecule.nelec if molecule.n_active_e is None else molecule.n_active_e

    if excitations is None:
        excitations = []
        for order in range(2, 0, -1):  # doubles first, then singles
            excitations += generate_excitations(order, range(0, nelec), range(nelec, nqubits))
    elif not all(len(ex) % 2 == 0 for ex in excitations):
        raise ValueError("Excitation with an odd number of orbitals found!")

    if thetas is None:
        thetas = np.array([mp2_amplitude(ex, molecule.eps, molecule.tei) for ex in excitations])
    elif len(thetas) != len(excitations):
        raise ValueError("Number of input parameters doesn't match the number of excitations")

    circuit = Circuit(nqubits, **kwargs)
    if include_hf:
        circuit += hf_circuit(nqubits, nelec, **kwargs)
    for excitation, theta in zip(excitations, thetas):
        circuit += builder(nqubits, excitation, theta, **kwargs)
    return circuit


def _check_excitation(excitation):
    n_orbitals = len(excitation)
    if not n_orbitals:
        raise ValueError("No excitations given")
    if n_orbitals % 2 != 0:
        raise ValueError(f"{excitation} must have an even number of items")
    return n_orbitals


def qeb_circuit(nqubits: int, excitation: Sequence[int], theta: float = 0.0, **kwargs) -> Circuit:
    """Qubit-excitation-based circuit of ONE excitation (reference: ``ansatz.py:353-398``; JW only).

    Same gate sequence as the reference -- a CNOT/X fan-in, one multi-controlled RY(2 theta), the fan-in reversed --
    but the controlled rotation is 2^k commuting Pauli rotations sharing one x-mask (one device pass), and the single
    trainable parameter is still the RY angle."""
    _check_excitation(excitation)
    n_tuples = len(excitation) // 2
    i_array, a_array = list(excitation[:n_tuples]), list(excitation[n_tuples:])
    fwd = [gates.CNOT(i_array[-1], i) for i in i_array[-2::-1]]
    fwd += [gates.CNOT(a_array[-1], a) for a in a_array[-2::-1]]
    fwd.append(gates.CNOT(a_array[-1], i_array[-1]))
    fwd += [gates.X(q) for q in excitation if q not in (i_array[-1], a_array[-1])]
    circuit = Circuit(nqubits, **kwargs)
    circuit.add(fwd)
    circuit.add(gates.RY(a_array[-1], 2.0 * theta).controlled_by(*excitation[:-1]))
    circuit.add(fwd[::-1])
    return circuit


def givens_circuit(nqubits: int, excitation: Sequence[int], theta: float = 0.0, **kwargs) -> Circuit:
    """Fermionic excitation as a Givens rotation (reference: ``ansatz.py:401-435``): ``GIVENS`` for two orbitals,
    ``GeneralizedRBS(in, out, -theta)`` otherwise; each is one fused run of commuting Pauli rotations here."""
    n_orbitals = _check_excitation(excitation)
    orbitals = sorted(excitation)
    q_in, q_out = orbitals[: n_orbitals // 2], orbitals[n_orbitals // 2 :]
    circuit = Circuit(nqubits, **kwargs)
    if n_orbitals == 2:
        circuit.add(gates.GIVENS(q_in[0], q_out[0], theta))
    else:
        circuit.add(gates.GeneralizedRBS(q_in, q_out, -theta))
    return circuit
