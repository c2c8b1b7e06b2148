"""
Fermion -> qubit maps for single ladder-operator products (what ``openfermion.jordan_wigner`` /
``bravyi_kitaev`` do for the reference at ``src/qibochem/ansatz/ansatz.py:325-336``), on bit masks.

Only what the UCC path needs: the image of ``T - T^dagger`` for ONE excitation (<= 8 Pauli strings).  The order of
the strings follows openfermion's term order (left-major expansion of the product of ladder operators, terms of T
first, cancelled terms removed), because that order is the parameter order of the circuit
(``examples/ucc_example1.py:17`` relies on it).
"""

from __future__ import annotations

from qibochem_b200.pauli import PauliSum

_CANCEL_TOL = 1e-8  # openfermion's EQ_TOLERANCE


def _prefix(p: int) -> int:
    return (1 << p) - 1


def jw_ladder(p: int, dagger: bool) -> PauliSum:
    """a_p = Z_0..Z_{p-1} (X_p + iY_p)/2 ;  a_p^dagger = Z_0..Z_{p-1} (X_p - iY_p)/2   (qubit-position masks)."""
    chain, site = _prefix(p), 1 << p
    return PauliSum({(site, chain): 0.5, (site, chain | site): (-0.5j if dagger else 0.5j)})


# ---- Bravyi-Kitaev on the binary-indexed (Fenwick) tree, the layout openfermion.bravyi_kitaev uses ------------
def _lowbit(j: int) -> int:
    return j & -j


def bk_sets(p: int, n_qubits: int) -> tuple[int, int, int]:
    """(update, parity, remainder) qubit masks of mode p.  Qubit j-1 (1-indexed node j) stores the occupation parity
    of modes (j - lowbit(j), j]:  update = ancestors of node p+1, parity = nodes that sum to the parity of modes < p,
    remainder = parity minus the children of node p+1."""
    j = p + 1
    update = 0
    a = j + _lowbit(j)
    while a <= n_qubits:
        update |= 1 << (a - 1)
        a += _lowbit(a)
    parity = 0
    k = j - 1
    while k > 0:
        parity |= 1 << (k - 1)
        k -= _lowbit(k)
    children = 0
    k, stop = j - 1, j - _lowbit(j)
    while k > stop:
        children |= 1 << (k - 1)
        k -= _lowbit(k)
    return update, parity, parity & ~children


def bk_ladder(p: int, dagger: bool, n_qubits: int) -> PauliSum:
    """a_p^(dagger) = (X_U X_p Z_P  -/+  i X_U Y_p Z_R) / 2."""
    update, parity, remainder = bk_sets(p, n_qubits)
    site = 1 << p
    return PauliSum({(update | site, parity): 0.5, (update | site, remainder | site): (-0.5j if dagger else 0.5j)})


def bk_occupation_mask(occupied_modes, n_qubits: int) -> int:
    """Qubit mask of the BK image of an occupation-number basis state (what ``hf_circuit(.., "bk")`` prepares,
    ``src/qibochem/ansatz/ansatz.py:279-280``): qubit j-1 holds the parity of modes (j - lowbit(j), j]."""
    mask = 0
    occ = set(occupied_modes)
    for j in range(1, n_qubits + 1):
        if sum(1 for m in range(j - _lowbit(j), j) if m in occ) & 1:
            mask |= 1 << (j - 1)
    return mask


def excitation_generator(excitation, ferm_qubit_map: str = "jw", n_qubits: int | None = None) -> PauliSum:
    """Qubit image of T - T^dagger, T = a_{p1}^dagger a_{p2}^dagger .. a_{q1} a_{q2} .. with the orbitals of
    `excitation` sorted in descending order (``ansatz.py:318-329``)."""
    orbitals = sorted(excitation, reverse=True)
    half = len(orbitals) // 2
    t_ops = [(p, True) for p in orbitals[:half]] + [(p, False) for p in orbitals[half:]]
    tdag_ops = [(p, not d) for p, d in reversed(t_ops)]
    if ferm_qubit_map == "jw":
        ladder = jw_ladder
    else:
        nq = max(orbitals) + 1 if n_qubits is None else n_qubits  # openfermion: count_qubits(operator)
        ladder = lambda p, d: bk_ladder(p, d, nq)  # noqa: E731
    total = PauliSum()
    for ops, sign in ((t_ops, 1.0), (tdag_ops, -1.0)):
        prod = PauliSum.identity(sign)
        for p, d in ops:
            prod = prod * ladder(p, d)
        total.iadd(prod, tol=_CANCEL_TOL)
    return total
