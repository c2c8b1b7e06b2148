"""
Molecular integrals -> second-quantised Hamiltonian -> qubit Hamiltonian, straight to bit-mask Pauli sums.

Reference: ``src/qibochem/driver/hamiltonian.py`` (``_fermionic_hamiltonian`` :12-26, ``_qubit_hamiltonian`` :29-48,
``_qubit_to_symbolic_hamiltonian`` :51-68), which delegates to openfermion's ``get_tensors_from_integrals``,
``InteractionOperator``, ``jordan_wigner`` / ``bravyi_kitaev`` and then builds a sympy expression term by term.
Here the same operator is assembled from the Majorana-pair images of ``a_p^dagger a_q`` on bit masks; no sympy, no
per-term Python objects in the result.  Pinned by the H2 tables of ``doc/source/tutorials/hamiltonian.rst:40,59``.
"""

from __future__ import annotations

import numpy as np

from qibochem_b200.ansatz._fermion import bk_ladder, jw_ladder
from qibochem_b200.pauli import PauliSum


def spin_orbital_tensors(oei: np.ndarray, tei: np.ndarray):
    """(one_body[2n,2n], two_body[2n,2n,2n,2n]) in openfermion's layout: spin-orbital 2p / 2p+1 = alpha / beta of
    spatial orbital p, ``H = const + sum T_pq a+_p a_q + sum V_pqrs a+_p a+_q a_r a_s`` with V already holding the 1/2."""
    n = oei.shape[0]
    one = np.zeros((2 * n, 2 * n))
    two = np.zeros((2 * n,) * 4)
    for s in (0, 1):
        one[s::2, s::2] = oei
        two[s::2, s::2, s::2, s::2] = 0.5 * tei  # same spin
        two[s::2, 1 - s :: 2, 1 - s :: 2, s::2] = 0.5 * tei  # mixed spin
    return one, two


def fermionic_terms(oei, tei, constant: float, tol: float = 1e-8) -> dict:
    """{((p, 1), (q, 0)): coeff, ...} like ``openfermion.FermionOperator.terms`` (1 = creation, 0 = annihilation)."""
    one, two = spin_orbital_tensors(np.asarray(oei, float), np.asarray(tei, float))
    terms = {(): float(constant)}
    for p, q in zip(*np.nonzero(one)):
        terms[((int(p), 1), (int(q), 0))] = float(one[p, q])
    for p, q, r, s in zip(*np.nonzero(two)):
        terms[((int(p), 1), (int(q), 1), (int(r), 0), (int(s), 0))] = float(two[p, q, r, s])
    return {k: v for k, v in terms.items() if abs(v) >= tol or k == ()}


def qubit_hamiltonian(oei, tei, constant: float, ferm_qubit_map: str = "jw", tol: float = 1e-8) -> PauliSum:
    """Qubit image of the molecular Hamiltonian as a ``PauliSum`` (identity term = constant part)."""
    if ferm_qubit_map not in ("jw", "bk"):
        raise KeyError("Unknown fermion->qubit mapping!")
    one, two = spin_orbital_tensors(np.asarray(oei, float), np.asarray(tei, float))
    n = one.shape[0]
    if ferm_qubit_map == "jw":
        create = [jw_ladder(p, True) for p in range(n)]
        destroy = [jw_ladder(p, False) for p in range(n)]
    else:
        create = [bk_ladder(p, True, n) for p in range(n)]
        destroy = [bk_ladder(p, False, n) for p in range(n)]
    total = PauliSum.identity(float(constant))
    for p, q in zip(*np.nonzero(one)):
        total.iadd((create[p] * destroy[q]) * float(one[p, q]))
    # a+_p a+_q a_r a_s = (a+_p a_s)(a+_q a_r) for p != q, r != s up to sign: build pair products once
    pair_cache: dict[tuple[int, int], PauliSum] = {}

    def cc(p, q):  # a+_p a+_q
        if (p, q) not in pair_cache:
            pair_cache[(p, q)] = create[p] * create[q]
        return pair_cache[(p, q)]

    dd_cache: dict[tuple[int, int], PauliSum] = {}

    def dd(r, s):  # a_r a_s
        if (r, s) not in dd_cache:
            dd_cache[(r, s)] = destroy[r] * destroy[s]
        return dd_cache[(r, s)]

    for p, q, r, s in zip(*np.nonzero(two)):
        if p == q or r == s:
            continue  # a+_p a+_p = 0
        total.iadd((cc(int(p), int(q)) * dd(int(r), int(s))) * float(two[p, q, r, s]))
    return total.compress(tol)
