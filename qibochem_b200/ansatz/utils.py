"""
Host-side combinatorics of the UCC ansatz: which excitations, in which order, with which starting amplitudes.

The ORDER is part of parity (excitations do not commute), so these functions reproduce the reference's results
exactly -- ``src/qibochem/ansatz/utils.py:11-44`` (``generate_excitations``), ``:47-114`` (``_sort_excitations``),
``:117-154`` (``mp2_amplitude``) -- and are checked against outputs of the reference's own module
(``tests/golden/excitations.json``).  The implementation is index-based (no repeated list scans), so the 4 424
excitations of a (30 qubit, 14 electron) register order in milliseconds.
"""

from __future__ import annotations

from collections.abc import Sequence
from itertools import combinations

import numpy as np


def _spin_allowed(excitation: list[int], order: int) -> bool:
    """Even index sum and as many odd (beta) spin-orbitals emptied as filled (utils.py:37-43)."""
    n_odd_from = sum(i & 1 for i in excitation[:order])
    n_odd_to = sum(i & 1 for i in excitation[order:])
    return sum(excitation) % 2 == 0 and n_odd_from == n_odd_to


def generate_excitations(
    order: int, excite_from: Sequence[int], excite_to: Sequence[int], conserve_spin: bool = True
) -> list[list[int]]:
    """All excitations of the given order between occupied and virtual spin-orbitals, sorted the reference way.

    Args:
        order: 1 == single, 2 == double
        excite_from: occupied orbitals
        excite_to: virtual orbitals
        conserve_spin: keep only spin-conserving excitations
    """
    if order > min(len(excite_from), len(excite_to)):
        return [[]]
    found = [[*occ, *virt] for occ in combinations(excite_from, order) for virt in combinations(excite_to, order)]
    if conserve_spin:
        found = [ex for ex in found if _spin_allowed(ex, order)]
    return _sort_excitations(found)


def _mo_spread(ex: Sequence[int], order: int, weighted: bool) -> int:
    """sum_i w_i |MO(ex[2i+1]) - MO(ex[2i])| with MO(p) = p // 2; w_i = order+1-i when weighted, else 1."""
    return sum(((order + 1 - i) if weighted else 1) * abs(ex[2 * i + 1] // 2 - ex[2 * i] // 2) for i in range(order))


def _sort_excitations(excitations: list[list[int]]) -> list[list[int]]:
    """Reference ordering (utils.py:47-114): (1) excitations whose electrons/holes sit in the same spatial orbitals
    come first, ascending; (2) the rest is ranked by a weighted MO spread (stable), and after each excitation taken
    from the head of that ranking its same-MO partners (spin-flipped source and/or target) follow immediately."""
    order = len(excitations[0]) // 2
    if order > 2:
        raise NotImplementedError("Can only handle single and double excitations")
    if any(len(ex) // 2 != order for ex in excitations):
        raise ValueError("Cannot handle excitations of different orders")

    pool = [list(ex) for ex in excitations]
    # (1) paired excitations
    paired = sorted(ex for ex in pool if _mo_spread(ex, order, weighted=False) == 0)
    result: list[list[int]] = []
    taken = [False] * len(pool)
    position: dict[tuple, list[int]] = {}
    for k, ex in enumerate(pool):
        position.setdefault(tuple(ex), []).append(k)

    def take(ex) -> bool:
        """Move the first not-yet-taken occurrence of `ex` to the result."""
        for k in position.get(tuple(ex), ()):
            if not taken[k]:
                taken[k] = True
                result.append(pool[k])
                return True
        return False

    for ex in paired:
        take(ex)
    # (2) stable ranking of what is left (singles keep plain ascending order)
    rest = [k for k in range(len(pool)) if not taken[k]]
    if order != 1:
        rest.sort(key=lambda k: _mo_spread(pool[k], order, weighted=True))
    else:
        rest.sort(key=lambda k: pool[k])
    flip = lambda orbs: [p + 1 if p % 2 == 0 else p - 1 for p in orbs]  # noqa: E731
    for k in rest:
        if taken[k]:
            continue
        taken[k] = True
        head = pool[k]
        result.append(head)
        src, dst = head[:order], head[order:]
        partners = sorted(sorted(list(s) + list(d)) for s in (src, flip(src)) for d in (dst, flip(dst)))
        for ex in partners:
            take(ex)
    return result


def mp2_amplitude(excitation: Sequence[int], orbital_energies: Sequence[float], tei: np.ndarray) -> float:
    r"""MP2 guess amplitude :math:`t_{ij}^{ab} = (g_{ijab} - g_{ijba}) / (e_i + e_j - e_a - e_b)` of a double
    excitation given in spin-orbital indices; 0.0 for a single excitation (utils.py:117-154).

    Args:
        excitation: 2 or 4 spin-orbital indices
        orbital_energies: eigenvalues of the Fock operator (per spatial orbital)
        tei: two-electron integrals in the MO basis, second-quantisation index order
    """
    if len(excitation) not in (2, 4):
        raise ValueError(f"{excitation} must have either 2 or 4 elements")
    if len(excitation) == 2:
        return 0.0
    i, j, a, b = excitation
    mi, mj, ma, mb = (p // 2 for p in excitation)
    same_spin = lambda p, q: (p + q) % 2 == 0  # noqa: E731
    direct = tei[mi, mj, ma, mb] if same_spin(i, b) and same_spin(j, a) else 0.0
    exchange = tei[mi, mj, mb, ma] if same_spin(i, a) and same_spin(j, b) else 0.0
    eps = np.asarray(orbital_energies)
    return (direct - exchange) / (eps[mi] + eps[mj] - eps[ma] - eps[mb])
