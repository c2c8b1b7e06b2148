"""
Grouping of Hamiltonian terms into simultaneously measurable sets and the measurement gates of each group
(reference: ``src/qibochem/measurement/optimization.py:56-104,198-220`` and ``measurement/util.py:22-79``).

Only ``grouping=None`` and ``grouping="qwc"`` (qubit-wise commuting) are on the B200 path; the general-commuting
schemes ("gc", "gc2") need Clifford basis-change circuits and are out of scope.

The commutation graph is built from bit masks instead of parsed strings, and coloured with the same rule as
``networkx.coloring.greedy_color`` (strategy "largest_first": nodes by decreasing degree, ties in input order; each
node takes the smallest colour unused by its neighbours), so small inputs reproduce the reference's groups
(``tests/golden/commuting_groups.json``).
"""

from __future__ import annotations

import numpy as np

from qibochem_b200 import gates
from qibochem_b200.pauli import PauliSum, masks_to_string, string_to_masks

_LETTER = {(1, 0): "X", (1, 1): "Y", (0, 1): "Z"}


def _term_to_string(x: int, z: int) -> str:
    return masks_to_string(x, z)


def _check_terms_commutativity(term1: str, term2: str, qubitwise: bool) -> bool:
    """Do the Pauli strings (e.g. "X0 Z1 Y3") commute -- qubit-wise, or in general?  (util.py:22-43)"""
    x1, z1 = string_to_masks(term1)
    x2, z2 = string_to_masks(term2)
    if qubitwise:
        common = (x1 | z1) & (x2 | z2)
        return ((x1 ^ x2) | (z1 ^ z2)) & common == 0
    return bin((x1 & z2) ^ (z1 & x2)).count("1") % 2 == 0


def _conflict_matrix(x: np.ndarray, z: np.ndarray, qubitwise: bool, rows: slice) -> np.ndarray:
    """bool[rows, T]: True where the two terms do NOT commute (edge of the complement graph)."""
    xr, zr = x[rows, None], z[rows, None]
    if qubitwise:
        common = (xr | zr) & (x | z)[None, :]
        return (((xr ^ x[None, :]) | (zr ^ z[None, :])) & common) != 0
    v = (xr & z[None, :]) ^ (zr & x[None, :])
    par = np.zeros(v.shape, dtype=np.uint64)
    for s in (32, 16, 8, 4, 2, 1):
        v = v ^ (v >> np.uint64(s))
    par = v & np.uint64(1)
    return par != 0


def group_commuting_masks(x, z, qubitwise: bool = True) -> list[list[int]]:
    """Greedy ("largest_first") colouring of the non-commutation graph.  Returns groups of term indices, each group
    in ascending index order, groups ordered by colour.  Host-native (``csrc/grouping.cpp``, OpenMP): 10^5 terms take
    seconds; ``_group_commuting_masks_numpy`` is the same algorithm in numpy (kept as the test oracle)."""
    import ctypes as C

    from qibochem_b200 import _native

    x = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    z = np.ascontiguousarray(np.asarray(z, dtype=np.uint64))
    if x.size == 0:
        return []
    colour = np.zeros(x.size, dtype=np.int32)
    ncol = C.c_int32(0)
    u64p = C.POINTER(C.c_uint64)
    _native.check(
        _native.load().qc_group_commuting(x.size, x.ctypes.data_as(u64p), z.ctypes.data_as(u64p), int(bool(qubitwise)),
                                          colour.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ncol)),
        "qc_group_commuting",
    )  # fmt: skip
    order = np.argsort(colour, kind="stable")
    bounds = np.searchsorted(colour[order], np.arange(ncol.value + 1))
    return [order[bounds[c] : bounds[c + 1]].tolist() for c in range(ncol.value)]


def _group_commuting_masks_numpy(x, z, qubitwise: bool = True, block: int = 2048) -> list[list[int]]:
    """numpy restatement of ``group_commuting_masks`` (test oracle for the native routine)."""
    x = np.asarray(x, dtype=np.uint64)
    z = np.asarray(z, dtype=np.uint64)
    T = x.size
    degree = np.zeros(T, dtype=np.int64)
    for r0 in range(0, T, block):
        degree[r0 : r0 + block] = _conflict_matrix(x, z, qubitwise, slice(r0, min(T, r0 + block))).sum(axis=1)
    order = np.argsort(-degree, kind="stable")
    colour = np.full(T, -1, dtype=np.int64)
    for u in order.tolist():
        conflicts = _conflict_matrix(x, z, qubitwise, slice(u, u + 1))[0]
        used = np.unique(colour[conflicts & (colour >= 0)])
        c = 0
        for v in used.tolist():  # smallest non-negative integer not used by a neighbour
            if v == c:
                c += 1
            elif v > c:
                break
        colour[u] = c
    return [np.nonzero(colour == c)[0].tolist() for c in range(int(colour.max()) + 1)] if T else []


def _group_commuting_terms(terms_list: list[str], qubitwise: bool) -> list[list[str]]:
    """String interface of the reference (util.py:46-79): groups of Pauli strings, every group and the list of
    groups sorted lexicographically."""
    terms_list = list(terms_list)
    masks = [string_to_masks(t) for t in terms_list]
    groups = group_commuting_masks([m[0] for m in masks], [m[1] for m in masks], qubitwise)
    return sorted(sorted(terms_list[k] for k in grp) for grp in groups)


def _qwc_measurement_gates(expression: PauliSum) -> list[gates.M]:
    """One ``M(q, basis=...)`` per qubit touched by a group of qubit-wise commuting terms, sorted by qubit; the first
    term that touches a qubit fixes its basis (optimization.py:56-77)."""
    basis: dict[int, str] = {}
    for x, z in expression.terms:
        q = 0
        m = x | z
        while m >> q:
            if (m >> q) & 1 and q not in basis:
                basis[q] = _LETTER[((x >> q) & 1, (z >> q) & 1)]
            q += 1
    return [gates.M(q, basis=basis[q]) for q in sorted(basis)]


def _measurement_basis_rotations(hamiltonian, grouping: str | None = None):
    """[(expression, measurement gates, extra rotation gates)] per group (optimization.py:198-220).
    ``expression`` is a ``PauliSum`` holding the group's terms with their coefficients."""
    form: PauliSum = hamiltonian.form
    keys = [k for k in form.terms if k != (0, 0)]
    if grouping is None:
        groups = [[k] for k in range(len(keys))]
    elif grouping == "qwc":
        idx_groups = group_commuting_masks([k[0] for k in keys], [k[1] for k in keys], qubitwise=True)
        # reference ordering: terms inside a group and the groups themselves sorted by their string form
        names = [masks_to_string(*k) for k in keys]
        groups = sorted((sorted(g, key=lambda t: names[t]) for g in idx_groups), key=lambda g: [names[t] for t in g])
    elif grouping in ("gc", "gc2"):
        raise NotImplementedError('general-commuting grouping ("gc", "gc2") is outside the B200 hot path; use "qwc"')
    else:
        raise NotImplementedError("Unknown Pauli term grouping method!")
    out = []
    for g in groups:
        expr = PauliSum({keys[t]: form.terms[keys[t]] for t in g})
        out.append((expr, _qwc_measurement_gates(expr), []))
    return out
