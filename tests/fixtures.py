"""Known-answer fixtures taken from the reference's docs/tests (values only; citations relative to the reference)."""

from oracle import oracle as O

# doc/source/tutorials/hamiltonian.rst:59 -- H2 / STO-3G at 0.74804 Angstrom, Jordan-Wigner, 16 significant figures
H2_JW_TERMS = {
    (): -0.10728041160866736,
    ((0, "Z"),): 0.17018261181714206,
    ((1, "Z"),): 0.17018261181714206,
    ((2, "Z"),): -0.21975065439248248,
    ((3, "Z"),): -0.21975065439248248,
    ((0, "Z"), (1, "Z")): 0.16830546187934975,
    ((0, "Z"), (2, "Z")): 0.1201637905291267,
    ((0, "Z"), (3, "Z")): 0.16557911534082753,
    ((1, "Z"), (2, "Z")): 0.16557911534082753,
    ((1, "Z"), (3, "Z")): 0.1201637905291267,
    ((2, "Z"), (3, "Z")): 0.1740435576141825,
    ((0, "X"), (1, "X"), (2, "Y"), (3, "Y")): -0.045415324811700825,
    ((0, "X"), (1, "Y"), (2, "Y"), (3, "X")): 0.045415324811700825,
    ((0, "Y"), (1, "X"), (2, "X"), (3, "Y")): 0.045415324811700825,
    ((0, "Y"), (1, "Y"), (2, "X"), (3, "X")): -0.045415324811700825,
}
H2_HF_ENERGY = -1.11628373627429  # doc/source/tutorials/ansatz.rst:73
H2_VACUUM_ENERGY = 0.7074183344740923  # doc/source/tutorials/hamiltonian.rst:40, ansatz.rst:80
H2_FCI_ENERGY = -1.1371622544371  # lowest eigenvalue of the 16x16 matrix of H2_JW_TERMS

# doc/source/tutorials/measurement.rst:290-325 -- 2-qubit BK-reduced H2 model; exact expectation -0.0171209
BK_H2_CONST = -0.4584
BK_H2_TERMS = [
    (0.3593, [(0, "Z")]),
    (-0.4826, [(1, "Z")]),
    (0.5818, [(0, "Z"), (1, "Z")]),
    (0.0896, [(0, "X"), (1, "X")]),
    (0.0896, [(0, "Y"), (1, "Y")]),
]


def doc_circuit_gates(n, f=0.1):
    """The RX/RZ/CNOT/RX/RZ circuit of measurement.rst:170-176 and :300-306 as an oracle gate list."""
    return (
        [("RX", (i,), f) for i in range(n)]
        + [("RZ", (i,), f) for i in range(n)]
        + [("CNOT", (i, i + 1), None) for i in range(n - 1)]
        + [("RX", (i,), 2 * f) for i in range(n)]
        + [("RZ", (i,), 2 * f) for i in range(n)]
    )


def masks_from_terms(n, terms):
    """{((q,'P'),...): c} -> (constant, xmasks, zmasks, coeffs)"""
    const = 0.0
    xs, zs, cs = [], [], []
    for key, c in terms.items():
        if not key:
            const = c
            continue
        x, z = O.pauli_masks(n, " ".join(f"{p}{q}" for q, p in key))
        xs.append(x), zs.append(z), cs.append(c)
    return const, xs, zs, cs


def term_list(terms):
    return [(c, list(k)) for k, c in terms.items() if k]


def h2_molecule():
    """H2 / STO-3G at 0.74804 Angstrom from the integrals printed at doc/source/tutorials/hamiltonian.rst:40
    (fermionic Hamiltonian: h_pp and V_pqrs = tei[p,q,r,s] / 2); MP2 denominators need orbital energies, which the
    docs do not print -- eps is a placeholder and only used where stated."""
    import numpy as np

    from qibochem_b200.driver import Molecule

    oei = np.diag([-1.248461959132892, -0.48007161818330846])
    tei = np.zeros((2, 2, 2, 2))
    v = {
        (0, 0, 0, 0): 0.3366109237586995, (0, 0, 1, 1): 0.09083064962340165, (0, 1, 0, 1): 0.09083064962340165,
        (0, 1, 1, 0): 0.33115823068165495, (1, 0, 0, 1): 0.3311582306816552, (1, 0, 1, 0): 0.09083064962340165,
        (1, 1, 0, 0): 0.09083064962340165, (1, 1, 1, 1): 0.348087115228365,
    }  # fmt: skip
    for k, val in v.items():
        tei[k] = 2 * val
    return Molecule.from_integrals(oei, tei, nelec=2, e_nuc=0.7074183344740923, eps=np.array([-0.58, 0.67]))


def synthetic_molecule(n_spatial: int, nelec: int, seed: int, scale: float = 0.1):
    """Molecule-shaped synthetic integrals: symmetric h_pq and g_pqrs with the 8-fold
    permutational symmetry of real orbitals, N(0, scale^2), seeded; plus a diagonal one-body ladder so that the HF
    determinant is a sensible reference.  NOT a real molecule (PySCF is unavailable): labelled '<X>-shaped'."""
    import numpy as np

    from qibochem_b200.driver import Molecule

    rng = np.random.default_rng(seed)
    h = rng.normal(scale=scale, size=(n_spatial, n_spatial))
    h = 0.5 * (h + h.T) + np.diag(np.linspace(-1.5, 0.5, n_spatial))
    g = rng.normal(scale=scale, size=(n_spatial,) * 4)  # chemist notation (pq|rs)
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    g = g / 8.0
    tei = np.einsum("pqrs->prsq", g)  # second-quantisation order used by the reference (molecule.py:190-195)
    return Molecule.from_integrals(h, tei, nelec=nelec, e_nuc=0.3, eps=np.diag(h).copy())
