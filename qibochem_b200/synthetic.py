"""
Synthetic workloads of the named benchmark configurations (BASELINE.json configs 4-5): a UCCSD-ordered
Pauli-rotation stream and a molecular-structured Hamiltonian of ~10^5 terms, as flat mask arrays in index-bit
positions (qubit q <-> bit n-1-q).  Deterministic given the seeds; no molecule, no integrals.
"""

from __future__ import annotations

from itertools import combinations

import numpy as np

from qibochem_b200.ansatz._fermion import excitation_generator
from qibochem_b200.ansatz.utils import generate_excitations
from qibochem_b200.pauli import reverse_bits


def uccsd_rotation_stream(n: int, nelec: int, max_excitations: int | None = None, seed: int = 6):
    """Rotations of ``circuit_ansatz(.., "ucc")`` for an (n qubit, nelec electron) register: doubles then singles in
    the reference's order, 8 / 2 JW strings each; theta_e ~ U(-0.1, 0.1).  Returns (xmask, zmask, angle, n_exc)."""
    excitations = []
    for order in (2, 1):
        excitations += generate_excitations(order, range(nelec), range(nelec, n))
    if max_excitations is not None:
        excitations = excitations[:max_excitations]
    thetas = np.random.default_rng(seed).uniform(-0.1, 0.1, size=len(excitations))
    xs, zs, angles = [], [], []
    for exc, theta in zip(excitations, thetas):
        for (x, z), coeff in excitation_generator(exc).terms.items():
            xs.append(reverse_bits(x, n))
            zs.append(reverse_bits(z, n))
            angles.append((2.0j * coeff * theta).real)  # RZ angle of the reference's program
    return (np.array(xs, dtype=np.uint64), np.array(zs, dtype=np.uint64), np.array(angles, dtype=np.float64), len(excitations))


def hf_index(n: int, nelec: int) -> int:
    """Index of the JW Hartree-Fock determinant: qubits 0..nelec-1 set, qubit 0 = most significant bit."""
    return ((1 << nelec) - 1) << (n - nelec)


def molecular_like_hamiltonian(n: int, n_terms: int = 100_000, seed_sets: int = 7, seed_coeff: int = 8):
    """JW-molecular-structured term list: n Z_p, C(n,2) Z_pZ_q, for every pair p<q the hopping strings X_p Z..Z X_q
    and Y_p Z..Z Y_q bare and decorated with one extra Z_r, and four-sets {p<q<r<s} drawn without replacement
    carrying XXYY / YYXX / XYYX / YXXY with Z chains inside (p,q) and (r,s); coefficients N(0, 0.05^2).
    n=30, n_terms=100000 gives 99 999 strings + 1 constant over 19 012 distinct x-masks.
    Returns (xmask, zmask, coeff, constant)."""
    bit = lambda q: 1 << (n - 1 - q)  # noqa: E731
    chain = lambda p, q: sum(bit(r) for r in range(p + 1, q))  # noqa: E731
    xs, zs = [], []
    for p in range(n):
        xs.append(0), zs.append(bit(p))
    for p, q in combinations(range(n), 2):
        xs.append(0), zs.append(bit(p) | bit(q))
    for p, q in combinations(range(n), 2):
        x = bit(p) | bit(q)
        for y in (0, x):  # XX.. and YY..
            base = chain(p, q) | y
            xs.append(x), zs.append(base)
            for r in range(n):
                if r not in (p, q):
                    xs.append(x), zs.append(base ^ bit(r))
    n_quads = (n_terms - 1 - len(xs)) // 4
    all_quads = list(combinations(range(n), 4))
    if n_quads > len(all_quads):
        raise ValueError(f"at most {1 + len(xs) + 4 * len(all_quads)} terms are available for n={n}")
    pick = np.random.default_rng(seed_sets).choice(len(all_quads), size=max(n_quads, 0), replace=False)
    for k in np.sort(pick):
        p, q, r, s = all_quads[k]
        x = bit(p) | bit(q) | bit(r) | bit(s)
        zc = chain(p, q) | chain(r, s)
        for ysites in ((r, s), (p, q), (q, r), (p, s)):  # XXYY, YYXX, XYYX, YXXY
            xs.append(x), zs.append(zc | bit(ysites[0]) | bit(ysites[1]))
    coeff = np.random.default_rng(seed_coeff).normal(scale=0.05, size=len(xs))
    constant = float(np.random.default_rng(seed_coeff + 1000).normal(scale=0.05))
    return np.array(xs, dtype=np.uint64), np.array(zs, dtype=np.uint64), coeff, constant
