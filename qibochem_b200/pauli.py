"""
Bit-mask Pauli algebra used on the host side.

A Pauli string on qubits 0..n-1 is a pair of integers (x, z) in QUBIT positions (bit q <-> qubit q):
X = (1,0), Z = (0,1), Y = (1,1).  ``PauliSum`` is an insertion-ordered map (x, z) -> complex coefficient; it plays the
role that sympy expressions of ``qibo.symbols`` play in the reference (``src/qibochem/driver/hamiltonian.py:51-68``)
but stays O(1) per term, so 10^5-term Hamiltonians are cheap to build and to hand to the device as flat arrays.

Index-bit convention of the device (qubit q <-> bit n-1-q, qibo's) is applied only in ``to_index_masks``.
"""

from __future__ import annotations

import numpy as np

_LETTER = {(0, 0): "I", (1, 0): "X", (1, 1): "Y", (0, 1): "Z"}
_BITS = {"I": (0, 0), "X": (1, 0), "Y": (1, 1), "Z": (0, 1)}
_I_POW = (1, 1j, -1, -1j)


def popcount(v: int) -> int:
    return bin(v).count("1")


def mul_strings(x1: int, z1: int, x2: int, z2: int) -> tuple[int, int, complex]:
    """(x1,z1)*(x2,z2) = phase * (x1^x2, z1^z2), with Hermitian strings P(x,z) = i^{|x&z|} X^x Z^z."""
    x, z = x1 ^ x2, z1 ^ z2
    e = popcount(x1 & z1) + popcount(x2 & z2) - popcount(x & z) + 2 * popcount(z1 & x2)
    return x, z, _I_POW[e % 4]


def string_to_masks(pauli_string: str) -> tuple[int, int]:
    """"X0 Y3 Z11" -> (x, z) in qubit positions."""
    x = z = 0
    for op in pauli_string.split():
        bx, bz = _BITS[op[0]]
        q = int(op[1:])
        x |= bx << q
        z |= bz << q
    return x, z


def masks_to_string(x: int, z: int) -> str:
    """(x, z) -> "X0 Y3 Z11" (ascending qubit order; identity -> "")."""
    out = []
    q = 0
    m = x | z
    while m >> q:
        if (m >> q) & 1:
            out.append(f"{_LETTER[((x >> q) & 1, (z >> q) & 1)]}{q}")
        q += 1
    return " ".join(out)


def reverse_bits(v: int, n: int) -> int:
    """qubit-position mask -> index-bit mask (bit q -> bit n-1-q)."""
    out = 0
    q = 0
    while v >> q:
        if (v >> q) & 1:
            out |= 1 << (n - 1 - q)
        q += 1
    return out


def reverse_bits_array(v: np.ndarray, n: int) -> np.ndarray:
    """Vectorised ``reverse_bits`` for uint64 arrays."""
    v = np.asarray(v, dtype=np.uint64)
    out = np.zeros_like(v)
    for q in range(n):
        out |= ((v >> np.uint64(q)) & np.uint64(1)) << np.uint64(n - 1 - q)
    return out


class PauliSum:
    """Linear combination of Pauli strings with complex coefficients (insertion ordered)."""

    __slots__ = ("terms",)
    __array_ufunc__ = None  # numpy scalars defer to __rmul__/__radd__ instead of broadcasting over the terms

    def __init__(self, terms: dict | None = None):
        self.terms: dict[tuple[int, int], complex] = dict(terms) if terms else {}

    # -- constructors -----------------------------------------------------------------------------------------
    @classmethod
    def identity(cls, coeff=1.0) -> "PauliSum":
        return cls({(0, 0): coeff})

    @classmethod
    def from_string(cls, pauli_string: str, coeff=1.0) -> "PauliSum":
        return cls({string_to_masks(pauli_string): coeff})

    # -- algebra ----------------------------------------------------------------------------------------------
    def _add_term(self, key, coeff, tol=0.0):
        if key in self.terms:
            new = self.terms[key] + coeff
            if tol and abs(new) < tol:
                del self.terms[key]  # a cancelled term leaves the map (like openfermion's SymbolicOperator)
            else:
                self.terms[key] = new
        else:
            self.terms[key] = coeff

    def iadd(self, other: "PauliSum", tol: float = 0.0) -> "PauliSum":
        for key, c in other.terms.items():
            self._add_term(key, c, tol)
        return self

    def __add__(self, other):
        if not isinstance(other, PauliSum):
            other = PauliSum.identity(other)
        return PauliSum(self.terms).iadd(other)

    __radd__ = __add__

    def __neg__(self):
        return PauliSum({k: -c for k, c in self.terms.items()})

    def __sub__(self, other):
        if not isinstance(other, PauliSum):
            other = PauliSum.identity(other)
        return self + (-other)

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        if not isinstance(other, PauliSum):
            return PauliSum({k: c * other for k, c in self.terms.items()})
        out = PauliSum()
        for (x1, z1), c1 in self.terms.items():  # left-major, like openfermion's __imul__
            for (x2, z2), c2 in other.terms.items():
                x, z, ph = mul_strings(x1, z1, x2, z2)
                out._add_term((x, z), c1 * c2 * ph)
        return out

    def __rmul__(self, other):
        return PauliSum({k: other * c for k, c in self.terms.items()})

    def __truediv__(self, other):
        return self * (1.0 / other)

    def __matmul__(self, other):
        return self * other

    def __pow__(self, k: int):
        out = PauliSum.identity()
        for _ in range(int(k)):
            out = out * self
        return out

    def dagger(self) -> "PauliSum":
        return PauliSum({k: np.conj(c) for k, c in self.terms.items()})

    def compress(self, tol: float = 1e-8) -> "PauliSum":
        """Drop |coeff| < tol and strip negligible imaginary parts (openfermion ``compress`` semantics)."""
        out = {}
        for k, c in self.terms.items():
            if abs(c) < tol:
                continue
            c = complex(c)
            if abs(c.imag) < tol:
                c = c.real
            elif abs(c.real) < tol:
                c = 1j * c.imag
            out[k] = c
        self.terms = out
        return self

    # -- inspection -------------------------------------------------------------------------------------------
    def __len__(self):
        return len(self.terms)

    def __iter__(self):
        return iter(self.terms.items())

    @property
    def constant(self) -> complex:
        return self.terms.get((0, 0), 0.0)

    @property
    def nqubits(self) -> int:
        m = 0
        for x, z in self.terms:
            m |= x | z
        return m.bit_length()

    def strings(self) -> list[tuple[str, complex]]:
        return [(masks_to_string(x, z), c) for (x, z), c in self.terms.items()]

    def __repr__(self):
        parts = []
        for (x, z), c in self.terms.items():
            s = masks_to_string(x, z).replace(" ", "*")
            parts.append(f"{c}" if not s else f"{c}*{s}")
        return " + ".join(parts) if parts else "0"

    def isclose(self, other: "PauliSum", tol: float = 1e-8) -> bool:
        keys = set(self.terms) | set(other.terms)
        return all(abs(self.terms.get(k, 0.0) - other.terms.get(k, 0.0)) <= tol for k in keys)


# ---- qibo.symbols look-alikes ---------------------------------------------------------------------------------
class _Symbol(PauliSum):
    """Single-qubit Pauli symbol; behaves as a one-term PauliSum and remembers its qubit (qibo.symbols.X/Y/Z)."""

    __slots__ = ("target_qubit",)
    letter = "I"

    def __init__(self, q: int):
        bx, bz = _BITS[self.letter]
        super().__init__({(bx << int(q), bz << int(q)): 1.0})
        self.target_qubit = int(q)

    @property
    def name(self) -> str:
        return f"{self.letter}{self.target_qubit}"


class X(_Symbol):
    letter = "X"


class Y(_Symbol):
    letter = "Y"


class Z(_Symbol):
    letter = "Z"


class I(_Symbol):  # noqa: E742
    letter = "I"
