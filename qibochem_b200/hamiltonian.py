"""
``SymbolicHamiltonian``: a sum of Pauli strings held as FLAT MASK ARRAYS (what the device consumes), with the slice
of ``qibo.hamiltonians.SymbolicHamiltonian`` that the reference's hot path calls:

* ``.expectation(circuit)``            -- tests/test_ansatz.py:131, examples/ucc_example1.py:45 of the reference
* ``.expectation_from_samples(freq, qubit_map)`` -- src/qibochem/measurement/result.py:54
* ``.form`` / ``.nqubits`` / ``.terms`` / ``+ - * @``

The reference builds the same object as a sympy expression (``src/qibochem/driver/hamiltonian.py:51-68``), which is
O(T) Python objects per use; here 10^5 terms are three numpy arrays.
"""

from __future__ import annotations

import numpy as np

from qibochem_b200.pauli import PauliSum, masks_to_string, reverse_bits_array


class SymbolicHamiltonian:
    def __init__(self, form=None, nqubits: int | None = None):
        if form is None:
            form = PauliSum()
        if not isinstance(form, PauliSum):
            form = PauliSum.identity(form)  # a plain number
        self._form: PauliSum | None = form
        need = form.nqubits
        self.nqubits = int(nqubits) if nqubits is not None else max(need, 1)
        if self.nqubits < need:
            raise ValueError(f"form acts on {need} qubits but nqubits={self.nqubits}")
        self._arrays = None  # (x, z, c) in index-bit positions, identity excluded
        self._constant = None
        self._plan = None

    # ---- construction from flat arrays (synthetic / large Hamiltonians) -------------------------------------------
    @classmethod
    def from_index_masks(cls, nqubits: int, xmask, zmask, coeff, constant: float = 0.0) -> "SymbolicHamiltonian":
        """Masks already in index-bit positions (qubit q <-> bit nqubits-1-q); no identity term among them."""
        self = cls.__new__(cls)
        self._form = None
        self.nqubits = int(nqubits)
        x = np.ascontiguousarray(np.asarray(xmask, dtype=np.uint64))
        z = np.ascontiguousarray(np.asarray(zmask, dtype=np.uint64))
        c = np.ascontiguousarray(np.asarray(coeff, dtype=np.float64))
        if not (x.shape == z.shape == c.shape and x.ndim == 1):
            raise ValueError("xmask, zmask, coeff must be 1-d arrays of equal length")
        if x.size and (int(x.max()) | int(z.max())) >> self.nqubits:
            raise ValueError("mask has bits outside the register")
        if np.any((x | z) == 0):
            raise ValueError("identity term found among the masks; pass it as `constant`")
        self._arrays = (x, z, c)
        self._constant = float(constant)
        self._plan = None
        return self

    # ---- views ------------------------------------------------------------------------------------------------
    @property
    def form(self) -> PauliSum:
        if self._form is None:
            n = self.nqubits
            x, z, c = self._arrays
            xq, zq = reverse_bits_array(x, n), reverse_bits_array(z, n)
            terms = {(0, 0): self._constant} if self._constant else {}
            for a, b, v in zip(xq.tolist(), zq.tolist(), c.tolist()):
                terms[(a, b)] = terms.get((a, b), 0.0) + v
            self._form = PauliSum(terms)
        return self._form

    def index_masks(self):
        """(xmask, zmask, coeff) numpy arrays in index-bit positions, identity term excluded, coefficients real
        (Re<psi|H|psi> only sees the real part of a coefficient because <psi|P|psi> is real)."""
        if self._arrays is None:
            n = self.nqubits
            keys = [(k, c) for k, c in self._form.terms.items() if k != (0, 0)]
            xq = np.array([k[0] for k, _ in keys], dtype=np.uint64)
            zq = np.array([k[1] for k, _ in keys], dtype=np.uint64)
            c = np.array([complex(v).real for _, v in keys], dtype=np.float64)
            self._arrays = (reverse_bits_array(xq, n), reverse_bits_array(zq, n), c)
            self._constant = float(complex(self._form.constant).real)
        return self._arrays

    @property
    def constant(self) -> float:
        self.index_masks()
        return self._constant

    @property
    def nterms(self) -> int:
        return self.index_masks()[0].size

    @property
    def terms(self) -> list[tuple[float, list[tuple[int, str]]]]:
        """[(coefficient, [(qubit, 'X'|'Y'|'Z'), ...]), ...] without the constant."""
        out = []
        for (x, z), c in self.form.terms.items():
            if (x, z) == (0, 0):
                continue
            factors = [(int(op[1:]), op[0]) for op in masks_to_string(x, z).split()]
            out.append((c, factors))
        return out

    # ---- algebra (host side, small Hamiltonians) ----------------------------------------------------------------
    def _binary(self, other, op):
        oform = other.form if isinstance(other, SymbolicHamiltonian) else other
        n = max(self.nqubits, other.nqubits) if isinstance(other, SymbolicHamiltonian) else self.nqubits
        return SymbolicHamiltonian(op(self.form, oform), nqubits=n)

    def __add__(self, other):
        return self._binary(other, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binary(other, lambda a, b: a - b)

    def __rsub__(self, other):
        return SymbolicHamiltonian(other - self.form, nqubits=self.nqubits)

    def __mul__(self, other):
        return self._binary(other, lambda a, b: a * b)

    __rmul__ = __mul__

    def __matmul__(self, other):
        return self._binary(other, lambda a, b: a * b)

    # ---- the hot path -------------------------------------------------------------------------------------------
    def expectation(self, circuit_or_state, nshots: int | None = None) -> float:
        """const + sum_k c_k Re<psi|P_k|psi> on the device (K2), psi = circuit|0..0> (K0 + K1)."""
        from qibochem_b200.circuit import Circuit, CircuitResult
        from qibochem_b200.engine import DeviceState
        from qibochem_b200.runtime import get_runtime
        from qibochem_b200.sharded import ShardedState

        if nshots is not None:
            from qibochem_b200.measurement.result import expectation_from_samples

            return expectation_from_samples(circuit_or_state, self, n_shots=nshots)
        if isinstance(circuit_or_state, Circuit):
            if circuit_or_state.nqubits != self.nqubits:
                raise ValueError(f"circuit has {circuit_or_state.nqubits} qubits, Hamiltonian {self.nqubits}")
            st = circuit_or_state.run_on_device()
        elif isinstance(circuit_or_state, CircuitResult):
            circuit_or_state._check_alive()
            st = circuit_or_state._state
        elif isinstance(circuit_or_state, (DeviceState, ShardedState)):
            st = circuit_or_state
        else:
            psi = np.asarray(circuit_or_state, dtype=np.complex128).reshape(-1)
            if psi.size != 1 << self.nqubits:
                raise ValueError("state vector has the wrong size")
            st = get_runtime().state(self.nqubits)
            st.from_numpy(psi)
        x, z, c = self.index_masks()
        # <psi|psi> multiplies the identity term; a circuit run from a basis state is unitary, so the extra reduction
        # (a launch and a device->host read per energy) is only paid for states the caller supplied
        prepared_here = isinstance(circuit_or_state, Circuit)
        norm = st.norm2() if (self._constant and not prepared_here) else 1.0
        return self._constant * norm + st.expectation(x, z, c)

    def expectation_from_samples(self, freq: dict, qubit_map=None) -> float:
        """Expectation of a Z-only Hamiltonian from measurement frequencies (qibo semantics): character k of a key is
        the outcome of qubit ``qubit_map[k]``.  The signed sums over keys are exact integers computed on the device
        (K3); only the final division by the number of shots is floating point."""
        from qibochem_b200.measurement.result import _frequencies_to_arrays, _z_terms_from_frequencies

        x, z, c = self.index_masks()
        if np.any(x != 0):
            raise NotImplementedError("expectation_from_samples needs a Hamiltonian of Z terms only (rotate first)")
        if qubit_map is None:
            qubit_map = list(range(self.nqubits))
        qubit_map = list(qubit_map)
        keys, counts, total = _frequencies_to_arrays(freq)
        # term masks in key-bit positions: key read as a binary number, character k <-> bit (len-1-k)
        nchar = len(next(iter(freq))) if freq else len(qubit_map)
        masks = np.zeros(z.size, dtype=np.uint64)
        n = self.nqubits
        for t, zm in enumerate(z.tolist()):
            m = 0
            for q in range(n):
                if (zm >> (n - 1 - q)) & 1:
                    m |= 1 << (nchar - 1 - qubit_map.index(q))
            masks[t] = m
        sums = _z_terms_from_frequencies(keys, counts, masks)
        return float(self._constant + np.dot(c, sums / total))

    def __repr__(self):
        return f"SymbolicHamiltonian(nqubits={self.nqubits}, nterms={self.nterms}, constant={self.constant})"
