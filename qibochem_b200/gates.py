"""
Gate objects of the circuit surface (the subset of ``qibo.gates`` the hot path and the reference's tests of it use).

Every gate lowers to Pauli rotations ``exp(-i phi P / 2)`` plus a global phase, so that ONE device kernel family (K1)
executes the whole circuit:  P = i exp(-i pi P/2) for Pauli gates, S = e^{i pi/4} RZ(pi/2),
H = e^{i pi/2} RZ(pi/2) RX(pi/2) RZ(pi/2), CNOT = e^{i pi/4} exp(-i pi/4 Z_c) exp(-i pi/4 X_t) exp(+i pi/4 Z_c X_t), ...
The global phase is tracked on the host and applied when amplitudes are copied out; expectation values and
sampling do not depend on it.
"""

from __future__ import annotations

import cmath
import math

from qibochem_b200.pauli import string_to_masks

_PI = math.pi


class Gate:
    """Base class.  ``lower()`` -> ([(x, z, angle, param_index|None, scale)], phase) in QUBIT-position masks."""

    name = "gate"
    n_params = 0

    def __init__(self, *qubits: int):
        self.qubits = tuple(int(q) for q in qubits)
        self.parameters: tuple = ()

    @property
    def target_qubits(self):
        return self.qubits[-1:] if len(self.qubits) > 1 else self.qubits

    @property
    def control_qubits(self):
        return self.qubits[:-1] if len(self.qubits) > 1 else ()

    def _rot(self, letters: str, angle: float, param: int | None = None, scale: float = 1.0):
        x, z = string_to_masks(letters)
        return (x, z, angle, param, scale)

    def lower(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def dagger(self) -> "Gate":
        return self  # self-inverse by default

    def __repr__(self):
        p = f", {self.parameters[0]}" if self.parameters else ""
        return f"{self.name}({', '.join(map(str, self.qubits))}{p})"


class _PauliGate(Gate):
    letter = "X"

    def __init__(self, q: int):
        super().__init__(q)
        self.name = self.letter.lower()

    def lower(self):
        return [self._rot(f"{self.letter}{self.qubits[0]}", _PI)], 1j


class X(_PauliGate):
    letter = "X"


class Y(_PauliGate):
    letter = "Y"


class Z(_PauliGate):
    letter = "Z"


class H(Gate):
    name = "h"

    def lower(self):
        q = self.qubits[0]
        half = _PI / 2
        return [self._rot(f"Z{q}", half), self._rot(f"X{q}", half), self._rot(f"Z{q}", half)], 1j


class _Phase(Gate):
    """diag(1, e^{i a}) = e^{i a/2} RZ(a)"""

    angle = 0.0

    def lower(self):
        return [self._rot(f"Z{self.qubits[0]}", self.angle)], cmath.exp(0.5j * self.angle)


class S(_Phase):
    name = "s"
    angle = _PI / 2

    def dagger(self):
        return SDG(*self.qubits)


class SDG(_Phase):
    name = "sdg"
    angle = -_PI / 2

    def dagger(self):
        return S(*self.qubits)


class T(_Phase):
    name = "t"
    angle = _PI / 4

    def dagger(self):
        return TDG(*self.qubits)


class TDG(_Phase):
    name = "tdg"
    angle = -_PI / 4

    def dagger(self):
        return T(*self.qubits)


class CNOT(Gate):
    name = "cx"

    def __init__(self, control: int, target: int):
        super().__init__(control, target)

    def lower(self):
        c, t = self.qubits
        half = _PI / 2
        return [self._rot(f"Z{c}", half), self._rot(f"X{t}", half), self._rot(f"Z{c} X{t}", -half)], cmath.exp(0.25j * _PI)


class CZ(Gate):
    name = "cz"

    def __init__(self, control: int, target: int):
        super().__init__(control, target)

    def lower(self):
        c, t = self.qubits
        half = _PI / 2
        return [self._rot(f"Z{c}", half), self._rot(f"Z{t}", half), self._rot(f"Z{c} Z{t}", -half)], cmath.exp(0.25j * _PI)


class SWAP(Gate):
    name = "swap"

    def __init__(self, q0: int, q1: int):
        super().__init__(q0, q1)

    @property
    def target_qubits(self):
        return self.qubits

    @property
    def control_qubits(self):
        return ()

    def lower(self):
        a, b = self.qubits
        half = _PI / 2
        return [self._rot(f"{p}{a} {p}{b}", half) for p in "XYZ"], cmath.exp(0.25j * _PI)


class _ParamRotation(Gate):
    """exp(-i theta P / 2) with one trainable parameter theta (qibo's RX/RY/RZ convention)."""

    n_params = 1
    letter = "Z"

    def __init__(self, q: int, theta: float = 0.0):
        super().__init__(q)
        self.parameters = (theta,)

    def lower(self):
        return [self._rot(f"{self.letter}{self.qubits[0]}", _real(self.parameters[0]), 0, 1.0)], 1.0

    def dagger(self):
        return type(self)(self.qubits[0], -self.parameters[0])


class RX(_ParamRotation):
    name = "rx"
    letter = "X"


class RY(_ParamRotation):
    name = "ry"
    letter = "Y"


class RZ(_ParamRotation):
    name = "rz"
    letter = "Z"


class PauliRotation(Gate):
    """exp(-i phi P / 2) for a whole Pauli string, e.g. ``PauliRotation("X0 Z1 Y2", phi)``.

    This is what ``ucc_circuit`` emits in place of the reference's basis-change/CNOT-ladder/RZ program
    (``src/qibochem/ansatz/_ansatz.py:55-92``): the single parameter is the RZ angle of that program."""

    name = "pauli_rotation"
    n_params = 1

    def __init__(self, pauli_string: str, phi=0.0):
        ops = sorted(((int(o[1:]), o[0]) for o in pauli_string.split() if o[0] != "I"), key=lambda t: t[0])
        super().__init__(*[q for q, _ in ops])
        self.pauli_string = " ".join(f"{p}{q}" for q, p in ops)
        self.parameters = (phi,)

    @property
    def target_qubits(self):
        return self.qubits

    @property
    def control_qubits(self):
        return ()

    def lower(self):
        return [self._rot(self.pauli_string, _real(self.parameters[0]), 0, 1.0)], 1.0

    def dagger(self):
        return PauliRotation(self.pauli_string, -self.parameters[0])

    def __repr__(self):
        return f"exp(-i*{self.parameters[0]}/2 * [{self.pauli_string}])"


class M(Gate):
    """Measurement of `qubits` in the X, Y or Z basis (``gates.M(q, basis=...)`` of qibo; the reference passes the
    gate class, ``src/qibochem/measurement/optimization.py:64-75``; the letters "X"/"Y"/"Z" are accepted too)."""

    name = "measure"

    def __init__(self, *qubits: int, basis=None):
        super().__init__(*qubits)
        if basis is None:
            basis = "Z"
        if isinstance(basis, (list, tuple)):
            letters = [_basis_letter(b) for b in basis]
            if len(letters) != len(self.qubits):
                raise ValueError("one basis per measured qubit expected")
        else:
            letters = [_basis_letter(basis)] * len(self.qubits)
        self.basis = tuple(letters)

    @property
    def target_qubits(self):
        return self.qubits

    @property
    def control_qubits(self):
        return ()

    def lower(self):
        return [], 1.0

    def __repr__(self):
        return f"M({', '.join(f'{b}{q}' for q, b in zip(self.qubits, self.basis))})"


def _basis_letter(basis) -> str:
    if isinstance(basis, str):
        letter = basis.upper()
    elif isinstance(basis, type) and issubclass(basis, _PauliGate):
        letter = basis.letter
    elif isinstance(basis, _PauliGate):
        letter = basis.letter
    else:
        raise NotImplementedError(f"Unsupported measurement basis {basis!r}")
    if letter not in ("X", "Y", "Z"):
        raise NotImplementedError(f"Unsupported measurement basis {basis!r}")
    return letter


def _real(v) -> float:
    """Parameters may arrive complex-typed with zero imaginary part (the reference's ucc_circuit does that,
    ``src/qibochem/ansatz/ansatz.py:347-349``)."""
    if isinstance(v, complex):
        return v.real
    return float(getattr(v, "real", v))
