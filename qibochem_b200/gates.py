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

    def controlled_by(self, *controls: int) -> "Gate":
        """Multi-controlled rotation (qibo's ``gate.controlled_by``; the reference uses it for the MCRY of
        ``qeb_circuit``, ``src/qibochem/ansatz/ansatz.py:395``).  With projectors (1 - Z_c)/2 on the controls,
        ``C^k-R_P(theta) = prod_S exp(-i theta (-1)^{|S|} / 2^k  Z_S P_t / 2)`` over the subsets S of the controls: 2^k
        commuting strings that share one x-mask, i.e. ONE fused pass on the device."""
        if not controls:
            return self
        t = self.qubits[0]
        if t in controls or len(set(controls)) != len(controls):
            raise ValueError("control and target qubits must be distinct")
        k = len(controls)
        strings, coeffs = [], []
        for subset in range(1 << k):
            zs = [controls[j] for j in range(k) if (subset >> j) & 1]
            strings.append(" ".join([f"Z{q}" for q in zs] + [f"{self.letter}{t}"]))
            coeffs.append((-1.0) ** len(zs) / (1 << k))
        gate = PauliRotationSum(strings, coeffs, self.parameters[0], name=f"c{self.name}")
        gate.qubits = tuple(controls) + (t,)
        return gate


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


class PauliRotationSum(Gate):
    """``exp(-i theta/2 * sum_j c_j P_j)`` for MUTUALLY COMMUTING Pauli strings with one shared parameter theta:
    lowers to one rotation per string with angle ``c_j * theta``.  Building block of the controlled rotations, the
    Givens / RBS gates and the qubit-excitation (QEB) circuits; strings that also share their x-mask fuse into a
    single device pass."""

    name = "pauli_rotation_sum"
    n_params = 1

    def __init__(self, strings, coeffs, theta=0.0, name: str | None = None):
        strings = [" ".join(sorted(s.split(), key=lambda o: int(o[1:]))) for s in strings]
        coeffs = [float(c) for c in coeffs]
        if len(strings) != len(coeffs) or not strings:
            raise ValueError("one real coefficient per Pauli string expected")
        masks = [string_to_masks(s) for s in strings]
        for i, (x1, z1) in enumerate(masks):
            for x2, z2 in masks[i + 1 :]:
                if (bin(x1 & z2).count("1") + bin(z1 & x2).count("1")) & 1:
                    raise ValueError("PauliRotationSum needs mutually commuting strings")
        qubits = sorted({int(o[1:]) for s in strings for o in s.split()})
        super().__init__(*qubits)
        self.strings, self.coeffs = strings, coeffs
        self.parameters = (theta,)
        if name:
            self.name = name

    @property
    def target_qubits(self):
        return self.qubits

    @property
    def control_qubits(self):
        return ()

    def lower(self):
        theta = _real(self.parameters[0])
        return [self._rot(s, c * theta, 0, c) for s, c in zip(self.strings, self.coeffs)], 1.0

    def dagger(self):
        out = PauliRotationSum(self.strings, self.coeffs, -self.parameters[0], name=self.name)
        out.qubits = self.qubits
        return out

    def __repr__(self):
        return f"{self.name}({', '.join(map(str, self.qubits))}, {self.parameters[0]})"


def _subspace_rotation_generator(qubits_in, qubits_out):
    """Pauli expansion of the Hermitian G with ``exp(-i theta G) = exp(theta (K - K^dagger))``,
    ``K = |in = 0..0, out = 1..1><in = 1..1, out = 0..0|``: the rotation that takes ``|1..1, 0..0>`` to
    ``cos(theta)|1..1, 0..0> + sin(theta)|0..0, 1..1>`` and leaves every other basis state alone.
    Returns (strings, real coefficients)."""
    from qibochem_b200.pauli import PauliSum, masks_to_string

    lower = lambda q: (PauliSum.from_string(f"X{q}") + PauliSum.from_string(f"Y{q}", 1j)) * 0.5  # noqa: E731  |0><1|
    raise_ = lambda q: (PauliSum.from_string(f"X{q}") + PauliSum.from_string(f"Y{q}", -1j)) * 0.5  # noqa: E731  |1><0|
    K = PauliSum.identity()
    for q in qubits_in:
        K = K * lower(q)
    for q in qubits_out:
        K = K * raise_(q)
    A = K - K.dagger()  # anti-Hermitian:  A = -i G
    strings, coeffs = [], []
    for (x, z), c in A.terms.items():
        g = 1j * complex(c)  # G = i A
        if abs(g) < 1e-14:
            continue
        if abs(g.imag) > 1e-12:
            raise RuntimeError("internal error: generator is not Hermitian")
        strings.append(masks_to_string(x, z))
        coeffs.append(g.real)
    return strings, coeffs


class GIVENS(PauliRotationSum):
    """qibo's ``gates.GIVENS(q0, q1, theta)`` [recalled: qibo is not installable here]: identity on |00>, |11> and
    ``[[cos, -sin], [sin, cos]]`` on (|01>, |10>).  Used by ``givens_circuit`` (``ansatz.py:430``)."""

    def __init__(self, q0: int, q1: int, theta=0.0):
        strings, coeffs = _subspace_rotation_generator([q0], [q1])
        # exp(-i theta G) takes |q0 = 1, q1 = 0> to cos|10> + sin|01>; qibo's matrix takes |10> to -sin|01> + cos|10>
        super().__init__(strings, [-2.0 * c for c in coeffs], theta, name="givens")
        self.qubits = (int(q0), int(q1))

    def dagger(self):
        return GIVENS(self.qubits[0], self.qubits[1], -self.parameters[0])


class GeneralizedRBS(PauliRotationSum):
    """qibo's ``gates.GeneralizedRBS(qubits_in, qubits_out, theta)`` with phi = 0 [recalled]: the rotation
    ``[[cos, -sin], [sin, cos]]`` on (``|in = 1..1, out = 0..0>``, ``|in = 0..0, out = 1..1>``), identity elsewhere.
    Used by ``givens_circuit`` for double excitations (``ansatz.py:432``)."""

    def __init__(self, qubits_in, qubits_out, theta=0.0, phi=0.0):
        if phi:
            raise NotImplementedError("GeneralizedRBS with phi != 0 is outside the hot path")
        qubits_in, qubits_out = [int(q) for q in qubits_in], [int(q) for q in qubits_out]
        strings, coeffs = _subspace_rotation_generator(qubits_in, qubits_out)
        super().__init__(strings, [2.0 * c for c in coeffs], theta, name="grbs")
        self.qubits_in, self.qubits_out = tuple(qubits_in), tuple(qubits_out)
        self.qubits = self.qubits_in + self.qubits_out

    def dagger(self):
        return GeneralizedRBS(self.qubits_in, self.qubits_out, -self.parameters[0])


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
