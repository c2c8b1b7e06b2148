"""
``Circuit``: the object ``hf_circuit`` / ``ucc_circuit`` / ``circuit_ansatz`` return, with the part of
``qibo.Circuit``'s protocol that the reference's hot path uses: ``add``, ``+``/``+=``, ``copy``, ``queue``,
``get_parameters`` / ``set_parameters`` (one RZ angle per Pauli string, ``README.md:45``, ``examples/ucc_example1.py``)
and ``__call__(initial_state=None, nshots=None)`` -> result with ``.state()`` / ``.frequencies(binary=True)``
(used at ``src/qibochem/measurement/result.py:106-114``).

Executing a circuit compiles its queue to a *Pauli-rotation program* -- flat ``(xmask, zmask, angle)`` arrays in
index-bit positions -- and runs it on the device with K0 (basis state) + K1 (fused rotations).
"""

from __future__ import annotations

from collections import Counter
from collections.abc import Iterable

import numpy as np

from qibochem_b200 import gates as _gates
from qibochem_b200.pauli import reverse_bits
from qibochem_b200.runtime import get_runtime


class Program:
    """Compiled form of a circuit (everything the device needs)."""

    __slots__ = ("nqubits", "basis_index", "xmask", "zmask", "angle", "param_rot", "param_scale", "param_owner", "n_param_gates", "phase",
                 "measured", "x_basis_mask", "y_basis_mask")  # fmt: skip


class Circuit:
    def __init__(self, nqubits: int, **kwargs):
        if kwargs.get("density_matrix"):
            raise NotImplementedError("density-matrix simulation is outside the B200 hot path")
        if kwargs.get("accelerators"):
            # qibo's multi-device argument, reached through the reference's **kwargs forwarding (ansatz.py:120, :341):
            # here one PROCESS per GPU (torchrun); the runtime then shards every state over the ranks
            get_runtime().configure_accelerators(kwargs["accelerators"])
        self.nqubits = int(nqubits)
        self.init_kwargs = dict(kwargs)
        self._queue: list[_gates.Gate] = []
        self._program: Program | None = None
        self._pending: np.ndarray | None = None  # parameters set through the fast path, not yet written to the gates

    @property
    def queue(self) -> list[_gates.Gate]:
        """The gate list (qibo's ``circuit.queue``).  Parameters that an optimiser loop pushed through the fast path of
        ``set_parameters`` are written back to the gate objects first."""
        self._sync_gates()
        return self._queue

    @queue.setter
    def queue(self, gates) -> None:
        self._queue = list(gates)
        self._pending = None
        self._program = None

    def _sync_gates(self) -> None:
        if self._pending is not None:
            flat, self._pending = self._pending, None
            for g, p in zip((g for g in self._queue if g.n_params), flat.tolist()):
                g.parameters = (p,)

    # ---- building ---------------------------------------------------------------------------------------------
    def add(self, gate) -> None:
        if isinstance(gate, _gates.Gate):
            if any(q < 0 or q >= self.nqubits for q in gate.qubits):
                raise ValueError(f"{gate!r} acts outside a {self.nqubits}-qubit circuit")
            self.queue.append(gate)
            self._program = None
        elif isinstance(gate, Iterable):
            for g in gate:
                self.add(g)
        else:
            raise TypeError(f"cannot add {type(gate)} to a circuit")

    def __iadd__(self, other: "Circuit") -> "Circuit":
        if other.nqubits != self.nqubits:
            raise ValueError("cannot add circuits with a different number of qubits")
        self.queue.extend(other.queue)
        self._program = None
        return self

    def __add__(self, other: "Circuit") -> "Circuit":
        out = self.copy()
        out += other
        return out

    def copy(self, deep: bool = False) -> "Circuit":
        out = Circuit(self.nqubits, **self.init_kwargs)
        if deep:
            import copy as _copy

            out.queue = [_copy.copy(g) for g in self.queue]
        else:
            out.queue = list(self.queue)
        return out

    def invert(self) -> "Circuit":
        out = Circuit(self.nqubits, **self.init_kwargs)
        out.queue = [g.dagger() for g in reversed(self.queue) if not isinstance(g, _gates.M)]
        return out

    @property
    def ngates(self) -> int:
        return len(self.queue)

    @property
    def parametrized_gates(self) -> list[_gates.Gate]:
        return [g for g in self.queue if g.n_params]

    @property
    def measurements(self) -> list[_gates.M]:
        return [g for g in self.queue if isinstance(g, _gates.M)]

    # ---- parameters -------------------------------------------------------------------------------------------
    def get_parameters(self) -> list[tuple]:
        """One 1-tuple per parametrized gate, in queue order (qibo's default ``format="list"``)."""
        if self._pending is not None:
            return [(p,) for p in self._pending.tolist()]
        return [tuple(g.parameters) for g in self.parametrized_gates]

    def set_parameters(self, parameters) -> None:
        """Flat sequence with one value per parametrized gate (or a list of 1-tuples as ``get_parameters`` gives)."""
        if self._program is not None and isinstance(parameters, np.ndarray) and parameters.ndim == 1 and parameters.dtype.kind in "fiu":
            # fast path of an optimiser loop (README.md:45 of the reference calls this once per energy): only the angle
            # vector of the compiled program changes; the gate objects are updated lazily (see `queue`)
            prog = self._program
            if parameters.size != prog.n_param_gates:
                raise ValueError(f"Given list of parameters has length {parameters.size} while the circuit contains {prog.n_param_gates} parametrized gates.")
            flat = parameters.astype(np.float64)  # a copy: the caller may reuse its array
            prog.angle[prog.param_rot] = prog.param_scale * flat[prog.param_owner]
            self._pending = flat
            return
        pg = self.parametrized_gates
        params = list(parameters)
        if len(params) != len(pg):
            raise ValueError(f"Given list of parameters has length {len(params)} while the circuit contains {len(pg)} parametrized gates.")
        flat = np.empty(len(pg), dtype=np.float64)
        for k, (g, p) in enumerate(zip(pg, params)):
            if isinstance(p, (tuple, list, np.ndarray)):
                p = p[0]
            g.parameters = (p,)
            flat[k] = _gates._real(p)
        if self._program is not None:  # fast path of an optimiser loop: only the angle vector changes
            prog = self._program
            # a gate may lower to several rotations that share its parameter (controlled RY, GIVENS, GeneralizedRBS)
            prog.angle[prog.param_rot] = prog.param_scale * flat[prog.param_owner]

    # ---- compilation ------------------------------------------------------------------------------------------
    def program(self) -> Program:
        if self._program is not None:
            return self._program
        n = self.nqubits
        prog = Program()
        prog.nqubits = n
        xs, zs, angles, prot, pscale, powner = [], [], [], [], [], []
        n_param_gates = 0
        phase = 1.0 + 0.0j
        basis_qubits = 0  # qubit-position mask of leading X gates (they only move the initial basis state: K0)
        leading = True
        measured: list[tuple[int, str]] = []
        for g in self.queue:
            if isinstance(g, _gates.M):
                for q, b in zip(g.qubits, g.basis):
                    if any(q == mq for mq, _ in measured):
                        raise ValueError(f"qubit {q} is measured twice")
                    measured.append((q, b))
                continue
            if measured:
                raise NotImplementedError("gates after a measurement are not supported")
            if leading and isinstance(g, _gates.X):
                basis_qubits ^= 1 << g.qubits[0]
                continue
            leading = False
            rots, ph = g.lower()
            phase *= ph
            for x, z, ang, pidx, scale in rots:
                if pidx is not None:
                    prot.append(len(xs))
                    pscale.append(scale)
                    powner.append(n_param_gates)
                xs.append(reverse_bits(x, n))
                zs.append(reverse_bits(z, n))
                angles.append(ang)
            if g.n_params:
                n_param_gates += 1
        prog.basis_index = reverse_bits(basis_qubits, n)
        prog.xmask = np.array(xs, dtype=np.uint64)
        prog.zmask = np.array(zs, dtype=np.uint64)
        prog.angle = np.array(angles, dtype=np.float64)
        prog.param_rot = np.array(prot, dtype=np.int64)
        prog.param_scale = np.array(pscale, dtype=np.float64)
        prog.param_owner = np.array(powner, dtype=np.int64)  # index into parametrized_gates / get_parameters()
        prog.n_param_gates = n_param_gates
        prog.phase = phase
        prog.measured = measured
        prog.x_basis_mask = sum(1 << (n - 1 - q) for q, b in measured if b == "X")
        prog.y_basis_mask = sum(1 << (n - 1 - q) for q, b in measured if b == "Y")
        self._program = prog
        return prog

    # ---- execution --------------------------------------------------------------------------------------------
    def run_on_device(self, initial_state=None):
        """K0 + K1 on the runtime's device state; returns the (still device-resident) ``DeviceState``."""
        prog = self.program()
        rt = get_runtime()
        st = rt.state(self.nqubits)
        if initial_state is None:
            if hasattr(st, "prepare_state"):  # one native call: K0 + K1 (a single launch on a small register)
                st.prepare_state(prog.basis_index, prog.xmask, prog.zmask, prog.angle)
                return st
            st.set_basis_state(prog.basis_index)
        else:
            psi = np.asarray(initial_state, dtype=np.complex128).reshape(-1)
            if psi.size != 1 << self.nqubits:
                raise ValueError("initial_state has the wrong size")
            if prog.basis_index:  # leading X gates act on the given state: permute amplitudes on the host copy
                psi = psi[np.arange(psi.size) ^ prog.basis_index]
            st.from_numpy(psi)
        if prog.xmask.size:
            st.apply_pauli_rotations(prog.xmask, prog.zmask, prog.angle)
        return st

    def __call__(self, initial_state=None, nshots: int | None = None, seed: int | None = None) -> "CircuitResult":
        st = self.run_on_device(initial_state)
        prog = self.program()
        res = CircuitResult(self, st, get_runtime().generation, prog.phase)
        if nshots and prog.measured:
            res._draw(int(nshots), seed)
        return res

    execute = __call__

    def draw(self) -> str:
        lines = [f"{q}: " + " ".join(repr(g) for g in self.queue if q in g.qubits) for q in range(self.nqubits)]
        out = "\n".join(lines)
        print(out)
        return out

    def summary(self) -> str:
        names = Counter(g.name for g in self.queue)
        return f"Circuit: nqubits={self.nqubits}, ngates={self.ngates}, " + ", ".join(f"{k}: {v}" for k, v in names.most_common())


_seed_counter = [0x9E3779B97F4A7C15]


def _next_seed() -> int:
    """Fresh Philox key per sampling call when the caller gives no seed (the reference does not seed either)."""
    import os

    _seed_counter[0] = (_seed_counter[0] * 6364136223846793005 + 1442695040888963407) % (1 << 64)
    seed = _seed_counter[0] ^ int.from_bytes(os.urandom(8), "little")
    if get_runtime().world > 1:  # every rank of a sharded run must draw from the SAME key: rank 0's
        import torch.distributed as dist

        box = [seed]
        dist.broadcast_object_list(box, src=0)
        seed = int(box[0])
    return seed


class CircuitResult:
    """What ``circuit()`` returns: amplitudes stay on the device until asked for."""

    def __init__(self, circuit: Circuit, state, generation: int, phase: complex):
        self.circuit = circuit
        self._state = state
        self._generation = generation
        self._phase = phase
        self._samples = None
        self.nshots = None

    def _check_alive(self):
        if get_runtime().generation != self._generation:
            raise RuntimeError("the device state of this result was overwritten by a later circuit execution")

    def state(self, numpy: bool = True) -> np.ndarray:
        self._check_alive()
        psi = self._state.to_numpy()
        return psi * self._phase if self._phase != 1.0 else psi

    def probabilities(self) -> np.ndarray:
        return np.abs(self.state()) ** 2

    def _draw(self, nshots: int, seed: int | None):
        prog = self.circuit.program()
        self._check_alive()
        self.nshots = nshots
        self._samples = self._state.sample(
            prog.x_basis_mask, prog.y_basis_mask, nshots, _next_seed() if seed is None else seed, on_device=True
        )

    def device_samples(self):
        """torch int64 CUDA tensor of full-register outcome indices (bit n-1-q <-> qubit q)."""
        return self._samples

    def samples(self, binary: bool = True):
        """Shots over the measured qubits (in gate order), as 0/1 arrays (binary) or integers."""
        prog = self.circuit.program()
        n = prog.nqubits
        idx = self._samples.cpu().numpy().astype(np.uint64)
        bits = np.stack([(idx >> np.uint64(n - 1 - q)) & np.uint64(1) for q, _ in prog.measured], axis=1).astype(np.int64)
        if binary:
            return bits
        m = bits.shape[1]
        return (bits << np.arange(m - 1, -1, -1)).sum(axis=1)

    def frequencies(self, binary: bool = True) -> Counter:
        """``Counter`` over the measured qubits; character k of a key belongs to the k-th measured qubit
        (qibo's ``frequencies(binary=True)``).  Counting happens on the device (K3 histogram)."""
        if self._samples is None:
            return Counter()
        prog = self.circuit.program()
        n = prog.nqubits
        qubits = [q for q, _ in prog.measured]
        keep = sum(1 << (n - 1 - q) for q in qubits)
        keys, counts = self._state.frequencies(self._samples, keep)
        # compact key bit j <-> j-th lowest kept index bit <-> j-th HIGHEST measured qubit
        desc = sorted(qubits, reverse=True)
        pos = {q: j for j, q in enumerate(desc)}
        out = Counter()
        for k, c in zip(keys.tolist(), counts.tolist()):
            chars = "".join("1" if (k >> pos[q]) & 1 else "0" for q in qubits)
            out[chars if binary else int(chars, 2)] += c
        return out
