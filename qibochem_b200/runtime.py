"""
Process-wide runtime: which GPU this process drives and the device-resident state vectors it keeps alive.

One process drives one GPU (multi-GPU runs use one process per GPU, see ``qibochem_b200.sharded``).  State vectors
are cached per qubit count so that a VQE loop re-uses the same HBM allocation for every energy evaluation.
"""

from __future__ import annotations

from qibochem_b200.engine import DeviceState

_LARGE = 26  # from 2^26 amplitudes (1 GiB) on, keep a single cached state


class Runtime:
    def __init__(self):
        self.device = 0
        self._states: dict[int, DeviceState] = {}
        self.generation = 0  # bumped every time a cached state is re-initialised (invalidates lazy results)

    def set_device(self, device: int) -> None:
        if device != self.device:
            self.release()
            self.device = int(device)

    def state(self, n_qubits: int) -> DeviceState:
        st = self._states.get(n_qubits)
        if st is None:
            if n_qubits >= _LARGE:
                self.release()
            st = DeviceState(n_qubits, device=self.device)
            self._states[n_qubits] = st
        self.generation += 1
        return st

    def release(self) -> None:
        for st in self._states.values():
            st.close()
        self._states.clear()


_runtime = Runtime()


def get_runtime() -> Runtime:
    return _runtime


def set_device(device: int) -> None:
    _runtime.set_device(device)
