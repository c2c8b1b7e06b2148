"""
Process-wide runtime: which GPU this process drives, whether it is one rank of a sharded run, and the device-resident
state vectors it keeps alive.

One process drives one GPU.  A multi-GPU run is ONE PROCESS PER GPU (``torchrun --nproc-per-node N script.py``, every
rank executes the same script): ``set_world()`` -- or the first ``Circuit(n, accelerators=...)``, the reference's only
multi-device hook (its builders forward ``**kwargs`` into ``qibo.Circuit``: ``src/qibochem/ansatz/ansatz.py:120``,
``:341``; ``_ansatz.py:81``) -- makes ``state(n)`` hand out a ``ShardedState`` over all ranks instead of a
``DeviceState``, and every public call (``circuit()``, ``hamiltonian.expectation(circuit)``,
``expectation_from_samples``, ``adjoint_gradient``) runs sharded with identical results on every rank.

State vectors are cached per qubit count so that a VQE loop re-uses the same HBM allocation for every evaluation.
"""

from __future__ import annotations

import os

from qibochem_b200.engine import DeviceState

_LARGE = 26  # from 2^26 amplitudes (1 GiB) on, keep a single cached state


class Runtime:
    def __init__(self):
        self.device = 0
        self.rank, self.world = 0, 1
        self._owns_process_group = False
        self._states: dict[int, object] = {}
        self._context: DeviceState | None = None
        self.generation = 0  # bumped every time a cached state is re-initialised (invalidates lazy results)

    # ---- configuration ----------------------------------------------------------------------------------------
    def set_device(self, device: int) -> None:
        if device != self.device:
            self.release()
            self.device = int(device)

    def set_world(self, rank: int | None = None, world: int | None = None, device: int | None = None) -> None:
        """Declare this process to be rank ``rank`` of ``world`` (a power of two) ranks, one GPU each.  Defaults come
        from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK); the NCCL process group is created if the script
        has not done so itself."""
        rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        device = int(os.environ.get("LOCAL_RANK", str(rank))) if device is None else int(device)
        if world < 1 or world & (world - 1):
            raise ValueError("the number of ranks must be a power of two")
        if (rank, world, device) == (self.rank, self.world, self.device):
            return
        self.release()
        self.rank, self.world, self.device = rank, world, device
        if world > 1:
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(device)
            if not dist.is_initialized():
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29512")
                dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{device}"))
                self._owns_process_group = True

    def configure_accelerators(self, accelerators) -> None:
        """``Circuit(n, accelerators={"/GPU:0": 1, "/GPU:1": 1})`` (qibo's multi-device argument): the total device count
        must equal the number of ranks this script was launched with."""
        wanted = sum(int(v) for v in accelerators.values()) if isinstance(accelerators, dict) else int(accelerators)
        env_world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world == 1 and env_world > 1:
            self.set_world()
        if wanted != self.world:
            raise RuntimeError(
                f"accelerators asks for {wanted} devices but this process is rank {self.rank} of {self.world}: multi-GPU "
                f"runs are one process per GPU -- launch the script with `torchrun --nproc-per-node {wanted}`"
            )

    # ---- states -----------------------------------------------------------------------------------------------
    def state(self, n_qubits: int):
        """The cached state of ``n_qubits`` qubits: a ``DeviceState``, or a ``ShardedState`` when world > 1.  Asking
        for a register of >= 26 qubits releases every other cached state first (one large allocation at a time);
        results that still point at a released state raise on use (``generation``)."""
        st = self._states.get(n_qubits)
        if st is None:
            if n_qubits >= _LARGE:
                self.release_states()
            st = self._new_state(n_qubits)
            self._states[n_qubits] = st
        self.generation += 1
        return st

    def _new_state(self, n_qubits: int):
        if self.world == 1:
            return DeviceState(n_qubits, device=self.device)
        from qibochem_b200.sharded import IpcExchanger, ShardedState

        g = self.world.bit_length() - 1
        if n_qubits - g < 1:
            raise ValueError(f"a {n_qubits}-qubit register cannot be sharded over {self.world} ranks")
        local = DeviceState(n_qubits - g, device=self.device)
        return ShardedState(n_qubits, self.rank, self.world, local, IpcExchanger(local, self.rank, self.world, self.device),
                            min_swap_pos=min(8, max(0, n_qubits - g - 4)))  # fmt: skip

    def adopt(self, state) -> None:
        """Make an existing ``DeviceState`` / ``ShardedState`` the cached state of its qubit count (a caller that
        allocated the shard itself, e.g. on its own stream, shares it with the public API this way)."""
        old = self._states.get(state.n_qubits)
        if old is not None and old is not state:
            old.close()
        self._states[state.n_qubits] = state
        self.generation += 1

    def forget(self, state) -> None:
        """Undo ``adopt`` without closing the state (its owner keeps it)."""
        if self._states.get(state.n_qubits) is state:
            del self._states[state.n_qubits]
            self.generation += 1

    def context_state(self) -> DeviceState:
        """A 1-qubit local state: device context for calls that only move host arrays through a kernel (K3 on
        frequencies the caller supplies)."""
        if self._context is None:
            self._context = DeviceState(1, device=self.device)
        return self._context

    def release_states(self) -> None:
        for st in self._states.values():
            st.close()
        self._states.clear()
        self.generation += 1

    def release(self) -> None:
        self.release_states()
        if self._context is not None:
            self._context.close()
            self._context = None


_runtime = Runtime()


def get_runtime() -> Runtime:
    return _runtime


def set_device(device: int) -> None:
    _runtime.set_device(device)


def set_world(rank: int | None = None, world: int | None = None, device: int | None = None) -> None:
    _runtime.set_world(rank, world, device)
