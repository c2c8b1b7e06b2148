"""CPU stand-ins used ONLY by the tests of the sharding host logic (gloo, no GPU): a numpy shard that executes the
DeviceState API with the oracle's arithmetic, and an exchanger that moves half-shards with gloo send/recv."""

import numpy as np
import torch
import torch.distributed as dist

from oracle import oracle as O


class OracleShard:
    def __init__(self, n_local: int):
        self.n = n_local
        self.psi = np.zeros(1 << n_local, dtype=np.complex128)
        self.last_passes = 0
        self.last_sweeps = 0

    def set_basis_state(self, index, has_one=True):
        self.psi[:] = 0
        if has_one:
            self.psi[index] = 1.0

    def apply_pauli_rotations(self, x, z, a):
        for xi, zi, ai in zip(np.asarray(x).tolist(), np.asarray(z).tolist(), np.asarray(a).tolist()):
            self.psi = O.apply_pauli_rotation(self.psi, int(xi), int(zi), float(ai))
        self.last_passes = len(np.asarray(x))
        return self.last_passes

    def expectation(self, x, z, c):
        self.last_sweeps = len(set(np.asarray(x).tolist()))
        return O.expectation_masks(self.psi, np.asarray(x).tolist(), np.asarray(z).tolist(), np.asarray(c).tolist())

    def norm2(self):
        return float(np.vdot(self.psi, self.psi).real)

    def to_numpy(self):
        return self.psi.copy()


class GlooExchanger:
    def __init__(self, shard: OracleShard, rank: int, world: int):
        self.shard, self.rank, self.world = shard, rank, world

    def barrier(self):
        dist.barrier()

    def swap(self, global_k: int, local_pos: int):
        partner = self.rank ^ (1 << global_k)
        side = (self.rank >> global_k) & 1
        idx = np.arange(self.shard.psi.size)
        sel = ((idx >> local_pos) & 1) == (1 - side)  # side 0 sends its bit=1 half, side 1 its bit=0 half
        send = torch.from_numpy(np.ascontiguousarray(self.shard.psi[sel]).view(np.float64))
        recv = torch.empty_like(send)
        if side == 0:
            dist.send(send, partner)
            dist.recv(recv, partner)
        else:
            dist.recv(recv, partner)
            dist.send(send, partner)
        self.shard.psi[sel] = recv.numpy().view(np.complex128)

    def allreduce_sum(self, value: float) -> float:
        t = torch.tensor([value], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def allgather_array(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.float64))
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return np.concatenate([o.numpy().view(np.complex128) for o in out])
