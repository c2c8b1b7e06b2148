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

    def sample(self, x_basis_mask, y_basis_mask, n_shots, seed, shot_offset=0, on_device=False, return_total=False):
        """DeviceState.sample's contract with the oracle's arithmetic: rotate a copy, inverse CDF on Philox uniforms
        whose counter is shot index + shot_offset, normalised by the SHARD's own mass."""
        n = self.n
        basis = {n - 1 - b: ("X" if (x_basis_mask >> b) & 1 else "Y") for b in range(n) if ((x_basis_mask | y_basis_mask) >> b) & 1}
        rot = O.run_gates(n, O.measurement_basis_gates(basis), self.psi.copy()) if basis else self.psi
        cdf = np.cumsum(np.abs(rot) ** 2)
        total = float(cdf[-1])
        u = O.philox_uniform53(int(n_shots), int(seed), shot_offset=int(shot_offset))
        idx = np.minimum(np.searchsorted(cdf, u * total, side="right"), (1 << n) - 1).astype(np.uint64)
        return (idx, total) if return_total else idx

    def parity_sums(self, samples, masks, counts=None):
        return O.parity_sums(np.asarray(samples, dtype=np.uint64), np.asarray(masks, dtype=np.uint64))

    def from_numpy(self, psi):
        self.psi = np.array(psi, dtype=np.complex128)

    # adjoint gradient stages (contract of qc_hpsi / qc_lambda_inner / qc_gradient_backward)
    def hpsi(self, x, z, c, accumulate=False):
        if not accumulate or not hasattr(self, "lam"):
            self.lam = np.zeros_like(self.psi)
        for xi, zi, ci in zip(np.asarray(x).tolist(), np.asarray(z).tolist(), np.asarray(c).tolist()):
            self.lam = self.lam + ci * O.apply_pauli(self.psi, int(xi), int(zi))

    def lambda_inner(self):
        return complex(np.vdot(self.lam, self.psi))

    def gradient_backward(self, x, z, a):
        x, z, a = np.asarray(x).tolist(), np.asarray(z).tolist(), np.asarray(a).tolist()
        grad = np.zeros(len(x))
        for r in range(len(x) - 1, -1, -1):
            grad[r] = np.vdot(self.lam, O.apply_pauli(self.psi, int(x[r]), int(z[r]))).imag
            self.psi = O.apply_pauli_rotation(self.psi, int(x[r]), int(z[r]), -float(a[r]))
            self.lam = O.apply_pauli_rotation(self.lam, int(x[r]), int(z[r]), -float(a[r]))
        return grad

    def to_numpy(self):
        return self.psi.copy()


class GlooExchanger:
    def __init__(self, shard: OracleShard, rank: int, world: int):
        self.shard, self.rank, self.world = shard, rank, world

    def barrier(self):
        dist.barrier()

    def swap(self, global_k: int, local_pos: int, with_scratch: bool = False):
        partner = self.rank ^ (1 << global_k)
        side = (self.rank >> global_k) & 1
        idx = np.arange(self.shard.psi.size)
        sel = ((idx >> local_pos) & 1) == (1 - side)  # side 0 sends its bit=1 half, side 1 its bit=0 half
        for name in ("psi", "lam") if with_scratch else ("psi",):
            arr = getattr(self.shard, name)
            send = torch.from_numpy(np.ascontiguousarray(arr[sel]).view(np.float64))
            recv = torch.empty_like(send)
            if side == 0:
                dist.send(send, partner)
                dist.recv(recv, partner)
            else:
                dist.recv(recv, partner)
                dist.send(send, partner)
            arr[sel] = recv.numpy().view(np.complex128)

    def allreduce_array(self, values):
        t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    def allreduce_sum(self, value: float) -> float:
        t = torch.tensor([value], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def allreduce_int64(self, values):
        t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.int64))
        dist.all_reduce(t)
        return t.numpy()

    def allgather_floats(self, value):
        t = torch.tensor([value], dtype=torch.float64)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def allgather_object(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def allgather_array(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.float64))
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return np.concatenate([o.numpy().view(np.complex128) for o in out])
