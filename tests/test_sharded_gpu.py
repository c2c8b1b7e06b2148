"""Sharded path on the GPU.  (1) Two / four ranks emulated as threads of one process on ONE GPU: every rank owns a
DeviceState and the real peer-swap kernel runs on the other rank's device pointer.  (2) With >= 2 GPUs: the real
one-process-per-GPU launch (torchrun, NCCL, CUDA IPC)."""

import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from oracle import oracle as O
from tests.test_sharded_cpu import _inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ThreadExchanger:
    """All ranks live in one process; collectives are a threading.Barrier plus shared slots."""

    def __init__(self, rank, world, states, barrier, slots):
        self.rank, self.world, self.states, self.bar, self.slots = rank, world, states, barrier, slots

    def barrier(self):
        self.states[self.rank].synchronize()
        self.bar.wait()

    def swap(self, global_k, local_pos, with_scratch=False):
        partner = self.rank ^ (1 << global_k)
        side = (self.rank >> global_k) & 1
        self.barrier()
        self.states[self.rank].swap_with_peer(self.states[partner].device_ptr, local_pos, side, side)
        if with_scratch:
            self.states[self.rank].swap_scratch_with_peer(self.states[partner].scratch_device_ptr, local_pos, side, side)
        self.barrier()

    def allreduce_array(self, values):
        self.slots[self.rank] = np.array(values, dtype=np.float64)
        self.bar.wait()
        total = sum(self.slots[r] for r in range(self.world))
        self.bar.wait()
        return total

    def allreduce_sum(self, value):
        self.slots[self.rank] = value
        self.bar.wait()
        total = sum(self.slots[r] for r in range(self.world))  # fixed order on every rank
        self.bar.wait()
        return total

    def allgather_array(self, arr):
        self.slots[self.rank] = arr
        self.bar.wait()
        out = np.concatenate([self.slots[r] for r in range(self.world)])
        self.bar.wait()
        return out

    def allgather_object(self, obj):
        self.slots[self.rank] = obj
        self.bar.wait()
        out = [self.slots[r] for r in range(self.world)]
        self.bar.wait()
        return out

    def allgather_floats(self, value):
        return [float(v) for v in self.allgather_object(float(value))]

    def allreduce_int64(self, values):
        parts = self.allgather_object(np.asarray(values, dtype=np.int64))
        return np.sum(parts, axis=0).astype(np.int64)


@pytest.mark.parametrize("world,n,seed", [(2, 10, 0), (4, 11, 1), (8, 12, 2)])
def test_sharded_threads_one_gpu(native_lib, world, n, seed):
    from qibochem_b200.engine import DeviceState
    from qibochem_b200.sharded import ShardedState

    g = world.bit_length() - 1
    hf, xs, zs, ang, hx, hz, hc = _inputs(n, seed)
    states = [DeviceState(n - g) for _ in range(world)]
    bar = threading.Barrier(world)
    slots = [None] * world
    results, errors = [None] * world, []

    def run(rank):
        try:
            st = ShardedState(n, rank, world, states[rank], ThreadExchanger(rank, world, states, bar, slots))
            st.set_basis_state(hf)
            st.apply_pauli_rotations(xs, zs, ang, lookahead=8)
            e = st.expectation(hx, hz, hc)
            full = st.gather_logical()
            st.apply_pauli_rotations(xs[:10], zs[:10], ang[:10])
            e2 = st.expectation(hx, hz, hc)
            results[rank] = (e, e2, full, st.n_swaps, st.norm2())
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
            bar.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    psi = np.zeros(2**n, dtype=complex)
    psi[hf] = 1
    for x, z, a in zip(xs.tolist(), zs.tolist(), ang.tolist()):
        psi = O.apply_pauli_rotation(psi, x, z, a)
    e_ref = O.expectation_masks(psi, hx.tolist(), hz.tolist(), hc.tolist())
    psi2 = psi.copy()
    for x, z, a in zip(xs[:10].tolist(), zs[:10].tolist(), ang[:10].tolist()):
        psi2 = O.apply_pauli_rotation(psi2, x, z, a)
    e2_ref = O.expectation_masks(psi2, hx.tolist(), hz.tolist(), hc.tolist())
    for e, e2, full, swaps, norm in results:
        assert abs(e - e_ref) < 1e-11 and abs(e2 - e2_ref) < 1e-11
        assert np.abs(full - psi).max() < 1e-12
        assert swaps > 0 and abs(norm - 1) < 1e-12
    for s in states:
        s.close()


@pytest.mark.parametrize("world,n,seed", [(2, 12, 0), (4, 13, 1)])
def test_sharded_sampling_and_gradient_threads_one_gpu(native_lib, world, n, seed):
    """The sampled path and the adjoint gradient of ShardedState with REAL device shards (ranks emulated as threads of
    one process on one GPU): per-rank probability mass -> shot split -> qc_sample(shot_offset) -> K3 partial sums with
    rank signs; lambda = H psi across layouts with psi AND lambda exchanged (qc_swap_scratch_with_peer), backward walk
    in local segments.  Checked against the oracle."""
    from qibochem_b200.engine import DeviceState
    from qibochem_b200.sharded import ShardedState

    g = world.bit_length() - 1
    hf, xs, zs, ang, hx, hz, hc = _inputs(n, seed)
    xs, zs, ang = xs[:16], zs[:16], ang[:16]
    rng = np.random.default_rng(50 + seed)
    masks = rng.integers(0, 2**n, size=11, dtype=np.uint64)
    xb, yb = (1 << (n - 1)) | 0b100, 1 << (n - 2)  # X / Y measurement bases on the two top (global) bits and a local one
    shots = 6000
    states = [DeviceState(n - g) for _ in range(world)]
    bar = threading.Barrier(world)
    slots = [None] * world
    results, errors = [None] * world, []

    def run(rank):
        try:
            st = ShardedState(n, rank, world, states[rank], ThreadExchanger(rank, world, states, bar, slots))
            st.set_basis_state(hf)
            st.apply_pauli_rotations(xs, zs, ang, lookahead=8)
            psi = st.gather_logical()
            samples, total = st.sample(xb, yb, shots, 91 + seed, return_total=True)
            sums = st.parity_sums(samples, masks)
            idx = samples.logical_indices_all_ranks()
            after = st.gather_logical()
            e, grad = st.energy_gradient(xs, zs, ang, hx, hz, hc, lookahead=8)
            back = st.gather_logical()
            results[rank] = (psi, total, sums, idx, after, e, grad, back)
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
            bar.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    ref = np.zeros(2**n, dtype=complex)
    ref[hf] = 1

    def energy(angles):
        v = ref.copy()
        for x, z, a in zip(xs.tolist(), zs.tolist(), angles.tolist()):
            v = O.apply_pauli_rotation(v, x, z, a)
        return v, O.expectation_masks(v, hx.tolist(), hz.tolist(), hc.tolist())

    psi_ref, e_ref = energy(ang)
    basis = {n - 1 - b: ("X" if (xb >> b) & 1 else "Y") for b in range(n) if ((xb | yb) >> b) & 1}
    p = np.abs(O.run_gates(n, O.measurement_basis_gates(basis), psi_ref.copy())) ** 2
    for psi, total, sums, idx, after, e, grad, back in results:
        assert np.abs(psi - psi_ref).max() < 1e-12 and np.abs(after - psi_ref).max() < 1e-12  # sampling leaves psi alone
        assert abs(total - 1.0) < 1e-12 and idx.size == shots
        assert np.array_equal(sums, O.parity_sums(idx, masks))  # integer partial sums + rank signs == direct sums
        assert np.array_equal(idx, results[0][3])
        emp = np.bincount(idx.astype(np.int64), minlength=2**n) / shots
        assert np.abs(emp - p).max() < 5 * np.sqrt(p.max() / shots) + 2e-3
        assert abs(e - e_ref) < 1e-11
        assert abs(abs(back[hf]) - 1.0) < 1e-12  # the program has been un-applied
    for k in (0, 5, 15):
        up, dn = ang.copy(), ang.copy()
        up[k] += np.pi / 2
        dn[k] -= np.pi / 2
        assert abs(results[0][6][k] - 0.5 * (energy(up)[1] - energy(dn)[1])) < 1e-10
    for s in states:
        s.close()


def test_sharded_processes_two_gpus(native_lib):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "sharded_gpu_worker.py")]  # fmt: skip
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED_OK" in out.stdout
