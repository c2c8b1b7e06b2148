"""Host logic of the sharded path on CPU: world_size 2 and 4 over gloo, numpy shards with the oracle's arithmetic.
Checks layout bookkeeping, Belady eviction, rank-sign folding and the all-reduce against the unsharded oracle."""

import os

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import oracle as O


def _inputs(n, seed):
    """Seeded inputs shared by the workers and the reference; x-masks pair on <= 4 bits like UCCSD / molecular terms."""
    rng = np.random.default_rng(seed)
    hf = int(rng.integers(0, 2**n))
    R, T = 40, 60

    def xmasks(count):
        out = np.zeros(count, dtype=np.uint64)
        for k in range(count):
            w = int(rng.integers(0, 5))
            for b in rng.choice(n, size=w, replace=False):
                out[k] |= np.uint64(1 << int(b))
        return out

    xs = xmasks(R)
    xs[1::7] |= np.uint64(1 << (n - 1))  # force pairing on the top (global) bit
    xs[1::7] &= ~np.uint64(1)
    zs = rng.integers(0, 2**n, size=R, dtype=np.uint64)
    ang = rng.uniform(-2, 2, size=R)
    hx = xmasks(T)
    hz = rng.integers(0, 2**n, size=T, dtype=np.uint64)
    hc = rng.normal(size=T)
    return hf, xs, zs, ang, hx, hz, hc


def _worker(rank, world, port, n, seed, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qibochem_b200.sharded import ShardedState
        from tests.sharded_helpers import GlooExchanger, OracleShard

        g = world.bit_length() - 1
        shard = OracleShard(n - g)
        st = ShardedState(n, rank, world, shard, GlooExchanger(shard, rank, world))
        hf, xs, zs, ang, hx, hz, hc = _inputs(n, seed)
        st.set_basis_state(hf)
        st.apply_pauli_rotations(xs, zs, ang, lookahead=8)
        norm = st.norm2()
        e = st.expectation(hx, hz, hc)
        full = st.gather_logical()
        # more rotations after the relayout done by the expectation
        st.apply_pauli_rotations(xs[:10], zs[:10], ang[:10])
        full2 = st.gather_logical()
        e2 = st.expectation(hx, hz, hc)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[e, e2, norm, st.n_swaps], full, full2]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,seed", [(2, 6, 0), (2, 9, 1), (4, 8, 2), (4, 7, 3)])
def test_sharded_matches_unsharded_oracle(tmp_path, world, n, seed):
    port = 29500 + (os.getpid() + seed * 17 + world) % 2000
    mp.spawn(_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    hf, xs, zs, ang, hx, hz, hc = _inputs(n, seed)
    psi = np.zeros(2**n, dtype=complex)
    psi[hf] = 1
    for x, z, a in zip(xs.tolist(), zs.tolist(), ang.tolist()):
        psi = O.apply_pauli_rotation(psi, x, z, a)
    e_ref = O.expectation_masks(psi, hx.tolist(), hz.tolist(), hc.tolist())
    psi2 = psi.copy()
    for x, z, a in zip(xs[:10].tolist(), zs[:10].tolist(), ang[:10].tolist()):
        psi2 = O.apply_pauli_rotation(psi2, x, z, a)
    e2_ref = O.expectation_masks(psi2, hx.tolist(), hz.tolist(), hc.tolist())
    dim = 2**n
    for r in range(world):
        data = np.load(os.path.join(str(tmp_path), f"r{r}.npy"))
        e, e2, norm, swaps = data[0].real, data[1].real, data[2].real, int(data[3].real)
        assert abs(e - e_ref) < 1e-12 and abs(e2 - e2_ref) < 1e-12
        assert abs(norm - 1.0) < 1e-13
        assert swaps > 0
        assert np.abs(data[4 : 4 + dim] - psi).max() < 1e-13
        assert np.abs(data[4 + dim :] - psi2).max() < 1e-13
