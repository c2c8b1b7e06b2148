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


def _sampling_worker(rank, world, port, n, seed, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qibochem_b200.sharded import ShardedState
        from tests.sharded_helpers import GlooExchanger, OracleShard

        g = world.bit_length() - 1
        shard = OracleShard(n - g)
        st = ShardedState(n, rank, world, shard, GlooExchanger(shard, rank, world))
        rng = np.random.default_rng(100 + seed)
        psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
        psi /= np.linalg.norm(psi)
        st.from_numpy(psi)
        xb = (1 << (n - 1)) | 0b10  # X basis on the TOP (global) bit and on a local bit: forces a re-layout
        yb = 1 << (n - 2) if world > 2 else 0b100
        shots = 4000
        samples, total = st.sample(xb, yb, shots, 77 + seed, return_total=True)
        masks = rng.integers(0, 2**n, size=9, dtype=np.uint64)
        sums = st.parity_sums(samples, masks)
        idx = samples.logical_indices_all_ranks()
        keys, counts = st.frequencies(samples, (1 << n) - 1)
        np.savez(os.path.join(out_dir, f"s{rank}.npz"), sums=sums, idx=idx, keys=keys, counts=counts, total=total,
                 masks=masks, psi=psi, swaps=st.n_swaps, gathered=st.gather_logical())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,seed", [(2, 7, 0), (4, 8, 1)])
def test_sharded_sampling_path(tmp_path, world, n, seed):
    """SURVEY 8e split of the sampled path: per-rank mass -> multinomial -> per-rank shots -> integer partial sums ->
    all-reduce.  Checks: the shots are valid draws of the rotated distribution (chi^2-free: exact parity sums of the
    gathered shots == all-reduced K3 partials, bit for bit), every rank agrees, the state is untouched, and the
    empirical distribution matches |<i|B|psi>|^2."""
    port = 29500 + (os.getpid() + 31 * seed + world + 7) % 2000
    mp.spawn(_sampling_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    data = [np.load(os.path.join(str(tmp_path), f"s{r}.npz")) for r in range(world)]
    d0 = data[0]
    psi, masks, idx = d0["psi"], d0["masks"], d0["idx"]
    assert idx.size == 4000 and abs(float(d0["total"]) - 1.0) < 1e-12
    for d in data:
        assert np.array_equal(d["sums"], d0["sums"]) and np.array_equal(d["idx"], idx)
        assert np.abs(d["gathered"] - psi).max() < 1e-15  # the re-layout moved amplitudes, not values
        assert int(d["swaps"]) > 0
    assert np.array_equal(d0["sums"], O.parity_sums(idx, masks))  # integer partials + rank signs == direct sums
    keys, counts = d0["keys"], d0["counts"]
    uk, uc = np.unique(idx, return_counts=True)
    assert np.array_equal(keys, uk) and np.array_equal(counts, uc)
    xb = (1 << (n - 1)) | 0b10
    yb = 1 << (n - 2) if world > 2 else 0b100
    basis = {n - 1 - b: ("X" if (xb >> b) & 1 else "Y") for b in range(n) if ((xb | yb) >> b) & 1}
    p = np.abs(O.run_gates(n, O.measurement_basis_gates(basis), psi.copy())) ** 2
    emp = np.bincount(idx.astype(np.int64), minlength=2**n) / idx.size
    assert np.abs(emp - p).max() < 5 * np.sqrt(p.max() / idx.size) + 1e-3


def _gradient_worker(rank, world, port, n, seed, out_dir):
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
        xs, zs, ang = xs[:14], zs[:14], ang[:14]
        st.set_basis_state(hf)
        st.apply_pauli_rotations(xs, zs, ang, lookahead=8)
        e, grad = st.energy_gradient(xs, zs, ang, hx, hz, hc, lookahead=8)
        back = st.gather_logical()  # the program has been un-applied: |hf> again
        np.save(os.path.join(out_dir, f"g{rank}.npy"), np.concatenate([[e, st.n_swaps], grad, back]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,seed", [(2, 6, 0), (4, 7, 3)])
def test_sharded_adjoint_gradient(tmp_path, world, n, seed):
    """lambda = H psi built across layouts, psi and lambda exchanged together, backward walk in local segments:
    equals the parameter-shift gradient of the unsharded oracle."""
    port = 29500 + (os.getpid() + 13 * seed + world + 91) % 2000
    mp.spawn(_gradient_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    hf, xs, zs, ang, hx, hz, hc = _inputs(n, seed)
    xs, zs, ang = xs[:14], zs[:14], ang[:14]

    def energy(angles):
        psi = np.zeros(2**n, dtype=complex)
        psi[hf] = 1
        for x, z, a in zip(xs.tolist(), zs.tolist(), angles.tolist()):
            psi = O.apply_pauli_rotation(psi, x, z, a)
        return O.expectation_masks(psi, hx.tolist(), hz.tolist(), hc.tolist())

    e_ref = energy(ang)
    g_ref = np.zeros(ang.size)
    for r in range(ang.size):
        up, dn = ang.copy(), ang.copy()
        up[r] += np.pi / 2
        dn[r] -= np.pi / 2
        g_ref[r] = 0.5 * (energy(up) - energy(dn))
    dim = 2**n
    for r in range(world):
        data = np.load(os.path.join(str(tmp_path), f"g{r}.npy"))
        assert abs(data[0].real - e_ref) < 1e-12
        assert np.abs(data[2 : 2 + ang.size].real - g_ref).max() < 1e-11
        back = data[2 + ang.size :]
        assert abs(abs(back[hf]) - 1.0) < 1e-12 and np.abs(np.delete(back, hf)).max() < 1e-12
        assert len(back) == dim
