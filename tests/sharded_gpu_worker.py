"""torchrun worker: one process per GPU, NCCL + CUDA IPC peer swaps, checked against the oracle (n = 14)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from oracle import oracle as O
    from qibochem_b200.engine import DeviceState
    from qibochem_b200.sharded import IpcExchanger, ShardedState
    from tests.test_sharded_cpu import _inputs

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    n = 14
    g = world.bit_length() - 1
    hf, xs, zs, ang, hx, hz, hc = _inputs(n, 5)
    local = DeviceState(n - g, device=local_rank)
    ex = IpcExchanger(local, rank, world, local_rank)
    st = ShardedState(n, rank, world, local, ex)
    st.set_basis_state(hf)
    st.apply_pauli_rotations(xs, zs, ang, lookahead=8)
    e = st.expectation(hx, hz, hc)
    full = st.gather_logical()
    psi = np.zeros(2**n, dtype=complex)
    psi[hf] = 1
    for x, z, a in zip(xs.tolist(), zs.tolist(), ang.tolist()):
        psi = O.apply_pauli_rotation(psi, x, z, a)
    e_ref = O.expectation_masks(psi, hx.tolist(), hz.tolist(), hc.tolist())
    assert abs(e - e_ref) < 1e-11, (e, e_ref)
    assert np.abs(full - psi).max() < 1e-12
    assert st.n_swaps > 0
    ex.close()
    local.close()
    public_api(rank, world, local_rank)
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_OK swaps", st.n_swaps)


def public_api(rank, world, local_rank):
    """The reference's own surface under torchrun: `accelerators` forwarded through the builders' **kwargs
    (src/qibochem/ansatz/ansatz.py:120, :341), exact and sampled expectation, frequencies, adjoint gradient."""
    import torch.distributed as dist

    from oracle import oracle as O
    from qibochem_b200 import SymbolicHamiltonian, gates, synthetic
    from qibochem_b200.ansatz import hf_circuit, ucc_circuit
    from qibochem_b200.derivative import adjoint_gradient
    from qibochem_b200.measurement import expectation_from_samples
    from qibochem_b200.runtime import get_runtime
    from qibochem_b200.sharded import ShardedState

    n, ne = 12, 4
    acc = {f"/GPU:{r}": 1 for r in range(world)}
    c = hf_circuit(n, ne, accelerators=acc)
    for exc, th in (([0, 1, 4, 5], 0.3), ([2, 3, 10, 11], -0.2), ([0, 11], 0.4), ([1, 3, 6, 9], 0.15)):
        c += ucc_circuit(n, exc, th, accelerators=acc)
    assert get_runtime().world == world
    hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, 700)
    ham = SymbolicHamiltonian.from_index_masks(n, hx, hz, hc, constant=const)
    prog = c.program()
    psi = np.zeros(2**n, dtype=complex)
    psi[prog.basis_index] = 1
    for x, z, a in zip(prog.xmask.tolist(), prog.zmask.tolist(), prog.angle.tolist()):
        psi = O.apply_pauli_rotation(psi, x, z, a)
    e_ref = O.expectation_masks(psi, hx.tolist(), hz.tolist(), hc.tolist(), const)
    e = ham.expectation(c)
    assert isinstance(get_runtime().state(n), ShardedState)
    assert abs(e - e_ref) < 1e-11, (e, e_ref)
    assert np.abs(c().state() - psi).max() < 1e-12
    # VQE-style loop: new parameters, same compiled program
    params = np.array([complex(p[0]).real for p in c.get_parameters()], dtype=float) * 0.7  # ucc parameters are complex-typed
    c.set_parameters(params)
    prog = c.program()
    psi2 = np.zeros(2**n, dtype=complex)
    psi2[prog.basis_index] = 1
    for x, z, a in zip(prog.xmask.tolist(), prog.zmask.tolist(), prog.angle.tolist()):
        psi2 = O.apply_pauli_rotation(psi2, x, z, a)
    e2_ref = O.expectation_masks(psi2, hx.tolist(), hz.tolist(), hc.tolist(), const)
    assert abs(ham.expectation(c) - e2_ref) < 1e-11
    # adjoint gradient == finite differences of the oracle energy (two parameters)
    e_adj, grad = adjoint_gradient(c, ham)
    assert abs(e_adj - e2_ref) < 1e-11
    for k in (0, len(params) - 1):
        def energy(shift):
            q = np.zeros(2**n, dtype=complex)
            q[prog.basis_index] = 1
            ang = prog.angle.copy()
            ang[prog.param_rot[prog.param_owner == k]] += shift * prog.param_scale[prog.param_owner == k]
            for x, z, a in zip(prog.xmask.tolist(), prog.zmask.tolist(), ang.tolist()):
                q = O.apply_pauli_rotation(q, x, z, a)
            return O.expectation_masks(q, hx.tolist(), hz.tolist(), hc.tolist(), const)
        fd = (energy(1e-5) - energy(-1e-5)) / 2e-5
        assert abs(grad[k] - fd) < 1e-7, (k, grad[k], fd)
    # sampled path: QWC groups, X/Y bases on global qubits included; statistical tolerance of the reference's tests
    small = np.sort(np.random.default_rng(3).choice(hx.size, size=60, replace=False))
    ham_s = SymbolicHamiltonian.from_index_masks(n, hx[small], hz[small], hc[small], constant=0.25)
    exact = ham_s.expectation(c)
    sampled = expectation_from_samples(c, ham_s, n_shots=40_000, grouping="qwc", seed=11)
    assert abs(sampled - exact) < 0.02, (sampled, exact)
    # frequencies over measured qubits (one of them global, in the X basis)
    cm = c.copy()
    cm.add(gates.M(0, basis=gates.X))
    cm.add(gates.M(5, 7))
    freq = cm(nshots=5000, seed=3).frequencies(binary=True)
    assert sum(freq.values()) == 5000 and all(len(k) == 3 for k in freq)
    basis_state = O.run_gates(n, O.measurement_basis_gates({0: "X"}), psi2.copy())
    p = (np.abs(basis_state) ** 2).reshape([2] * n).sum(axis=tuple(q for q in range(n) if q not in (0, 5, 7))).reshape(-1)
    emp = np.array([freq.get(format(k, "03b"), 0) for k in range(8)]) / 5000
    assert np.abs(emp - p).max() < 0.03, (emp, p)
    box = [None] * world
    dist.all_gather_object(box, (e, sampled, dict(freq)))
    assert all(b == box[0] for b in box), "ranks disagree"
    get_runtime().release()
    if rank == 0:
        print("PUBLIC_API_OK", e, sampled, exact)


if __name__ == "__main__":
    main()
