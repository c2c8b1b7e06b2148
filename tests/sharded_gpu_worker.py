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
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_OK swaps", st.n_swaps)


if __name__ == "__main__":
    main()
