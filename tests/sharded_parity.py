"""Multi-GPU parity checks run by `bench_sharded.py` (bench.py's N > 1 arm) BEFORE anything is timed, and by nothing
else: the sharded path against the C oracle and against the single-GPU energy.  Kept under tests/ because it is test
infrastructure -- the only place outside the CPU-baseline leg where a bench process touches `oracle/`, and only as the
checker of the product path, never as the thing measured."""

from __future__ import annotations

import time

import numpy as np

ENERGY_30Q_1GPU = 1.2250109420586086  # C4 workload (30 q, 1024 rotations, 100 000 terms) on one GPU: BENCH_r01.json "energy"


def parity_checks(rank: int, world: int, local_rank: int, stream) -> dict:
    """Multi-GPU parity on the record, before anything is timed (SURVEY 8d gates):
    (a) a 22-qubit sharded HF + 256-rotation + 5000-term energy against the C oracle (rank 0, all host threads), <= 1e-10;
    (b) the 30-qubit C4 workload sharded over the same ranks against the deterministic 1-GPU energy, <= 1e-12.
    The oracle is the checker here, never the thing measured."""
    import torch
    import torch.distributed as dist

    from bench import workload
    from qibochem_b200 import synthetic
    from qibochem_b200.engine import DeviceState
    from qibochem_b200.sharded import IpcExchanger, ShardedState

    g = world.bit_length() - 1
    out = {}
    t_start = time.perf_counter()

    def sharded_energy(n, wl):
        local = DeviceState(n - g, device=local_rank, stream=stream.cuda_stream)
        ex = IpcExchanger(local, rank, world, local_rank)
        st = ShardedState(n, rank, world, local, ex, min_swap_pos=min(8, n - g - 4))
        st.set_basis_state(wl["hf"])
        st.apply_pauli_rotations(wl["rx"], wl["rz"], wl["ra"])
        e = wl["const"] + st.expectation(wl["hx"], wl["hz"], wl["hc"])
        swaps = st.n_swaps
        ex.close()
        local.close()
        return e, swaps

    # (a) oracle check
    n_a, ne_a = 22, 10
    wl = workload(n_a, ne_a, 256, 5000)
    e_a, swaps_a = sharded_energy(n_a, wl)
    ref = torch.zeros(1, dtype=torch.float64, device=f"cuda:{local_rank}")
    if rank == 0:
        from oracle import oracle_c as OC

        OC.set_num_threads(0)
        psi = np.zeros(1 << n_a, dtype=np.complex128)
        psi[wl["hf"]] = 1.0
        for x, z, a in zip(wl["rx"].tolist(), wl["rz"].tolist(), wl["ra"].tolist()):
            OC.pauli_rotation(psi, n_a, x, z, a)
        ref[0] = wl["const"] + OC.expectation_masks(psi, n_a, wl["hx"], wl["hz"], wl["hc"])
        del psi
    dist.broadcast(ref, src=0)
    err_a = abs(e_a - float(ref[0]))
    out["oracle_22q"] = {"n": n_a, "rotations": 256, "terms": 5001, "energy": e_a, "oracle_energy": float(ref[0]),
                         "abs_err": err_a, "tol": 1e-10, "global_bit_swaps": swaps_a, "ok": bool(err_a <= 1e-10)}  # fmt: skip
    # (b) 1-GPU twin of the 30-qubit headline workload
    wl = workload(30, 14, 1024, 100_000)
    e_b, swaps_b = sharded_energy(30, wl)
    err_b = abs(e_b - ENERGY_30Q_1GPU)
    out["twin_30q"] = {"n": 30, "rotations": 1024, "terms": 100_000, "energy": e_b, "one_gpu_energy": ENERGY_30Q_1GPU,
                       "abs_err": err_b, "tol": 1e-12, "global_bit_swaps": swaps_b, "ok": bool(err_b <= 1e-12)}  # fmt: skip
    spread = torch.tensor([e_a, -e_a, e_b, -e_b], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(spread, op=dist.ReduceOp.MAX)
    out["rank_spread"] = float(max(spread[0] + spread[1], spread[2] + spread[3]))  # max - min over ranks
    out["seconds"] = time.perf_counter() - t_start
    out["ok"] = bool(out["oracle_22q"]["ok"] and out["twin_30q"]["ok"] and out["rank_spread"] == 0.0)
    out["n"], out["abs_err"] = 30, max(err_a, err_b)
    return out
