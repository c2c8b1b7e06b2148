"""Host-only census of the sharded rotation phase: how many global-bit exchanges and local passes the FULL UCCSD stream
of a register needs under the Belady eviction rule (no GPU, no data: the layout logic of ShardedState is replayed
with a stand-in shard).  usage: python tools/count_swaps.py <qubits> <electrons> <ranks> [lookahead]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from qibochem_b200 import synthetic  # noqa: E402
from qibochem_b200.sharded import ShardedState  # noqa: E402


class NullShard:
    def __init__(self):
        self.calls = 0
        self.rotations = 0

    def set_basis_state(self, index, has_one=True):
        pass

    def apply_pauli_rotations(self, x, z, a):
        self.calls += 1
        self.rotations += len(x)
        return 0


class NullExchanger:
    def __init__(self):
        self.swaps = 0

    def swap(self, global_k, local_pos, with_scratch=False):
        self.swaps += 1


def main():
    n, ne, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    look = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
    rx, rz, ra, n_exc = synthetic.uccsd_rotation_stream(n, ne)
    g = world.bit_length() - 1
    touching = int(np.count_nonzero(rx[::1] >> np.uint64(n - g)))
    shard, ex = NullShard(), NullExchanger()
    st = ShardedState(n, 0, world, shard, ex, min_swap_pos=8)
    st._plan_stats = None
    t0 = time.time()
    st.set_basis_state(synthetic.hf_index(n, ne))
    st.apply_pauli_rotations(rx, rz, ra, lookahead=look)
    print(f"n={n} ne={ne} ranks={world}: {n_exc} excitations, {rx.size} rotations, {touching} rotations pair on a default-global bit; "
          f"{ex.swaps} exchanges, {shard.calls} local segments (lookahead {look}); host time {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
