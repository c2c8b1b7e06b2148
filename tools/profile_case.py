"""Short single-GPU case for ncu captures: a few K1 passes and K2 sweeps at full size (30 qubits by default)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from qibochem_b200 import synthetic  # noqa: E402
from qibochem_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
nelec = 14 if n == 30 else 2 * ((n * 14 // 30) // 2)
rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, nelec, max_excitations=6)
hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, 100_000 if n >= 28 else 3000)
st = DeviceState(n)
st.set_basis_state(synthetic.hf_index(n, nelec))
st.apply_pauli_rotations(rx, rz, ra)
# 6 x-masks with 4 terms (quads), 2 hopping x-masks with 58 terms, the diagonal group
ux = np.unique(hx)
pick = list(ux[-6:]) + [x for x in ux if bin(int(x)).count("1") == 2][:2] + [0]
sel = np.isin(hx, np.array(pick, dtype=np.uint64))
e = st.expectation(hx[sel], hz[sel], hc[sel])
print("passes", st.last_passes, "sweeps", st.last_sweeps, "energy", e, "norm2", st.norm2())
