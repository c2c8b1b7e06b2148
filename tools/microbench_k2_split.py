"""Where does the tiled expectation spend its time? Quad groups only vs hopping/diagonal groups only vs everything."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from qibochem_b200 import synthetic  # noqa: E402
from qibochem_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
nelec = 14 if n == 30 else 2 * ((n * 14 // 30) // 2)
rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, nelec, max_excitations=32)
hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, T)
pc = np.array([bin(int(x)).count("1") for x in hx])
st = DeviceState(n)
st.set_basis_state(synthetic.hf_index(n, nelec))
st.apply_pauli_rotations(rx, rz, ra)
st.profile_enable(True)
st.set_option("k2_mode", 2)
only = sys.argv[3] if len(sys.argv) > 3 else ""
for name, sel in (("all", pc >= 0), ("quads", pc == 4), ("pairs+diag", pc < 4)):
    if only and name != only:
        continue
    x, z, c = hx[sel], hz[sel], hc[sel]
    st.expectation(x, z, c)
    st.profile_reset()
    e = st.expectation(x, z, c)
    p = st.profile()["k2_tiled"]
    print(f"n={n} {name:11s}: terms {x.size:6d} x-masks {np.unique(x).size:6d}  passes {p['launches']:4d}  {p['ms']:9.1f} ms  E={e:.10f}", flush=True)
