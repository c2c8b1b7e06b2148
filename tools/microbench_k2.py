"""K2 sweeps vs tiles on the benchmark Hamiltonian (host wall clock around the call)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from qibochem_b200 import synthetic  # noqa: E402
from qibochem_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
nelec = 14 if n == 30 else 2 * ((n * 14 // 30) // 2)
rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, nelec, max_excitations=32)
hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, T)
st = DeviceState(n)
st.set_basis_state(synthetic.hf_index(n, nelec))
st.apply_pauli_rotations(rx, rz, ra)
st.profile_enable(True)
only = sys.argv[3] if len(sys.argv) > 3 else ""
for mode, name in ((2, "tiles"), (1, "sweeps")):
    if (mode == 1 and n >= 28 and only != "all") or (only not in ("", "all") and only != name):
        continue
    st.set_option("k2_mode", mode)
    t0 = time.perf_counter(); e = st.expectation(hx, hz, hc); t1 = time.perf_counter()
    st.profile_reset()
    t2 = time.perf_counter(); e = st.expectation(hx, hz, hc); t3 = time.perf_counter()
    p = st.profile()
    print(f"n={n} T={hx.size} {name:7s}: first {t1-t0:8.3f}s  second {t3-t2:8.3f}s  launches/passes={st.last_sweeps}  E={e:.12f}", flush=True)
    print("   ", {k: (round(v['ms'], 2), v['launches']) for k, v in p.items() if v['launches']}, flush=True)
