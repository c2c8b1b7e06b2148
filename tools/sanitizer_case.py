"""One small pass through every shared-memory kernel, for `compute-sanitizer --tool memcheck|racecheck`:
rotation phases (K1O), rotation tiles (K1T), the single-launch small-register program, tiled expectation (K2T: quad,
HOP and generic records), tiled H|psi> (K6T) + backward walk, the sampler, and the global<->local bit exchange kernel.
usage: compute-sanitizer --tool racecheck python tools/sanitizer_case.py [n]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from qibochem_b200 import synthetic  # noqa: E402
from qibochem_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ne = 6
rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, ne, max_excitations=12)
hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, 2500)
hf = synthetic.hf_index(n, ne)
energies = []
with DeviceState(n) as st, DeviceState(n) as other:
    for k1 in (0, 2, 1):
        st.set_option("k1_mode", k1)
        st.set_basis_state(hf)
        st.apply_pauli_rotations(rx, rz, ra)
        st.set_option("k2_mode", 2)
        energies.append(st.expectation(hx, hz, hc))
    st.set_option("k2_mode", 1)
    energies.append(st.expectation(hx[:300], hz[:300], hc[:300]))
    st.set_option("k2_mode", 2)
    e, g = st.energy_gradient(rx, rz, ra, hx, hz, hc)
    energies.append(e)
    st.apply_pauli_rotations(rx, rz, ra)
    samples = st.sample(0b1010, 0b0100, 2000, 7, on_device=True)
    sums = st.parity_sums(samples, np.array([3, 5, 2**n - 1], dtype=np.uint64))
    other.set_basis_state(1)
    st.swap_local_pair(other, 9)
    print("norms after exchange", st.norm2() + other.norm2())
with DeviceState(12) as small:
    rx, rz, ra, _ = synthetic.uccsd_rotation_stream(12, 4, max_excitations=20)
    small.set_basis_state(synthetic.hf_index(12, 4))
    print("small-register passes", small.apply_pauli_rotations(rx, rz, ra), "norm", small.norm2())
print("energies", energies, "parity sums", sums.tolist(), "|g|max", float(np.abs(g).max()))
assert abs(energies[0] - energies[1]) < 1e-10 and abs(energies[0] - energies[2]) < 1e-10 and abs(energies[0] - energies[4]) < 1e-10
print("SANITIZER_CASE_OK")
