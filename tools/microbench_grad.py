"""Adjoint gradient vs the cost of parameter-shift (2 energy evaluations per parameter) on the benchmark workload."""
import sys
import time

sys.path.insert(0, ".")
from qibochem_b200 import synthetic  # noqa: E402
from qibochem_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
nelec = 14 if n == 30 else 2 * ((n * 14 // 30) // 2)
rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, nelec, max_excitations=max(1, R // 8))
rx, rz, ra = rx[:R], rz[:R], ra[:R]
hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, T)
st = DeviceState(n)
st.profile_enable(True)
hf = synthetic.hf_index(n, nelec)
for rep in range(2):
    st.set_basis_state(hf)
    st.apply_pauli_rotations(rx, rz, ra)
    t0 = time.perf_counter()
    e_fwd = st.expectation(hx, hz, hc)
    t_e = time.perf_counter() - t0
    st.profile_reset()
    t0 = time.perf_counter()
    e, g = st.energy_gradient(rx, rz, ra, hx, hz, hc)
    t_g = time.perf_counter() - t0
    prof = {k: (round(v["ms"], 1), v["launches"]) for k, v in st.profile().items() if v["launches"]}
print(f"n={n} R={R} T={hx.size}: energy (tiled) {t_e:.3f}s  adjoint gradient of all {R} angles {t_g:.3f}s  "
      f"|E_adj - E_tiled| = {abs(e - e_fwd):.2e}  |g|max = {abs(g).max():.4f}")
print(f"   parameter-shift would need 2 x {R} energy evaluations ~ {2 * R * t_e:.0f}s -> x{2 * R * t_e / t_g:.0f}")
print("  ", prof)
