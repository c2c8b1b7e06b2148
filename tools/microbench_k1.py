"""K1: one HBM pass per fused run vs tile-batched passes on the benchmark rotation stream."""
import sys
import time

sys.path.insert(0, ".")
from qibochem_b200 import synthetic  # noqa: E402
from qibochem_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nelec = 14 if n == 30 else 2 * ((n * 14 // 30) // 2)
rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, nelec, max_excitations=max(1, R // 8))
rx, rz, ra = rx[:R], rz[:R], ra[:R]
st = DeviceState(n)
st.profile_enable(True)
for mode, name in ((1, "per-run"), (0, "batched")):
    st.set_option("k1_mode", mode)
    for rep in range(2):
        st.set_basis_state(synthetic.hf_index(n, nelec))
        st.profile_reset()
        t0 = time.perf_counter()
        passes = st.apply_pauli_rotations(rx, rz, ra)
        st.synchronize()
        dt = time.perf_counter() - t0
        p = st.profile()["k1_rotation"]
    gbs = 32.0 * 2.0**n * passes / 1e9 / (p["ms"] / 1e3)
    print(f"n={n} R={R} {name:8s}: passes={passes} device {p['ms']:.1f} ms ({p['ms']/passes:.2f} ms/pass, {gbs:.0f} GB/s actual r+w)"
          f"  wall {dt*1e3:.1f} ms  norm2={st.norm2():.12f}", flush=True)
