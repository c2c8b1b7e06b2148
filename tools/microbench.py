"""Kernel micro-benchmarks on a dense state (host wall clock around a stream sync; kernels are ms-scale)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from qibochem_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
st = DeviceState(n)
st.set_basis_state(0)
rng = np.random.default_rng(0)
# dense product state: one Y rotation per qubit
t0 = time.perf_counter()
st.apply_pauli_rotations([1 << b for b in range(n)], [1 << b for b in range(n)], rng.uniform(0.3, 2.8, size=n))
st.synchronize()
print(f"n={n} dense init {time.perf_counter()-t0:.3f}s norm2={st.norm2():.15f}", flush=True)
B = 16.0 * 2**n


def timeit(label, fn, reps=5, bytes_per=2 * B):
    fn(); st.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); st.synchronize(); ts.append(time.perf_counter() - t)
    t = min(ts)
    print(f"{label:50s} {t*1e3:9.3f} ms  {bytes_per/t/1e9:8.1f} GB/s", flush=True)


top = 1 << (n - 1)
for label, x, z in [
    ("K1 single string x=top bit", top, 0),
    ("K1 single string x=top|bit7, z chain", top | 128, (top - 1) & ~255),
    ("K1 x = 4 spread bits (double-like)", top | (1 << (n - 2)) | 8 | 4, 0),
    ("K1 x = bit 5", 32, 3),
    ("K1 x = bit 0 (low-x path)", 1, 2),
    ("K1 diagonal", 0, top | 5),
]:
    timeit(label, lambda: st.apply_pauli_rotations([x], [z], [0.1]))
# fused double (8 strings one pass)
x = top | (1 << (n - 2)) | 8 | 4
sites = [n - 1, n - 2, 3, 2]
zs = []
for pat in range(16):
    if bin(pat).count("1") % 2 == 1:
        zs.append(sum(1 << sites[k] for k in range(4) if (pat >> k) & 1))
timeit("K1 fused 8-string double (1 pass)", lambda: st.apply_pauli_rotations([x] * 8, zs, [0.01 * k for k in range(8)]))
assert st.last_passes == 1
for m in (1, 4, 16, 58):
    zz = rng.integers(0, 2**n, size=m, dtype=np.uint64)
    zz = [int(v) & ~int(x) for v in zz]  # even nY
    timeit(f"K2 one x-mask sweep, {m} z-masks", lambda: st.expectation([x] * m, zz, np.ones(m)), bytes_per=B)
zz = [int(v) for v in rng.integers(0, 2**n, size=465, dtype=np.uint64)]
timeit("K2 diagonal group, 465 z-masks", lambda: st.expectation([0] * 465, zz, np.ones(465)), bytes_per=B)
xs = [int(v) | top for v in rng.integers(0, 2**n, size=16, dtype=np.uint64)]
timeit("K2 16 x-masks x 4 terms (16 sweeps)", lambda: st.expectation(np.repeat(xs, 4), [0, 3, 5, 6] * 16, np.ones(64)), bytes_per=16 * B)
print("norm2 after", st.norm2())
