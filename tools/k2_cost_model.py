"""Fits the per-pass cost model of the tiled expectation to an ncu launch list.

usage: python tools/k2_cost_model.py profiles/r1_launches_bench_30q_v13.csv [n_qubits] [n_terms]
Needs tools/k2_plan_dump (build line in tools/k2_plan_dump.cpp); no GPU."""
import csv
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, ".")
from qibochem_b200 import synthetic  # noqa: E402

launches = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
hx, hz, hc, _ = synthetic.molecular_like_hamiltonian(n, T)
with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
    f.write(struct.pack("i", n) + struct.pack("Q", hx.size) + hx.tobytes() + hz.tobytes() + hc.tobytes())
out = subprocess.run([os.path.join(os.path.dirname(__file__), "k2_plan_dump"), f.name], capture_output=True, text=True, check=True)
os.unlink(f.name)
plan = np.array([[float(v) for v in line.split()] for line in out.stdout.splitlines()])
rows = list(csv.reader(open(launches)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
t = np.array([float(r[vi].replace(",", "")) / 1e6 for r in rows[h + 1 :] if len(r) > vi and "expectation_tiled" in r[ki]])[: len(plan)]
A = np.column_stack([np.ones(len(plan)), plan[:, 1], plan[:, 2], plan[:, 3], plan[:, 4]])
coef = np.linalg.lstsq(A, t, rcond=None)[0]
pred = A @ coef
print(f"{len(plan)} passes, {t.sum():.0f} ms measured")
print("ms: per pass %.2f, per visit %.3f, per quad %.4f, per hop %.3f, per generic record %.3f" % tuple(coef))
print("rms residual %.2f ms; shares: passes %.0f ms, visits %.0f ms, quads %.0f ms, hops %.0f ms, generic %.0f ms"
      % (np.sqrt(np.mean((pred - t) ** 2)), coef[0] * len(plan), coef[1] * plan[:, 1].sum(), coef[2] * plan[:, 2].sum(),
         coef[3] * plan[:, 3].sum(), coef[4] * plan[:, 4].sum()))
