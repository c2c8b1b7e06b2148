"""
README example of the reference (README.md:25-48) on the B200 path: H2 / STO-3G, HF + UCCD, VQE with BFGS.

PySCF is not available in this image, so the molecule comes from the integrals the reference's docs print
(doc/source/tutorials/hamiltonian.rst:40, 0.74804 Angstrom).  The loss is a plain Python callable; its gradient comes
from one adjoint pass (`qibochem_b200.derivative.adjoint_gradient`) instead of two energy evaluations per parameter.

    python examples/h2_vqe.py          # needs a B200; FCI of this Hamiltonian: -1.1371622544371 Ha
"""

import os
import sys

import numpy as np
from scipy.optimize import minimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from qibochem_b200.ansatz import hf_circuit, ucc_circuit  # noqa: E402
from qibochem_b200.derivative import adjoint_gradient  # noqa: E402
from qibochem_b200.driver import Molecule  # noqa: E402

oei = np.diag([-1.248461959132892, -0.48007161818330846])
tei = np.zeros((2, 2, 2, 2))
for key, val in {
    (0, 0, 0, 0): 0.3366109237586995, (0, 0, 1, 1): 0.09083064962340165, (0, 1, 0, 1): 0.09083064962340165,
    (0, 1, 1, 0): 0.33115823068165495, (1, 0, 0, 1): 0.3311582306816552, (1, 0, 1, 0): 0.09083064962340165,
    (1, 1, 0, 0): 0.09083064962340165, (1, 1, 1, 1): 0.348087115228365,
}.items():  # fmt: skip
    tei[key] = 2 * val
h2 = Molecule.from_integrals(oei, tei, nelec=2, e_nuc=0.7074183344740923)
hamiltonian = h2.hamiltonian()

circuit = hf_circuit(h2.nso, h2.nelec)
circuit += ucc_circuit(h2.nso, [0, 1, 2, 3])
print(f"HF energy: {hamiltonian.expectation(circuit):.10f} Ha   ({len(circuit.get_parameters())} parameters)")


def loss_and_gradient(params):
    circuit.set_parameters(params)
    return adjoint_gradient(circuit, hamiltonian)


x0 = np.random.default_rng(0).uniform(-0.1, 0.1, size=len(circuit.get_parameters()))
best = minimize(loss_and_gradient, x0, jac=True, method="BFGS")
print(f"VQE energy: {best.fun:.10f} Ha after {best.nit} BFGS iterations (FCI -1.1371622544 Ha)")
