"""
Gradients of ``E(params) = <psi(params)|H|psi(params)>`` for the optimiser loop.

Reference: the docs differentiate ansatz circuits with ``qibo.derivative.parameter_shift(circuit, hamiltonian,
parameter_index)`` (``doc/source/tutorials/adaptive.rst:46-61``, ``:118-141``, ``:225-233``): two full energy
evaluations per parameter.  ``parameter_shift`` below keeps that call (same arguments, same result) on the B200
kernels; ``adjoint_gradient`` returns ALL partial derivatives from one forward execution, ``H|psi>`` and one
backward walk over the circuit (``csrc/gradient.cu``), i.e. O(1) energy evaluations instead of O(parameters).
"""

from __future__ import annotations

import numpy as np

from qibochem_b200 import gates as _gates


def parameter_shift(circuit, hamiltonian, parameter_index: int, initial_state=None, scale_factor: float = 1.0) -> float:
    """dE/dtheta_k by the parameter-shift rule (qibo's signature without the sampling arguments).

    Valid for gates ``exp(-i theta P / 2)`` with ``P^2 = 1`` -- every parametrized gate of this package
    (RX / RY / RZ / PauliRotation): ``dE/dtheta = (E(theta + pi/2) - E(theta - pi/2)) / 2``.
    """
    params = [p[0] for p in circuit.get_parameters()]
    if parameter_index < 0 or parameter_index >= len(params):
        raise ValueError(f"This index is out of bounds: the circuit has {len(params)} parameters.")
    gate = circuit.parametrized_gates[parameter_index]
    if not isinstance(gate, (_gates.RX, _gates.RY, _gates.RZ, _gates.PauliRotation)):
        raise ValueError("parameter_shift needs a gate of the form exp(-i theta P / 2)")
    original = params[parameter_index]
    shift = np.pi / (2.0 * scale_factor)
    energies = []
    for s in (+shift, -shift):
        params[parameter_index] = original + s
        circuit.set_parameters(params)
        energies.append(hamiltonian.expectation(circuit if initial_state is None else circuit(initial_state)))
    params[parameter_index] = original
    circuit.set_parameters(params)
    return float(0.5 * scale_factor * (energies[0] - energies[1]))


def adjoint_gradient(circuit, hamiltonian, initial_state=None):
    """``(E, dE/dparams)`` with one partial derivative per parametrized gate (the order of ``get_parameters()``).

    One forward execution (K0 + K1), ``lambda = H psi``, then a backward walk that un-applies the fused rotation runs
    on ``psi`` and ``lambda`` and reads ``dE/dphi_r = Im<lambda|P_r|psi>`` for every rotation on the way.
    """
    prog = circuit.program()
    st = circuit.run_on_device(initial_state)
    hx, hz, hc = hamiltonian.index_masks()
    e, g_rot = st.energy_gradient(prog.xmask, prog.zmask, prog.angle, hx, hz, hc)
    grad = np.zeros(len(circuit.parametrized_gates), dtype=np.float64)
    # a parametrized gate may lower to several rotations that share its parameter: chain rule angle = scale * param
    if prog.param_rot.size:
        np.add.at(grad, prog.param_owner, prog.param_scale * g_rot[prog.param_rot])
    return float(e + hamiltonian.constant), grad
