"""
Generates the golden fixtures under tests/golden/ by importing the REFERENCE's own host-side modules from
/root/reference (by file path, with a stub `qibo` package: qibo itself is not installed in this image).
Run in the build container only; the GPU box has no /root/reference and reads the committed JSON files.

    python tests/golden/make_golden.py
"""

import importlib.util
import json
import os
import sys
import types

REF = "/root/reference/src/qibochem"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub_qibo(recorder):
    qibo = types.ModuleType("qibo")
    config = types.ModuleType("qibo.config")

    def raise_error(exc, msg=""):
        raise exc(msg)

    config.raise_error = raise_error
    gates = types.ModuleType("qibo.gates")

    class _Gate:
        def __init__(self, *qubits, **kw):
            self.name = type(self).__name__
            self.qubits = tuple(q for q in qubits if isinstance(q, int))
            self.params = tuple(q for q in qubits if not isinstance(q, int))
            self.dag = False

        def dagger(self):
            g = type(self)(*self.qubits, *self.params)
            g.dag = not self.dag
            return g

    for name in ("X", "H", "S", "CNOT", "RZ", "CZ", "SWAP", "M", "RX", "RY", "Y", "Z", "SDG"):
        setattr(gates, name, type(name, (_Gate,), {}))
    gates.Gate = _Gate

    class Circuit:
        def __init__(self, nqubits, **kw):
            self.nqubits = nqubits
            self.queue = []

        def add(self, g):
            if isinstance(g, _Gate):
                self.queue.append(g)
                recorder.append(g)
            else:
                for x in g:
                    self.add(x)

    qibo.Circuit = Circuit
    qibo.gates = gates
    qibo.config = config
    sys.modules.update({"qibo": qibo, "qibo.config": config, "qibo.gates": gates})


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    rec = []
    _stub_qibo(rec)
    utils = _load("ref_ansatz_utils", f"{REF}/ansatz/utils.py")
    mutil = _load("ref_measurement_util", f"{REF}/measurement/util.py")
    shots = _load("ref_shot_allocation", f"{REF}/measurement/shot_allocation.py")
    # _ansatz.py does `from qibochem.ansatz.utils import generate_excitations`: alias the module loaded above
    for name in ("qibochem", "qibochem.ansatz"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["qibochem.ansatz.utils"] = utils
    _ansatz = _load("ref__ansatz", f"{REF}/ansatz/_ansatz.py")
    for name in ("qibochem", "qibochem.ansatz", "qibochem.ansatz.utils"):
        del sys.modules[name]

    # ---- excitation generation / ordering (ansatz/utils.py:11-114; circuit_ansatz order ansatz.py:123-127)
    exc = {"generate": {}, "circuit_ansatz_order": {}}
    for n, ne in [(4, 2), (6, 2), (8, 4), (12, 4), (14, 6), (10, 6)]:
        for order in (1, 2):
            exc["generate"][f"{order}_{n}_{ne}"] = utils.generate_excitations(order, range(0, ne), range(ne, n))
        full = []
        for order in range(2, 0, -1):
            full += utils.generate_excitations(order, range(0, ne), range(ne, n))
        exc["circuit_ansatz_order"][f"{n}_{ne}"] = full
    exc["counts_30_14"] = [
        len(utils.generate_excitations(2, range(0, 14), range(14, 30))),
        len(utils.generate_excitations(1, range(0, 14), range(14, 30))),
    ]
    exc["no_spin"] = utils.generate_excitations(1, [0, 1], [2, 3], conserve_spin=False)
    json.dump(exc, open(f"{OUT}/excitations.json", "w"), indent=0)

    # ---- _expi_pauli gate programs (ansatz/_ansatz.py:55-92)
    progs = {}
    for s in ["Z1", "Z0 Z1", "X0 X1", "Y0 Y1", "X0 X1 Y2 X3", "Y0 Z1 X2", "X3 Z1 Y0", "Y5 X2 Z3 Z4 Y0"]:
        circ = _ansatz._expi_pauli(6, s, 0.25)
        progs[s] = [[g.name + ("_dagger" if g.dag else ""), list(g.qubits), list(g.params)] for g in circ.queue]
    json.dump(progs, open(f"{OUT}/expi_pauli_programs.json", "w"), indent=0)

    # ---- BK image of the HF occupation vector (ansatz.py:276-281 with _ansatz._bk_matrix)
    import numpy as _np

    bk = {}
    for n, ne in [(4, 2), (6, 2), (7, 3), (8, 4), (12, 4), (13, 7), (16, 9)]:
        occ = _np.concatenate((_np.ones(ne, dtype=_np.int8), _np.zeros(n - ne, dtype=_np.int8)))
        bk[f"{n}_{ne}"] = [int(v) for v in (_ansatz._bk_matrix(n) @ occ) % 2]
    json.dump(bk, open(f"{OUT}/bk_occupation.json", "w"))

    # ---- QWC / general commuting groups (measurement/util.py:22-79), deterministic given the node order
    import numpy as np

    rng = np.random.default_rng(0)
    cases = []
    fixed = [
        ["X0 Z1", "X0", "Z0", "Z0 Z1"],
        ["Z0", "Z1", "Z0 Z1", "X0 X1", "Y0 Y1"],
        ["Z0", "X0 Y1", "Z0 Y2"],
        ["X0 X1 Y2", "X0 Y1 Y2", "Y0 X1 X2", "Y0 Y1 X2"],
    ]
    for terms in fixed:
        cases.append(terms)
    for n, t in [(4, 12), (5, 25), (6, 40), (8, 60)]:
        seen = []
        while len(seen) < t:
            ops = [(q, "IXYZ"[rng.integers(0, 4)]) for q in range(n)]
            s = " ".join(f"{p}{q}" for q, p in ops if p != "I")
            if s and s not in seen:
                seen.append(s)
        cases.append(seen)
    groups = []
    for terms in cases:
        groups.append(
            {
                "terms": terms,
                "qwc": mutil._group_commuting_terms(terms, True),
                "gc": mutil._group_commuting_terms(terms, False),
                "commute_qwc": [[bool(mutil._check_terms_commutativity(a, b, True)) for b in terms] for a in terms],
                "commute_gc": [[bool(mutil._check_terms_commutativity(a, b, False)) for b in terms] for a in terms],
            }
        )
    json.dump(groups, open(f"{OUT}/commuting_groups.json", "w"), indent=0)

    # ---- variance-based shot allocation (measurement/shot_allocation.py:102-125)
    va = []
    for total, trial, var, method in [(100, 10, [0.0, 1.0, 4.0], "vmsa"), (100, 10, [0.0, 1.0, 4.0], "vpsr"), (1000, 50, [0.5, 0.1, 0.9, 0.0], "vmsa"), (1000, 50, [0.5, 0.1, 0.9, 0.3], "vpsr")]:
        va.append({"args": [total, trial, var, method], "out": shots.allocate_shots_by_variance(total, trial, var, method)})
    json.dump(va, open(f"{OUT}/shot_allocation_variance.json", "w"), indent=0)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
