"""ctypes wrapper of the C oracle (oracle/oracle_c.c).  TEST / BASELINE INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_c.so")
KIND = {"X": 0, "H": 1, "S": 2, "SDG": 3, "CNOT": 4, "RZ": 5, "Y": 6, "Z": 7}
PAULI = {"X": 0, "Y": 1, "Z": 2}
_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "oracle_c.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_LIB)
        _lib.oc_expectation_terms.restype = C.c_double
        _lib.oc_expectation_masks.restype = C.c_double
        _lib.oc_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return load().oc_num_threads()


def set_num_threads(n: int = 0) -> int:
    """Explicit OpenMP thread count (0 = every logical core), independent of an inherited OMP_NUM_THREADS."""
    load().oc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def encode_gates(gates):
    """oracle gate list [(name, qubits, param)] -> flat arrays"""
    kinds = np.array([KIND[g[0]] for g in gates], dtype=np.int32)
    q0 = np.array([g[1][0] for g in gates], dtype=np.int32)
    q1 = np.array([g[1][1] if len(g[1]) > 1 else -1 for g in gates], dtype=np.int32)
    params = np.array([0.0 if g[2] is None else float(np.real(g[2])) for g in gates], dtype=np.float64)
    return kinds, q0, q1, params


def run_gates(state: np.ndarray, n: int, gates) -> np.ndarray:
    """In-place gate-by-gate execution (all host threads)."""
    assert state.dtype == np.complex128 and state.flags.c_contiguous and state.size == 1 << n
    kinds, q0, q1, params = encode_gates(gates)
    load().oc_run_gates(
        state.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int64(len(gates)), kinds.ctypes.data_as(C.c_void_p),
        q0.ctypes.data_as(C.c_void_p), q1.ctypes.data_as(C.c_void_p), params.ctypes.data_as(C.c_void_p),
    )  # fmt: skip
    return state


def expectation_terms(state: np.ndarray, n: int, terms, tmp: np.ndarray | None = None) -> float:
    """terms = [(coeff, [(qubit, 'X'|'Y'|'Z'), ...]), ...]; qibo-style copy + factor passes + dot."""
    if tmp is None:
        tmp = np.empty_like(state)
    coeffs = np.array([float(np.real(c)) for c, _ in terms], dtype=np.float64)
    offsets = np.zeros(len(terms) + 1, dtype=np.int64)
    qs, ps = [], []
    for k, (_, factors) in enumerate(terms):
        for q, p in factors:
            qs.append(q), ps.append(PAULI[p])
        offsets[k + 1] = len(qs)
    qs, ps = np.array(qs, dtype=np.int32), np.array(ps, dtype=np.int32)
    return load().oc_expectation_terms(
        state.ctypes.data_as(C.c_void_p), tmp.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int64(len(terms)),
        coeffs.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), qs.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p),
    )  # fmt: skip


def pauli_rotation(state: np.ndarray, n: int, x: int, z: int, phi: float) -> np.ndarray:
    load().oc_pauli_rotation(state.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_uint64(x), C.c_uint64(z), C.c_double(phi))
    return state


def expectation_masks(state: np.ndarray, n: int, xs, zs, coeffs) -> float:
    xs = np.ascontiguousarray(xs, dtype=np.uint64)
    zs = np.ascontiguousarray(zs, dtype=np.uint64)
    cs = np.ascontiguousarray(coeffs, dtype=np.float64)
    return load().oc_expectation_masks(
        state.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int64(xs.size), xs.ctypes.data_as(C.c_void_p),
        zs.ctypes.data_as(C.c_void_p), cs.ctypes.data_as(C.c_void_p),
    )  # fmt: skip


def expectation_masks_per_term(state: np.ndarray, n: int, xs, zs) -> np.ndarray:
    xs = np.ascontiguousarray(xs, dtype=np.uint64)
    zs = np.ascontiguousarray(zs, dtype=np.uint64)
    out = np.zeros(xs.size, dtype=np.float64)
    load().oc_expectation_masks_per_term(
        state.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int64(xs.size), xs.ctypes.data_as(C.c_void_p),
        zs.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
    )  # fmt: skip
    return out
