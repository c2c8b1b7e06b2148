"""
ctypes binding of ``libqibochem_b200.so`` (the C ABI declared in ``include/qibochem_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is raised.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqibochem_b200.so")

_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol of include/qibochem_b200.h
SIGNATURES = {
    "qc_last_error": (C.c_char_p, []),
    "qc_version": (C.c_int, []),
    "qc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "qc_create": (C.c_int, [C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    "qc_destroy": (C.c_int, [_vp]),
    "qc_num_qubits": (C.c_int, [_vp]),
    "qc_state_device_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "qc_synchronize": (C.c_int, [_vp]),
    "qc_set_basis_state": (C.c_int, [_vp, C.c_uint64, C.c_int]),
    "qc_apply_pauli_rotations": (C.c_int, [_vp, C.c_size_t, _u64p, _u64p, _f64p, C.POINTER(C.c_int)]),
    "qc_prepare_state": (C.c_int, [_vp, C.c_uint64, C.c_int, C.c_size_t, _u64p, _u64p, _f64p, C.POINTER(C.c_int)]),
    "qc_expectation": (C.c_int, [_vp, C.c_size_t, _u64p, _u64p, _f64p, _f64p, _f64p, C.POINTER(C.c_int)]),
    "qc_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "qc_k2_plan_stats": (C.c_int, [C.c_int, C.c_size_t, _u64p, _u64p, _f64p, _u64p, _f64p]),
    "qc_energy_gradient": (C.c_int, [_vp, C.c_size_t, _u64p, _u64p, _f64p, C.c_size_t, _u64p, _u64p, _f64p, _f64p, _f64p]),
    "qc_hpsi": (C.c_int, [_vp, C.c_size_t, _u64p, _u64p, _f64p, C.c_int]),
    "qc_lambda_inner": (C.c_int, [_vp, _f64p, _f64p]),
    "qc_gradient_backward": (C.c_int, [_vp, C.c_size_t, _u64p, _u64p, _f64p, _f64p]),
    "qc_hpsi_plan_stats": (C.c_int, [C.c_int, C.c_size_t, _u64p, _u64p, _f64p, _u64p]),
    "qc_norm2": (C.c_int, [_vp, _f64p]),
    "qc_sample": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_uint64, C.c_uint64, _vp, C.c_int, _f64p]),
    "qc_parity_sums": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int, _u64p, C.c_size_t, _i64p]),
    "qc_frequencies": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_uint64, _u64p, _i64p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qc_outcome_variance": (C.c_int, [_vp, _u64p, _i64p, C.c_size_t, _u64p, _f64p, C.c_size_t, C.c_double, _f64p, _f64p]),
    "qc_group_commuting": (C.c_int, [C.c_size_t, _u64p, _u64p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "qc_profile_enable": (C.c_int, [_vp, C.c_int]),
    "qc_profile_reset": (C.c_int, [_vp]),
    "qc_transfer_bytes": (C.c_int, [_vp, _u64p, _u64p, C.c_int]),
    "qc_profile_get": (C.c_int, [_vp, C.c_int, _f64p, _u64p, _f64p]),
    "qc_copy_state_to_host": (C.c_int, [_vp, _f64p, C.c_uint64, C.c_uint64]),
    "qc_copy_state_from_host": (C.c_int, [_vp, _f64p, C.c_uint64, C.c_uint64]),
    "qc_ipc_export": (C.c_int, [_vp, C.c_char_p]),
    "qc_ipc_export_scratch": (C.c_int, [_vp, C.c_char_p]),
    "qc_scratch_device_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "qc_ipc_open": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "qc_ipc_close": (C.c_int, [_vp, _vp]),
    "qc_swap_with_peer": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int]),
    "qc_swap_scratch_with_peer": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int]),
    "qc_swap_local_pair": (C.c_int, [_vp, _vp, C.c_int]),
}  # fmt: skip

_lib = None


def load() -> C.CDLL:
    """Load the library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m qibochem_b200.build_native` "
            "(or __graft_entry__.build()). qibochem_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().qc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"qibochem_b200 native call failed{(' in ' + what) if what else ''}: {msg}")
