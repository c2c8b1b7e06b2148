"""
Thin Python owner of one device shard (``qc_state`` of the C ABI).

This is the only module that talks to the native library.  Masks passed here are in index-bit positions of the
shard (qubit q of an n-qubit register <-> bit n-1-q, qibo's convention; see ``include/qibochem_b200.h``).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from qibochem_b200 import _native


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def device_count() -> int:
    lib = _native.load()
    n = C.c_int(0)
    _native.check(lib.qc_device_count(C.byref(n)), "qc_device_count")
    return n.value


class DeviceState:
    """complex128 state vector of ``n_qubits`` (local) qubits resident in the HBM of one B200."""

    def __init__(self, n_qubits: int, device: int = 0, stream: int | None = None):
        self._lib = _native.load()
        self._h = C.c_void_p()
        self.n_qubits = int(n_qubits)
        self.device = int(device)
        _native.check(
            self._lib.qc_create(self.device, self.n_qubits, C.c_void_p(stream) if stream else None, C.byref(self._h)),
            "qc_create",
        )
        self.last_passes = 0
        self.last_sweeps = 0

    # -- life cycle -------------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.qc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def dim(self) -> int:
        return 1 << self.n_qubits

    @property
    def device_ptr(self) -> int:
        p = C.c_void_p()
        _native.check(self._lib.qc_state_device_ptr(self._h, C.byref(p)), "qc_state_device_ptr")
        return p.value

    def synchronize(self) -> None:
        _native.check(self._lib.qc_synchronize(self._h), "qc_synchronize")

    # -- K0 ---------------------------------------------------------------------------------------------------
    def set_basis_state(self, index: int, has_one: bool = True) -> None:
        _native.check(self._lib.qc_set_basis_state(self._h, C.c_uint64(int(index)), int(bool(has_one))), "qc_set_basis_state")

    # -- K1 ---------------------------------------------------------------------------------------------------
    def apply_pauli_rotations(self, xmask, zmask, angle) -> int:
        x, z, a = _u64(xmask), _u64(zmask), _f64(angle)
        if not (x.shape == z.shape == a.shape and x.ndim == 1):
            raise ValueError("xmask, zmask and angle must be 1-d arrays of equal length")
        passes = C.c_int(0)
        _native.check(
            self._lib.qc_apply_pauli_rotations(
                self._h, x.size, _ptr(x, C.c_uint64), _ptr(z, C.c_uint64), _ptr(a, C.c_double), C.byref(passes)
            ),
            "qc_apply_pauli_rotations",
        )
        self.last_passes = passes.value
        return passes.value

    def prepare_state(self, index: int, xmask, zmask, angle, has_one: bool = True) -> int:
        """K0 + K1 in one native call: basis state ``index``, then the rotation program (one launch on a small register)."""
        x, z, a = _u64(xmask), _u64(zmask), _f64(angle)
        if not (x.shape == z.shape == a.shape and x.ndim == 1):
            raise ValueError("xmask, zmask and angle must be 1-d arrays of equal length")
        passes = C.c_int(0)
        _native.check(
            self._lib.qc_prepare_state(self._h, C.c_uint64(int(index)), int(bool(has_one)), x.size, _ptr(x, C.c_uint64),
                                       _ptr(z, C.c_uint64), _ptr(a, C.c_double), C.byref(passes)),
            "qc_prepare_state",
        )  # fmt: skip
        self.last_passes = passes.value
        return passes.value

    # -- K2 ---------------------------------------------------------------------------------------------------
    def expectation(self, xmask, zmask, coeff, per_term: bool = False):
        x, z, c = _u64(xmask), _u64(zmask), _f64(coeff)
        if not (x.shape == z.shape == c.shape and x.ndim == 1):
            raise ValueError("xmask, zmask and coeff must be 1-d arrays of equal length")
        e = C.c_double(0.0)
        sweeps = C.c_int(0)
        pt = np.zeros(x.size, dtype=np.float64) if per_term else None
        _native.check(
            self._lib.qc_expectation(
                self._h, x.size, _ptr(x, C.c_uint64), _ptr(z, C.c_uint64), _ptr(c, C.c_double), C.byref(e),
                _ptr(pt, C.c_double) if per_term else None, C.byref(sweeps),
            ),
            "qc_expectation",
        )  # fmt: skip
        self.last_sweeps = sweeps.value
        return (e.value, pt) if per_term else e.value

    def energy_gradient(self, rxmask, rzmask, angle, hxmask, hzmask, coeff):
        """Adjoint-mode gradient: the state must hold U|psi0> for the given rotation program.  Returns
        ``(sum_k c_k Re<psi|P_k|psi>, dE/d angle[R])`` and leaves |psi0> in the state."""
        rx, rz, ra = _u64(rxmask), _u64(rzmask), _f64(angle)
        hx, hz, hc = _u64(hxmask), _u64(hzmask), _f64(coeff)
        if not (rx.shape == rz.shape == ra.shape and hx.shape == hz.shape == hc.shape and rx.ndim == hx.ndim == 1):
            raise ValueError("mask / angle / coefficient arrays must be 1-d and of matching lengths")
        e = C.c_double(0.0)
        grad = np.zeros(rx.size, dtype=np.float64)
        _native.check(
            self._lib.qc_energy_gradient(
                self._h, rx.size, _ptr(rx, C.c_uint64), _ptr(rz, C.c_uint64), _ptr(ra, C.c_double),
                hx.size, _ptr(hx, C.c_uint64), _ptr(hz, C.c_uint64), _ptr(hc, C.c_double), C.byref(e), _ptr(grad, C.c_double),
            ),
            "qc_energy_gradient",
        )  # fmt: skip
        return e.value, grad

    def hpsi(self, hxmask, hzmask, coeff, accumulate: bool = False) -> None:
        """lambda (+)= sum_k c_k P_k psi into the handle's scratch array (first stage of the adjoint gradient)."""
        hx, hz, hc = _u64(hxmask), _u64(hzmask), _f64(coeff)
        if not (hx.shape == hz.shape == hc.shape and hx.ndim == 1):
            raise ValueError("xmask, zmask and coeff must be 1-d arrays of equal length")
        _native.check(
            self._lib.qc_hpsi(self._h, hx.size, _ptr(hx, C.c_uint64), _ptr(hz, C.c_uint64), _ptr(hc, C.c_double), int(bool(accumulate))),
            "qc_hpsi",
        )

    def lambda_inner(self) -> complex:
        """<lambda|psi> over this shard."""
        re, im = C.c_double(0.0), C.c_double(0.0)
        _native.check(self._lib.qc_lambda_inner(self._h, C.byref(re), C.byref(im)), "qc_lambda_inner")
        return complex(re.value, im.value)

    def gradient_backward(self, rxmask, rzmask, angle) -> np.ndarray:
        """This shard's part of dE/d angle[r] for the given rotations (walked last to first); psi and lambda are left
        with the rotations un-applied."""
        rx, rz, ra = _u64(rxmask), _u64(rzmask), _f64(angle)
        if not (rx.shape == rz.shape == ra.shape and rx.ndim == 1):
            raise ValueError("xmask, zmask and angle must be 1-d arrays of equal length")
        grad = np.zeros(rx.size, dtype=np.float64)
        _native.check(
            self._lib.qc_gradient_backward(self._h, rx.size, _ptr(rx, C.c_uint64), _ptr(rz, C.c_uint64), _ptr(ra, C.c_double), _ptr(grad, C.c_double)),
            "qc_gradient_backward",
        )
        return grad

    def norm2(self) -> float:
        v = C.c_double(0.0)
        _native.check(self._lib.qc_norm2(self._h, C.byref(v)), "qc_norm2")
        return v.value

    # -- K5 / K3 ----------------------------------------------------------------------------------------------
    def sample(self, x_basis_mask: int, y_basis_mask: int, n_shots: int, seed: int, shot_offset: int = 0,
               on_device: bool = False, return_total: bool = False):  # fmt: skip
        """Draw shots of all local qubits in the product basis.  Returns uint64 indices (numpy array, or a torch
        int64 CUDA tensor holding the same bits when ``on_device``)."""
        total = C.c_double(0.0)
        if on_device:
            import torch  # device memory plumbing only

            out = torch.empty(max(int(n_shots), 1), dtype=torch.int64, device=f"cuda:{self.device}")
            ptr = C.c_void_p(out.data_ptr())
        else:
            out = np.zeros(int(n_shots), dtype=np.uint64)
            ptr = C.c_void_p(out.ctypes.data)
        _native.check(
            self._lib.qc_sample(
                self._h, C.c_uint64(int(x_basis_mask)), C.c_uint64(int(y_basis_mask)), int(n_shots),
                C.c_uint64(int(seed) & (2**64 - 1)), C.c_uint64(int(shot_offset)), ptr, int(on_device), C.byref(total),
            ),
            "qc_sample",
        )  # fmt: skip
        if on_device:
            out = out[: int(n_shots)]
        return (out, total.value) if return_total else out

    @staticmethod
    def _buf(samples):
        """(pointer, n, on_device, keepalive) for a numpy uint64 array or a torch CUDA int64 tensor."""
        if isinstance(samples, np.ndarray) or not hasattr(samples, "data_ptr"):
            arr = _u64(samples)
            return C.c_void_p(arr.ctypes.data), arr.size, 0, arr
        t = samples.contiguous()
        return C.c_void_p(t.data_ptr()), t.numel(), 1, t

    def parity_sums(self, samples, masks, counts=None) -> np.ndarray:
        """S_j = sum_s w_s (-1)^{popcount(samples[s] & masks[j])}  (exact int64)."""
        ptr, n, on_dev, keep = self._buf(samples)
        m = _u64(masks)
        out = np.zeros(m.size, dtype=np.int64)
        cptr, ckeep = None, None
        if counts is not None:
            if on_dev:
                ckeep = counts.contiguous()
                cptr = C.c_void_p(ckeep.data_ptr())
            else:
                ckeep = np.ascontiguousarray(np.asarray(counts, dtype=np.int64))
                if ckeep.size != n:
                    raise ValueError("counts and samples differ in length")
                cptr = C.c_void_p(ckeep.ctypes.data)
        _native.check(
            self._lib.qc_parity_sums(self._h, ptr, cptr, n, on_dev, _ptr(m, C.c_uint64), m.size, _ptr(out, C.c_int64)),
            "qc_parity_sums",
        )
        del keep, ckeep
        return out

    def outcome_variance(self, keys, counts, masks, coeffs, mean: float) -> tuple[float, float]:
        """(sum_u (v_u - mean)^2, sum_u count_u (v_u - mean)^2) over distinct outcomes, v_u = sum_j c_j (-1)^{popc(key_u & m_j)}."""
        k, m = _u64(keys), _u64(masks)
        c = np.ascontiguousarray(np.asarray(counts, dtype=np.int64))
        w = _f64(coeffs)
        if k.shape != c.shape or m.shape != w.shape:
            raise ValueError("keys/counts and masks/coeffs must have matching lengths")
        a, b = C.c_double(0.0), C.c_double(0.0)
        _native.check(
            self._lib.qc_outcome_variance(self._h, _ptr(k, C.c_uint64), _ptr(c, C.c_int64), k.size, _ptr(m, C.c_uint64),
                                          _ptr(w, C.c_double), m.size, C.c_double(float(mean)), C.byref(a), C.byref(b)),
            "qc_outcome_variance",
        )  # fmt: skip
        return a.value, b.value

    def frequencies(self, samples, keep_mask: int):
        """Histogram of the samples restricted to ``keep_mask`` bits -> (keys ascending, counts)."""
        ptr, n, on_dev, keep = self._buf(samples)
        cap = max(1, min(n, 1 << min(bin(int(keep_mask)).count("1"), 62)))
        keys = np.zeros(cap, dtype=np.uint64)
        cnts = np.zeros(cap, dtype=np.int64)
        nu = C.c_size_t(0)
        _native.check(
            self._lib.qc_frequencies(
                self._h, ptr, n, on_dev, C.c_uint64(int(keep_mask)), _ptr(keys, C.c_uint64), _ptr(cnts, C.c_int64), cap, C.byref(nu)
            ),
            "qc_frequencies",
        )  # fmt: skip
        del keep
        return keys[: nu.value].copy(), cnts[: nu.value].copy()

    # -- host copies ------------------------------------------------------------------------------------------
    def to_numpy(self, offset: int = 0, count: int | None = None) -> np.ndarray:
        count = self.dim - offset if count is None else count
        out = np.empty(count, dtype=np.complex128)
        _native.check(
            self._lib.qc_copy_state_to_host(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(offset), C.c_uint64(count)),
            "qc_copy_state_to_host",
        )  # fmt: skip
        return out

    def from_numpy(self, state: np.ndarray, offset: int = 0) -> None:
        arr = np.ascontiguousarray(np.asarray(state, dtype=np.complex128))
        _native.check(
            self._lib.qc_copy_state_from_host(self._h, arr.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(offset), C.c_uint64(arr.size)),
            "qc_copy_state_from_host",
        )  # fmt: skip

    # -- profiler ----------------------------------------------------------------------------------------------
    KERNEL_CLASSES = ("k0_init", "k1_rotation", "k2_expectation", "reduce", "k5_sampling", "k3_shots", "k4_exchange", "k2_tiled", "gradient")

    def set_option(self, key: str, value: int) -> None:
        """e.g. ``set_option("k2_mode", 1)``: 0 auto, 1 sweeps (one HBM pass per x-mask), 2 tiles."""
        _native.check(self._lib.qc_set_option(self._h, key.encode(), int(value)), "qc_set_option")

    def took_real_path(self) -> bool:
        """Whether the last tiled expectation ran the real-amplitude kernels (option ``k2_real``)."""
        return self._lib.qc_set_option(self._h, b"k2_real_taken", 0) == 0

    def transfer_bytes(self, reset: bool = False) -> tuple[int, int]:
        """(host->device, device->host) bytes moved by the compute entry points since the last reset."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        _native.check(self._lib.qc_transfer_bytes(self._h, C.byref(a), C.byref(b), int(reset)), "qc_transfer_bytes")
        return a.value, b.value

    def profile_enable(self, enabled: bool = True) -> None:
        _native.check(self._lib.qc_profile_enable(self._h, int(enabled)), "qc_profile_enable")

    def profile_reset(self) -> None:
        _native.check(self._lib.qc_profile_reset(self._h), "qc_profile_reset")

    def profile(self) -> dict:
        """{class: {"ms": device time (0 unless enabled), "launches": n, "bytes": algorithmic bytes}}"""
        out = {}
        for k, name in enumerate(self.KERNEL_CLASSES):
            ms, n, b = C.c_double(0.0), C.c_uint64(0), C.c_double(0.0)
            _native.check(self._lib.qc_profile_get(self._h, k, C.byref(ms), C.byref(n), C.byref(b)), "qc_profile_get")
            out[name] = {"ms": ms.value, "launches": n.value, "bytes": b.value}
        return out

    # -- K4 ---------------------------------------------------------------------------------------------------
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        _native.check(self._lib.qc_ipc_export(self._h, buf), "qc_ipc_export")
        return buf.raw

    def ipc_export_scratch(self) -> bytes:
        buf = C.create_string_buffer(64)
        _native.check(self._lib.qc_ipc_export_scratch(self._h, buf), "qc_ipc_export_scratch")
        return buf.raw

    @property
    def scratch_device_ptr(self) -> int:
        p = C.c_void_p()
        _native.check(self._lib.qc_scratch_device_ptr(self._h, C.byref(p)), "qc_scratch_device_ptr")
        return p.value

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        _native.check(self._lib.qc_ipc_open(self._h, C.create_string_buffer(handle, 64), C.byref(p)), "qc_ipc_open")
        return p.value

    def ipc_close(self, peer_ptr: int) -> None:
        _native.check(self._lib.qc_ipc_close(self._h, C.c_void_p(peer_ptr)), "qc_ipc_close")

    def swap_with_peer(self, peer_ptr: int, local_bit: int, my_side: int, part: int | None = None) -> None:
        part = my_side if part is None else part
        _native.check(
            self._lib.qc_swap_with_peer(self._h, C.c_void_p(peer_ptr), int(local_bit), int(my_side), int(part)),
            "qc_swap_with_peer",
        )

    def swap_scratch_with_peer(self, peer_scratch_ptr: int, local_bit: int, my_side: int, part: int | None = None) -> None:
        part = my_side if part is None else part
        _native.check(
            self._lib.qc_swap_scratch_with_peer(self._h, C.c_void_p(peer_scratch_ptr), int(local_bit), int(my_side), int(part)),
            "qc_swap_scratch_with_peer",
        )

    def swap_local_pair(self, other: "DeviceState", local_bit: int) -> None:
        """self holds the global-bit-0 shard, ``other`` the global-bit-1 shard (same process)."""
        _native.check(self._lib.qc_swap_local_pair(self._h, other._h, int(local_bit)), "qc_swap_local_pair")
