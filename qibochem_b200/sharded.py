"""
State sharding over G = 2^g GPUs, one process per GPU (the reference is single-process and has no
counterpart).

The amplitude index is split into ``n - g`` LOCAL bits (inside a shard) and ``g`` GLOBAL bits (the rank).  Which
*logical* index bit (logical bit b <-> qubit n-1-b) sits at which *physical* position is a permutation owned by the
host; masks are translated per call.  Consequences for a Pauli string (x, z):

* z on a global bit  -> a per-rank constant sign, folded into the rotation angle / term coefficient (no traffic);
* x on a global bit  -> the partner shard's amplitudes are needed: the global bit is first exchanged with a local
  bit that the upcoming strings do not pair on (``qc_swap_with_peer``: in-place peer-memory swap of half a shard
  over NVLink), after which every local kernel runs unchanged.  Evictions follow Belady's rule (farthest next use).

The exact energy finishes with one scalar all-reduce.  ``ShardedState`` only needs a *local shard* (``DeviceState``
API) and an *exchanger* (swap + all-reduce + barrier), so the host logic is testable on CPU with gloo.
"""

from __future__ import annotations

import numpy as np


def _planner_stats_fn():
    """``f(n_local, x, z, c) -> uint64[16]`` over the native planner (host-only), or None when the library is absent
    (CPU tests of the host logic with stand-in shards)."""
    try:
        import ctypes as C

        from qibochem_b200 import _native

        lib = _native.load()
    except Exception:
        return None
    u64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_double)

    def stats(n_local, x, z, c):
        x, z = np.ascontiguousarray(x, dtype=np.uint64), np.ascontiguousarray(z, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        out = np.zeros(16, dtype=np.uint64)
        _native.check(lib.qc_k2_plan_stats(int(n_local), x.size, x.ctypes.data_as(u64p), z.ctypes.data_as(u64p), c.ctypes.data_as(f64p),
                                           out.ctypes.data_as(u64p), None), "qc_k2_plan_stats")  # fmt: skip
        return out

    return stats


def _bits(mask: int):
    b = 0
    while mask >> b:
        if (mask >> b) & 1:
            yield b
        b += 1


class ShardedState:
    """n-qubit state over ``world`` = 2^g ranks.

    Args:
        n_qubits: total qubits
        rank, world: this process' rank and the (power of two) number of ranks
        local: object with the ``DeviceState`` API holding 2^(n-g) amplitudes
        exchanger: object with ``swap(global_k, local_pos)``, ``allreduce_sum(float) -> float``, ``barrier()``
        min_swap_pos: lowest local physical position eligible for an exchange (keeps peer accesses in long runs)
    """

    def __init__(self, n_qubits: int, rank: int, world: int, local, exchanger, min_swap_pos: int = 0):
        g = world.bit_length() - 1
        if 1 << g != world:
            raise ValueError("the number of ranks must be a power of two")
        if n_qubits - g < 1:
            raise ValueError("too many ranks for this register")
        self.n, self.g, self.nl = n_qubits, g, n_qubits - g
        self.rank, self.world = rank, world
        self.local, self.ex = local, exchanger
        self.min_swap_pos = min(min_swap_pos, self.nl - 1)
        self.pos = list(range(n_qubits))  # pos[logical bit] = physical position (>= nl: global bit pos-nl of the rank)
        self.n_swaps = 0
        self.last_passes = 0
        self.last_sweeps = 0
        self._layout_cache: dict = {}
        self._plan_stats = _planner_stats_fn()
        # "usage" (default): the g bits on which the fewest remaining terms pair become global.  "planner": candidates
        # priced with the tile planner.  Measured on the hardware the planner rule is neutral to slightly worse
        # (33 q / 2 GPUs: 302 vs 295 passes; 34 q / 4: 369 vs 305, 22.67 vs 22.5 s; 35 q / 8: 353 vs 341), so it is opt-in.
        import os

        self._use_planner = os.environ.get("QC_SHARD_LAYOUT", "usage") == "planner" and self._plan_stats is not None
        self._lambda_live = False  # adjoint gradient in progress: the scratch array (lambda) follows every re-layout
        self.n_qubits = n_qubits  # DeviceState-compatible surface: Circuit / SymbolicHamiltonian / measurement use it
        self.device = getattr(local, "device", 0)

    @property
    def dim(self) -> int:
        return 1 << self.n

    def synchronize(self) -> None:
        self.local.synchronize()

    def close(self) -> None:
        if hasattr(self.ex, "close"):
            self.ex.close()
        if hasattr(self.local, "close"):
            self.local.close()

    # ---- layout ---------------------------------------------------------------------------------------------
    def _is_global(self, logical_bit: int) -> bool:
        return self.pos[logical_bit] >= self.nl

    def _global_mask(self) -> int:
        return sum(1 << b for b in range(self.n) if self._is_global(b))

    def _to_physical(self, mask: int) -> tuple[int, int]:
        """logical mask -> (local physical mask, global part as a mask over rank bits)"""
        loc = glob = 0
        for b in _bits(mask):
            p = self.pos[b]
            if p < self.nl:
                loc |= 1 << p
            else:
                glob |= 1 << (p - self.nl)
        return loc, glob

    def _rank_sign(self, global_part: int) -> float:
        return -1.0 if bin(self.rank & global_part).count("1") & 1 else 1.0

    def _swap(self, global_logical: int, local_logical: int) -> None:
        """Exchange a global logical bit with a local one (data movement + permutation update)."""
        pg, pl = self.pos[global_logical], self.pos[local_logical]
        assert pg >= self.nl > pl
        self.ex.swap(pg - self.nl, pl, with_scratch=self._lambda_live)
        self.pos[global_logical], self.pos[local_logical] = pl, pg
        self.n_swaps += 1

    def _localise(self, needed_mask: int, future_x: list[int], protect_mask: int = 0) -> None:
        """Make every logical bit of `needed_mask` local, evicting the local bits whose next pairing use in
        `future_x` is farthest away (never one of `needed_mask | protect_mask`)."""
        todo = [b for b in _bits(needed_mask) if self._is_global(b)]
        if not todo:
            return
        banned = needed_mask | protect_mask
        candidates = [b for b in range(self.n) if not self._is_global(b) and not (banned >> b) & 1 and self.pos[b] >= self.min_swap_pos]
        if len(candidates) < len(todo):
            candidates = [b for b in range(self.n) if not self._is_global(b) and not (needed_mask >> b) & 1]
        if len(candidates) < len(todo):
            raise ValueError("a single Pauli string pairs on more bits than fit in one shard")
        next_use = {}
        for b in candidates:
            nu = len(future_x)
            for k, fx in enumerate(future_x):
                if (fx >> b) & 1:
                    nu = k
                    break
            next_use[b] = nu
        candidates.sort(key=lambda b: (-next_use[b], -self.pos[b]))
        for gb, lb in zip(todo, candidates):
            self._swap(gb, lb)

    # ---- K0 ---------------------------------------------------------------------------------------------------
    def set_basis_state(self, index: int) -> None:
        # The state is rewritten, so the bit permutation can be reset for free: every energy evaluation then walks
        # through the SAME sequence of layouts and the device-side plans of the local shard are reused.
        self.pos = list(range(self.n))
        loc, glob = self._to_physical(index)
        self.local.set_basis_state(loc, has_one=(glob == self.rank))

    # ---- K1 ---------------------------------------------------------------------------------------------------
    def apply_pauli_rotations(self, xmask, zmask, angle, lookahead: int = 4096) -> int:
        xs = [int(v) for v in np.asarray(xmask, dtype=np.uint64).tolist()]
        zs = [int(v) for v in np.asarray(zmask, dtype=np.uint64).tolist()]
        ang = np.asarray(angle, dtype=np.float64)
        R = len(xs)
        passes = 0
        i = 0
        while i < R:
            gm = self._global_mask()
            if xs[i] & gm:
                self._localise(xs[i], xs[i + 1 : i + 1 + lookahead])
                gm = self._global_mask()
            j = i
            while j < R and not (xs[j] & gm):
                j += 1
            px = np.empty(j - i, dtype=np.uint64)
            pz = np.empty(j - i, dtype=np.uint64)
            pa = np.empty(j - i, dtype=np.float64)
            for k in range(i, j):
                xl, _ = self._to_physical(xs[k])
                zl, zg = self._to_physical(zs[k])
                px[k - i], pz[k - i] = xl, zl
                pa[k - i] = ang[k] * self._rank_sign(zg)
            passes += self.local.apply_pauli_rotations(px, pz, pa)
            i = j
        self.last_passes = passes
        return passes

    # ---- K2 ---------------------------------------------------------------------------------------------------
    def expectation(self, xmask, zmask, coeff) -> float:
        """sum_k c_k Re<psi|P_k|psi> over the whole register (identical on every rank after the all-reduce)."""
        xs = np.asarray(xmask, dtype=np.uint64)
        zs = np.asarray(zmask, dtype=np.uint64)
        cs = np.asarray(coeff, dtype=np.float64)
        remaining = np.ones(xs.size, dtype=bool)
        partial = 0.0
        sweeps = 0
        first = True
        while remaining.any():
            gm = np.uint64(self._global_mask())
            ready = remaining & ((xs & gm) == 0)
            if not ready.any() or (first and self._use_planner):
                # the rotations leave an arbitrary layout: the first one is chosen like every later one
                self._relayout_for_terms(xs[remaining], zs[remaining], cs[remaining])
                gm = np.uint64(self._global_mask())
                ready = remaining & ((xs & gm) == 0)
                if not ready.any():
                    raise RuntimeError("could not find a layout that makes any remaining term local")
            first = False
            idx = np.nonzero(ready)[0]
            px, pz, pc = self._translate_terms(xs[idx], zs[idx], cs[idx])
            partial += self.local.expectation(px, pz, pc)
            sweeps += self.local.last_sweeps
            remaining[idx] = False
        self.last_sweeps = sweeps
        return self.ex.allreduce_sum(partial)

    def _translate_terms(self, xs, zs, cs):
        """Vectorised logical -> physical translation of term masks + rank signs."""
        px = np.zeros(xs.size, dtype=np.uint64)
        pz = np.zeros(xs.size, dtype=np.uint64)
        par = np.zeros(xs.size, dtype=np.uint64)
        one = np.uint64(1)
        for b in range(self.n):
            p = self.pos[b]
            xb = (xs >> np.uint64(b)) & one
            zb = (zs >> np.uint64(b)) & one
            if p < self.nl:
                px |= xb << np.uint64(p)
                pz |= zb << np.uint64(p)
            elif (self.rank >> (p - self.nl)) & 1:
                par ^= zb
        return px, pz, np.where(par == 1, -cs, cs)

    def _relayout_for_terms(self, xs_remaining: np.ndarray, zs_remaining=None, cs_remaining=None) -> None:
        """Choose the next set of g global logical bits for the remaining terms and swap it in.

        Default rule: the g bits on which the fewest remaining terms pair (most terms become local).  With
        ``QC_SHARD_LAYOUT=planner`` and a shard large enough for the tiled expectation, the candidates -- every g-subset of the 2g least-used bits, plus the
        current layout -- are priced with the tile planner itself (``qc_k2_plan_stats`` is host-only): passes, sub-cube
        visits and records of the terms that the candidate makes local, through the per-pass cost model fitted to the
        kernel, plus the exchanges it needs; the cheapest per served x-mask wins.  The choice is cached
        (every energy evaluation walks through the same sequence of layouts)."""
        usage = [(int(((xs_remaining >> np.uint64(b)) & np.uint64(1)).sum()), b) for b in range(self.n)]
        usage.sort()
        cur_global = [b for b in range(self.n) if self._is_global(b)]
        want_global = [b for _, b in usage[: self.g]]
        if self._use_planner and zs_remaining is not None and self.nl >= 16 and xs_remaining.size >= 512 and self.g <= 3:
            key = (hash(xs_remaining.tobytes()), hash(zs_remaining.tobytes()), tuple(cur_global))
            cached = self._layout_cache.get(key)
            if cached is None:
                cached = self._price_layouts(xs_remaining, zs_remaining, cs_remaining, [b for _, b in usage[: 2 * self.g]], cur_global)
                self._layout_cache[key] = cached
            want_global = list(cached)
        incoming = [b for b in cur_global if b not in want_global]  # become local
        outgoing = [b for b in want_global if b not in cur_global]  # become global
        for gb, lb in zip(incoming, outgoing):
            self._swap(gb, lb)

    def _price_layouts(self, xs, zs, cs, candidates, cur_global):
        from itertools import combinations

        best = None
        sets = [tuple(sorted(c)) for c in combinations(candidates, self.g)]
        if tuple(sorted(cur_global)) not in sets:
            sets.append(tuple(sorted(cur_global)))
        for G in sets:
            gm = np.uint64(sum(1 << b for b in G))
            sel = (xs & gm) == 0
            if not sel.any():
                continue
            # masks with the global bits squeezed out (any bijection onto the local positions prices the same plan shape)
            local_bits = [b for b in range(self.n) if b not in G]
            px = np.zeros(int(sel.sum()), dtype=np.uint64)
            pz = np.zeros_like(px)
            for p, b in enumerate(local_bits):
                px |= ((xs[sel] >> np.uint64(b)) & np.uint64(1)) << np.uint64(p)
                pz |= ((zs[sel] >> np.uint64(b)) & np.uint64(1)) << np.uint64(p)
            st = self._plan_stats(self.nl, px, pz, cs[sel])
            n_x, passes, visits, dots = int(st[0]), int(st[2]), int(st[3]), int(st[4])
            swaps = len(set(G) - set(cur_global))
            cost = (2.69 * passes + 0.446 * visits + 0.139 * dots + 14.0 * swaps) / max(1, n_x)  # ms at 30 local qubits
            if best is None or cost < best[0]:
                best = (cost, G)
        return best[1] if best else tuple(cur_global)

    def norm2(self) -> float:
        return self.ex.allreduce_sum(self.local.norm2())

    # ---- adjoint gradient (reference caller: qibo.derivative.parameter_shift, doc/source/tutorials/adaptive.rst:46-61)
    def energy_gradient(self, rxmask, rzmask, angle, hxmask, hzmask, coeff, lookahead: int = 4096):
        """``(sum_k c_k Re<psi|P_k|psi>, dE/d angle[R])`` with psi = U psi0 resident (the given rotation program was
        applied by ``apply_pauli_rotations``); leaves psi0.  lambda = H psi is built layout by layout like
        ``expectation``; afterwards every exchange moves psi AND lambda, and the rotations are walked backwards in
        maximal segments whose x-masks are local."""
        rxs = [int(v) for v in np.asarray(rxmask, dtype=np.uint64).tolist()]
        rzs = [int(v) for v in np.asarray(rzmask, dtype=np.uint64).tolist()]
        ang = np.asarray(angle, dtype=np.float64)
        xs = np.asarray(hxmask, dtype=np.uint64)
        zs = np.asarray(hzmask, dtype=np.uint64)
        cs = np.asarray(coeff, dtype=np.float64)
        R = len(rxs)
        grad = np.zeros(R, dtype=np.float64)
        if xs.size == 0:
            return 0.0, grad
        self.local.hpsi(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0), accumulate=False)  # lambda = 0
        self._lambda_live = True
        try:
            remaining = np.ones(xs.size, dtype=bool)
            while remaining.any():
                gm = np.uint64(self._global_mask())
                ready = remaining & ((xs & gm) == 0)
                if not ready.any():
                    self._relayout_for_terms(xs[remaining])
                    gm = np.uint64(self._global_mask())
                    ready = remaining & ((xs & gm) == 0)
                    if not ready.any():
                        raise RuntimeError("could not find a layout that makes any remaining term local")
                idx = np.nonzero(ready)[0]
                px, pz, pc = self._translate_terms(xs[idx], zs[idx], cs[idx])
                self.local.hpsi(px, pz, pc, accumulate=True)
                remaining[idx] = False
            energy = self.local.lambda_inner().real
            j = R
            while j > 0:
                gm = self._global_mask()
                if rxs[j - 1] & gm:
                    self._localise(rxs[j - 1], rxs[max(0, j - 1 - lookahead) : j - 1][::-1])
                    gm = self._global_mask()
                i = j
                while i > 0 and not (rxs[i - 1] & gm):
                    i -= 1
                px = np.empty(j - i, dtype=np.uint64)
                pz = np.empty(j - i, dtype=np.uint64)
                pa = np.empty(j - i, dtype=np.float64)
                sg = np.empty(j - i, dtype=np.float64)
                for k in range(i, j):
                    xl, _ = self._to_physical(rxs[k])
                    zl, zg = self._to_physical(rzs[k])
                    px[k - i], pz[k - i] = xl, zl
                    sg[k - i] = self._rank_sign(zg)
                    pa[k - i] = ang[k] * sg[k - i]
                grad[i:j] = sg * self.local.gradient_backward(px, pz, pa)  # chain rule: local angle = sign * angle
                j = i
        finally:
            self._lambda_live = False
        total = self.ex.allreduce_array(np.concatenate([[energy], grad]))
        return float(total[0]), total[1:]

    # ---- host copies (small registers / tests) -------------------------------------------------------------------
    def from_numpy(self, state: np.ndarray) -> None:
        """Scatter a full LOGICAL-order state: the layout is reset to the identity, rank r keeps slice r."""
        psi = np.ascontiguousarray(np.asarray(state, dtype=np.complex128)).reshape(-1)
        if psi.size != 1 << self.n:
            raise ValueError("state vector has the wrong size")
        self.pos = list(range(self.n))
        self.local.from_numpy(psi[self.rank << self.nl : (self.rank + 1) << self.nl])

    def to_numpy(self) -> np.ndarray:
        return self.gather_logical()

    # ---- K5 / K3: sampled path (per-rank mass -> host multinomial -> per-rank shots -> integer partial
    #      sums -> all-reduce).  Reference call site: src/qibochem/measurement/result.py:105-117 ---------------------
    def sample(self, x_basis_mask: int, y_basis_mask: int, n_shots: int, seed: int, shot_offset: int = 0,
               on_device: bool = True, return_total: bool = False):  # fmt: skip
        """Draw ``n_shots`` shots of the whole register in the product basis, split over the ranks: rank r draws
        Multinomial(n_shots, mass)[r] shots from ITS shard's conditional distribution, with Philox counters offset by
        the shots of the lower ranks (disjoint streams).  Returns a ``ShardedSamples`` (this rank's part)."""
        rot = int(x_basis_mask) | int(y_basis_mask)
        if bin(rot).count("1") > self.nl:
            raise NotImplementedError("more X/Y-measured qubits than local index bits of a shard")
        # a basis rotation on a global bit would mix shards: make the rotated bits local first
        self._localise(rot, [], protect_mask=0)
        xl, xg = self._to_physical(int(x_basis_mask))
        yl, yg = self._to_physical(int(y_basis_mask))
        assert xg == 0 and yg == 0
        _, mass = self.local.sample(xl, yl, 0, seed, on_device=on_device, return_total=True)
        masses = np.array(self.ex.allgather_floats(mass), dtype=np.float64)
        p = np.clip(masses, 0.0, None)
        p = p / p.sum()
        # identical on every rank: same generator, same probabilities (bit-identical after the all-gather)
        split = np.random.default_rng([int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF, 0x5A17]).multinomial(int(n_shots), p)
        mine = int(split[self.rank])
        offset = int(shot_offset) + int(split[: self.rank].sum())
        local = self.local.sample(xl, yl, mine, seed, shot_offset=offset, on_device=on_device) if mine else None
        out = ShardedSamples(self, local, mine, list(self.pos), int(n_shots))
        return (out, float(masses.sum())) if return_total else out

    def _masks_to_physical(self, masks, pos):
        """logical masks -> (local physical masks, parity of this rank's bits under the global part)."""
        m = np.asarray(masks, dtype=np.uint64)
        loc = np.zeros(m.size, dtype=np.uint64)
        par = np.zeros(m.size, dtype=np.uint64)
        one = np.uint64(1)
        for b in range(self.n):
            bit = (m >> np.uint64(b)) & one
            p = pos[b]
            if p < self.nl:
                loc |= bit << np.uint64(p)
            elif (self.rank >> (p - self.nl)) & 1:
                par ^= bit
        return loc, par

    def parity_sums(self, samples: "ShardedSamples", masks, counts=None) -> np.ndarray:
        """Exact int64 sums S_j = sum_shots (-1)^{popcount(shot & mask_j)} over ALL ranks' shots (K3 partials, the
        global bits of a mask contribute a per-rank sign, one integer all-reduce)."""
        if counts is not None:
            raise NotImplementedError("weighted sums over sharded samples")
        loc, par = self._masks_to_physical(masks, samples.pos)
        if samples.n_local:
            part = self.local.parity_sums(samples.local, loc)
        else:
            part = np.zeros(loc.size, dtype=np.int64)
        part = np.where(par == 1, -part, part).astype(np.int64)
        return self.ex.allreduce_int64(part)

    def frequencies(self, samples: "ShardedSamples", keep_mask: int):
        """Histogram over the ``keep_mask`` bits (logical index-bit positions) of all ranks' shots -> (keys compacted
        like DeviceState.frequencies, ascending; counts)."""
        idx = samples.logical_indices_all_ranks()
        kept = [b for b in range(self.n) if (int(keep_mask) >> b) & 1]
        comp = np.zeros(idx.size, dtype=np.uint64)
        for j, b in enumerate(kept):
            comp |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(j)
        keys, counts = np.unique(comp, return_counts=True)
        return keys.astype(np.uint64), counts.astype(np.int64)

    def outcome_variance(self, keys, counts, masks, coeffs, mean: float):
        return self.local.outcome_variance(keys, counts, masks, coeffs, mean)  # host arrays in, rank-independent

    # ---- tests / small registers ------------------------------------------------------------------------------
    def gather_logical(self) -> np.ndarray:
        """Full state in LOGICAL index order on every rank (small n only; used by tests)."""
        shard = self.local.to_numpy()
        full_phys = self.ex.allgather_array(shard)  # physical order: rank-major
        idx = np.arange(1 << self.n, dtype=np.uint64)
        phys = np.zeros_like(idx)
        for b in range(self.n):
            phys |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(self.pos[b])
        return full_phys[phys]


class ShardedSamples:
    """This rank's part of a sharded measurement: local physical outcome indices + the layout they were drawn in."""

    def __init__(self, state: ShardedState, local, n_local: int, pos: list[int], n_total: int):
        self.state, self.local, self.n_local, self.pos, self.n_total = state, local, n_local, pos, n_total

    def _local_numpy(self) -> np.ndarray:
        if not self.n_local:
            return np.zeros(0, dtype=np.uint64)
        loc = self.local
        if hasattr(loc, "cpu"):
            loc = loc.cpu().numpy()
        return np.asarray(loc).astype(np.uint64)

    def logical_indices_all_ranks(self) -> np.ndarray:
        """All shots of the measurement as LOGICAL full-register indices, rank-major (identical on every rank)."""
        st = self.state
        phys = self._local_numpy() | (np.uint64(st.rank) << np.uint64(st.nl))
        logical = np.zeros(phys.size, dtype=np.uint64)
        for b in range(st.n):
            logical |= ((phys >> np.uint64(self.pos[b])) & np.uint64(1)) << np.uint64(b)
        parts = st.ex.allgather_object(logical)
        return np.concatenate(parts) if parts else logical

    def cpu(self):
        return self

    def numpy(self) -> np.ndarray:
        return self.logical_indices_all_ranks()


class IpcExchanger:
    """Exchanger for one-process-per-GPU runs: torch.distributed (NCCL) for barriers / all-reduce, CUDA IPC peer
    pointers + the in-place swap kernel for the data movement."""

    def __init__(self, state, rank: int, world: int, device: int):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.state, self.rank, self.world, self.device = state, rank, world, device
        handles = [None] * world
        dist.all_gather_object(handles, state.ipc_export())
        self.peer = {r: state.ipc_open(handles[r]) for r in range(world) if r != rank}
        self.peer_scratch: dict[int, int] = {}  # opened lazily: only the adjoint gradient moves the scratch arrays
        self._buf = torch.zeros(1, dtype=torch.float64, device=f"cuda:{device}")
        self.swap_seconds = 0.0
        self.swap_bytes = 0

    def barrier(self) -> None:
        self.state.synchronize()
        self.dist.all_reduce(self._buf)  # NCCL barrier on torch's stream
        self.torch.cuda.current_stream().synchronize()

    def _open_scratch_peers(self) -> None:
        handles = [None] * self.world
        self.dist.all_gather_object(handles, self.state.ipc_export_scratch())
        self.peer_scratch = {r: self.state.ipc_open(handles[r]) for r in range(self.world) if r != self.rank}

    def swap(self, global_k: int, local_pos: int, with_scratch: bool = False) -> None:
        import time

        if with_scratch and not self.peer_scratch:
            self._open_scratch_peers()
        partner = self.rank ^ (1 << global_k)
        side = (self.rank >> global_k) & 1
        self.barrier()  # both shards idle and consistent
        t0 = time.perf_counter()
        self.state.swap_with_peer(self.peer[partner], local_pos, side, side)
        if with_scratch:
            self.state.swap_scratch_with_peer(self.peer_scratch[partner], local_pos, side, side)
        self.barrier()  # both halves of the exchange have landed
        self.swap_seconds += time.perf_counter() - t0
        self.swap_bytes += self.state.dim * 16 // 2 * (2 if with_scratch else 1)  # bytes leaving (= entering) this GPU

    def allreduce_sum(self, value: float) -> float:
        t = self.torch.tensor([value], dtype=self.torch.float64, device=f"cuda:{self.device}")
        self.dist.all_reduce(t)
        return float(t.item())

    def allreduce_array(self, values: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).to(f"cuda:{self.device}")
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def allreduce_int64(self, values: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(values, dtype=np.int64)).to(f"cuda:{self.device}")
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def allgather_floats(self, value: float) -> list[float]:
        t = self.torch.tensor([value], dtype=self.torch.float64, device=f"cuda:{self.device}")
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def allgather_object(self, obj) -> list:
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def allgather_array(self, arr: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.float64)).to(f"cuda:{self.device}")
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.concatenate([o.cpu().numpy().view(np.complex128) for o in out])

    def close(self) -> None:
        self.barrier()
        for p in list(self.peer.values()) + list(self.peer_scratch.values()):
            self.state.ipc_close(p)
        self.peer, self.peer_scratch = {}, {}
