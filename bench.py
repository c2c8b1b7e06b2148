#!/usr/bin/env python
"""
bench.py -- UCCSD energy evaluations per second on the named synthetic configuration (BASELINE.json configs[3]/[4]).

One STEP = one energy evaluation of the hot path: HF basis state (K0) -> R UCCSD Pauli rotations (K1) -> expectation
of a T-term molecular-structured Hamiltonian (K2) [-> all-reduce over shards when N > 1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--qubits n] [--rotations R] [--terms T]

N = 1: 30 qubits (17.18 GB state >> 126 MB L2, so no L2 flush is needed between steps), R = 1024 rotations (the
first 128 double excitations of the (30, 14) UCCSD order), T = 100 000 terms over 19 012 x-masks.
N > 1 (launched by torchrun, one rank per GPU): weak-scaling ladder 33 q @ 2, 34 q @ 4, 35 q @ 8 (2^32 amplitudes per
GPU), reported in 30-qubit-equivalent evaluations per second (x 2^(n-30)) so that the value is extensive.

The JSON line carries the driver contract keys plus `roofline` (dominant kernel, device time from CUDA events on the
launching stream), `cpu_baseline` (C/OpenMP port of the reference's gate-by-gate algorithm on the host cores, bounded
sample) and `e2e` (the same step through the public Python API with host arrays in, energy out).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "uccsd_energy_evals_per_sec"
DEFAULT_QUBITS = {1: 30, 2: 33, 4: 34, 8: 35}


def unit_for(n: int) -> str:
    """The metric is quoted per 30-qubit evaluation; a larger register counts 2^(n-30) of them (extensive value)."""
    return "evals/s" if n == 30 else "evals/s (30-qubit equivalent: x 2^(n-30))"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=None)
    ap.add_argument("--electrons", type=int, default=None)
    ap.add_argument("--rotations", type=int, default=1024, help="UCCSD rotation prefix (8 per double excitation)")
    ap.add_argument("--terms", type=int, default=100_000)
    ap.add_argument("--cpu-sample-qubits", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the oracle / 1-GPU-twin parity checks")
    ap.add_argument("--no-gradient", action="store_true", help="skip the adjoint-gradient measurement (one call, ~30 s at 30 qubits)")
    ap.add_argument("--no-ladder-point", action="store_true", help="skip the 32-qubit single-GPU point of the weak-scaling ladder")
    ap.add_argument("--no-small-configs", action="store_true", help="skip BASELINE configs 1-3 (4 / 12 / 14 qubits)")
    ap.add_argument("--no-full-stream", action="store_true", help="N = 1: skip the extra full-UCCSD-stream measurement")
    ap.add_argument("--full-stream", action="store_true", help="N > 1: add one step with the complete UCCSD rotation stream (60-80 s)")
    return ap.parse_args()


def workload(n: int, nelec: int, rotations: int, terms: int):
    from qibochem_b200 import synthetic

    rx, rz, ra, n_exc = synthetic.uccsd_rotation_stream(n, nelec, max_excitations=max(1, rotations // 8))
    rx, rz, ra = rx[:rotations], rz[:rotations], ra[:rotations]
    hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, terms)
    return {"n": n, "nelec": nelec, "rx": rx, "rz": rz, "ra": ra, "hx": hx, "hz": hz, "hc": hc, "const": const,
            "hf": synthetic.hf_index(n, nelec), "n_xmasks": int(np.unique(hx).size)}  # fmt: skip


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=self.tmp, stderr=subprocess.DEVNULL,
            )  # fmt: skip
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), smax.append(float(parts[1])), power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}  # fmt: skip


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's algorithm (gate-by-gate _expi_pauli programs + per-term expectation), C/OpenMP port
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(wl_full: dict, sample_qubits: int, budget_s: float = 20.0) -> dict:
    """Times a bounded sample of the SAME workload shape at `sample_qubits` and scales to the full register:
    every gate / term factor is one pass over 2^n amplitudes, so time scales with 2^n (memory-bound; stated in the
    `sample` string).  Returns evals/s of the full workload on this host."""
    from oracle import oracle as O
    from oracle import oracle_c as OC
    from qibochem_b200.ansatz._fermion import excitation_generator
    from qibochem_b200.ansatz.utils import generate_excitations
    from qibochem_b200.pauli import masks_to_string

    OC.set_num_threads(0)  # every logical core: torchrun's OMP_NUM_THREADS=1 must not leak into the baseline
    n_full, ns = wl_full["n"], min(sample_qubits, wl_full["n"])
    nelec_s = max(2, round(wl_full["nelec"] * ns / n_full))
    nelec_s -= nelec_s % 2
    # sample rotations: the first double excitations of the (ns, nelec_s) UCCSD order, as reference gate programs
    excs = generate_excitations(2, range(nelec_s), range(nelec_s, ns))
    state = np.zeros(1 << ns, dtype=np.complex128)
    state[0] = 1.0
    OC.run_gates(state, ns, O.hf_gates(ns, nelec_s))  # also warms the OpenMP pool
    t_rot, n_rot, n_gates = 0.0, 0, 0
    for exc in excs:
        gates = O.ucc_gates(exc, 0.05)
        t0 = time.perf_counter()
        OC.run_gates(state, ns, gates)
        t_rot += time.perf_counter() - t0
        n_rot += 8
        n_gates += len(gates)
        if t_rot > budget_s / 2:
            break
    # sample terms: molecular-structured strings on ns qubits, qibo-style (copy + one pass per factor + dot)
    from qibochem_b200 import synthetic
    from qibochem_b200.pauli import reverse_bits

    hx, hz, hc, _ = synthetic.molecular_like_hamiltonian(ns, min(wl_full["hx"].size, 2000))
    pick = np.random.default_rng(0).choice(hx.size, size=min(hx.size, 400), replace=False)
    tmp = np.empty_like(state)
    t_term, n_term = 0.0, 0
    for k in pick:
        xq, zq = reverse_bits(int(hx[k]), ns), reverse_bits(int(hz[k]), ns)
        factors = [(int(op[1:]), op[0]) for op in masks_to_string(xq, zq).split()]
        t0 = time.perf_counter()
        OC.expectation_terms(state, ns, [(float(hc[k]), factors)], tmp)
        t_term += time.perf_counter() - t0
        n_term += 1
        if t_term > budget_s / 2:
            break
    # average factor count differs between n: scale per-term cost by (weight+2) passes (copy, factors, dot)
    scale = float(2 ** (n_full - ns))
    # reference gates per rotation grow with n (longer CNOT ladders): use the exact gate census of the full stream
    full_gates_per_rot = reference_gates_per_rotation(wl_full)
    t_gate = t_rot / max(n_gates, 1)
    w_sample = np.mean([bin(int(hx[k]) | int(hz[k])).count("1") + 2 for k in pick[:n_term]])
    w_full = float(np.mean([bin(int(a) | int(b)).count("1") + 2 for a, b in zip(wl_full["hx"][::97], wl_full["hz"][::97])]))
    t_eval = scale * (t_gate * full_gates_per_rot * wl_full["rx"].size + (t_term / max(n_term, 1)) * (w_full / w_sample) * wl_full["hx"].size)
    return {
        "value": 1.0 / t_eval,
        "unit": "evals/s",
        "cores": OC.num_threads(),
        "kind": "port",
        "sample": (
            f"C/OpenMP port of the reference algorithm (in-place gate passes of _expi_pauli programs + qibo-style "
            f"per-term expectation) timed at {ns} qubits: {n_gates} gate passes ({n_rot} rotations) in {t_rot:.2f}s, "
            f"{n_term} terms in {t_term:.2f}s; scaled x2^{n_full - ns} (every pass is memory-bound, linear in 2^n), to "
            f"{full_gates_per_rot:.1f} gates/rotation x {wl_full['rx'].size} rotations and {wl_full['hx'].size} terms of mean "
            f"{w_full:.1f} passes"
        ),
        "seconds_per_eval_extrapolated": t_eval,
        # the value is an EXTRAPOLATION of the measured sample, never a measured full-size run
        "extrapolated": True,
        "sample_qubits": ns,
        "scale": scale,
        "sample_seconds_measured": t_rot + t_term,
        "host_logical_cores": os.cpu_count(),
    }


def cpu_scaling_points(gate_budget_s: float = 12.0) -> dict:
    """per-gate-pass time of the two CPU restatements at n = 20, 22, 24, 26 (MEASURED), the effective
    bytes/s they give (one gate pass reads and writes every amplitude: 32 B x 2^n), and the 30-qubit figure they
    extrapolate to.  "qibojit-equivalent" = in-place C/OpenMP passes on every host core (oracle/oracle_c.c);
    "qibo-numpy-equivalent" = out-of-place single-threaded numpy passes (oracle/oracle.py, what qibo's default backend
    does [recalled])."""
    from oracle import oracle as O
    from oracle import oracle_c as OC

    threads = OC.set_num_threads(0)
    rows = []
    for n in (20, 22, 24, 26):
        gates = O.expi_pauli_gates(f"X0 Z1 Z2 Y5 X9 Y{n - 1}", 0.1)  # one string of a double: H/S/CNOT ladder/RZ mix
        st = np.zeros(1 << n, dtype=np.complex128)
        st[1] = 1.0
        OC.run_gates(st, n, gates[:2])  # page the buffer in, warm the pool
        t0 = time.perf_counter()
        OC.run_gates(st, n, gates)
        t_c = (time.perf_counter() - t0) / len(gates)
        row = {"qubits": n, "c_openmp_s_per_pass": t_c, "c_openmp_GBps": 32.0 * 2.0**n / t_c / 1e9}
        if n <= 24 and gate_budget_s > 0:
            k = 8 if n <= 22 else 3
            t0 = time.perf_counter()
            O.run_gates(n, gates[:k], st)
            t_np = (time.perf_counter() - t0) / k
            row.update(numpy_1thread_s_per_pass=t_np, numpy_1thread_GBps=32.0 * 2.0**n / t_np / 1e9)
        rows.append(row)
        del st
    bw_c = float(np.median([r["c_openmp_GBps"] for r in rows[1:]]))  # n >= 22: beyond the last-level cache
    bw_np = float(np.median([r["numpy_1thread_GBps"] for r in rows if "numpy_1thread_GBps" in r and r["qubits"] >= 22]))
    return {"measured": rows, "threads": threads,
            "fitted_GBps": {"c_openmp_all_cores": bw_c, "numpy_single_thread": bw_np},
            "extrapolated_s_per_pass_30q": {"c_openmp_all_cores": 32.0 * 2.0**30 / bw_c / 1e9, "numpy_single_thread": 32.0 * 2.0**30 / bw_np / 1e9},
            "note": "measured and extrapolated values are in separate keys; time(30 q evaluation) = gate passes x 32 B x 2^30 / fitted bytes/s"}  # fmt: skip


def reference_gates_per_rotation(wl: dict) -> float:
    """Gate count of the reference's `_expi_pauli` program per string (2(nX+2nY) + 2(w-1) + 1, _ansatz.py:73-91)."""
    tot = 0
    for x, z in zip(wl["rx"].tolist(), wl["rz"].tolist()):
        nx, ny, w = bin(x & ~z).count("1"), bin(x & z).count("1"), bin(x | z).count("1")
        tot += 2 * (nx + 2 * ny) + 2 * (w - 1) + 1
    return tot / max(1, wl["rx"].size)


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, n, nelec, out):
    """--impl reference: the reference's own CPU algorithm on the host cores (oracle port; qibo is not installable)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(n, nelec, args.rotations, args.terms)
    vals, last = [], None
    for _ in range(args.warmup):
        cpu_reference_sample(wl, min(args.cpu_sample_qubits, 22), budget_s=2.0)
    for _ in range(args.steps):
        last = cpu_reference_sample(wl, args.cpu_sample_qubits, budget_s=8.0)
        vals.append(last["value"])
    v = float(np.mean(vals)) * float(2 ** (n - 30))
    extra = {k: last[k] for k in ("extrapolated", "sample_qubits", "scale", "sample_seconds_measured", "host_logical_cores", "seconds_per_eval_extrapolated")}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": unit_for(n), "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / (v / 2 ** (n - 30)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": config_dict(wl, args),
        "cpu_baseline": {**{k: last[k] for k in ("unit", "cores", "kind", "sample")}, "value": v, **extra},
        "e2e": {"value": v, "unit": unit_for(n), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        # value / ms_per_step are EXTRAPOLATED from a measured sample at `sample_qubits` (x `scale` in 2^n and the gate
        # census of the full stream); each timed step really ran `sample_seconds_measured` of CPU work
        "extrapolated": True, "sample_qubits": last["sample_qubits"], "scale": last["scale"],
        "measured_seconds_per_step": last["sample_seconds_measured"],
    }  # fmt: skip
    out.emit(line)


def config_dict(wl, args):
    return {
        "workload": f"synthetic {wl['n']}-qubit UCCSD-like energy: HF({wl['nelec']}e) + {wl['rx'].size} Pauli rotations "
        f"({wl['rx'].size // 8} JW doubles, reference order) + {wl['hx'].size + 1}-term molecular-structured Hamiltonian "
        f"({wl['n_xmasks']} x-masks), complex128",
        "qubits": wl["n"], "rotations": int(wl["rx"].size), "terms": int(wl["hx"].size + 1), "x_masks": wl["n_xmasks"],
        "l2": "state (>= 17 GB per GPU) exceeds the 126 MB L2; no flush needed between steps",
        "sharding": "index bits" if args.gpus > 1 else "none",
    }  # fmt: skip


class JsonStdout:
    """The driver reads ONE JSON line from stdout: everything else a library prints there (NCCL's version banner,
    ...) is sent to stderr by pointing fd 1 at fd 2 until the line is written."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: dict) -> None:
        sys.stdout.flush()
        os.write(self.saved, (json.dumps(line) + "\n").encode())


def main():
    args = parse_args()
    n = args.qubits if args.qubits is not None else DEFAULT_QUBITS.get(args.gpus, 30)
    nelec = args.electrons if args.electrons is not None else {30: 14}.get(n, 2 * ((n * 14 // 30) // 2))
    out = JsonStdout()
    if args.impl == "reference":
        run_reference(args, n, nelec, out)
        return
    if args.gpus > 1:
        from bench_sharded import main_sharded

        main_sharded(args, n, nelec, out)
        return

    import torch

    from qibochem_b200.engine import DeviceState

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(0)
    wl = workload(n, nelec, args.rotations, args.terms)
    stream = torch.cuda.Stream()
    st = DeviceState(n, device=0, stream=stream.cuda_stream)

    def step():
        st.set_basis_state(wl["hf"])
        st.apply_pauli_rotations(wl["rx"], wl["rz"], wl["ra"])
        return wl["const"] + st.expectation(wl["hx"], wl["hz"], wl["hc"])

    energies = [step() for _ in range(args.warmup)]
    # ---- timed region: K steps, device time from CUDA events on the launching stream
    st.profile_reset()
    st.profile_enable(True)
    clocks = ClockSampler(0)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        energies.append(step())
    ev1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clock_info = clocks.stop()
    dev_ms = ev0.elapsed_time(ev1)
    prof = st.profile()
    st.profile_enable(False)
    passes, sweeps = st.last_passes, st.last_sweeps
    ms_per_step = dev_ms / args.steps
    value = 1e3 / ms_per_step

    # ---- e2e: the same step through the public Python API (Circuit / SymbolicHamiltonian objects, host arrays in)
    e2e = measure_e2e(wl, args, st)

    # ---- roofline of the dominant kernel (by device time inside the timed region)
    roofline = build_roofline(wl, prof, args, dev_ms)
    launches = int(sum(v["launches"] for v in prof.values()))

    def extra(fn, *a):
        """The keys below are context for the headline: a failure in one of them (say, no room for the 32-qubit state)
        is recorded in its key and must not cost the line itself."""
        try:
            return fn(*a)
        except Exception as exc:  # noqa: BLE001
            return {"error": f"{type(exc).__name__}: {exc}"}

    full = None if (args.no_full_stream or args.rotations != 1024) else extra(measure_full_stream, wl, st, stream)
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_reference_sample(wl, args.cpu_sample_qubits)
        cpu["fit"] = extra(cpu_scaling_points)
    real_path = extra(measure_real_path, wl, st, stream) if n == 30 else None
    gradient = None if args.no_gradient else extra(measure_gradient, wl, st, stream, ms_per_step)
    small = None if args.no_small_configs else extra(measure_small_configs)
    ladder = None
    if not args.no_ladder_point and n == 30 and args.rotations == 1024 and torch.cuda.get_device_properties(0).total_memory > 100e9:
        st.close()
        ladder = extra(measure_ladder_point, 32, args, stream)
    line = {
        "metric": METRIC, "value": value * 2.0 ** (n - 30), "unit": unit_for(n), "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (complex128)", "data": "synthetic", "config": config_dict(wl, args),
        "energy": energies[-1], "energy_spread": float(np.ptp(energies)),
        "passes_per_step": {"k1_rotation_passes": passes, "k2_xmask_sweeps": sweeps},
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "full_uccsd_stream": full, "real_amplitude_fast_path": real_path, "gradient_30q" if n == 30 else "gradient": gradient,
        "small_configs": small, "weak_ladder_32q_1gpu": ladder,
    }  # fmt: skip
    out.emit(line)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(path)) if os.path.exists(path) else None


# FP64 issue peak of a B200 SM: 63.4 DFMA/clk/SM, measured on this pool's B200 by tools/ubench/fp64_pipe.cu
# (profiles/r1_fp64_pipe_b200.txt).  MEASURED_PEAKS.json carries HBM and bf16 figures only -- it has NO FP64 entry.
FP64_DFMA_PER_CLK_SM = 63.4
# dram__bytes_read.sum + dram__bytes_write.sum per K2T launch (30 q): CONSTANT taken from the committed ncu --set full
# capture of this command, not a live counter of the run that prints it (see "traffic_source")
K2T_TRAFFIC_30Q = {"bytes": 17.187e9, "source": "ncu --set full capture, profiles/r2_ncu_full_k2t_30q.csv (r1: profiles/r1_ncu_full_k2t_30q_v13.csv): 17.18-17.19 GB read + ~5 MB written per pass"}


def build_roofline(wl, prof, args, dev_ms):
    """Names the resource that binds the dominant kernel.  K2T (tiled expectation): FP64 issue -- one HBM read of the
    shard serves ~80 x-masks, so the per-x-mask algorithmic bytes of SURVEY 8d are NOT a traffic floor for it; they are
    kept as `algorithmic_hbm_equiv`.  Every other kernel class: HBM."""
    peaks = load_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max((k for k in prof if prof[k]["launches"]), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    secs = d["ms"] / 1e3
    algo_gbps = d["bytes"] / 1e9 / secs if secs else 0.0
    b_n = 16.0 * 2.0 ** wl["n"]
    hbm_actual = {"k2_tiled": b_n, "k1_rotation": 2 * b_n}.get(dom)  # bytes a launch really moves
    hbm_actual_gbps = (hbm_actual * d["launches"] / 1e9 / secs) if (hbm_actual and secs) else None
    common = {
        "kernel": dom, "launches": d["launches"], "avg_launch_ms": d["ms"] / max(d["launches"], 1),
        "share_of_step": d["ms"] / dev_ms,
        "by_kernel": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                          "algorithmic_GBps": (v["bytes"] / 1e9 / (v["ms"] / 1e3)) if v["ms"] else None}
                      for k, v in prof.items() if v["launches"]},
    }  # fmt: skip
    fp64 = k2_fp64_roofline(wl, prof, args) if dom == "k2_tiled" else None
    if fp64 is not None:
        tf = fp64["achieved"] * 2.0 / 1e12  # DFMA-rate: 2 flop per FP64 instruction
        tf_peak = fp64["peak"] * 2.0 / 1e12
        traffic = K2T_TRAFFIC_30Q if wl["n"] == 30 else {"bytes": None, "source": None}
        return {
            **common, "bound": "fp64_issue", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak,
            "peak_source": fp64["peak_source"] + "; MEASURED_PEAKS.json has no FP64 entry (HBM and bf16 only)",
            "note": "FP64 instructions per evaluation counted from the plan (52 per DOT record and thread, 3 per LIN/HOP entry) x 2 flop / "
                    "device time of the K2T launches, against the measured DFMA issue rate",
            "traffic": traffic["bytes"], "traffic_source": traffic["source"],
            "hbm_actual": {"GBps": hbm_actual_gbps, "peak": hbm_peak, "frac": hbm_actual_gbps / hbm_peak if hbm_actual_gbps else None,
                           "bytes_per_launch": hbm_actual, "peak_source": hbm_src,
                           "note": "one pass reads the shard once; this is what crosses the memory bus"},
            "algorithmic_hbm_equiv": {"GBps": algo_gbps, "bytes_per_launch": d["bytes"] / max(d["launches"], 1),
                                      "note": "SURVEY 8d figure (16 B x 2^n per distinct x-mask) / device time: what one-sweep-per-x-mask "
                                              "would have to move; exceeds the HBM peak because a tile serves many x-masks -- not a bound"},
            "fp64": fp64,
        }  # fmt: skip
    return {
        **common, "bound": "hbm", "achieved": algo_gbps, "peak": hbm_peak, "unit": "GB/s", "frac": algo_gbps / hbm_peak,
        "peak_source": hbm_src, "note": "achieved = algorithmic bytes / device time", "traffic": None,
        "hbm_actual": {"GBps": hbm_actual_gbps}, "algorithmic_bytes_per_launch": d["bytes"] / max(d["launches"], 1),
    }  # fmt: skip


def measure_real_path(wl, st, stream) -> dict:
    """EXTRA key, never the headline (which stays complex128 arithmetic): the benchmark state -- a HF determinant evolved
    by UCC rotations -- is exactly real, and with option k2_real the tiled expectation runs real-amplitude kernels
    (one multiplication per pair product, 64 instead of 128 registers per sub-cube, three CTAs per SM).  One warm-up +
    one timed step; the realness check (one read of the shard) is inside the timed region."""
    import torch

    st.set_option("k2_real", 1)
    ms = e = taken = None
    for _ in range(2):
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ev0.record(stream)
        st.set_basis_state(wl["hf"])
        st.apply_pauli_rotations(wl["rx"], wl["rz"], wl["ra"])
        ev1.record(stream)
        e = wl["const"] + st.expectation(wl["hx"], wl["hz"], wl["hc"])
        ev2.record(stream)
        torch.cuda.synchronize()
        ms, ms_k2 = ev0.elapsed_time(ev2), ev1.elapsed_time(ev2)
        taken = st.took_real_path()
    st.set_option("k2_real", 0)
    return {"taken": bool(taken), "ms_per_step": ms, "expectation_ms": ms_k2, "evals_per_sec": 1e3 / ms, "energy": e,
            "note": "same workload as the headline; identical energy to 1e-12; not the headline because the arithmetic is real, not complex128"}  # fmt: skip


def measure_gradient(wl, st, stream, energy_ms: float) -> dict:
    """Adjoint gradient of the benchmark energy with respect to ALL rotation angles (reference caller:
    qibo.derivative.parameter_shift, doc/source/tutorials/adaptive.rst:46-61, two energy evaluations per parameter):
    forward execution, lambda = H psi through tiles (K6T), backward walk over the fused runs.  ONE timed call (the plan of
    the tiled H|psi> is built inside it, ~0.5 s of host time)."""
    import torch

    st.set_basis_state(wl["hf"])
    st.apply_pauli_rotations(wl["rx"], wl["rz"], wl["ra"])
    st.profile_reset()
    st.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    e, grad = st.energy_gradient(wl["rx"], wl["rz"], wl["ra"], wl["hx"], wl["hz"], wl["hc"])
    ev1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    prof = st.profile()
    st.profile_enable(False)
    ms = ev0.elapsed_time(ev1)
    return {"parameters": int(grad.size), "ms": ms, "wall_ms": 1e3 * wall, "in_units_of_one_energy": ms / energy_ms,
            "parameter_shift_would_need_energies": 2 * int(grad.size), "energy": wl["const"] + e,
            "max_abs_gradient": float(np.abs(grad).max()),
            "device_ms": {k: round(v["ms"], 1) for k, v in prof.items() if v["launches"]}}  # fmt: skip


def measure_ladder_point(n: int, args, stream) -> dict:
    """SURVEY 7's weak-scaling ladder is 32 q @ 1, 33 q @ 2, 34 q @ 4, 35 q @ 8 (2^32 amplitudes per GPU): the 1-GPU
    point, with the SAME generators as the N > 1 arms (the n-qubit Hamiltonian has more hopping strings per x-mask than
    the 30-qubit one, so per-x-mask costs across the ladder compare like with like).  One warm-up + one timed step."""
    import torch

    from qibochem_b200.engine import DeviceState

    nelec = 2 * ((n * 14 // 30) // 2)
    wl = workload(n, nelec, args.rotations, args.terms)
    with DeviceState(n, device=0, stream=stream.cuda_stream) as st:
        ms = e = None
        for _ in range(2):
            st.profile_reset()
            st.profile_enable(True)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            st.set_basis_state(wl["hf"])
            st.apply_pauli_rotations(wl["rx"], wl["rz"], wl["ra"])
            e = wl["const"] + st.expectation(wl["hx"], wl["hz"], wl["hc"])
            ev1.record(stream)
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            prof = st.profile()
        passes = st.last_sweeps
    k2 = prof["k2_tiled"]["ms"]
    return {"qubits": n, "ms_per_step": ms, "thirty_qubit_equivalent_evals_per_sec": 2.0 ** (n - 30) * 1e3 / ms, "energy": e,
            "x_masks": wl["n_xmasks"], "k2_tiled_ms": k2, "k2_passes": int(passes),
            "k2_ms_per_xmask_per_2^32_amplitudes": k2 / wl["n_xmasks"] * 2.0 ** (32 - n), "k1_ms": prof["k1_rotation"]["ms"]}  # fmt: skip


def measure_small_configs() -> dict:
    """BASELINE configs 1-3 (the launch-bound regime) through the public API, bounded to a few seconds:
    H2 (4 q, doc fixture) and LiH-shaped (12 q) exact energies per `set_parameters` + `hamiltonian.expectation`, and the
    BeH2-shaped (14 q) QWC-grouped sampled energy at 10^6 shots per group (`expectation_from_samples`).  LiH / BeH2 use
    molecule-SHAPED synthetic integrals (no PySCF in the image): sizes and term structure are the molecule's."""
    from qibochem_b200.ansatz import circuit_ansatz
    from qibochem_b200.measurement import expectation_from_samples
    from qibochem_b200.measurement.optimization import _measurement_basis_rotations
    from qibochem_b200.runtime import get_runtime
    from tests.fixtures import h2_molecule, synthetic_molecule

    out = {}
    for key, mol in (("h2_4q", h2_molecule()), ("lih_shaped_12q", synthetic_molecule(6, 4, 1))):
        ham = mol.hamiltonian()
        circuit = circuit_ansatz(mol, "ucc")
        params = np.array([complex(p[0]).real for p in circuit.get_parameters()], dtype=float)
        e = ham.expectation(circuit)  # warm-up: compiles the program
        reps = 100
        t0 = time.perf_counter()
        for k in range(reps):
            circuit.set_parameters(params * (1.0 + 1e-3 * k))
            e = ham.expectation(circuit)
        dt = (time.perf_counter() - t0) / reps
        st = get_runtime().state(mol.nso)
        st.profile_reset()
        st.profile_enable(True)
        circuit.set_parameters(params)
        e = ham.expectation(circuit)
        prof = st.profile()
        st.profile_enable(False)
        out[key] = {"qubits": mol.nso, "rotations": int(params.size), "terms": int(ham.nterms), "energy": e,
                    "us_per_energy_e2e": 1e6 * dt, "energies_per_sec": 1.0 / dt,
                    "device_us": {k: round(1e3 * v["ms"], 2) for k, v in prof.items() if v["launches"]},
                    "launches_per_energy": int(sum(v["launches"] for v in prof.values()))}  # fmt: skip
    mol = synthetic_molecule(7, 6, 3)
    ham = mol.hamiltonian()
    circuit = circuit_ansatz(mol, "ucc", thetas=np.random.default_rng(4).uniform(-0.1, 0.1, size=204))
    t0 = time.perf_counter()
    groups = _measurement_basis_rotations(ham, "qwc")
    t_group = time.perf_counter() - t0
    exact = ham.expectation(circuit)
    shots = 1_000_000
    st = get_runtime().state(mol.nso)
    e = dt = None
    for rep in range(2):
        st.profile_reset()
        st.profile_enable(True)
        t0 = time.perf_counter()
        e = expectation_from_samples(circuit, ham, n_shots=shots, grouping="qwc", seed=5 + rep)
        dt = time.perf_counter() - t0
    prof = st.profile()
    st.profile_enable(False)
    n_masks = int(ham.nterms)
    out["beh2_shaped_14q_sampled"] = {
        "qubits": mol.nso, "rotations": len(circuit.get_parameters()), "terms": n_masks, "qwc_groups": len(groups),
        "shots_per_group": shots, "s_per_sampled_energy": dt, "Gshots_per_sec": len(groups) * shots / dt / 1e9,
        "exact": exact, "sampled": e, "abs_diff": abs(e - exact), "grouping_s": t_group,
        "device_ms": {k: round(v["ms"], 2) for k, v in prof.items() if v["launches"]},
        # rooflines of the two sampling-path kernels (per group: one K5 draw of `shots` shots, one K3 pass over them)
        "k3_parity_sums": {"algorithmic_GBps": (prof["k3_shots"]["bytes"] / 1e9 / (prof["k3_shots"]["ms"] / 1e3)) if prof["k3_shots"]["ms"] else None,
                           "note": "8 B per shot and mask tile (uint64 samples), integer ALU / L2 bound: the samples of a group (8 MB) stay in L2"},
        "k5_sampling": {"ms": prof["k5_sampling"]["ms"], "shots_per_sec": len(groups) * shots / (prof["k5_sampling"]["ms"] / 1e3) if prof["k5_sampling"]["ms"] else None,
                        "note": "one warp per shot: Philox + binary search over chunk prefix sums + scan inside a 128-amplitude chunk; latency bound (L2)"},
    }  # fmt: skip
    get_runtime().release_states()
    return out


def measure_full_stream(wl, st, stream):
    """Context for the headline number: the same evaluation with the COMPLETE UCCSD rotation stream of the register
    (all doubles, then all singles, reference order) instead of the 1024-rotation prefix; one warm-up + one timed step."""
    import torch

    from qibochem_b200 import synthetic

    rx, rz, ra, n_exc = synthetic.uccsd_rotation_stream(wl["n"], wl["nelec"])
    ms = None
    for _ in range(2):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        st.set_basis_state(wl["hf"])
        passes = st.apply_pauli_rotations(rx, rz, ra)
        e = wl["const"] + st.expectation(wl["hx"], wl["hz"], wl["hc"])
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
    return {"rotations": int(rx.size), "excitations": int(n_exc), "k1_passes": int(passes), "ms_per_step": ms,
            "evals_per_sec": 1e3 / ms, "energy": e}


def k2_fp64_roofline(wl, prof, args):
    """What really bounds the tiled expectation: FP64 issue.  FP64 instructions per pass are counted from the plan
    (DOT record: 48 DMUL/DFMA + 4 DADD per thread; LIN / HOP entry: counted as 3, the HOP figure, and a HOP head as
    one DOT record -- both on the low side) and compared with the DFMA rate measured on this
    pool's B200 by tools/ubench/fp64_pipe.cu (profiles/r1_fp64_pipe_b200.txt: 63.4 DFMA/clk/SM = 34.9 TFLOP/s)."""
    import ctypes as C

    from qibochem_b200 import _native

    lib = _native.load()
    stats = np.zeros(16, dtype=np.uint64)
    u64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    hx, hz, hc = (np.ascontiguousarray(wl[k]) for k in ("hx", "hz", "hc"))
    rc = lib.qc_k2_plan_stats(wl["n"], hx.size, hx.ctypes.data_as(u64p), hz.ctypes.data_as(u64p), hc.ctypes.data_as(f64p),
                              stats.ctypes.data_as(u64p), None)  # fmt: skip
    if rc != 0 or not prof["k2_tiled"]["ms"]:
        return None
    n_pass, n_cubes, n_dot, n_lin = (int(stats[k]) for k in (2, 3, 4, 5))
    threads = 2.0 ** wl["n"] / 32.0  # one thread per 32-amplitude sub-cube
    fp64_instr = (52.0 * n_dot + 3.0 * n_lin) * threads  # per energy evaluation
    seconds = prof["k2_tiled"]["ms"] / 1e3 / args.steps
    peak_instr = FP64_DFMA_PER_CLK_SM * 148 * 1.965e9
    return {"unit": "FP64 instr/s (DFMA-rate)", "achieved": fp64_instr / seconds, "peak": peak_instr, "frac": fp64_instr / seconds / peak_instr,
            "peak_source": "tools/ubench/fp64_pipe.cu on this pool's B200: 63.4 DFMA/clk/SM x 148 SMs x 1.965 GHz (34.9 TFLOP/s)",
            "operand_bound": {"clk_per_record_and_warp": 146.0, "frac": fp64_instr / 52.0 / 32.0 * 146.0 / (seconds * 148 * 4 * 1.965e9),
                              "note": "three-register-operand DFMAs take 2.85-3.05 clk per SM sub-partition on B200 and a weight fed by a "
                                      "broadcast LDS.128 ~4 clk (profiles/r1_fp64_operands_b200.txt): a DOT record costs 142-150 clk per "
                                      "warp at best; frac = warp-level record-equivalents x 146 clk / (592 SM sub-partitions x clock x time)"},
            "plan": {"hbm_passes": n_pass, "sub_cube_gathers": n_cubes, "dot_records": n_dot, "lin_entries": n_lin},
            "shared_memory_gather_bytes_per_eval": n_cubes * 16.0 * 2.0 ** wl["n"]}  # fmt: skip


def measure_e2e(wl, args, st):
    """Energy through the public API: a `Circuit` of PauliRotation gates + `SymbolicHamiltonian.expectation`; every step
    receives a fresh host parameter vector (`set_parameters`), ships the program to the device and reads the energy
    back."""
    from qibochem_b200 import Circuit, SymbolicHamiltonian, gates
    from qibochem_b200.pauli import masks_to_string, reverse_bits
    from qibochem_b200.runtime import get_runtime

    n = wl["n"]
    rt = get_runtime()
    rt.adopt(st)  # share the already allocated state with the public API's runtime
    circuit = Circuit(n)
    circuit.add(gates.X(q) for q in range(wl["nelec"]))
    for x, z, a in zip(wl["rx"].tolist(), wl["rz"].tolist(), wl["ra"].tolist()):
        circuit.add(gates.PauliRotation(masks_to_string(reverse_bits(x, n), reverse_bits(z, n)), a))
    ham = SymbolicHamiltonian.from_index_masks(n, wl["hx"], wl["hz"], wl["hc"], constant=wl["const"])
    rng = np.random.default_rng(0)
    params = [wl["ra"] * (1.0 + 1e-3 * rng.normal()) for _ in range(args.steps + 1)]
    circuit.set_parameters(params[0])
    ham.expectation(circuit)  # warm-up (compiles the program)
    st.transfer_bytes(reset=True)
    t0 = time.perf_counter()
    for k in range(args.steps):
        circuit.set_parameters(params[k + 1])
        ham.expectation(circuit)
    dt = (time.perf_counter() - t0) / args.steps
    rt.forget(st)
    h2d, d2h = (v // args.steps for v in st.transfer_bytes())  # counted by the library at every copy it issues
    return {"value": 2.0 ** (n - 30) / dt, "unit": unit_for(n), "ms_per_step": 1e3 * dt, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "note": "the Hamiltonian's tiled plan (3.4 MB) is uploaded once and cached on the device by content hash; per step "
                    "only the rotation tables, the per-launch kernel parameters and the energy cross the PCIe bus",
            "api": "Circuit.set_parameters(host array) + SymbolicHamiltonian.expectation(circuit) -> float"}  # fmt: skip


if __name__ == "__main__":
    main()
