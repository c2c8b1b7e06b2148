"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line with the agreed keys."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
         "--cpu-sample-qubits", "16", "--terms", "20000", "--rotations", "64"],
        capture_output=True, text=True, timeout=600, cwd=ROOT,
    )  # fmt: skip
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "uccsd_energy_evals_per_sec" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["qubits"] == 30
    # the value is an extrapolation of a measured sample and says so in fields, not in prose
    assert d["extrapolated"] is True and d["sample_qubits"] == 16 and d["scale"] == 2.0 ** (30 - 16)
    assert 0 < d["measured_seconds_per_step"] < 60
    assert d["cpu_baseline"]["extrapolated"] is True and d["cpu_baseline"]["sample_qubits"] == 16


def test_reference_arm_sets_its_own_thread_count():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU baseline must not inherit it (round 1 reported
    `cores: 1` at N > 1 and inflated the GPU / CPU ratio 16-32 x)."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    proc = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
         "--cpu-sample-qubits", "14", "--terms", "20000", "--rotations", "64"],
        capture_output=True, text=True, timeout=600, cwd=ROOT, env=env,
    )  # fmt: skip
    assert proc.returncode == 0, proc.stderr[-2000:]
    d = json.loads([ln for ln in proc.stdout.splitlines() if ln.strip()][0])
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert d["cpu_baseline"]["cores"] == usable  # every core this process may run on, not the inherited 1


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
        capture_output=True, text=True, timeout=300, cwd=ROOT, env=env,
    )  # fmt: skip
    assert proc.returncode == 0 and proc.stdout.strip() == ""
