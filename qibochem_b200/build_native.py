"""Builds the native library in-tree with nvcc for sm_100a (the only target)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libqibochem_b200.so")
SOURCES = ["state.cu", "rotation.cu", "expectation.cu", "expectation_tiled.cu", "k2_plan.cpp", "gradient.cu", "gradient_tiled.cu", "grouping.cpp", "sampling.cu", "exchange.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp",
    "-Xptxas", "-v",
]  # fmt: skip


def extra_flags() -> list[str]:
    """Extra -D flags for kernel A/B experiments (tools/): QC_NVCC_DEFINES="QC_K2T_PACE=8 FOO=1"."""
    return [f"-D{d}" for d in os.environ.get("QC_NVCC_DEFINES", "").split()]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the native library cannot be built")
    return nvcc


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "qibochem_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    tmp = LIB + f".tmp{os.getpid()}"  # built aside and renamed: a reader never sees a half-written library
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra_flags(), "-shared", "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("nvcc failed building libqibochem_b200.so")
    os.replace(tmp, LIB)
    with open(os.path.join(HERE, "csrc", "ptxas_info.txt"), "w") as f:  # registers / spills / shared memory per kernel
        f.write("".join(line for line in proc.stderr.splitlines(keepends=True) if "Compile time" not in line))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
