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
    cmd = [find_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode:
        raise RuntimeError("nvcc failed building libqibochem_b200.so")
    with open(os.path.join(HERE, "csrc", "ptxas_info.txt"), "w") as f:  # registers / spills / shared memory per kernel
        f.write("".join(line for line in proc.stderr.splitlines(keepends=True) if "Compile time" not in line))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
