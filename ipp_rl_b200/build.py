"""Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libipp_b200.so")
SOURCES = ["ipp_engine.cu", "mcts.cu", "grf.cu", "observe.cu", "experience.cu", "kalman_blocks.cu", "fields.cu"]
HEADERS = ["step_kernel.cuh", "step_async.cuh", "step_bulk.cuh", "quad_math.cuh", "rollout_kernel.cuh", "engine_internal.h", os.path.join("..", "..", "include", "ipp_mcts.h"), os.path.join("..", "..", "include", "ipp_b200.h"),
           os.path.join("..", "..", "include", "ipp_experience.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-ldl",
    "-diag-suppress", "128",  # "loop is not reachable" in constant-folded template instantiations
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build the sm_100a extension")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))
