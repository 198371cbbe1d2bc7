"""Build libqcsim_b200.so (sm_100a only) in-tree with nvcc.

The built library lives next to this file so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libqcsim_b200.so")

SOURCES = ["api.cu", "engine.cu", "fusion.cu", "dist.cu", "multi.cu", "qft_pipe.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newest_source_mtime() -> float:
    m = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def needs_build() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest_source_mtime()


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra_flags: list[str] | None = None) -> str:
    """Compile the CUDA extension if it is missing or stale; returns the library path.
    `out` / `extra_flags` build an experimental variant (e.g. -DQCSIM_PIPE_GROUPS=3) next to the shipped library."""
    if out is not None:
        return _compile(out, verbose, extra_flags or [])
    if not force and not needs_build():
        return LIB
    return _compile(LIB, verbose, [])


def _compile(LIB: str, verbose: bool, extra_flags: list[str]) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libqcsim_b200.so (there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, *extra_flags]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-o", LIB + ".tmp", "-lnccl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
