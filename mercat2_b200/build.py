"""Build libmercat2_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmercat2_b200.so"
INCLUDE = PKG.parent / "include"

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inl")) + [INCLUDE / "mercat2_b200.h"])


def needs_build() -> bool:
    if not LIB.exists():
        return True
    built = LIB.stat().st_mtime
    return any(src.stat().st_mtime > built for src in sources())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmercat2_b200.so")
    cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-o", str(LIB), str(CSRC / "mc2.cu"), "-lz"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
