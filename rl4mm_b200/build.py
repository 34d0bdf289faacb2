"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

* ``rl4mm_b200/_native/liblobsim.so``  -- CUDA kernels + C ABI (include/lobsim.h), nvcc, sm_100a only
* ``rl4mm_b200/_native/libsynth.so``   -- synthetic LOBSTER stream generator (host C++)
* ``rl4mm_b200/_native/libingest.so``  -- LOBSTER CSV reader for the packer (host C++)
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "_native"
INCLUDE = HERE.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _newer(srcs, out: Path) -> bool:
    return not out.exists() or any(Path(s).stat().st_mtime > out.stat().st_mtime for s in srcs)


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def build_synth(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    out = OUT / "libsynth.so"
    srcs = [CSRC / "synth_lobster.cpp", INCLUDE / "lobsim.h"]
    if force or _newer(srcs, out):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(out), str(srcs[0])], check=True)
    return out


def build_ingest(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    out = OUT / "libingest.so"
    srcs = [CSRC / "lobster_ingest.cpp"]
    if force or _newer(srcs, out):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(out), str(srcs[0])], check=True)
    return out


def build_lobsim(force: bool = False, verbose: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    out = OUT / "liblobsim.so"
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [INCLUDE / "lobsim.h"]
    if force or _newer(srcs, out):
        cmd = [nvcc_path(), *NVCC_FLAGS, "-I", str(INCLUDE), "-o", str(out)] + [str(s) for s in sorted(CSRC.glob("*.cu"))]
        res = subprocess.run(cmd, capture_output=True, text=True)
        (OUT / "ptxas.log").write_text(res.stderr)
        if verbose or res.returncode:
            print(res.stdout, res.stderr)
        if res.returncode:
            raise RuntimeError("nvcc failed")
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_synth(force)
    build_ingest(force)
    build_lobsim(force, verbose)


if __name__ == "__main__":
    import sys

    build_all(force="--force" in sys.argv, verbose=True)
