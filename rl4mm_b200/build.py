"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

* ``rl4mm_b200/_native/liblobsim.so``  -- CUDA kernels + C ABI (include/lobsim.h), nvcc, sm_100a only.  One translation
  unit for the host side + the general kernel (csrc/lobsim.cu) and two per compiled book layout (csrc/fast_layout.cu x
  csrc/layouts.h), compiled in parallel and linked into one library.  The sha256 of the sources is stamped into the
  library (``lobsim_source_hash()``) and checked at load time by ``_lib.lib()``.
* ``rl4mm_b200/_native/libsynth.so``   -- synthetic LOBSTER stream generator (host C++)
* ``rl4mm_b200/_native/libingest.so``  -- LOBSTER CSV reader + packer (host C++)
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "_native"
OBJ = OUT / "obj"
INCLUDE = HERE.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _newer(srcs, out: Path) -> bool:
    return not out.exists() or any(Path(s).stat().st_mtime > out.stat().st_mtime for s in srcs)


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def lobsim_sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [INCLUDE / "lobsim.h"]


def source_hash() -> str:
    """sha256 over the names and contents of everything liblobsim.so is built from."""
    h = hashlib.sha256()
    for p in lobsim_sources():
        h.update(p.name.encode() + b"\0" + p.read_bytes() + b"\0")
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


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
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", str(out), str(srcs[0])], check=True)
    return out


def n_fast_layouts() -> int:
    m = re.search(r"#define\s+LOBSIM_N_FAST_LAYOUTS\s+(\d+)", (CSRC / "layouts.h").read_text())
    return int(m.group(1))


def build_lobsim(force: bool = False, verbose: bool = False, variant: str = None, extra_flags=()) -> Path:
    """`variant` + `extra_flags`: an experimental build (A/B of -D switches) into _native/variants/liblobsim_<variant>.so,
    loaded instead of the product library when LOBSIM_NATIVE_LIB points at it (tools/gpu_ab.sh)."""
    global OBJ
    OUT.mkdir(exist_ok=True)
    out, stamp = OUT / "liblobsim.so", OUT / "liblobsim.hash"
    want = source_hash()
    if variant:
        (OUT / "variants").mkdir(exist_ok=True)
        out, stamp = OUT / "variants" / f"liblobsim_{variant}.so", OUT / "variants" / f"liblobsim_{variant}.hash"
        OBJ = OUT / "variants" / f"obj_{variant}"
        want = hashlib.sha256((want + " ".join(extra_flags)).encode()).hexdigest()
    if not force and out.exists() and stamp.exists() and stamp.read_text().strip() == want:
        return out
    OBJ.mkdir(exist_ok=True)
    nvcc = nvcc_path()
    common = [nvcc, *NVCC_FLAGS, *extra_flags, "-I", str(INCLUDE), "-I", str(CSRC)]
    jobs = [("lobsim", common + [f'-DLOBSIM_SOURCE_HASH="{source_hash()}"', "-c", str(CSRC / "lobsim.cu")])]
    for i in range(n_fast_layouts()):
        for part in (0, 1):
            jobs.append((f"fast{i}_{part}", common + [f"-DLOBSIM_LAYOUT_INDEX={i}", f"-DLOBSIM_TU_PART={part}", "-c", str(CSRC / "fast_layout.cu")]))

    def run(job):
        name, cmd = job
        res = subprocess.run(cmd + ["-o", str(OBJ / f"{name}.o")], capture_output=True, text=True)
        return name, res

    logs, failed = [], False
    with ThreadPoolExecutor(max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for name, res in ex.map(run, jobs):
            logs.append(f"==== {name}\n{res.stderr}")
            if verbose or res.returncode:
                print(f"==== {name}\n{res.stdout}{res.stderr}")
            failed = failed or res.returncode != 0
    (OUT / "ptxas.log").write_text("\n".join(logs))
    if failed:
        raise RuntimeError("nvcc failed")
    objs = [str(OBJ / f"{name}.o") for name, _ in jobs]
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", str(out)] + objs,
                         capture_output=True, text=True)
    if res.returncode:
        print(res.stdout, res.stderr)
        raise RuntimeError("link failed")
    stamp.write_text(want + "\n")
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_synth(force)
    build_ingest(force)
    build_lobsim(force, verbose)


if __name__ == "__main__":
    import sys

    if "--variant" in sys.argv:      # python -m rl4mm_b200.build --variant NAME -DFOO=1 ...
        i = sys.argv.index("--variant")
        print(build_lobsim(force="--force" in sys.argv, variant=sys.argv[i + 1], extra_flags=[a for a in sys.argv[i + 2:] if a != "--force"]))
    else:
        build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
