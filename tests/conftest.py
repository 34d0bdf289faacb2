import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "replay_path: GPU test that goes through lobsim_replay (kept for selection with -m)")
    config.addinivalue_line("markers", "fast_only: GPU test that is about the straight-line kernels only (not repeated "
                                       "on the general kernel family)")


def pytest_generate_tests(metafunc):
    """Every GPU test runs on three independent implementations of the order semantics: "fast" = the straight-line static-layout
    kernels with every book that fits in the flat order pools of book_flat.cuh (in shared memory and in HBM); "sorted" = the same
    kernels on the sorted level arrays only (LOBSIM_REPLAY_FLAT=0, LOBSIM_FLAT_BLOBS=0); "general" = the runtime-layout kernel
    (LOBSIM_FORCE_GENERAL=1).  Tests marked `replay_path` run a fourth time, "hybrid": the hot-pool / cold-level-array book of
    book_hybrid.cuh forced on for every layout that can hold it (LOBSIM_REPLAY_HYBRID=1; by default only the NO >= 1024 layouts)."""
    if "kernel_family" in metafunc.fixturenames and metafunc.definition.get_closest_marker("gpu"):
        fams = ["fast"] if metafunc.definition.get_closest_marker("fast_only") else ["fast", "sorted", "general"]
        if metafunc.definition.get_closest_marker("replay_path") and len(fams) > 1:
            fams.append("hybrid")
        metafunc.parametrize("kernel_family", fams, indirect=True)


@pytest.fixture(autouse=True)
def kernel_family(request, monkeypatch):
    fam = getattr(request, "param", None)
    if fam is not None:
        monkeypatch.setenv("LOBSIM_FORCE_GENERAL", "1" if fam == "general" else "0")
        monkeypatch.setenv("LOBSIM_REPLAY_FLAT", "0" if fam == "sorted" else "1")
        monkeypatch.setenv("LOBSIM_FLAT_BLOBS", "0" if fam == "sorted" else "1")
        if fam == "hybrid":
            monkeypatch.setenv("LOBSIM_REPLAY_HYBRID", "1")
        else:
            monkeypatch.delenv("LOBSIM_REPLAY_HYBRID", raising=False)
    return fam


@pytest.fixture(scope="session")
def fixture_stream():
    from parity_helpers import load_fixture_stream

    return load_fixture_stream("reference")
