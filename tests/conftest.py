import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "replay_path: GPU test that goes through lobsim_replay -- also run with "
                                       "LOBSIM_REPLAY_FLAT=0 (every book on the sorted level arrays of k_replay_fast instead "
                                       "of the flat order pools of k_replay_flat)")
    config.addinivalue_line("markers", "fast_only: GPU test that is about the straight-line kernels only (not repeated "
                                       "on the general kernel family)")


def pytest_generate_tests(metafunc):
    """Every GPU test runs twice: on the straight-line static-layout kernels ("fast") and, with LOBSIM_FORCE_GENERAL=1,
    on the general runtime-layout kernel ("general") -- two independent implementations of the same semantics.  Tests
    marked `replay_path` run a third time on the sorted-array replay kernel ("sorted": LOBSIM_REPLAY_FLAT=0), since
    "fast" replays on the flat order pools (book_flat.cuh) -- a third implementation of the order semantics."""
    if "kernel_family" in metafunc.fixturenames and metafunc.definition.get_closest_marker("gpu"):
        fams = ["fast"] if metafunc.definition.get_closest_marker("fast_only") else ["fast", "general"]
        if metafunc.definition.get_closest_marker("replay_path"):
            fams.insert(1, "sorted")
        metafunc.parametrize("kernel_family", fams, indirect=True)


@pytest.fixture(autouse=True)
def kernel_family(request, monkeypatch):
    fam = getattr(request, "param", None)
    if fam is not None:
        monkeypatch.setenv("LOBSIM_FORCE_GENERAL", "1" if fam == "general" else "0")
        monkeypatch.setenv("LOBSIM_REPLAY_FLAT", "0" if fam == "sorted" else "1")
    return fam


@pytest.fixture(scope="session")
def fixture_stream():
    from parity_helpers import load_fixture_stream

    return load_fixture_stream("reference")
