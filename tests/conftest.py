import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _reference_libraries():
    """Where the reference tree is mounted (the build container), make sure oracle/_ref/*.so and tests/_build/ are there and up to date before any test looks
    for them -- `make` rebuilds only what changed, so after __graft_entry__.build() this costs a fraction of a second.  Elsewhere (the GPU box) the prebuilt
    files that travelled with the snapshot are used as they are, and tests that need missing ones skip."""
    if os.path.isdir("/root/reference/ball2d"):
        import subprocess
        from scisim_b200 import build
        build.build_library()          # the plugin example links against the product library
        res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert res.returncode == 0, "oracle/Makefile.ref failed:\n" + res.stdout[-4000:]
    yield


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_binding
    return oracle_binding.load()


@pytest.fixture(scope="session")
def gpu_ctx():
    import scisim_b200
    ctx = scisim_b200.Context(0)
    yield ctx
    ctx.close()
