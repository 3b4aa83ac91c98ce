"""CPU-only: the C-ABI shared library builds for sm_100a, loads, and exports every symbol the header declares."""
import ctypes
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_declared_symbols():
    from scisim_b200 import _lib, build
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _lib.declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert missing == []
    # every bound signature corresponds to a declared symbol and vice versa
    bound = _lib.load()
    for n in names:
        assert getattr(bound, n).argtypes is not None, n


def test_library_carries_sm100a_code_only():
    import subprocess
    from scisim_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build_library()], stdout=subprocess.PIPE, text=True).stdout
    archs = set(l.split(".")[-2] for l in out.splitlines() if l.strip().endswith(".cubin"))
    assert archs == {"sm_100a"}, out


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import scisim_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(scisim_b200.SciSimB200Error):
        scisim_b200.Context(0)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing in the package or the library sources may reference it."""
    pkg = os.path.join(ROOT, "scisim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "oracle_binding" not in txt and "oracle/" not in txt, os.path.join(dirpath, f)
