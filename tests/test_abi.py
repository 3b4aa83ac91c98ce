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


def _header_prototypes():
    """name -> ( return kind, [ parameter kinds ] ) parsed from include/scisim_b200.h; kinds: 'ptr', 'int', 'f64'."""
    import re
    hdr = open(os.path.join(ROOT, "include", "scisim_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    hdr = re.sub(r"#[^\n]*", " ", hdr)
    protos = {}

    def kind(decl):
        decl = decl.strip()
        if "*" in decl:
            return "ptr"
        if re.match(r"^(const\s+)?double\b", decl):
            return "f64"
        if re.match(r"^(const\s+)?(unsigned\s+)?(int|long|u?int\d+_t|size_t)\b", decl):
            return "int"
        if decl == "void":
            return "void"
        raise AssertionError("unrecognised declaration: %r" % decl)

    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(sg_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", hdr):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        ps = [p for p in (q.strip() for q in params.split(",")) if p and p != "void"]
        protos[name] = (kind(ret), [kind(p) for p in ps])
    return protos


def test_python_binding_signatures_match_the_header():
    """Every ctypes signature in scisim_b200/_lib.py has the parameter count and parameter kinds ( pointer / integer / double ) of its
    prototype in include/scisim_b200.h; integer widths agree where the header uses fixed-width types."""
    from scisim_b200 import _lib
    protos = _header_prototypes()
    assert sorted(protos) == _lib.declared_symbols()
    bound = _lib.load()

    def ckind(t):
        if t is None:
            return "void"
        if t in (ctypes.c_double,):
            return "f64"
        if t in (ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint):
            return "int"
        return "ptr"  # c_void_p, c_char_p, POINTER( ... )

    for name, (ret, params) in protos.items():
        fn = getattr(bound, name)
        assert ckind(fn.restype) == ret, name
        assert [ckind(t) for t in fn.argtypes] == params, name


def test_struct_layouts_match_the_header():
    """sg_contacts / sg_pairs / sg_teleported as ctypes structures have the size a C compiler gives the header's structs."""
    import subprocess
    import tempfile
    from scisim_b200 import _lib
    src = '#include "scisim_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main( void ) { printf( "%zu %zu %zu %zu %zu\\n", sizeof( sg_contacts ), sizeof( sg_pairs ), sizeof( sg_teleported ), ' \
          'offsetof( sg_teleported, delta1 ), offsetof( sg_teleported, portal0 ) ); return 0; }\n'
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "layout.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "layout")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, c], check=True)
        sizes = [int(v) for v in subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True).stdout.split()]
    assert sizes[0] == ctypes.sizeof(_lib.SgContacts)
    assert sizes[1] == ctypes.sizeof(_lib.SgPairs)
    assert sizes[2] == ctypes.sizeof(_lib.SgTeleported)
    assert sizes[3] == _lib.SgTeleported.delta1.offset and sizes[4] == _lib.SgTeleported.portal0.offset
