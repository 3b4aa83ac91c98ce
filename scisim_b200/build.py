"""Builds libscisim_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

`python -m scisim_b200.build` or `scisim_b200.build.build_library()`; __graft_entry__.build() calls the latter.
nvcc cross-compiles for sm_100a without a GPU.  -fmad=false: the reference is built without FP contraction
(ISO C++ mode, CMakeLists.txt:53-56), and every parity-relevant expression is FP64.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscisim_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # no FMA contraction: bit parity with the reference's FP64 arithmetic
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-diag-suppress", "549",
    "-shared",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "scisim_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, res.returncode, res.stdout


def build_library(force=False, verbose=False):
    """One object per .cu (compiled side by side, rebuilt only when older than a dependency), then one link."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if not f.endswith(".cu"))
    hdr_t = max(hdr_t, os.path.getmtime(os.path.join(HERE, "..", "include", "scisim_b200.h")), os.path.getmtime(__file__))
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(src)):
            jobs.append((nvcc, src, obj, verbose))
    log = ""
    with ThreadPoolExecutor(max_workers=max(1, min(8, len(jobs) or 1))) as ex:
        for src, rc, out in ex.map(_compile_one, jobs):
            log += out
            if rc != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log += res.stdout
    if verbose:
        sys.stderr.write(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed linking libscisim_b200.so:\n" + res.stdout)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
