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


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libscisim_b200.so:\n" + res.stdout)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
