"""Builds csrc/ into libbeluga_b200.so for sm_100a (in-tree, so the .so travels to the GPU box)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libbeluga_b200.so")
SRCS = [os.path.join(HERE, "csrc", "beluga_b200.cu")]
INC = os.path.join(os.path.dirname(HERE), "include")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
]


def nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(SO):
        return True
    deps = SRCS + [os.path.join(INC, "beluga_b200.h")] + [
        os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith((".cuh", ".h"))
    ]
    return any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    cmd = [nvcc()] + NVCC_FLAGS + ["-I", INC, "-o", SO] + SRCS
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
