"""Builds libLICHeE_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "lichee_search.cu")
SO = os.path.join(HERE, "libLICHeE_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def build(force=False, verbose=False):
    deps = [SRC, os.path.join(HERE, "..", "include", "lichee_b200.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    cmd = [NVCC] + FLAGS + ["-o", SO, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libLICHeE_b200.so")
    return SO


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(SO)
