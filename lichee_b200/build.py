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


CLI = os.path.join(HERE, "lichee_b200_build")
HOST_SRC = [os.path.join(HERE, "csrc", f) for f in ("lichee_host.cpp", "lichee_cli.cpp")]


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(target) < os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False):
    hdr = os.path.join(HERE, "..", "include", "lichee_b200.h")
    if force or _stale(SO, [SRC, hdr]):
        res = subprocess.run([NVCC] + FLAGS + ["-o", SO, SRC], capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libLICHeE_b200.so")
    # C++ host mirror of the reference's side of the boundary + the `-build` command line
    if force or _stale(CLI, HOST_SRC + [os.path.join(HERE, "csrc", "lichee_host.hpp"), hdr, SO]):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", CLI] + HOST_SRC + [SO, "-Wl,-rpath,$ORIGIN"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("g++ failed building lichee_b200_build")
    return SO


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(SO)
