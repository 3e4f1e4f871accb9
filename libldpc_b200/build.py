"""Builds libldpc_b200/libldpc.so (C ABI + CUDA kernels, sm_100a) and the ldpcsim CLI with nvcc.

In-tree, explicit nvcc invocation (no JIT cache): the built files travel with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libldpc.so")
CLI = os.path.join(HERE, "ldpcsim")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]
SOURCES = ["engine.cu", "code.cpp", "sim_driver.cpp", "shim.cpp"]
HEADERS = ["engine.hpp", "code.hpp", "kernels.cuh", "bec_kernel.cuh", "../../include/ldpc_b200.h"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if force or _stale(LIB, deps):
        cmd = [NVCC] + ARCH + COMMON + ["-shared", "-cudart", "static", "-o", LIB] + srcs
        if verbose:
            cmd += ["-Xptxas", "-v"]
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    cli_src = os.path.join(CSRC, "cli_main.cpp")
    if force or _stale(CLI, [cli_src, LIB]):
        cmd = [NVCC] + ARCH + COMMON + ["-o", CLI, cli_src, LIB, "-Xlinker", "-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
