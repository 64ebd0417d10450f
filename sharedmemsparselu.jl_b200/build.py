"""Build libsmslu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsmslu.so")
SOURCES = ["api.cu", "kernels.cu", "symbolic.cpp"]
HEADERS = ["kernels.cuh", "symbolic.hpp", os.path.join("..", "..", "include", "smslu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=default", "-shared"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    out = os.environ.get("SMSLU_BUILD_OUT") or LIB
    tmp = out + ".%d.tmp" % os.getpid()
    trace = ["-DSMSLU_TRACE"] if os.environ.get("SMSLU_TRACE") == "1" else []
    if trace and os.environ.get("SMSLU_TRACE_ROWS"):
        trace.append("-DSMSLU_TRACE_ROWS=" + os.environ["SMSLU_TRACE_ROWS"])
    trace += os.environ.get("SMSLU_NVCC_FLAGS", "").split()          # A/B builds: extra -D switches
    cmd = [nvcc()] + NVCC_FLAGS + trace + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
