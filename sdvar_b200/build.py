"""In-tree nvcc build of libsdvar_b200.so (sm_100a only; see __graft_entry__.build)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libsdvar_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(obj: str, src: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "sdvar_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, src):
            log = open(obj + ".log", "w")
            procs.append((src, log, subprocess.Popen([NVCC, *FLAGS, "-c", src, "-o", obj], stdout=log, stderr=subprocess.STDOUT)))
    failed = False
    for src, log, p in procs:
        rc = p.wait()
        log.close()
        out = open(log.name).read()
        if rc != 0 or verbose:
            print(f"---- {os.path.basename(src)} (rc={rc})\n{out}", file=sys.stderr)
        failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(LIB):
        # cudart is linked statically; the driver API (cuTensorMapEncodeTiled) is resolved at run time through
        # cudaGetDriverEntryPoint so the library loads on a machine without libcuda.so.1 (CPU-only CI).
        subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xlinker", "--no-undefined",
                               "-lpthread", "-ldl", "-lrt"])
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
