"""Build libtmgcn_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m tmgcn_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtmgcn_b200.so")
SOURCES = ["api.cu", "sparse.cu", "stencil.cu", "spmm.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_dw_tc.cu", "gemm.cu",
           "edge.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


def _nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "tmgcn.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu into objects (parallel) and link the shared library."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libtmgcn_b200.so")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        s = os.path.join(CSRC, src)
        headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
        headers.append(os.path.join(HERE, "..", "include", "tmgcn.h"))
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(s)
                and all(os.path.getmtime(obj) > os.path.getmtime(h) for h in headers)):
            procs.append((src, obj, None))
            continue
        cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        if p is not None:
            out, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{out}")
            if verbose or out.strip():
                print(f"[{src}]\n{out}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
