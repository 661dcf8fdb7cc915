"""Build libpgb.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libpgb.so")
SOURCES = ["pgb_api.cu", "pgb_pack.cu", "pgb_local.cu", "pgb_csr.cu", "pgb_rowsort.cu", "pgb_rowthread.cu",
           "pgb_rowhash.cu", "pgb_rowins.cu", "pgb_rowlong.cu", "pgb_krylov.cu", "pgb_geometry.cu", "pgb_system.cu", "pgb_dist.cu", "pgb_numeric_bulk.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", f"-I{INCLUDE}", f"-I{CSRC}",
    *os.environ.get("PGB_NVCC_EXTRA", "").split(),  # A/B builds of a tuning macro (-DPGB_...)
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in ("pgb_common.cuh", "pgb_rows.cuh", "pgb_rt1_table.cuh", "pgb_krylov.cuh")]
    headers.append(os.path.join(INCLUDE, "pgb.h"))
    srcs = [os.path.join(CSRC, src) for src in SOURCES]
    if not force and not _stale(LIB, srcs + headers):
        return LIB  # e.g. on a GPU box: the library travels, the object files do not
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode()}")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout.decode()}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
