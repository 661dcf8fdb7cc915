"""Where does the end-to-end (host arrays in, scipy matrix out) time go?  (dev tool)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sps
import torch

import pygeon_b200 as pg
from pygeon_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
for mat in (sd.cell_faces, sd.face_nodes):
    mat.indices = torch.from_numpy(mat.indices).pin_memory().numpy()
    mat.data = torch.from_numpy(mat.data).pin_memory().numpy()
sd.nodes = torch.from_numpy(sd.nodes).pin_memory().numpy()
disc = pg.Lagrange1()


def t(f, reps=3):
    out = None
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, out


ms, raw = t(lambda: engine.RawGrid.upload(sd, dev))
print(f"upload raw grid ({raw.h2d_bytes / 1e6:.0f} MB): {ms:.1f} ms")
ms, pack = t(lambda: engine.GridPack.from_raw(raw))
print(f"pack: {ms:.1f} ms")
ms, plan = t(lambda: engine.symbolic(engine.GridPack.from_raw(raw), "lagrange1"))
print(f"pack+symbolic: {ms:.1f} ms")
ms, A = t(lambda: (engine.clear_cache(sd), disc.assemble_stiff_matrix_device(sd))[1])
print(f"assemble_device (upload+all kernels): {ms:.1f} ms")
ms, parts = t(lambda: (A.indptr.cpu(), A.indices.cpu(), A.data.cpu()))
print(f"D2H pageable .cpu() ({sum(p.numel() * p.element_size() for p in parts) / 1e6:.0f} MB): {ms:.1f} ms")
pin = [torch.empty(p.shape, dtype=p.dtype).pin_memory() for p in parts]
ms, _ = t(lambda: [pp.copy_(src, non_blocking=True) for pp, src in zip(pin, (A.indptr, A.indices, A.data))])
print(f"D2H into pre-pinned buffers: {ms:.1f} ms")
ms, _ = t(lambda: [torch.empty(p.shape, dtype=p.dtype).pin_memory() for p in parts], reps=2)
print(f"allocating pinned buffers: {ms:.1f} ms")
ip, ix, da = (p.numpy() for p in parts)
ms, M = t(lambda: sps.csc_array((da, ix, ip), shape=A.shape))
print(f"scipy csc_array constructor: {ms:.1f} ms")
ms, M = t(lambda: engine._fast_csc(da, ix, ip, A.shape) if hasattr(engine, "_fast_csc") else None)
print(f"fast csc wrap: {ms:.1f} ms")
ms, _ = t(lambda: (engine.clear_cache(sd), disc.assemble_stiff_matrix(sd))[1])
print(f"full public call: {ms:.1f} ms")
