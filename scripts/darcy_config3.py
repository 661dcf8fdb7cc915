"""BASELINE config #3 on one GPU: 3D mixed Darcy RT0/P0 saddle system on the n^3 tet cube
(n = 256: 100 663 296 tets, 302 M unknowns), one-pass assembly + MINRES iterations.

    python scripts/darcy_config3.py [n] [minres_iters]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygeon_b200 as pg
from pygeon_b200 import engine
from pygeon_b200.solvers import minres_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
peak = 6525.9
t0 = time.perf_counter()
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
raw = engine.RawGrid.upload(sd, dev)
raw.check = False
torch.cuda.synchronize()
t_grid = time.perf_counter() - t0


def ev():
    return torch.cuda.Event(enable_timing=True)


# cold: pack + symbolic (new plan each time); warm: K1 + K2 into preallocated buffers
res = {}
for rep in range(2):
    e = [ev() for _ in range(2)]
    e[0].record()
    pack = engine.GridPack.from_raw(raw)
    plan = engine.symbolic(pack, "darcy")
    e[1].record()
    torch.cuda.synchronize()
    res["pack+symbolic_ms"] = e[0].elapsed_time(e[1])
    if rep == 0:
        del pack, plan
L = 5 if True else 0
pitch = (L + 3) & ~3 if plan.mode == 0 else L
vals = torch.empty((sd.num_cells, L * pitch), dtype=torch.float64, device=dev)
data = torch.empty((plan.nnz,), dtype=torch.float64, device=dev)
for rep in range(3):
    e = [ev() for _ in range(3)]
    e[0].record()
    engine.local_matrices(pack, "darcy_saddle", out=vals, plan=plan)
    e[1].record()
    engine.numeric(plan, vals, out=data)
    e[2].record()
    torch.cuda.synchronize()
    res["local_ms"], res["numeric_ms"] = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
del vals
nc, nr, nnz = sd.num_cells, plan.n_rows, plan.nnz
A = engine.DeviceCSR((nr, nr), plan.indptr, plan.indices, data)
del pack
torch.manual_seed(0)
b = torch.randn(nr, dtype=torch.float64, device=dev)
minres_device(A, b, rtol=0.0, maxiter=3)
torch.cuda.synchronize()
c0, c1 = ev(), ev()
c0.record()
x, info = minres_device(A, b, rtol=0.0, maxiter=iters)
c1.record()
torch.cuda.synchronize()
ms_it = c0.elapsed_time(c1) / max(info.iterations, 1)
ipb = 8 if A.idx64 else 4
byt = 12 * nnz + ipb * (nr + 1) + 112 * nr
cold = sum(res.values())
out = {
    "workload": f"mixed Darcy RT0/P0 saddle [[M,-B^T],[-B,0]] on {n}^3 tets", "cells": nc, "unknowns": nr, "nnz": nnz,
    "indptr": "int64" if A.idx64 else "int32", **{k: round(v, 3) for k, v in res.items()},
    "cold_cells_per_s": nc / (cold * 1e-3), "warm_cells_per_s": nc / ((res["local_ms"] + res["numeric_ms"]) * 1e-3),
    "minres_iters": info.iterations, "minres_ms_per_iter": ms_it, "minres_bytes_per_iter": byt,
    "minres_gbs": byt / ms_it / 1e6, "minres_frac_of_measured_peak": byt / ms_it / 1e6 / peak,
    "residual_first_last": [float(info.residuals[0]), float(info.residuals[-1])],
    "grid_setup_s": round(t_grid, 2), "max_mem_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1),
}
print(json.dumps(out))
