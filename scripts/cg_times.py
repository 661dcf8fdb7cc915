"""CG / MINRES time per iteration on assembled operators (dev tool; library under test = pygeon_b200/libpgb.so).
usage: python scripts/cg_times.py [n] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygeon_b200 as pg
from pygeon_b200 import engine
from pygeon_b200.solvers import cg_device, minres_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
for space, kinds in (("lagrange1", ("l1_stiff", "l1_mass")), ("rt0", ("rt0_mass",))):
    raw = engine.RawGrid.upload(sd, dev)
    raw.check = False
    pack = engine.GridPack.from_raw(raw)
    plan = engine.symbolic(pack, space)
    data = engine.numeric(plan, engine.local_matrices(pack, kinds[0], plan=plan))
    for k in kinds[1:]:
        data += engine.numeric(plan, engine.local_matrices(pack, k, plan=plan))
    A = engine.DeviceCSR((plan.n_rows, plan.n_rows), plan.indptr, plan.indices, data)
    nr, nnz = plan.n_rows, plan.nnz
    b = torch.ones(nr, dtype=torch.float64, device=dev)
    for name, fn, vecs in (("cg", cg_device, 88), ("minres", minres_device, 112)):
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            x, info = fn(A, b, rtol=0.0, maxiter=iters)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / max(info.iterations, 1))
        byt = 12 * nnz + 4 * nr + vecs * nr
        print(f"{space:10s} n={n} rows={nr} {name:6s} {info.iterations} its: {best * 1e3:7.1f} us/it  {byt / best / 1e6:6.0f} GB/s "
              f"resid[-1]={info.residuals[-1]:.6e}", flush=True)
    del A, data, plan, pack, raw, b, x
    torch.cuda.empty_cache()
