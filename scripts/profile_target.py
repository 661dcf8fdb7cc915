"""Small fixed program for ncu: `reps` cold assemblies + a few CG iterations on a n^3 tet grid.

    python scripts/profile_target.py [n] [space] [reps] [cg_iters]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygeon_b200 as pg
from pygeon_b200 import engine
from pygeon_b200.solvers import cg_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
space = sys.argv[2] if len(sys.argv) > 2 else "lagrange1"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cg_iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
kind = {"lagrange1": "l1_stiff", "rt0": "rt0_mass", "nedelec0": "n0_mass", "bdm1": "bdm1_mass", "darcy": "darcy_saddle", "lagrange2": "l2_stiff"}[space]
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
raw = engine.RawGrid.upload(sd, dev, ridges=(space in ("nedelec0", "lagrange2")))
raw.check = False
import time
for _ in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pack = engine.GridPack.from_raw(raw)
    plan = engine.symbolic(pack, space)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    vals = engine.local_matrices(pack, kind, plan=plan)
    data = engine.numeric(plan, vals)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"cells {sd.num_cells} rows {plan.n_rows} nnz {plan.nnz}: pack+symbolic {1e3 * (t1 - t0):.2f} ms, "
          f"local+numeric {1e3 * (t2 - t1):.2f} ms -> cold {sd.num_cells / (t2 - t0):.3g} cells/s, "
          f"warm {sd.num_cells / (t2 - t1):.3g} cells/s, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    del vals
torch.cuda.synchronize()
if cg_iters and space == "lagrange1":
    A = engine.DeviceCSR((plan.n_rows, plan.n_rows), plan.indptr, plan.indices, data)
    A.data += engine.numeric(plan, engine.local_matrices(pack, "l1_mass", plan=plan))
    b = torch.ones(plan.n_rows, dtype=torch.float64, device=dev)
    x, info = cg_device(A, b, rtol=0.0, maxiter=cg_iters)
    torch.cuda.synchronize()
print("done", space, n, plan.nnz)
