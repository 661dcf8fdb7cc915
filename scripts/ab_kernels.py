"""A/B dev tool: K1 / K2-numeric / SpMV times per space on one GPU (median of 7 after 3 warm-ups).
usage: python scripts/ab_kernels.py [n]   (the library under test is pygeon_b200/libpgb.so)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pygeon_b200 as pg
from pygeon_b200 import _lib, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
st = torch.cuda.current_stream().cuda_stream
for space, kind in (("lagrange1", "l1_stiff"), ("rt0", "rt0_mass"), ("nedelec0", "n0_mass"), ("darcy", "darcy_saddle")):
    raw = engine.RawGrid.upload(sd, dev, ridges=(space == "nedelec0"))
    raw.check = False
    pack = engine.GridPack.from_raw(raw)
    plan = engine.symbolic(pack, space)
    t1, t2, t3 = [], [], []
    nr, nnz = plan.n_rows, plan.nnz
    avg = nnz / nr
    tpr = 2 if avg <= 9 else 4 if avg <= 18 else 8 if avg <= 40 else 16
    x = torch.randn(nr, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    for rep in range(10):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        e[0].record()
        vals = engine.local_matrices(pack, kind, plan=plan)
        e[1].record()
        data = engine.numeric(plan, vals)
        e[2].record()
        for _ in range(10):
            _lib.check(_lib.lib.pgb_spmv(nr, plan.indptr.data_ptr(), 0, plan.indices.data_ptr(), data.data_ptr(),
                                         x.data_ptr(), y.data_ptr(), tpr, st))
        e[3].record()
        torch.cuda.synchronize()
        if rep >= 3:
            t1.append(e[0].elapsed_time(e[1]))
            t2.append(e[1].elapsed_time(e[2]))
            t3.append(e[2].elapsed_time(e[3]) / 10)
        del vals
    print(f"{space:10s} n={n} K1 {np.median(t1):.4f} ms  K2 {np.median(t2):.4f} ms  spmv {np.median(t3) * 1e3:.1f} us "
          f"({(12 * nnz + 20 * nr) / np.median(t3) / 1e6:.0f} GB/s)", flush=True)
    del pack, plan, data, raw, x, y
    torch.cuda.empty_cache()
