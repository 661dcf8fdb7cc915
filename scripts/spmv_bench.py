"""Time pgb_spmv for every threads-per-row choice on assembled operators (dev tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygeon_b200 as pg
from pygeon_b200 import _lib, engine

dev = torch.device("cuda", 0)
for space, kind, n in (("lagrange1", "l1_stiff", 128), ("rt0", "rt0_mass", 96), ("nedelec0", "n0_mass", 64)):
    sd = pg.structured_grid(3, n, device=dev)
    sd.compute_connectivity()
    raw = engine.RawGrid.upload(sd, dev, ridges=(space == "nedelec0"))
    pack = engine.GridPack.from_raw(raw)
    plan = engine.symbolic(pack, space)
    data = engine.numeric(plan, engine.local_matrices(pack, kind, plan=plan))
    nr, nnz = plan.n_rows, plan.nnz
    x = torch.randn(nr, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    byt = 12 * nnz + 4 * (nr + 1) + 16 * nr
    for tpr in (2, 4, 8, 16, 32):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            _lib.check(_lib.lib.pgb_spmv(nr, plan.indptr.data_ptr(), 0, plan.indices.data_ptr(), data.data_ptr(),
                                         x.data_ptr(), y.data_ptr(), tpr, st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(_lib.lib.pgb_spmv(nr, plan.indptr.data_ptr(), 0, plan.indices.data_ptr(), data.data_ptr(),
                                         x.data_ptr(), y.data_ptr(), tpr, st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{space:10s} n={n} rows={nr} nnz/row={nnz / nr:.1f} tpr={tpr:2d}: {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s")
    del sd, raw, pack, plan, data
