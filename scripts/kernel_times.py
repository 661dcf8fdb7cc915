import sys
sys.path.insert(0, '/root/repo')
import torch, time
import pygeon_b200 as pg
from pygeon_b200 import engine
dev = torch.device("cuda", 0)
for space, kind, n in (("nedelec0", "n0_mass", 64), ("rt0", "rt0_mass", 96), ("bdm1", "bdm1_mass", 48), ("darcy", "darcy_saddle", 96)):
    sd = pg.structured_grid(3, n, device=dev); sd.compute_connectivity()
    raw = engine.RawGrid.upload(sd, dev, ridges=(space == "nedelec0")); raw.check = False
    pack = engine.GridPack.from_raw(raw)
    plan = engine.symbolic(pack, space)
    for rep in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e[0].record(); vals = engine.local_matrices(pack, kind, plan=plan); e[1].record()
        data = engine.numeric(plan, vals); e[2].record(); torch.cuda.synchronize()
        print(space, n, "caps", plan.max_row_nnz, "local %.3f ms numeric %.3f ms wall %.3f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), 1e3*(time.perf_counter()-t0)))
