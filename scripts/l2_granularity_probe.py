import os, sys
sys.path.insert(0, '/root/repo')
import torch
import pygeon_b200 as pg
from pygeon_b200 import engine, _lib
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, 128, device=dev); sd.compute_connectivity()
raw = engine.RawGrid.upload(sd, dev); raw.check = False
pack = engine.GridPack.from_raw(raw)
plan = engine.symbolic(pack, "lagrange1")
vals = engine.local_matrices(pack, "l1_stiff", plan=plan)
out = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
def t(f, n=20):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("default granularity", _lib.lib.pgb_get_l2_fetch_granularity())
for g in (64, 32, 128, 64):
    _lib.check(_lib.lib.pgb_set_l2_fetch_granularity(g))
    print(g, "->", _lib.lib.pgb_get_l2_fetch_granularity(), "numeric %.3f ms" % t(lambda: engine.numeric(plan, vals, out=out)),
          "local %.3f ms" % t(lambda: engine.local_matrices(pack, "l1_stiff", out=vals, plan=plan)))
