"""One cold symbolic phase of the P1 pattern on n^3 Kuhn cubes (for ncu captures of the row kernels).
usage: python scripts/symbolic_once.py [n] [repeats]"""
import sys

sys.path.insert(0, "/root/repo")
import torch

import pygeon_b200 as pg
from pygeon_b200 import engine
from pygeon_b200._lib import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sd = pg.structured_grid(3, n, device="cuda")
pack = engine.get_pack(sd)
for _ in range(rep):
    pack.plans.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan = engine.symbolic(pack, "lagrange1")
    e1.record()
    torch.cuda.synchronize()
    print(f"symbolic {e0.elapsed_time(e1):.3f} ms  nnz {plan.nnz}  row kernels {lib.pgb_csr_last_row_kernel().decode()}")
