"""Cold / warm assembly time and the symbolic phase split for every space (one GPU).
usage: python scripts/space_times.py [n]"""
import ctypes as C
import sys

sys.path.insert(0, "/root/repo")
import numpy as np
import torch

import pygeon_b200 as pg
from pygeon_b200 import _lib, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
for space, kind in (("lagrange1", "l1_stiff"), ("rt0", "rt0_mass"), ("nedelec0", "n0_mass"), ("bdm1", "bdm1_mass"),
                    ("darcy", "darcy_saddle"), ("lagrange2", "l2_stiff"), ("rt1", "rt1_mass")):
    if len(sys.argv) > 2 and space not in sys.argv[2:]:
        continue
    raw = engine.RawGrid.upload(sd, dev, ridges=(space in ("nedelec0", "lagrange2")))
    raw.check = False
    cold, warm = [], []
    ph = np.zeros(4)
    for rep in range(5):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        _lib.lib.pgb_set_profile(1 if rep >= 2 else 0)
        torch.cuda.synchronize()
        e[0].record()
        pack = engine.GridPack.from_raw(raw)
        plan = engine.symbolic(pack, space)
        data = engine.numeric(plan, engine.local_matrices(pack, kind, plan=plan))
        e[1].record()
        data = engine.numeric(plan, engine.local_matrices(pack, kind, plan=plan))
        e[2].record()
        torch.cuda.synchronize()
        if rep >= 2:
            cold.append(e[0].elapsed_time(e[1]))
            warm.append(e[1].elapsed_time(e[2]))
            buf = (C.c_double * 4)()
            _lib.lib.pgb_get_phase_ms(buf, 4)
            ph += np.array(list(buf)) / 3
    _lib.lib.pgb_set_profile(0)
    print(f"{space:10s} n={n} cells={sd.num_cells} rows={plan.n_rows} nnz={plan.nnz} L={plan.L} caps={plan.max_row_nnz} "
          f"cold {np.mean(cold):.3f} ms warm {np.mean(warm):.3f} ms | transpose {ph[0]:.3f} place/canon {ph[1]:.3f} "
          f"rowsort {ph[2]:.3f} scan {ph[3]:.3f}", flush=True)
    del pack, plan, data, raw
    torch.cuda.empty_cache()
