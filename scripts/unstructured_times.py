"""Cold / warm assembly on an unstructured Delaunay tet mesh (varying node valence).  usage: [npts]"""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import pygeon_b200 as pg
from pygeon_b200 import _lib, engine
from pgb_helpers import delaunay_grid

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
t0 = time.perf_counter()
sd = delaunay_grid(3, npts, seed=1)
sd.compute_connectivity()
print(f"mesh: {sd.num_cells} tets, {sd.num_nodes} nodes, {sd.num_faces} faces, {sd.num_edges} edges in {time.perf_counter() - t0:.1f} s", flush=True)
val = np.bincount(sd.cell_node_table().ravel())
print("node valence: mean %.1f max %d; share of nodes > 32: %.3f" % (val.mean(), val.max(), (val > 32).mean()))
dev = torch.device("cuda", 0)
for space, kind in (("lagrange1", "l1_stiff"), ("rt0", "rt0_mass"), ("nedelec0", "n0_mass"), ("lagrange2", "l2_stiff")):
    raw = engine.RawGrid.upload(sd, dev, ridges=(space in ("nedelec0", "lagrange2")))
    raw.check = False
    cold, warm, ph = [], [], np.zeros(4)
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
            cold.append(e[0].elapsed_time(e[1])); warm.append(e[1].elapsed_time(e[2]))
            buf = (C.c_double * 4)(); _lib.lib.pgb_get_phase_ms(buf, 4); ph += np.array(list(buf)) / 3
    _lib.lib.pgb_set_profile(0)
    print(f"{space:10s} rows={plan.n_rows} nnz={plan.nnz} caps={plan.max_row_nnz} mode={plan.mode} cold {np.mean(cold):.3f} ms "
          f"warm {np.mean(warm):.3f} ms | count {ph[0]:.3f} place {ph[1]:.3f} rows {ph[2]:.3f} | {_lib.lib.pgb_csr_last_row_kernel().decode()}", flush=True)
