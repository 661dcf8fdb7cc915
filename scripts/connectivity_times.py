"""Device connectivity / geometry (pgb_ridges3d, pgb_grid_geometry, pgb_edge_geometry) against the numpy
path of SimplexGrid on the raw arrays of a structured tet grid (dev tool).
usage: python scripts/connectivity_times.py [n_device] [n_host]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pygeon_b200 as pg
from pygeon_b200 import connectivity, engine
from pygeon_b200.grid import SimplexGrid

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_host = int(sys.argv[2]) if len(sys.argv) > 2 else 48
dev = torch.device("cuda", 0)


def raw_copy(sd):
    return SimplexGrid(sd.dim, sd.nodes, sd.face_nodes, sd.cell_faces, "copy")


# host numpy vs whole device call (uploads and downloads included), same grid
sd = pg.structured_grid(3, n_host, device=dev)
a, b = raw_copy(sd), raw_copy(sd)
t0 = time.perf_counter()
a.compute_geometry()
t_host = time.perf_counter() - t0
b.compute_geometry(device=dev)  # warm-up (allocator, module load)
b = raw_copy(sd)
torch.cuda.synchronize()
t0 = time.perf_counter()
b.compute_geometry(device=dev)
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
assert np.array_equal(a.face_ridges.indices, b.face_ridges.indices) and np.allclose(a.cell_volumes, b.cell_volumes)
print(f"n={n_host} cells={sd.num_cells}: numpy compute_geometry {t_host:.2f} s, device (host in / host out) {t_dev:.3f} s "
      f"-> {t_host / t_dev:.0f}x", flush=True)
del a, b, sd

# kernels alone at the bench size
sd = pg.structured_grid(3, n_dev, device=dev)
raw = engine.RawGrid.upload(sd, dev, signs=True)
pack = engine.GridPack.from_raw(raw)
nc, nf, nn = raw.nc, raw.nf, raw.nn
fn = raw.fn_idx.view(nf, 3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(3):
    torch.cuda.synchronize()
    ev[0].record()
    fr, fs, rp = connectivity.ridges3d(fn, nn)
    ev[1].record()
    geo = connectivity.geometry(3, raw.nodes, raw.cf_idx.view(nc, 4), raw.cf_sgn.view(nc, 4), fn, pack.cell_nodes)
    ev[2].record()
    t, ln = connectivity.edge_geometry(raw.nodes, rp)
    ev[3].record()
    torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
gb_geo = (nf * (12 + 7 * 8) + nc * (32 + 4 * 8 + 4 * 56)) / 1e9
print(f"n={n_dev} cells={nc} faces={nf} ridges={rp.shape[0]}: pgb_ridges3d {ms[0]:.2f} ms ({3 * nf / ms[0] / 1e6:.2f} G keys/s), "
      f"pgb_grid_geometry {ms[1]:.3f} ms (~{gb_geo / ms[1] * 1e3:.0f} GB/s), pgb_edge_geometry {ms[2]:.3f} ms")
