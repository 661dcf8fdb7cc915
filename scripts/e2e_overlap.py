"""e2e public call with / without the copy overlaps (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pygeon_b200 as pg
from pygeon_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
for mat in (sd.cell_faces, sd.face_nodes):
    mat.indices = torch.from_numpy(mat.indices).pin_memory().numpy()
    mat.data = torch.from_numpy(mat.data).pin_memory().numpy()
sd.nodes = torch.from_numpy(sd.nodes).pin_memory().numpy()
disc = pg.Lagrange1()
for up, down in ((False, False), (True, False), (False, True), (True, True), (False, False)):
    engine.overlap_upload, engine.overlap_download = up, down
    ts = []
    for rep in range(7):
        engine.clear_cache(sd)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        A = disc.assemble_stiff_matrix(sd)
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"upload overlap {up} download overlap {down}: " + " ".join(f"{t:.1f}" for t in ts), flush=True)
