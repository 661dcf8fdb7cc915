"""BASELINE config #3 on ONE GPU, the system build around the assembled matrix: one-pass
saddle assembly on n^3 cubes, essential flux conditions on the boundary faces of one side eliminated in HBM
(DeviceLinearSystem.reduce_system: pgb_mask_to_map + pgb_csr_extract_*), block-Jacobi diagonal
(pgb_darcy_block_jacobi), a few preconditioned MINRES iterations.  Prints times and the peak device memory.
usage: python scripts/darcy_config3_system.py [n]"""
import json
import sys
import time

sys.path.insert(0, "/root/repo")
import numpy as np
import torch

import pygeon_b200 as pg
from pygeon_b200 import darcy, engine
from pygeon_b200.device_system import DeviceLinearSystem
from pygeon_b200.solvers import minres_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
torch.cuda.synchronize()


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t) * 1e3


A, t_asm = timed(lambda: darcy.darcy_saddle_device(sd))
engine.clear_cache(sd)  # the plan / pack are not needed any more
torch.cuda.empty_cache()
nf, nc = sd.num_faces, sd.num_cells
# essential flux condition on the boundary faces whose nodes all lie in the plane z = 0
fn = sd.face_nodes.indices.reshape(nf, 3)
bottom = (fn < (n + 1) ** 2).all(axis=1)
ess = np.zeros(nf + nc, dtype=bool)
ess[:nf] = bottom
b = torch.zeros(nf + nc, dtype=torch.float64, device=dev)
b[nf:] = -1.0
dls = DeviceLinearSystem(A, b)
dls.flag_ess_bc(ess, np.zeros(nf + nc))
(A0, b0, keep), t_red = timed(dls.reduce_system)
del dls, A
torch.cuda.empty_cache()
nf0 = int((keep < nf).sum().item())
dinv, t_jac = timed(lambda: darcy.darcy_block_jacobi(A0, nf0))
(x, info), t_mr = timed(lambda: minres_device(A0, b0, rtol=0.0, maxiter=20, dinv=dinv))
print(json.dumps({
    "workload": f"Darcy saddle system on {n}^3 cubes ({6 * n ** 3} tets), one GPU", "unknowns": nf + nc, "nnz": A0.nnz,
    "essential_dofs": int(ess.sum()), "reduced_unknowns": int(keep.numel()),
    "assembly_ms": t_asm, "reduce_system_ms": t_red, "block_jacobi_ms": t_jac, "minres_ms_per_iter": t_mr / max(info.iterations, 1),
    "residual_first_last": [float(info.residuals[0]), float(info.residuals[-1])],
    "peak_device_GiB": torch.cuda.max_memory_allocated() / 2 ** 30}))
