"""BASELINE config #1: 2D mixed Darcy on the 256x256 structured triangle grid (131 072 cells):
RT0 mass + div + PwConstants.  GPU public API next to the reference algorithm (oracle port) on the
host, full size; the saddle system is solved with scipy spsolve (the reference's solver) and with
the GPU MINRES, and the two solutions are compared.

    python scripts/config1_2d.py [n]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla
import torch

import pygeon_b200 as pg
from oracle import fem_oracle as fo
from pygeon_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sd = pg.structured_grid(2, n)
sd.compute_geometry()
nf, nc = sd.num_faces, sd.num_cells


def best(f, reps=3):
    ts, out = [], None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def gpu_blocks():
    engine.clear_cache(sd)
    M = pg.RT0().assemble_mass_matrix(sd)
    B = pg.PwConstants().assemble_mass_matrix(sd) @ pg.RT0().assemble_diff_matrix(sd)
    return M, B


def cpu_blocks():
    M = fo.port_mass("rt0", sd)
    B = fo.p0_mass(sd) @ fo.diff_matrix("rt0", sd)
    return M, B


gpu_blocks()
t_gpu, (M, B) = best(gpu_blocks)
t_cpu, (Mc, Bc) = best(cpu_blocks, reps=2)
scale = abs(Mc).max()
err_M = abs(sps.csr_array(M) - sps.csr_array(Mc)).max() / scale
err_B = abs(B - Bc).max() / abs(Bc).max()
t_dev, Ad = best(lambda: (engine.clear_cache(sd), pg.darcy_saddle_device(sd))[1])
t_dev_warm, Ad = best(lambda: pg.darcy_saddle_device(sd))

# p = x on the boundary as natural condition (tests/patch_tests/test_laplacian_mixed.py:52-60)
bf = sd.tags["domain_boundary_faces"]
sign = np.asarray(sd.cell_faces.sum(axis=1)).ravel()
rhs = np.zeros(nf + nc)
rhs[:nf][bf] = -sd.face_centers[0, bf] * sign[bf]
spp = sps.block_array([[Mc, -Bc.T], [Bc, None]]).tocsc()  # the reference's (non-symmetric) system
t_sp = time.perf_counter()
x_ref = spla.spsolve(spp, rhs)
t_sp = time.perf_counter() - t_sp
t_mr = time.perf_counter()
dinv = pg.darcy_block_jacobi(Ad, nf)
x_gpu, info = pg.minres(Ad, rhs, rtol=1e-13, maxiter=200000, M=dinv)  # symmetrised system, same rhs (f_p = 0)
torch.cuda.synchronize()
t_mr = time.perf_counter() - t_mr
err_x = np.abs(x_gpu - x_ref).max() / np.abs(x_ref).max()
err_p = np.abs(x_gpu[nf:] / sd.cell_volumes - sd.cell_centers[0]).max()
print(json.dumps({
    "workload": f"2D mixed Darcy RT0/P0 on {n}x{n} squares ({nc} triangles, {nf} faces)",
    "assembly_gpu_public_api_s": t_gpu, "assembly_cpu_port_s": t_cpu, "speedup_e2e": t_cpu / t_gpu,
    "cells_per_s_gpu_e2e": nc / t_gpu, "cells_per_s_cpu": nc / t_cpu,
    "saddle_one_pass_device_cold_s": t_dev, "saddle_one_pass_device_warm_s": t_dev_warm,
    "rel_err_M_vs_port": err_M, "rel_err_B_vs_port": err_B,
    "spsolve_cpu_s": t_sp, "minres_gpu_s": t_mr, "minres_iters": pg.minres.last_info.iterations, "minres_info": info,
    "rel_diff_solution_gpu_minres_vs_spsolve": err_x, "max_err_pressure_vs_exact_linear": err_p}))
