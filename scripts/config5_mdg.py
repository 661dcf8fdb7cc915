"""BASELINE config #5 in miniature on one GPU: cube of n^3 Kuhn cubes cut by a fracture plane (3D matrix + 2D
fracture + conforming mortar grid), mixed-dimensional RT0/P0 Darcy system with the mortar coupling:
M = pg.face_mass(mdg) (GPU assembly per subdomain + mortar trace term), B = pg.cell_mass(mdg) @ pg.div(mdg)
(jump operator included), symmetrised saddle system solved by the GPU MINRES with a block-Jacobi diagonal.  Two routes to the same system: the host operators
(pg.face_mass / pg.cell_mass / pg.div + sps.block_array, the reference's calls) and the device-resident one
(pg.mdg_darcy_saddle_device: add_diagonal / transpose / block-stacking kernels, matrix never leaves HBM).
usage: python scripts/config5_mdg.py [n]"""
import json
import sys
import time

sys.path.insert(0, "/root/repo")
import numpy as np
import scipy.sparse as sps
import torch

import pygeon_b200 as pg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
t0 = time.perf_counter()
mdg = pg.fractured_unit_cube(n)
mdg.compute_geometry()
t_grid = time.perf_counter() - t0
sd3, sd2 = mdg.subdomains()
intf = mdg.interfaces()[0]
for sd, d in mdg.subdomains(return_data=True):
    d.update({pg.PARAMETERS: {"flow": {pg.SECOND_ORDER_TENSOR: pg.SecondOrderTensor(np.full(sd.num_cells, 1.0 if sd.dim == 3 else 1e-2))}}})
mdg.interface_data(intf).update({pg.PARAMETERS: {"flow": {pg.NORMAL_DIFFUSIVITY: np.full(intf.num_cells, 1e2)}}})
torch.cuda.synchronize()
t0 = time.perf_counter()
M = pg.face_mass(mdg, keyword="flow")
C = pg.cell_mass(mdg, keyword="flow")
torch.cuda.synchronize()
t_asm = time.perf_counter() - t0
t0 = time.perf_counter()
B = (C @ pg.div(mdg)).tocsr()
sym = sps.block_array([[M, -B.T], [-B, None]]).tocsr()
t_stack = time.perf_counter() - t0
nf, nc = M.shape[0], C.shape[0]
rhs = np.zeros(nf + nc)
rhs[nf + sd3.num_cells:] = -1.0  # unit source in the fracture cells (second block row negated)
dinv = 1.0 / np.concatenate([M.diagonal(), np.asarray(B.multiply(B) @ (1.0 / M.diagonal())).ravel()])
torch.cuda.synchronize()
t0 = time.perf_counter()
x, info = pg.minres(sym, rhs, rtol=1e-8, maxiter=5000, M=dinv)
torch.cuda.synchronize()
t_solve = time.perf_counter() - t0
si = pg.minres.last_info
res = np.linalg.norm(sym @ x - rhs) / np.linalg.norm(rhs)
# device-resident route (the subdomain plans are cached: this is the warm cost of a re-assembly)
torch.cuda.synchronize()
t0 = time.perf_counter()
A, nfd, ncd = pg.mdg_darcy_saddle_device(mdg, keyword="flow")
dinv_d = pg.darcy_block_jacobi(A, nfd)
torch.cuda.synchronize()
t_dev_build = time.perf_counter() - t0
b_d = torch.from_numpy(rhs).cuda()
torch.cuda.synchronize()
t0 = time.perf_counter()
xd, info_d = pg.minres_device(A, b_d, rtol=1e-8, maxiter=5000, dinv=dinv_d)
torch.cuda.synchronize()
t_dev_solve = time.perf_counter() - t0
# exact comparison with the host stack of the same blocks (M enters transposed: the engine's CSR arrays are the
# CSC arrays of the reference's matrix, equal up to the last bit of the separately computed a[i][j] / a[j][i])
sym_t = sps.block_array([[sps.csr_array(M.T), -B.T], [-B, None]]).tocsr()
sym_t.sort_indices()
same = bool(A.nnz == sym_t.nnz and np.array_equal(A.indptr.cpu().numpy(), sym_t.indptr)
            and np.array_equal(A.indices.cpu().numpy(), sym_t.indices) and np.array_equal(A.data.cpu().numpy(), sym_t.data))
print(json.dumps({
    "workload": f"cube {n}^3 cut by a fracture plane: {sd3.num_cells} tets + {sd2.num_cells} fracture triangles + {intf.num_cells} mortar cells",
    "unknowns": nf + nc, "nnz": int(sym.nnz), "grid_build_host_s": t_grid, "gpu_block_assembly_s": t_asm,
    "host_block_stacking_s": t_stack, "minres_iterations": si.iterations, "minres_info": int(info), "minres_s": t_solve,
    "relative_residual": float(res), "mean_fracture_pressure": float(x[nf + sd3.num_cells:].mean() * sd2.num_cells),
    "device_route": {"build_s (assembly + trace term + B upload + transpose + stacking + block Jacobi)": t_dev_build,
                     "matrix_identical_to_host_stack": same, "minres_iterations": info_d.iterations,
                     "minres_s": t_dev_solve, "max_abs_diff_x": float(np.abs(xd.cpu().numpy() - x).max())}}))
