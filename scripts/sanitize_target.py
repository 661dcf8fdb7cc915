"""Small program for compute-sanitizer: every kernel family once on tiny grids."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sps
import torch

import pygeon_b200 as pg
from pygeon_b200 import engine

for dim, n in ((2, 7), (3, 3)):
    sd = pg.structured_grid(dim, n, perturb=0.1)
    sd.compute_geometry()
    rng = np.random.default_rng(0)
    K = rng.normal(size=(3, 3, sd.num_cells))
    K = np.einsum("iac,jac->ijc", K, K) + np.eye(3)[:, :, None]
    sot = pg.SecondOrderTensor(np.ones(sd.num_cells))
    sot.values = K
    data = pg.initialize_data({}, "k", {pg.SECOND_ORDER_TENSOR: sot, pg.WEIGHT: rng.uniform(0.5, 2, sd.num_cells)})
    for cls in (pg.Lagrange1, pg.RT0, pg.BDM1, pg.Nedelec0):
        d = cls("k")
        d.assemble_mass_matrix(sd, data)
        d.assemble_stiff_matrix(sd, data)
    pg.PwConstants("k").assemble_mass_matrix(sd, data)
    pg.darcy_saddle_matrix(sd, data, "k")
    for mode in (1, 2):
        pack = engine.GridPack.from_grid(sd)
        plan = engine.symbolic(pack, "nedelec0", mode=mode)
        engine.numeric(plan, engine.local_matrices(pack, "n0_mass", plan=plan))
    S = pg.Lagrange1().assemble_stiff_matrix(sd) + pg.Lagrange1().assemble_mass_matrix(sd)
    b = np.ones(sd.num_nodes)
    pg.cg(S, b, rtol=1e-10, maxiter=200)
    pg.cg(S, b, rtol=1e-10, maxiter=200, M=1.0 / S.diagonal())
    M = pg.RT0().assemble_mass_matrix(sd)
    B = pg.PwConstants().assemble_mass_matrix(sd) @ pg.RT0().assemble_diff_matrix(sd)
    A = sps.block_array([[M, -B.T], [-B, None]]).tocsc()
    pg.minres(A, np.ones(A.shape[0]), rtol=1e-8, maxiter=300)
torch.cuda.synchronize()
print("sanitize target done")
