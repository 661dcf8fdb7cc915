"""Generate tests/golden_mdg/*.npz: the UNMODIFIED reference's mixed-dimensional block operators on the cube
cut by a fracture plane (pygeon_b200.fractured_unit_cube).  TEST INFRASTRUCTURE ONLY; build container only.

    python -m oracle.gen_golden_mdg

Stored: the reference's pg.cell_mass / face_mass / ridge_mass / peak_mass / face_stiff (numerics/innerproducts.py,
stiffness.py) with a per-cell permeability on both subdomains and a normal diffusivity on the interface, and the
solution of the mixed-dimensional Darcy system built from them (SuperLU)."""

from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pygeon_b200 as pgb  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.ref_mdg import ref_mdg  # noqa: E402

CASES = [("fcube2", 2, 0.0), ("fcube4_pert", 4, 0.15)]


def coefficients(nc, seed):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(3, 3, nc))
    return np.einsum("iac,jac->ijc", A, A) + np.eye(3)[:, :, None], rng.uniform(0.5, 2.0, nc)


def main():
    rpg, pp = ref_loader.load()
    out_dir = os.path.join(ROOT, "tests", "golden_mdg")
    os.makedirs(out_dir, exist_ok=True)
    for name, n, pert in CASES:
        mdg = pgb.fractured_unit_cube(n, perturb=pert, seed=3)
        mdg.compute_geometry()
        rm, table = ref_mdg(mdg)
        store = {"n": n, "perturb": pert}
        for i, (sd, d_sd) in enumerate(rm.subdomains(return_data=True)):
            K, w = coefficients(sd.num_cells, 10 + i)
            sot = pp.SecondOrderTensor(np.ones(sd.num_cells))
            sot.values = K
            d_sd.update({pp.PARAMETERS: {"flow": {rpg.SECOND_ORDER_TENSOR: sot, rpg.WEIGHT: w}}})
            store[f"K{i}"], store[f"w{i}"] = K, w
        intf, d_intf = rm.interfaces(return_data=True)[0]
        kn = np.linspace(0.5, 2.0, intf.num_cells)
        d_intf.update({pp.PARAMETERS: {"flow": {rpg.NORMAL_DIFFUSIVITY: kn}}})
        store["kn"] = kn
        mats = {"cell_mass": rpg.cell_mass(rm, keyword="flow"), "face_mass": rpg.face_mass(rm, keyword="flow"),
                "ridge_mass": rpg.ridge_mass(rm, keyword="flow"), "peak_mass": rpg.peak_mass(rm, keyword="flow"),
                "face_stiff": rpg.face_stiff(rm, keyword="flow"), "lumped_face_mass": rpg.lumped_face_mass(rm, keyword="flow"),
                "div": rpg.div(rm)}
        for k, M in mats.items():
            M = sps.csr_array(M)
            M.sum_duplicates()
            M.sort_indices()
            store[k + "_indptr"], store[k + "_indices"], store[k + "_data"] = M.indptr, M.indices, M.data
            store[k + "_shape"] = np.array(M.shape)
        # mixed-dimensional Darcy: [[M, -B^T], [B, 0]] (p, q) with a unit source in the fracture cells
        M, B = mats["face_mass"], mats["cell_mass"] @ mats["div"]
        spp = sps.block_array([[M, -B.T], [B, None]]).tocsc()
        rhs = np.zeros(spp.shape[0])
        nf, nc3 = M.shape[0], rm.subdomains()[0].num_cells
        rhs[nf + nc3:] = 1.0
        store["darcy_rhs"], store["darcy_sol"] = rhs, sps.linalg.spsolve(spp, rhs)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **store)
        print(f"{name}: wrote {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
