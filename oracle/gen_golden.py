"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden

Each fixture stores one grid (nodes, face_nodes, cell_faces, and for 3D the reference's
face_ridges / ridge_peaks), one coefficient set, and for every (space, form) on the hot path the
reference's matrix put on the structural pattern: indptr / indices /
data of the canonicalised CSR plus the reference's own nnz.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fem_oracle as fo  # noqa: E402
from oracle import ref_loader  # noqa: E402
from pygeon_b200.grid import reference_element, structured_grid  # noqa: E402

CASES = [
    # name, dim, n (0 = reference element), perturb, coefficient mode
    ("ref2d", 2, 0, 0.0, "default"),
    ("ref3d", 3, 0, 0.0, "default"),
    ("tri4_pert_sym", 2, 4, 0.2, "sym"),
    ("tri4_pert_nonsym", 2, 4, 0.2, "nonsym"),
    ("tri8_lattice", 2, 8, 0.0, "default"),
    ("tet2_pert_sym", 3, 2, 0.2, "sym"),
    ("tet2_pert_nonsym", 3, 2, 0.2, "nonsym"),
    ("tet3_lattice", 3, 3, 0.0, "default"),
    ("tet3_lattice_sym", 3, 3, 0.0, "sym"),
    # 2D grid tilted in 3D: non-identity rotation_matrix (grids/grid.py:355-370, vec_l2.py:113-114, hdiv.py:493)
    ("tri4_tilted_nonsym", 2, 4, 0.2, "nonsym", True),
    ("tri3_tilted", 2, 3, 0.15, "default", True),
]


def tilt(sd):
    """Rigid motion of a planar grid into a generic plane of R^3 (fixed Euler angles + shift)."""
    a, b, c = 0.4, -0.7, 0.3
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    sd.nodes = Rz @ Ry @ Rx @ sd.nodes + np.array([[0.1], [0.2], [0.3]])
    return sd
FORMS = [
    ("lagrange1", "mass"), ("lagrange1", "stiff"), ("lagrange1", "gradgrad"),
    ("rt0", "mass"), ("rt0", "stiff"), ("bdm1", "mass"), ("bdm1", "stiff"),
    ("nedelec0", "mass"), ("nedelec0", "stiff"),
    ("lagrange1", "lumped"), ("rt0", "lumped"), ("bdm1", "lumped"),
    ("lagrange2", "mass"), ("lagrange2", "stiff"), ("rt1", "mass"),
]


def coefficients(nc, mode, seed=0):
    if mode == "default":
        return None, None
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(3, 3, nc))
    K = np.einsum("iac,jac->ijc", A, A) + np.eye(3)[:, :, None]
    if mode == "nonsym":
        K = K + 0.3 * rng.normal(size=(3, 3, nc))
    w = rng.uniform(0.5, 2.0, nc)
    return w, K


def main():
    pg, pp = ref_loader.load()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    cls = {"lagrange1": pg.Lagrange1, "rt0": pg.RT0, "bdm1": pg.BDM1, "nedelec0": pg.Nedelec0,
           "lagrange2": pg.Lagrange2, "rt1": pg.RT1}
    only = sys.argv[1:]
    for name, dim, n, pert, mode, *tilted in CASES:
        if only and name not in only:
            continue
        sd = reference_element(dim) if n == 0 else structured_grid(dim, n, perturb=pert, seed=1)
        if tilted and tilted[0]:
            tilt(sd)
        sd.compute_geometry()
        g = ref_loader.ref_grid(sd)
        w, K = coefficients(sd.num_cells, mode)
        data = None
        if mode != "default":
            sot = pp.SecondOrderTensor(np.ones(sd.num_cells))
            sot.values = K
            data = {pp.PARAMETERS: {"key": {pg.SECOND_ORDER_TENSOR: sot, pg.WEIGHT: w}}}
        store = dict(
            dim=dim,
            nodes=sd.nodes,
            fn_indices=g.face_nodes.indices, fn_indptr=g.face_nodes.indptr,
            cf_indices=g.cell_faces.indices, cf_indptr=g.cell_faces.indptr, cf_data=g.cell_faces.data,
            fr_indices=g.face_ridges.indices, fr_indptr=g.face_ridges.indptr, fr_data=g.face_ridges.data,
            rp_indices=g.ridge_peaks.indices, rp_indptr=g.ridge_peaks.indptr, rp_data=g.ridge_peaks.data,
            num_ridges=g.num_ridges, num_edges=g.num_edges,
            cell_volumes=g.cell_volumes,
        )
        if mode != "default":
            store["w"], store["K"] = w, K
        tabs = fo.cell_tables(g)
        for space, form in FORMS:
            d = cls[space]("key")
            if form == "mass":
                ref = d.assemble_mass_matrix(g, data)
                pat = fo.closed_form(space, "mass", g, data, "key", tabs)
            elif form == "lumped":
                ref = d.assemble_lumped_matrix(g, data)
                pat = fo.closed_form(space, "lumped", g, data, "key", tabs)
            elif form == "gradgrad":
                ref = d.assemble_grad_grad_matrix(g, data)
                pat = fo.closed_form(space, "stiff", g, data, "key", tabs, rot2d=False)
            else:
                ref = d.assemble_stiff_matrix(g, data)
                pat = fo.closed_form(space, "stiff", g, data, "key", tabs)
            refS = fo.canonicalise(ref, pat)
            key = f"{space}_{form}"
            store[key + "_indptr"] = refS.indptr.astype(np.int32)
            store[key + "_indices"] = refS.indices.astype(np.int32)
            store[key + "_data"] = refS.data
            store[key + "_refnnz"] = sps.csr_array(ref).nnz
        # RT1 stiffness (range space PwLinears) on the structural pattern of the RT1 dofs; broken-P1 mass
        ref = rt1_stiff = pg.RT1("key").assemble_stiff_matrix(g, data)
        refS = fo.canonicalise(ref, fo.closed_form("rt1", "mass", g, data, "key", tabs))
        store["rt1_stiff_indptr"], store["rt1_stiff_indices"] = refS.indptr.astype(np.int32), refS.indices.astype(np.int32)
        store["rt1_stiff_data"], store["rt1_stiff_refnnz"] = refS.data, sps.csr_array(ref).nnz
        for nm, mat in (("pwlinears_mass", pg.PwLinears("key").assemble_mass_matrix(g, data)),
                        ("pwlinears_lumped", pg.PwLinears("key").assemble_lumped_matrix(g, data))):
            mat = sps.csr_array(mat)
            mat.sum_duplicates()
            mat.sort_indices()
            store[nm + "_indptr"], store[nm + "_indices"], store[nm + "_data"] = (
                mat.indptr.astype(np.int32), mat.indices.astype(np.int32), mat.data)
        # error_l2 of a perturbed interpolant against the analytical function (discretization.py:304-358)
        fs = lambda x: np.sin(x[0]) + x[1] * x[2] + 2.0  # noqa: E731
        fv = lambda x: np.array([x[0] + 1.0, x[1] * x[0], x[2] - x[0]])  # noqa: E731
        for nm, d, f in (("lagrange1", pg.Lagrange1("key"), fs), ("rt0", pg.RT0("key"), fv),
                         ("bdm1", pg.BDM1("key"), fv), ("nedelec0", pg.Nedelec0("key"), fv),
                         ("p0", pg.PwConstants("key"), fs)):
            if nm == "nedelec0" and dim == 2:
                continue  # the reference's Nedelec0.interpolate reads ridge_peaks, empty in 2D (hcurl.py:160)
            u = d.interpolate(g, f)
            u = u * (1.0 + 0.05 * np.cos(np.arange(u.size)))
            store["errl2_" + nm] = np.array([d.error_l2(g, u, f, data=data), d.error_l2(g, u, f, relative=False, data=data),
                                             d.error_l2(g, u, f, poly_order=None, data=data)])
            store["errl2_u_" + nm] = u
        # RT1 pieces assembled on the host around the path: divergence, lumped matrix
        rt1 = pg.RT1("key")
        for nm, mat in (("rt1_diff", rt1.assemble_diff_matrix(g)), ("rt1_lumped", rt1.assemble_lumped_matrix(g, data))):
            mat = sps.csr_array(mat)
            mat.sum_duplicates()
            mat.sort_indices()
            store[nm + "_indptr"], store[nm + "_indices"], store[nm + "_data"] = mat.indptr, mat.indices, mat.data
            store[nm + "_shape"] = np.array(mat.shape)
        store["nedelec0_lumped_diag"] = pg.Nedelec0("key").assemble_lumped_matrix(g, data).diagonal()
        d0 = pg.PwConstants("key").assemble_mass_matrix(g, data)
        store["p0_mass_diag"] = d0.diagonal()
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **store)
        print(f"{name}: cells {sd.num_cells}, wrote {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
