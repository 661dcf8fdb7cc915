"""Broken (element-wise) polynomial spaces and the inclusions into them.

Reference: ``discretizations/fem/l2.py`` (``PwPolynomials``, ``PwLinears``), ``fem/vec_l2.py``
(``VecPwConstants``, ``VecPwLinears``), ``discretizations/poly_projection.py`` and the
``proj_to_PwPolynomials`` methods of the conforming spaces (``fem/h1.py:218-237``, ``fem/hdiv.py:255-274,
450-509``, ``fem/hcurl.py:44-61,215-291``).  The reference uses these inclusions ``Pi`` to *assemble*
(``Pi.T @ M @ Pi``); here assembly runs through the per-cell CUDA kernels, and the inclusions serve what
surrounds the hot path: ``eval_at_cell_centers``, ``error_l2``, ``assemble_broken_grad_matrix`` and
``RT1``'s range space.  They are O(cells) sparse matrices built from per-cell tables (one row of local
quantities per cell, no sparse fancy indexing); the broken-P1 mass matrices come from the K1/K2 engine
(space ``pwlinears``).

Dof numbering (``docs/reports_src/dofs.md``): ``PwLinears`` dof of (slot k, cell c) = ``k * nc + c`` with
the slots in ascending node order of the cell; the vector spaces stack ``dim`` scalar copies.
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import scipy.sparse as sps

from . import engine
from .constants import AMBIENT_DIM, MATRIX, SCALAR, SECOND_ORDER_TENSOR, UNITARY_DATA, VECTOR, WEIGHT
from .discretization import Discretization, PwConstants, _eval_points
from .params import coefficient


# --------------------------------------------------------------------------------------------------
# per-cell tables
# --------------------------------------------------------------------------------------------------
def cell_tables(sd) -> dict:
    """faces / signs / opposite node per face slot, ascending cell nodes, slot of every face node."""
    d, nc, nf = sd.dim, sd.num_cells, sd.num_faces
    cf, fn = sd.cell_faces, sd.face_nodes
    if cf.indices.size != nc * (d + 1) or fn.indices.size != nf * d:
        raise NotImplementedError("Grid is not simplicial; cannot compute opposite node.")
    faces = cf.indices.reshape(nc, d + 1).astype(np.int64)
    signs = np.asarray(cf.data, dtype=float).reshape(nc, d + 1)
    fnodes = fn.indices.reshape(nf, d).astype(np.int64)[faces]  # (nc, d+1, d)
    flat = np.sort(fnodes.reshape(nc, -1), axis=1)
    cnodes = flat[:, ::d]
    opp = cnodes.sum(axis=1)[:, None] - fnodes.sum(axis=2)
    return {"faces": faces, "signs": signs, "fnodes": fnodes, "cnodes": cnodes, "opp": opp}


def slot_of(cnodes: np.ndarray, nodes: np.ndarray) -> np.ndarray:
    """Position of ``nodes[c, ...]`` inside the ascending ``cnodes[c]``."""
    shape = nodes.shape
    nodes = nodes.reshape(shape[0], -1)
    return (nodes[:, :, None] > cnodes[:, None, :]).sum(axis=2).reshape(shape)


def plane_coords(sd) -> np.ndarray:
    """(dim, nn): node coordinates in the grid's own plane, ``rotation_matrix @ nodes``."""
    R = np.asarray(getattr(sd, "rotation_matrix", np.eye(sd.dim, AMBIENT_DIM)), dtype=float)
    return R @ np.asarray(sd.nodes, dtype=float)


def barycentric_gradients(sd, cnodes: np.ndarray):
    """(|c| (nc,), grad lambda (nc, d+1, d)) from the coordinates of the cell's nodes (ascending order)."""
    d = sd.dim
    X = plane_coords(sd)[:, cnodes]  # (d, nc, d+1)
    E = np.transpose(X[:, :, 1:] - X[:, :, :1], (1, 2, 0))  # (nc, d, d): rows = edge vectors from node 0
    vol = np.abs(np.linalg.det(E)) / float(np.prod(np.arange(1, d + 1)))
    Ginv = np.linalg.inv(E)  # columns = gradients of lambda_1..lambda_d
    G = np.empty((cnodes.shape[0], d + 1, d))
    G[:, 1:, :] = np.transpose(Ginv, (0, 2, 1))
    G[:, 0, :] = -G[:, 1:, :].sum(axis=1)
    return vol, G


# --------------------------------------------------------------------------------------------------
# scalar broken spaces
# --------------------------------------------------------------------------------------------------
class PwPolynomials(Discretization):
    """``fem/l2.py:12-282``."""

    tensor_order = SCALAR

    def ndof_per_cell(self, sd) -> int:
        raise NotImplementedError

    def ndof(self, sd) -> int:
        return sd.num_cells * self.ndof_per_cell(sd)

    def local_dofs_of_cell(self, sd, c: int) -> np.ndarray:
        return sd.num_cells * np.arange(self.ndof_per_cell(sd)) + c

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        return sps.csc_array((0, self.ndof(sd)))

    def assemble_stiff_matrix(self, sd, _data=None) -> sps.csc_array:
        n = self.ndof(sd)
        return sps.csc_array((n, n))

    def assemble_nat_bc(self, sd, _func, _b_faces) -> np.ndarray:
        return np.zeros(self.ndof(sd))

    def get_range_discr_class(self, dim: int):
        raise NotImplementedError("There's no zero discretization in PyGeoN (yet)")

    def proj_to_PwPolynomials(self, sd) -> sps.csc_array:
        return sps.eye_array(self.ndof(sd)).tocsc()


class PwLinears(PwPolynomials):
    """``fem/l2.py:435-643``: d+1 dofs per cell (the nodal values, slots in ascending node order)."""

    poly_order = 1
    _space = "pwlinears"

    def ndof_per_cell(self, sd) -> int:
        return sd.dim + 1

    @staticmethod
    def assemble_local_mass(dim: int) -> np.ndarray:
        return (np.ones((dim + 1, dim + 1)) + np.eye(dim + 1)) / ((dim + 1) * (dim + 2))

    @staticmethod
    def assemble_local_lumped_mass(dim: int) -> np.ndarray:
        return np.eye(dim + 1) / (dim + 1)

    # mass / lumped mass: K1 kinds l1_mass / l1_lumped over the dof table (k * nc + c) -----------------
    def assemble_mass_matrix_device(self, sd, data: dict | None = None) -> engine.DeviceCSR:
        return engine.assemble_device("pwlinears", "mass", sd, coefficient(sd, data, self.keyword, WEIGHT), None)

    def assemble_mass_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        return engine.assemble_host("pwlinears", "mass", sd, coefficient(sd, data, self.keyword, WEIGHT), None)

    def assemble_lumped_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        A = engine.assemble_device("pwlinears", "lumped", sd, coefficient(sd, data, self.keyword, WEIGHT), None)
        return sps.diags_array(A.diagonal().cpu().numpy()).tocsc()

    # operators -------------------------------------------------------------------------------------------
    def eval_at_cell_centers(self, sd) -> sps.csc_array:
        nc, m = sd.num_cells, sd.dim + 1
        rows = np.tile(np.arange(nc), m)
        return sps.csc_array((np.full(nc * m, 1.0 / m), (rows, np.arange(nc * m))), shape=(nc, nc * m))

    def proj_to_lower_PwPolynomials(self, sd) -> sps.csc_array:
        """Cell integrals (the P0 dofs) of a broken-P1 function (``l2.py:529-547``)."""
        nc, m = sd.num_cells, sd.dim + 1
        rows = np.tile(np.arange(nc), m)
        vals = np.tile(np.asarray(sd.cell_volumes) / m, m)
        return sps.csc_array((vals, (rows, np.arange(nc * m))), shape=(nc, nc * m))

    def proj_to_higher_PwPolynomials(self, sd):
        raise NotImplementedError("PwQuadratics is not part of this package")

    def get_dof_lookup_array(self, sd) -> sps.csc_array:
        """``L[node, cell] = dof`` (``l2.py:598-612``)."""
        nc, m = sd.num_cells, sd.dim + 1
        cn = cell_tables(sd)["cnodes"]
        dofs = (np.arange(m)[None, :] * nc + np.arange(nc)[:, None]).ravel()
        return sps.csc_array((dofs, cn.ravel(), np.arange(0, nc * m + 1, m)), shape=(sd.num_nodes, nc))

    def interpolate(self, sd, func: Callable) -> np.ndarray:
        """Values at the d+1 Gauss points ``alpha x_node + (1 - alpha) x_cell``, alpha = 1/sqrt(d+2),
        extrapolated to the nodes (``l2.py:487-527``)."""
        nc, m = sd.num_cells, sd.dim + 1
        cn = cell_tables(sd)["cnodes"]
        alpha = 1.0 / np.sqrt(sd.dim + 2)
        centers = np.asarray(sd.cell_centers)
        pts = alpha * np.asarray(sd.nodes)[:, cn.T.ravel()] + (1 - alpha) * np.tile(centers, m)  # slot-major
        vals = np.array(_eval_points(func, pts), dtype=float).reshape(m, nc)
        mean = vals.mean(axis=0)
        return (mean[None, :] + (vals - mean[None, :]) / alpha).ravel()

    def assemble_broken_grad_matrix(self, sd) -> sps.csc_array:
        """Element-wise gradient into the vector piecewise constants (``l2.py:614-643``); rows
        ``k * nc + c`` hold component k (in the grid's plane) of the gradient on cell c, as a
        PwConstants dof, i.e. integrated over the cell: ``|c| grad lambda``."""
        d, nc = sd.dim, sd.num_cells
        cn = cell_tables(sd)["cnodes"]
        vol, G = barycentric_gradients(sd, cn)  # (nc, d+1, d)
        G = G * vol[:, None, None]
        rows = (np.arange(d)[None, None, :] * nc + np.arange(nc)[:, None, None]) + np.zeros((1, d + 1, 1), dtype=np.int64)
        cols = (np.arange(d + 1)[None, :, None] * nc + np.arange(nc)[:, None, None]) + np.zeros((1, 1, d), dtype=np.int64)
        return sps.csc_array((G.ravel(), (rows.ravel(), cols.ravel())), shape=(d * nc, (d + 1) * nc))

    def error_l2(self, sd, num_sol, ana_sol, relative=True, poly_order=None, data=None) -> float:
        return _interp_error(self, sd, num_sol, ana_sol, relative, data)


def _interp_error(discr, sd, num_sol, ana_sol, relative, data) -> float:
    """``sqrt((u - I f)^T M (u - I f) / (I f)^T M (I f))`` (``discretization.py:330-338``)."""
    interp = discr.interpolate(sd, ana_sol)
    M = discr.assemble_mass_matrix(sd, data)
    diff = np.asarray(num_sol) - interp
    norm = interp @ M @ interp if relative else 1.0
    return float(np.sqrt(diff @ M @ diff / norm))


# --------------------------------------------------------------------------------------------------
# vector broken spaces: dim stacked copies of the scalar space (components in the grid's plane)
# --------------------------------------------------------------------------------------------------
class VecPwPolynomials(Discretization):
    """``discretizations/vec_discretization.py`` + ``fem/vec_l2.py:12-285``."""

    tensor_order = VECTOR
    base_cls: type = PwLinears

    def __init__(self, keyword: str = UNITARY_DATA) -> None:
        super().__init__(keyword)
        self.base_discr = self.base_cls(keyword)

    @staticmethod
    def vectorize(dim: int, matrix) -> sps.csc_array:
        return sps.kron(sps.eye_array(dim), matrix).tocsc()

    def ndof(self, sd) -> int:
        return sd.dim * self.base_discr.ndof(sd)

    def ndof_per_cell(self, sd) -> int:
        return sd.dim * self.base_discr.ndof_per_cell(sd)

    def assemble_weighting_matrix(self, sd, K) -> sps.csc_array:
        """Block matrix of the diagonals of ``R K R^T`` acting on the stacked components
        (``vec_l2.py:97-126``); K: (3, 3, nc) or None (identity)."""
        d, nb = sd.dim, self.base_discr.ndof_per_cell(sd)
        if K is None:
            return sps.eye_array(self.ndof(sd)).tocsc()
        R = np.asarray(getattr(sd, "rotation_matrix", np.eye(d, AMBIENT_DIM)), dtype=float)
        Kr = np.einsum("ia,abc,jb->jic", R, K, R)  # the reference's contraction yields R K^T R^T (vec_l2.py:113-114)
        blocks = [[sps.diags_array(np.tile(Kr[i, j], nb)) for j in range(d)] for i in range(d)]
        return sps.block_array(blocks).tocsc()

    def assemble_mass_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        M0 = self.vectorize(sd.dim, self.base_discr.assemble_mass_matrix(sd, data))
        K = coefficient(sd, data, self.keyword, SECOND_ORDER_TENSOR)
        return (M0 @ self.assemble_weighting_matrix(sd, K)).tocsc()

    def assemble_lumped_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        M0 = self.vectorize(sd.dim, self.base_discr.assemble_lumped_matrix(sd, data))
        K = coefficient(sd, data, self.keyword, SECOND_ORDER_TENSOR)
        return (M0 @ self.assemble_weighting_matrix(sd, K)).tocsc()

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        return sps.csc_array((0, self.ndof(sd)))

    def assemble_stiff_matrix(self, sd, _data=None) -> sps.csc_array:
        n = self.ndof(sd)
        return sps.csc_array((n, n))

    def assemble_nat_bc(self, sd, _func, _b_faces) -> np.ndarray:
        return np.zeros(self.ndof(sd))

    def get_range_discr_class(self, dim: int):
        return self.base_discr.get_range_discr_class(dim)

    def proj_to_PwPolynomials(self, sd) -> sps.csc_array:
        return sps.eye_array(self.ndof(sd)).tocsc()

    def proj_to_lower_PwPolynomials(self, sd) -> sps.csc_array:
        return self.vectorize(sd.dim, self.base_discr.proj_to_lower_PwPolynomials(sd))

    def proj_to_higher_PwPolynomials(self, sd) -> sps.csc_array:
        return self.vectorize(sd.dim, self.base_discr.proj_to_higher_PwPolynomials(sd))

    def interpolate(self, sd, func: Callable) -> np.ndarray:
        R = np.asarray(getattr(sd, "rotation_matrix", np.eye(sd.dim, AMBIENT_DIM)), dtype=float)
        return np.hstack([self.base_discr.interpolate(sd, lambda x, k=k: (R @ np.asarray(func(x)).ravel())[k])
                          for k in range(sd.dim)])

    def eval_at_cell_centers(self, sd) -> sps.csc_array:
        """Cell-centre values as ambient 3-vectors: (3 nc) x ndof (``vec_l2.py:247-259``)."""
        P = self.vectorize(sd.dim, self.base_discr.eval_at_cell_centers(sd))
        R = np.asarray(getattr(sd, "rotation_matrix", np.eye(sd.dim, AMBIENT_DIM)), dtype=float)
        return (sps.kron(R.T, sps.eye_array(sd.num_cells), "csc") @ P).tocsc()

    def assemble_broken_grad_matrix(self, sd) -> sps.csc_array:
        return self.vectorize(sd.dim, self.base_discr.assemble_broken_grad_matrix(sd))

    def error_l2(self, sd, num_sol, ana_sol, relative=True, poly_order=None, data=None) -> float:
        return _interp_error(self, sd, num_sol, ana_sol, relative, data)


class VecPwLinears(VecPwPolynomials):
    poly_order = 1
    base_cls = PwLinears


class _PwConstantsBase(PwConstants):
    """PwConstants with the broken-space methods the vector wrapper needs."""

    def ndof_per_cell(self, sd) -> int:
        return 1


class VecPwConstants(VecPwPolynomials):
    poly_order = 0
    base_cls = _PwConstantsBase


def get_PwPolynomials(poly_order: int, tensor_order: int):
    """``poly_projection.py:8-42`` restricted to the spaces of this package."""
    table = {(0, SCALAR): PwConstants, (1, SCALAR): PwLinears, (0, VECTOR): VecPwConstants, (1, VECTOR): VecPwLinears}
    if (poly_order, tensor_order) not in table:
        if tensor_order in (SCALAR, VECTOR, MATRIX) and poly_order in (0, 1, 2):
            raise NotImplementedError(f"piecewise polynomials of order {poly_order}, tensor order {tensor_order} "
                                      "are not part of this package")
        raise KeyError("Unsupported polynomial order {:} and tensor order {:}.".format(poly_order, tensor_order))
    return table[(poly_order, tensor_order)]


def proj_to_PwPolynomials(discr, sd, poly_order: int) -> sps.csc_array:
    """Inclusion of ``discr`` into the broken space of order ``poly_order`` by raising / lowering one
    order at a time (``poly_projection.py:45-79``)."""
    pi = discr.proj_to_PwPolynomials(sd)
    cur = discr.poly_order
    while cur < poly_order:
        pi = get_PwPolynomials(cur, discr.tensor_order)(discr.keyword).proj_to_higher_PwPolynomials(sd) @ pi
        cur += 1
    while cur > poly_order:
        pi = get_PwPolynomials(cur, discr.tensor_order)(discr.keyword).proj_to_lower_PwPolynomials(sd) @ pi
        cur -= 1
    return pi


# --------------------------------------------------------------------------------------------------
# inclusions of the conforming spaces into the broken P1 spaces
# --------------------------------------------------------------------------------------------------
def lagrange1_to_pwlinears(sd) -> sps.csc_array:
    """Ones at (dof of (slot, cell), node) (``fem/h1.py:218-237``)."""
    nc, m = sd.num_cells, sd.dim + 1
    cn = cell_tables(sd)["cnodes"]
    rows = (np.arange(m)[None, :] * nc + np.arange(nc)[:, None]).ravel()
    return sps.csc_array((np.ones(nc * m), (rows, cn.ravel())), shape=(nc * m, sd.num_nodes))


def _vec_rows(sd, slots: np.ndarray, cells: np.ndarray) -> np.ndarray:
    """Rows of the d components of the VecPwLinears dof at (slot, cell): (d, ...)."""
    nc, m, d = sd.num_cells, sd.dim + 1, sd.dim
    base = slots * nc + cells
    return base[None, ...] + (np.arange(d) * m * nc).reshape((d,) + (1,) * base.ndim)


def bdm1_to_vecpwlinears(sd) -> sps.csc_array:
    """BDM1 function of (face a, its node j) = lambda_j (x_j - x_a) / (s_a d |c|) on every cell of the
    face (``fem/hdiv.py:450-509``)."""
    d, nc, nf = sd.dim, sd.num_cells, sd.num_faces
    t = cell_tables(sd)
    X = plane_coords(sd)
    vol = np.asarray(sd.cell_volumes, dtype=float)
    tang = X[:, t["fnodes"]] - X[:, t["opp"]][:, :, :, None]  # (d, nc, d+1, d)
    vals = tang / (t["signs"] * vol[:, None] * d)[None, :, :, None]
    cols = t["faces"][:, :, None] + np.arange(d)[None, None, :] * nf  # pos = order in face_nodes
    cells = np.broadcast_to(np.arange(nc)[:, None, None], cols.shape)
    rows = _vec_rows(sd, slot_of(t["cnodes"], t["fnodes"]), cells)
    cols = np.broadcast_to(cols[None], rows.shape)
    return sps.csc_array((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(d * (d + 1) * nc, d * nf))


def rt0_to_vecpwlinears(sd) -> sps.csc_array:
    """``BDM1 inclusion @ vstack([I] * d)`` (``fem/hdiv.py:255-274,338-351``)."""
    nf = sd.num_faces
    return (bdm1_to_vecpwlinears(sd) @ sps.vstack([sps.eye_array(nf)] * sd.dim)).tocsc()


def cell_edges(sd, cnodes: np.ndarray):
    """(edge ids (nc, E), local node pairs (E, 2)) with E = d(d+1)/2; edge (p, q) is the global edge
    whose nodes are cnodes[:, p], cnodes[:, q]."""
    d = sd.dim
    pairs = np.array([(p, q) for p in range(d + 1) for q in range(p + 1, d + 1)])
    en = (sd.ridge_peaks if d == 3 else sd.face_ridges).indices.reshape(sd.num_edges, 2).astype(np.int64)
    key = np.minimum(en[:, 0], en[:, 1]) * sd.num_nodes + np.maximum(en[:, 0], en[:, 1])
    order = np.argsort(key)
    want = cnodes[:, pairs[:, 0]] * sd.num_nodes + cnodes[:, pairs[:, 1]]
    pos = np.searchsorted(key[order], want)
    if np.any(key[order][np.minimum(pos, key.size - 1)] != want):
        raise ValueError("a cell edge is missing from the grid's edge list")
    return order[pos], pairs, en


def nedelec0_to_vecpwlinears(sd) -> sps.csc_array:
    """Whitney form of edge e = (n0, n1) (the order stored in ``ridge_peaks`` / ``face_ridges``):
    ``lambda_n0 grad lambda_n1 - lambda_n1 grad lambda_n0`` (``fem/hcurl.py:44-61,166-180,215-291``)."""
    d, nc = sd.dim, sd.num_cells
    cn = cell_tables(sd)["cnodes"]
    _, G = barycentric_gradients(sd, cn)  # (nc, d+1, d)
    edges, pairs, en = cell_edges(sd, cn)
    E = pairs.shape[0]
    # is the stored first node of the edge the lower local node p?
    first_is_p = en[edges, 0] == cn[:, pairs[:, 0]]  # (nc, E)
    sgn = np.where(first_is_p, 1.0, -1.0)
    cells = np.broadcast_to(np.arange(nc)[:, None], (nc, E))
    rows, cols, vals = [], [], []
    for at, other, s in ((0, 1, 1.0), (1, 0, -1.0)):  # value at local node `at` is +-grad lambda_other
        slots = np.broadcast_to(pairs[None, :, at], (nc, E))
        r = _vec_rows(sd, slots, cells)  # (d, nc, E)
        g = np.transpose(G[:, pairs[:, other], :], (2, 0, 1))  # (d, nc, E)
        rows.append(r.ravel())
        cols.append(np.broadcast_to(edges[None], r.shape).ravel())
        vals.append((s * sgn[None] * g).ravel())
    return sps.csc_array((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(d * (d + 1) * nc, sd.num_edges))
