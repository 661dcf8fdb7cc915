"""``pg.Discretization`` operator API on top of the CUDA engine.

Same class names, constructor and method signatures as the reference
(``src/pygeon/discretizations/discretization.py:14-358`` and the ``fem/`` classes named in each
docstring), restricted to the rows of the hot-path scope table: mass and
stiffness matrices of ``Lagrange1``, ``RT0``, ``BDM1``, ``Nedelec0`` and ``PwConstants`` on 2D
triangle / 3D tetrahedral grids, plus the differential (incidence) matrices.

Every ``assemble_*`` returns a ``scipy.sparse.csc_array`` like the reference; the ``*_device``
variants return a :class:`~pygeon_b200.engine.DeviceCSR` that stays in HBM for the GPU solvers.
Matrices carry the *structural* pattern (all pairs of dofs sharing a cell, explicit zeros kept,
sorted indices); the reference's scipy SpGEMM drops entries whose floating-point sum is exactly
0.0, which is value dependent.
"""

from __future__ import annotations

import abc
from typing import Type

import numpy as np
import scipy.sparse as sps

from . import engine
from .constants import SCALAR, SECOND_ORDER_TENSOR, UNITARY_DATA, VECTOR, WEIGHT
from .params import coefficient


class Discretization(abc.ABC):
    """Abstract base (``discretization.py:14-58``)."""

    poly_order: int
    tensor_order: int
    _space: str

    def __init__(self, keyword: str = UNITARY_DATA) -> None:
        self.keyword = keyword

    def __repr__(self) -> str:
        s = f"Discretization of type {self.__class__.__name__}"
        if hasattr(self, "keyword"):
            s += f" with keyword {self.keyword}"
        return s

    @abc.abstractmethod
    def ndof(self, sd) -> int: ...

    # ---- coefficients ------------------------------------------------------------------
    def _coefficients(self, sd, data):
        w = coefficient(sd, data, self.keyword, WEIGHT)
        K = coefficient(sd, data, self.keyword, SECOND_ORDER_TENSOR)
        return w, K

    # ---- mass --------------------------------------------------------------------------
    def assemble_mass_matrix_device(self, sd, data: dict | None = None) -> engine.DeviceCSR:
        w, K = self._coefficients(sd, data)
        if self.tensor_order == SCALAR:
            K = None  # the scalar broken-P1 mass ignores the tensor (fem/l2.py:63-86)
        return engine.assemble_device(self._space, "mass", sd, w, K)

    def assemble_mass_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """``(Pi.T @ M @ Pi).tocsc()`` of ``discretization.py:65-89,112-136``."""
        if self.ndof(sd) == 0:
            return sps.csc_array((0, 0))
        w, K = self._coefficients(sd, data)
        if self.tensor_order == SCALAR:
            K = None
        return engine.assemble_host(self._space, "mass", sd, w, K)

    # ---- lumped mass -------------------------------------------------------------------
    _lumped_diagonal = True

    def assemble_lumped_matrix_device(self, sd, data: dict | None = None) -> engine.DeviceCSR:
        w, K = self._coefficients(sd, data)
        if self.tensor_order == SCALAR:
            K = None
        return engine.assemble_device(self._space, "lumped", sd, w, K)

    def assemble_lumped_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """Lumped mass (``discretization.py:91-110``; ``RT0`` override ``fem/hdiv.py:140-175``)."""
        if self.ndof(sd) == 0:
            return sps.csc_array((0, 0))
        if self._lumped_diagonal:
            A = self.assemble_lumped_matrix_device(sd, data)
            return sps.diags_array(A.diagonal().cpu().numpy()).tocsc()
        w, K = self._coefficients(sd, data)
        return engine.assemble_host(self._space, "lumped", sd, w, K)

    # ---- stiffness ---------------------------------------------------------------------
    def _stiff_form(self, sd) -> str:
        return "stiff"

    def assemble_stiff_matrix_device(self, sd, data: dict | None = None) -> engine.DeviceCSR:
        w, K = self._coefficients(sd, data)
        form = self._stiff_form(sd)
        if form in ("stiff2d",) or self._space in ("rt0", "bdm1"):
            K = None  # range space is PwConstants: only the weight enters (fem/l2.py:335-353)
        return engine.assemble_device(self._space, form, sd, w, K)

    def assemble_stiff_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """``(B.T @ A @ B).tocsc()`` of ``discretization.py:154-183``."""
        self.get_range_discr_class(sd.dim)  # raises like the reference when there is none
        w, K = self._coefficients(sd, data)
        form = self._stiff_form(sd)
        if form in ("stiff2d",) or self._space in ("rt0", "bdm1"):
            K = None
        return engine.assemble_host(self._space, form, sd, w, K)

    def source_term(self, sd, func) -> np.ndarray:
        """``assemble_mass_matrix(sd) @ interpolate(sd, func)`` (``discretization.py:221-235``)."""
        return self.assemble_mass_matrix(sd) @ self.interpolate(sd, func)

    # ---- inclusion into the broken polynomials and what is built on it ---------------------------
    def proj_to_PwPolynomials(self, sd) -> sps.csc_array:
        """Inclusion into the broken polynomial space of the same order (``discretization.py:270-282``)."""
        raise NotImplementedError(f"{type(self).__name__}: the broken space of order {self.poly_order} is not "
                                  "part of this package")

    def eval_at_cell_centers(self, sd) -> sps.csc_array:
        """Matrix evaluating a discrete function at the cell centres (``discretization.py:200-219``)."""
        from . import broken

        if self.ndof(sd) == 0:
            return sps.csc_array((3 ** self.tensor_order, 0))
        space = broken.get_PwPolynomials(self.poly_order, self.tensor_order)(self.keyword)
        return (space.eval_at_cell_centers(sd) @ self.proj_to_PwPolynomials(sd)).tocsc()

    def assemble_broken_grad_matrix(self, sd) -> sps.csc_array:
        """Element-wise gradient (``discretization.py:284-302``)."""
        from . import broken

        space = broken.get_PwPolynomials(self.poly_order, self.tensor_order)(self.keyword)
        return (space.assemble_broken_grad_matrix(sd) @ self.proj_to_PwPolynomials(sd)).tocsc()

    def error_l2(self, sd, num_sol, ana_sol, relative: bool = True, poly_order: int | None = -1,
                 data: dict | None = None) -> float:
        """L2 error against an analytical solution (``discretization.py:304-358``): by default both are
        compared in the broken polynomials of order max(poly_order of the space, 1); ``poly_order=None``
        compares with the interpolant of the space itself."""
        from . import broken

        if poly_order == -1:
            poly_order = max(self.poly_order, 1)
        if poly_order is None:
            return broken._interp_error(self, sd, num_sol, ana_sol, relative, data)
        space = broken.get_PwPolynomials(poly_order, self.tensor_order)(self.keyword)
        proj_sol = broken.proj_to_PwPolynomials(self, sd, poly_order) @ num_sol
        return space.error_l2(sd, proj_sol, ana_sol, relative, poly_order=None, data=data)

    def interpolate(self, sd, func) -> np.ndarray:
        raise NotImplementedError

    def assemble_nat_bc(self, sd, func, b_faces) -> np.ndarray:
        raise NotImplementedError

    @abc.abstractmethod
    def assemble_diff_matrix(self, sd) -> sps.csc_array: ...

    @abc.abstractmethod
    def get_range_discr_class(self, dim: int) -> Type["Discretization"]: ...


def _eval_points(func, pts: np.ndarray) -> list:
    """``[func(x) for x in pts.T]`` — the reference evaluates user callables point by point
    (e.g. ``fem/hdiv.py:208-212``); the results are returned as the reference would see them."""
    return [func(x) for x in pts.T]


def _boundary_faces(b_faces) -> np.ndarray:
    b_faces = np.asarray(b_faces)
    return np.where(b_faces)[0] if b_faces.dtype == bool else b_faces


class PwConstants(Discretization):
    """``fem/l2.py:284-353``: one dof per cell (the integral), mass = diag(w/|c|)."""

    poly_order = 0
    tensor_order = SCALAR
    _space = "p0"

    def ndof(self, sd) -> int:
        return sd.num_cells

    def assemble_mass_diagonal_device(self, sd, data: dict | None = None):
        w = coefficient(sd, data, self.keyword, WEIGHT)
        return engine.cell_volumes_inverse_weighted(sd, w)

    def assemble_mass_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        if self.ndof(sd) == 0:
            return sps.csc_array((0, 0))
        diag = self.assemble_mass_diagonal_device(sd, data).cpu().numpy()
        return sps.diags_array(diag).tocsc()

    def assemble_lumped_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        return self.assemble_mass_matrix(sd, data)

    def assemble_stiff_matrix(self, sd, data=None):
        raise NotImplementedError("There's no zero discretization in PyGeoN")

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        return sps.csc_array((0, self.ndof(sd)))

    def interpolate(self, sd, func) -> np.ndarray:
        """Cell integrals ``func(cell centre) * |c|`` (``fem/l2.py:390-406``)."""
        return np.array([f * vol for f, vol in zip(_eval_points(func, sd.cell_centers), sd.cell_volumes)])

    def assemble_nat_bc(self, sd, _func, _b_faces) -> np.ndarray:
        """Discontinuous functions have no boundary trace: zero (``fem/l2.py:169-190``)."""
        return np.zeros(self.ndof(sd))

    def get_range_discr_class(self, dim: int):
        raise NotImplementedError("There's no zero discretization in PyGeoN")

    def ndof_per_cell(self, sd) -> int:
        return 1

    def proj_to_PwPolynomials(self, sd) -> sps.csc_array:
        return sps.eye_array(self.ndof(sd)).tocsc()

    def eval_at_cell_centers(self, sd) -> sps.csc_array:
        """The dof is the cell integral: value = dof / |c| (``fem/l2.py:422-432``)."""
        return sps.diags_array(1.0 / np.asarray(sd.cell_volumes)).tocsc()

    def proj_to_higher_PwPolynomials(self, sd) -> sps.csc_array:
        """Into the broken P1: the cell value at every node slot (``fem/l2.py:408-420``)."""
        return sps.vstack([self.eval_at_cell_centers(sd)] * (sd.dim + 1)).tocsc()

    def assemble_broken_grad_matrix(self, sd) -> sps.csc_array:
        return sps.csc_array((3 * self.ndof(sd), self.ndof(sd)))


class Lagrange1(Discretization):
    """``fem/h1.py:13-324``: nodal P1."""

    poly_order = 1
    tensor_order = SCALAR
    _space = "lagrange1"

    def ndof(self, sd) -> int:
        return sd.num_nodes

    def _stiff_form(self, sd) -> str:
        # 2D: B^T M_RT0 B = (K rot grad u, rot grad v), see fem/h1.py:46-48
        return "stiff_rot" if sd.dim == 2 else "stiff"

    def assemble_grad_grad_matrix_device(self, sd, data: dict | None = None) -> engine.DeviceCSR:
        w, K = self._coefficients(sd, data)
        return engine.assemble_device(self._space, "stiff", sd, w, K)

    def assemble_grad_grad_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """``(K grad u, grad v)`` (``fem/h1.py:39-63``)."""
        w, K = self._coefficients(sd, data)
        return engine.assemble_host(self._space, "stiff", sd, w, K)

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        """``fem/h1.py:152-179``."""
        match sd.dim:
            case 3:
                return sd.ridge_peaks.T.tocsc()
            case 2:
                return sd.face_ridges.T.tocsc()
            case 1:
                return sd.cell_faces.T.tocsc()
            case 0:
                return sps.csc_array((0, 0))
            case _:
                raise ValueError("Dimension must be 0, 1, 2, or 3.")

    def interpolate(self, sd, func) -> np.ndarray:
        """Nodal values (``fem/h1.py:255-269``)."""
        return np.array(_eval_points(func, sd.nodes))

    def proj_to_PwPolynomials(self, sd) -> sps.csc_array:
        """``fem/h1.py:218-237``."""
        from . import broken

        return broken.lagrange1_to_pwlinears(sd)

    def eval_at_cell_centers(self, sd) -> sps.csc_array:
        """Average of the nodal values of every cell (``fem/h1.py:239-253``)."""
        from . import broken

        return (broken.PwLinears().eval_at_cell_centers(sd) @ broken.lagrange1_to_pwlinears(sd)).tocsc()

    def assemble_nat_bc(self, sd, func, b_faces) -> np.ndarray:
        """``(v, g)`` on the boundary faces by the midpoint value spread over the face's nodes
        (``fem/h1.py:271-302``); accumulated face by face in the order of ``b_faces``."""
        b_faces = _boundary_faces(b_faces)
        vals = np.zeros(self.ndof(sd))
        if b_faces.size == 0:
            return vals
        fn = sd.face_nodes
        g = np.array(_eval_points(func, sd.face_centers[:, b_faces]), dtype=float).reshape(b_faces.size)
        cnt = fn.indptr[b_faces + 1] - fn.indptr[b_faces]
        contrib = np.repeat(g * sd.face_areas[b_faces] / cnt, cnt)
        starts = np.repeat(fn.indptr[b_faces], cnt)
        within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        np.add.at(vals, fn.indices[starts + within], contrib)
        return vals

    def get_range_discr_class(self, dim: int):
        match dim:
            case 3:
                return Nedelec0
            case 2:
                return RT0
            case 1:
                return PwConstants
            case _:
                raise NotImplementedError("There's no zero discretization in PyGeoN")


class Lagrange2(Discretization):
    """``fem/h1.py:326-800``: quadratic Lagrange elements, one dof per node and per edge
    (edge dofs numbered ``num_nodes + edge id``).  The reference assembles the stiffness matrix in a
    Python loop over the cells (``fem/h1.py:545-599``); here it goes through the same K1/K2 engine
    as the lowest-order spaces (kinds ``PGB_L2_MASS`` / ``PGB_L2_STIFF``)."""

    poly_order = 2
    tensor_order = SCALAR
    _space = "lagrange2"

    def ndof(self, sd) -> int:
        return sd.num_nodes + sd.num_edges

    def assemble_stiff_matrix_device(self, sd, data: dict | None = None) -> engine.DeviceCSR:
        _, K = self._coefficients(sd, data)
        return engine.assemble_device(self._space, "stiff", sd, None, K)  # pg.WEIGHT is not read (h1.py:560-562)

    def assemble_stiff_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """``(K grad u, grad v)`` (``fem/h1.py:545-599``)."""
        _, K = self._coefficients(sd, data)
        return engine.assemble_host(self._space, "stiff", sd, None, K)

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        """Gradient into Nedelec1 (two dofs per edge), ``fem/h1.py:601-674`` (2D and 3D)."""
        match sd.dim:
            case 2:
                edge_nodes, num_edges, second = sd.face_ridges, sd.num_faces, 1
            case 3:
                edge_nodes, num_edges, second = sd.ridge_peaks, sd.num_ridges, -1
            case _:
                raise ValueError("Dimension must be 2 or 3.")
        edge_nodes = sps.csc_array(edge_nodes).astype(float)
        d0 = edge_nodes.copy().T
        d0.data[edge_nodes.data == -1] = -3
        d0.data[edge_nodes.data == 1] = -1
        diff_0 = sps.hstack((d0, 4 * sps.eye_array(num_edges)))
        d1 = edge_nodes.copy().T
        d1.data[edge_nodes.data == 1] = 3
        d1.data[edge_nodes.data == -1] = 1
        diff_1 = second * sps.hstack((d1, -4 * sps.eye_array(num_edges)))
        return sps.vstack((diff_0, diff_1)).tocsc()

    def interpolate(self, sd, func) -> np.ndarray:
        """Values at the nodes, then at the edge midpoints (``fem/h1.py:722-753``)."""
        if sd.dim == 2:
            edge_coords = sd.face_centers
        else:
            edge_coords = sd.nodes @ abs(sd.ridge_peaks.astype(float)) / 2
        return np.array(_eval_points(func, np.hstack((sd.nodes, edge_coords))))

    def get_range_discr_class(self, dim: int):
        raise NotImplementedError("Nedelec1 (the range of Lagrange2) is not part of this package")


class RT0(Discretization):
    """``fem/hdiv.py:13-274``: one dof per face."""

    poly_order = 1
    tensor_order = VECTOR
    _space = "rt0"

    def ndof(self, sd) -> int:
        return sd.num_faces

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        """``fem/hdiv.py:177-192``."""
        return sps.csc_array(sd.cell_faces.T)

    def interpolate(self, sd, func) -> np.ndarray:
        """Normal fluxes ``func(face centre) . n_f`` (``fem/hdiv.py:194-212``)."""
        f = _eval_points(func, sd.face_centers)
        return np.array([np.inner(np.asarray(v).flatten(), n) for v, n in zip(f, sd.face_normals.T)])

    def assemble_nat_bc(self, sd, func, b_faces) -> np.ndarray:
        """``(u . n, g)`` on the boundary: ``g(face centre)`` times the orientation of the face
        w.r.t. its only cell (``fem/hdiv.py:214-241``)."""
        b_faces = _boundary_faces(b_faces)
        vals = np.zeros(self.ndof(sd))
        if b_faces.size == 0:
            return vals
        sign = np.asarray(sd.cell_faces.tocsr()[b_faces, :].sum(axis=1)).ravel()
        g = np.array(_eval_points(func, sd.face_centers[:, b_faces]), dtype=float).reshape(b_faces.size)
        vals[b_faces] = g * sign
        return vals

    def get_range_discr_class(self, _dim: int):
        return PwConstants

    def proj_to_PwPolynomials(self, sd) -> sps.csc_array:
        """``fem/hdiv.py:255-274``."""
        from . import broken

        return broken.rt0_to_vecpwlinears(sd)


class BDM1(Discretization):
    """``fem/hdiv.py:277-509``: d dofs per face, dof (f, pos) -> f + pos*num_faces."""

    poly_order = 1
    tensor_order = VECTOR
    _space = "bdm1"

    _lumped_diagonal = False  # couples the d dofs that sit at the same node of a cell

    def ndof(self, sd) -> int:
        return sd.face_nodes.nnz

    def proj_to_RT0(self, sd) -> sps.csc_array:
        return (sps.hstack([sps.eye_array(sd.num_faces)] * sd.dim) / sd.dim).tocsc()

    def proj_from_RT0(self, sd) -> sps.csc_array:
        return sps.vstack([sps.eye_array(sd.num_faces)] * sd.dim).tocsc()

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        """``fem/hdiv.py:353-371``."""
        return (sps.csc_array(sd.cell_faces.T) @ self.proj_to_RT0(sd)).tocsc()

    def interpolate(self, sd, func) -> np.ndarray:
        """``n_f . func(node)`` for the d nodes of every face (``fem/hdiv.py:373-395``)."""
        d, nf = sd.dim, sd.num_faces
        fn = sd.face_nodes.indices.reshape(nf, d)
        fvals = np.array([np.asarray(v).flatten() for v in _eval_points(func, sd.nodes)])  # (nn, 3)
        vals = np.einsum("if,fji->jf", sd.face_normals, fvals[fn])  # (d, nf): pos-major
        return vals.reshape(-1)

    def assemble_nat_bc(self, sd, func, b_faces) -> np.ndarray:
        """``(q . n, g)`` on the boundary with the P1 mass of the face (``fem/hdiv.py:397-436``)."""
        b_faces = _boundary_faces(b_faces)
        d, nf = sd.dim, sd.num_faces
        vals = np.zeros(self.ndof(sd))
        if b_faces.size == 0:
            return vals
        local_mass = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))  # PwLinears.assemble_local_mass(d - 1)
        signs = sd.cell_faces @ np.ones(sd.num_cells)
        fn = sd.face_nodes.indices.reshape(nf, d)[b_faces]  # (nb, d)
        g = np.array(_eval_points(func, sd.nodes[:, fn.ravel()]), dtype=float).reshape(b_faces.size, d)
        loc = signs[b_faces, None] * (g @ local_mass.T)  # (nb, d)
        vals[(b_faces[:, None] + np.arange(d)[None, :] * nf).ravel()] = loc.ravel()
        return vals

    def get_range_discr_class(self, _dim: int):
        return PwConstants

    def proj_to_PwPolynomials(self, sd) -> sps.csc_array:
        """``fem/hdiv.py:450-509``."""
        from . import broken

        return broken.bdm1_to_vecpwlinears(sd)


class RT1(Discretization):
    """``fem/hdiv.py:512-1028``: Raviart-Thomas elements of order 1, d dofs per face (numbered like
    BDM1: ``f + k * num_faces``) and d per cell (``d * num_faces + k * num_cells + c``).  The reference
    assembles the mass matrix in a Python loop over the cells (``fem/hdiv.py:555-621``); here it goes
    through the K1/K2 engine (kind ``PGB_RT1_MASS``, local 8 x 8 / 15 x 15 blocks).  The divergence,
    the lumped cell block and the interpolation are O(num_cells) vectorised host code."""

    poly_order = 2
    tensor_order = VECTOR
    _space = "rt1"
    _lumped_diagonal = False

    def ndof(self, sd) -> int:
        return sd.dim * (sd.num_faces + sd.num_cells)

    # ---- tables in the reference's local order (reorder_faces, fem/hdiv.py:629-674) -------------
    @staticmethod
    def _rank_tables(sd):
        """Per cell: node ids ascending ``(nc, d+1)``, and faces / signs listed by DESCENDING opposite
        node (``faces_loc``, ``signs_loc`` of the reference)."""
        d, nc = sd.dim, sd.num_cells
        opp = np.asarray(sd.compute_opposite_nodes().data).reshape(nc, d + 1)
        faces = sd.cell_faces.indices.reshape(nc, d + 1)
        signs = np.asarray(sd.cell_faces.data, dtype=float).reshape(nc, d + 1)
        sorter = np.argsort(opp, axis=1)[:, ::-1]
        take = lambda a: np.take_along_axis(a, sorter, axis=1)  # noqa: E731
        return np.sort(opp, axis=1), take(faces), take(signs)

    def _cell_dofs(self, sd, faces_loc):
        """``local_dofs_of_cell`` for all cells (``fem/hdiv.py:534-552``): (nc, d(d+2))."""
        d, nc, nf = sd.dim, sd.num_cells, sd.num_faces
        loc_face = np.tile(faces_loc, (1, d)) + np.repeat(np.arange(d), d + 1)[None, :] * nf
        loc_cell = d * nf + nc * np.arange(d)[None, :] + np.arange(nc)[:, None]
        return np.hstack((loc_face, loc_cell))

    # ---- mass ------------------------------------------------------------------------------------
    def assemble_mass_matrix_device(self, sd, data: dict | None = None) -> engine.DeviceCSR:
        _, K = self._coefficients(sd, data)
        return engine.assemble_device(self._space, "mass", sd, None, K)  # pg.WEIGHT is not read (hdiv.py:575-577)

    def assemble_mass_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """``(K^-1 u, v)`` (``fem/hdiv.py:555-621``)."""
        _, K = self._coefficients(sd, data)
        return engine.assemble_host(self._space, "mass", sd, None, K)

    # ---- divergence into PwLinears ---------------------------------------------------------------
    @staticmethod
    def compute_local_div_matrix(dim: int) -> np.ndarray:
        """Nodal values (at the d+1 cell nodes, rows) of the divergence of the local basis (columns),
        before the factor ``s / (d |c|)``: the face function at node i of the face opposite node j has
        divergence ``1 + (d+1)(lambda_i - lambda_j)``, the cell function of node k
        ``(d+1) lambda_k - 1`` (same values as ``fem/hdiv.py:867-892``)."""
        n = dim + 1
        q = np.arange(dim * n)
        i, j = q // dim, (dim * n - 1 - q) % n  # local dof q <-> (node i, opposite node j)
        r = np.arange(n)[:, None]
        face = 1.0 + n * ((r == i[None, :]).astype(float) - (r == j[None, :]).astype(float))
        cell = n * (r == np.arange(dim)[None, :]).astype(float) - 1.0
        return np.hstack((face, cell))

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        """Divergence RT1 -> PwLinears (``fem/hdiv.py:813-865``), all cells at once."""
        d, nc = sd.dim, sd.num_cells
        loc_div = self.compute_local_div_matrix(d)  # (d+1, L)
        _, faces_loc, signs_loc = self._rank_tables(sd)
        L = d * (d + 2)
        signs = np.ones((nc, L))
        signs[:, : d * (d + 1)] = np.tile(signs_loc, (1, d))
        div = loc_div[None, :, :] * (signs / (d * sd.cell_volumes[:, None]))[:, None, :]  # (nc, d+1, L)
        cols = np.broadcast_to(self._cell_dofs(sd, faces_loc)[:, None, :], div.shape)
        ran = nc * np.arange(d + 1)[None, :] + np.arange(nc)[:, None]  # PwLinears.local_dofs_of_cell (l2.py:38-49)
        rows = np.broadcast_to(ran[:, :, None], div.shape)
        return sps.csc_array((div.ravel(), (rows.ravel(), cols.ravel())))

    # ---- lumped mass (Egger & Radu 2020; fem/hdiv.py:970-1028) --------------------------------------
    def _basis_at_center(self, sd, nodes_loc):
        """``eval_basis_functions_at_center`` for all cells: (nc, 3, d)."""
        d = sd.dim
        X = sd.nodes[:, nodes_loc]  # (3, nc, d+1)
        basis = (X.sum(axis=2)[:, :, None] - (d + 1) * X[:, :, :d]) / ((d + 1) ** 2)
        return np.transpose(basis, (1, 0, 2)) / (d * sd.cell_volumes)[:, None, None]

    def assemble_lumped_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """Egger & Radu lumping (``fem/hdiv.py:970-1028``): the BDM1 lumped block / (d+2) on the face dofs and
        ``|c| (d+1)/(d+2) P^T K P`` on the d cell dofs of every cell - both through K1/K2 (kinds
        ``PGB_BDM1_LUMPED`` and ``PGB_RT1_LUMPED_CELL``)."""
        d = sd.dim
        bdm1_lumped = BDM1(self.keyword).assemble_lumped_matrix(sd, data) / (d + 2)
        _, K = self._coefficients(sd, data)
        cell_block = engine.assemble_host("rt1cell", "lumped", sd, None, K)
        return sps.csc_array(sps.block_diag((bdm1_lumped, cell_block)))

    # ---- vectors -----------------------------------------------------------------------------------
    def interpolate(self, sd, func) -> np.ndarray:
        """Face dofs as BDM1, cell dofs from d x d systems at the cell centres (``fem/hdiv.py:894-934``)."""
        d, nc = sd.dim, sd.num_cells
        interp_faces = BDM1().interpolate(sd, func)
        nodes_loc = sd.cell_nodes().indices.reshape(nc, d + 1)
        P = self._basis_at_center(sd, nodes_loc)
        f = np.array([np.asarray(v, dtype=float).flatten() for v in _eval_points(func, sd.cell_centers)])  # (nc, 3)
        coeff = np.linalg.solve(P[:, :d, :], f[:, :d, None])[:, :, 0]  # (nc, d)
        return np.hstack((interp_faces, coeff.T.ravel()))

    def assemble_nat_bc(self, sd, func, b_faces) -> np.ndarray:
        """``fem/hdiv.py:936-957``: the BDM1 boundary term on the face dofs."""
        vals = np.zeros(self.ndof(sd))
        bdm1 = BDM1()
        vals[: bdm1.ndof(sd)] = bdm1.assemble_nat_bc(sd, func, b_faces)
        return vals

    def assemble_stiff_matrix_device(self, sd, data: dict | None = None) -> engine.DeviceCSR:
        w, _ = self._coefficients(sd, data)
        return engine.assemble_device(self._space, "stiff", sd, w, None)

    def assemble_stiff_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """``div^T M_PwLinears div`` (``discretization.py:154-183`` with ``hdiv.py:813-865,958`` and the
        broken-P1 mass ``l2.py:63-86``): K1 kind ``PGB_RT1_DIVDIV`` (only the weight enters)."""
        w, _ = self._coefficients(sd, data)
        return engine.assemble_host(self._space, "stiff", sd, w, None)

    def get_range_discr_class(self, _dim: int):
        from . import broken

        return broken.PwLinears


class Nedelec0(Discretization):
    """``fem/hcurl.py:11-180``: one dof per edge (ridge in 3D, face in 2D)."""

    poly_order = 1
    tensor_order = VECTOR
    _space = "nedelec0"

    def ndof(self, sd) -> int:
        return sd.num_edges

    def _stiff_form(self, sd) -> str:
        return "stiff2d" if sd.dim == 2 else "stiff"

    def assemble_diff_matrix(self, sd) -> sps.csc_array:
        """``fem/hcurl.py:82-106``."""
        match sd.dim:
            case 3:
                diff = sd.face_ridges.T
            case 2:
                diff = sd.cell_faces.T
            case _:
                diff = sps.csc_array((0, self.ndof(sd)))
        return diff.tocsc()

    def interpolate(self, sd, func) -> np.ndarray:
        """Tangential components at the edge midpoints (``fem/hcurl.py:146-164``)."""
        edge_nodes = sd.ridge_peaks if sd.dim == 3 else sd.face_ridges
        mid = sd.nodes @ abs(edge_nodes.astype(float)) / 2
        f = _eval_points(func, mid)
        return np.array([np.inner(np.asarray(v).flatten(), t) for v, t in zip(f, sd.edge_tangents.T)])

    def assemble_nat_bc(self, sd, func, b_faces) -> np.ndarray:
        raise NotImplementedError  # as the reference, fem/hcurl.py:108-126

    def proj_to_PwPolynomials(self, sd) -> sps.csc_array:
        """``fem/hcurl.py:44-61``."""
        from . import broken

        return broken.nedelec0_to_vecpwlinears(sd)

    def assemble_lumped_matrix(self, sd, data: dict | None = None) -> sps.csc_array:
        """Row sums of the mass matrix on the diagonal (``fem/hcurl.py:63-80``)."""
        diag_mass = self.assemble_mass_matrix(sd, data).sum(axis=0)
        return sps.diags_array(np.asarray(diag_mass).flatten()).tocsc()

    def get_range_discr_class(self, dim: int):
        match dim:
            case 2:
                return PwConstants
            case 3:
                return RT0
            case _:
                raise NotImplementedError
