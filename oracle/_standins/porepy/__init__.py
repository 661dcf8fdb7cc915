"""Minimal stand-in for the `porepy` surface PyGeoN's FEM hot path touches.

TEST INFRASTRUCTURE ONLY.  porepy is a third-party dependency of the reference
(pyproject.toml:26-28, pinned only to the moving `develop` branch) and is not
installed in this image.  None of the hot path's arithmetic lives in porepy: it
supplies the grid container (`pp.Grid`), `pp.SecondOrderTensor` and a few names.
This module provides exactly that surface so that the UNMODIFIED reference
source under /root/reference/src can be imported to (a) pin the numpy oracle
and (b) generate the golden fixtures under tests/golden/.

Never imported by the product package `pygeon_b200`.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps

PARAMETERS = "parameters"
DISCRETIZATION_MATRICES = "discretization_matrices"


def initialize_data(data: dict, keyword: str, specified_parameters: dict | None = None):
    """Same contract as porepy's helper: data[PARAMETERS][keyword] = params."""
    data.setdefault(PARAMETERS, {})
    data.setdefault(DISCRETIZATION_MATRICES, {})
    data[PARAMETERS].setdefault(keyword, {})
    data[PARAMETERS][keyword].update(specified_parameters or {})
    data[DISCRETIZATION_MATRICES].setdefault(keyword, {})
    return data


class SecondOrderTensor:
    """Symmetric 3x3 tensor per cell, `.values` of shape (3, 3, nc)."""

    def __init__(self, kxx, kyy=None, kzz=None, kxy=None, kxz=None, kyz=None):
        kxx = np.atleast_1d(np.asarray(kxx, dtype=float))
        nc = kxx.size
        z = np.zeros(nc)
        kyy = kxx if kyy is None else np.asarray(kyy, dtype=float)
        kzz = kxx if kzz is None else np.asarray(kzz, dtype=float)
        kxy = z if kxy is None else np.asarray(kxy, dtype=float)
        kxz = z if kxz is None else np.asarray(kxz, dtype=float)
        kyz = z if kyz is None else np.asarray(kyz, dtype=float)
        self.values = np.array(
            [[kxx, kxy, kxz], [kxy, kyy, kyz], [kxz, kyz, kzz]], dtype=float
        )

    def copy(self):
        out = SecondOrderTensor(np.ones(self.values.shape[2]))
        out.values = self.values.copy()
        return out


class Grid:
    """Simplicial-grid container with the attributes the hot path reads."""

    def __init__(self, dim, nodes, face_nodes, cell_faces, name="grid", **_):
        self.dim = int(dim)
        self.nodes = np.asarray(nodes, dtype=float)
        self.face_nodes = sps.csc_array(face_nodes)
        self.cell_faces = sps.csc_array(cell_faces)
        self.name = name
        self.num_nodes = self.nodes.shape[1]
        self.num_faces = self.face_nodes.shape[1]
        self.num_cells = self.cell_faces.shape[1]
        self.tags: dict = {}
        bnd = np.abs(self.cell_faces).sum(axis=1) == 1
        bnd = np.asarray(bnd).ravel()
        self.tags["domain_boundary_faces"] = bnd
        self.tags["tip_faces"] = np.zeros(self.num_faces, dtype=bool)
        self.tags["fracture_faces"] = np.zeros(self.num_faces, dtype=bool)
        self.tags["domain_boundary_nodes"] = (
            (self.face_nodes.astype(bool) @ bnd.astype(int)) > 0
        )
        self.tags["tip_nodes"] = np.zeros(self.num_nodes, dtype=bool)
        self.tags["fracture_nodes"] = np.zeros(self.num_nodes, dtype=bool)

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def cell_nodes(self):
        mat = (self.face_nodes.astype(int) @ np.abs(self.cell_faces).astype(int)) > 0
        return sps.csc_array(mat)

    def num_cell_nodes(self):
        return np.asarray(self.cell_nodes().sum(axis=0)).ravel()

    def compute_geometry(self):
        d = self.dim
        x = self.nodes
        fn = self.face_nodes.indices.reshape(self.num_faces, -1)
        self.face_centers = x[:, fn].mean(axis=2)
        cn = self.cell_nodes()
        cn_idx = cn.indices.reshape(self.num_cells, d + 1)
        self.cell_centers = x[:, cn_idx].mean(axis=2)
        if d == 1:
            self.face_normals = np.zeros((3, self.num_faces))
            self.face_normals[0] = 1.0
            self.face_areas = np.ones(self.num_faces)
            e = x[:, cn_idx[:, 1]] - x[:, cn_idx[:, 0]]
            self.cell_volumes = np.linalg.norm(e, axis=0)
        elif d == 2:
            # planar grid, possibly tilted in 3D: in-plane normal t x n_plane ((t_y, -t_x, 0) in the xy-plane)
            t = x[:, fn[:, 1]] - x[:, fn[:, 0]]
            n_plane = _plane_normal(x, cn_idx) if np.any(x[2]) else np.array([0.0, 0.0, 1.0])
            self.face_normals = np.cross(t, n_plane[:, None], axis=0)
            self.face_areas = np.linalg.norm(t, axis=0)
            a = x[:, cn_idx[:, 1]] - x[:, cn_idx[:, 0]]
            b = x[:, cn_idx[:, 2]] - x[:, cn_idx[:, 0]]
            self.cell_volumes = 0.5 * np.linalg.norm(np.cross(a, b, axis=0), axis=0)
        else:
            a = x[:, fn[:, 1]] - x[:, fn[:, 0]]
            b = x[:, fn[:, 2]] - x[:, fn[:, 0]]
            self.face_normals = 0.5 * np.cross(a, b, axis=0)
            self.face_areas = np.linalg.norm(self.face_normals, axis=0)
            p = x[:, cn_idx]  # (3, nc, 4)
            e1, e2, e3 = (p[:, :, k] - p[:, :, 0] for k in (1, 2, 3))
            self.cell_volumes = np.abs(np.einsum("ic,ic->c", e1, np.cross(e2, e3, axis=0))) / 6.0
        if d >= 2:
            # invariant: cell_faces[f, c] = +1  <=>  face_normals[:, f] points out of c
            f, c, s = sps.find(self.cell_faces)
            out = np.einsum(
                "if,if->f", self.face_normals[:, f], self.face_centers[:, f] - self.cell_centers[:, c]
            )
            if not np.all(np.sign(out) == np.sign(s)):
                raise ValueError("cell_faces signs inconsistent with node-order normals")

    def copy(self):
        import copy as _copy

        return _copy.copy(self)


def _plane_normal(x, cn_idx):
    """Unit normal of the plane of a 2D grid embedded in 3D, oriented to the upper half space."""
    a = x[:, cn_idx[0, 1]] - x[:, cn_idx[0, 0]]
    b = x[:, cn_idx[0, 2]] - x[:, cn_idx[0, 0]]
    n = np.cross(a, b)
    n /= np.linalg.norm(n)
    return -n if n[2] < 0 else n


class map_geometry:  # namespace like porepy's module of that name
    """`pp.map_geometry.map_grid`, restated from PorePy's published algorithm (PorePy itself is absent):
    the rotation about `n x e_z` by the angle between the plane normal `n` and `e_z` (Rodrigues' formula);
    the reference reads `R` and `keep_dims` from the tuple (grids/grid.py:366-368)."""

    @staticmethod
    def map_grid(g):
        cn_idx = g.cell_nodes().indices.reshape(g.num_cells, g.dim + 1)
        n = _plane_normal(g.nodes, cn_idx)
        ez = np.array([0.0, 0.0, 1.0])
        axis = np.cross(n, ez)
        s, c = np.linalg.norm(axis), float(n @ ez)
        if s < 1e-15:
            R = np.eye(3)
        else:
            axis = axis / s
            W = np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]])
            R = c * np.eye(3) + s * W + (1.0 - c) * np.outer(axis, axis)
        keep = np.array([True, True, False])
        return None, None, None, None, R, keep, None


class MortarGrid:
    """Conforming mortar grid (the subset of `pp.MortarGrid` the reference reads, grids/mortar_grid.py): one
    mortar cell per (primary face, secondary cell) pair, 0/1 projections, cell measures of the secondary
    cells.  Restated from PorePy's documented behaviour for matching grids; PorePy itself is absent."""

    def __init__(self, dim, face_up, cell_down, num_primary_faces, num_secondary_cells, cell_volumes, name="mortar"):
        self.dim = int(dim)
        self.name = name
        self._face_up = np.asarray(face_up)
        self._cell_down = np.asarray(cell_down)
        self.num_cells = self._face_up.size
        self._nf, self._nc = int(num_primary_faces), int(num_secondary_cells)
        self._vol = np.asarray(cell_volumes, dtype=float)
        self.tags = {}

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def compute_geometry(self):
        self.cell_volumes = self._vol[self._cell_down].copy()

    def primary_to_mortar_int(self):
        m = np.arange(self.num_cells)
        return sps.csc_array((np.ones(self.num_cells), (m, self._face_up)), shape=(self.num_cells, self._nf))

    def secondary_to_mortar_int(self):
        m = np.arange(self.num_cells)
        return sps.csc_array((np.ones(self.num_cells), (m, self._cell_down)), shape=(self.num_cells, self._nc))


class MixedDimensionalGrid:
    """Container of subdomains and interfaces with data dictionaries (the subset of
    `pp.MixedDimensionalGrid` the reference's operator layer calls)."""

    def __init__(self):
        self._sds, self._intfs = [], []

    def add_subdomains(self, sds):
        for sd in sds if isinstance(sds, (list, tuple)) else [sds]:
            self._sds.append((sd, {}))
        self._sds.sort(key=lambda t: -t[0].dim)

    def add_interface(self, intf, sd_pair, _map=None):
        up, down = sd_pair if sd_pair[0].dim > sd_pair[1].dim else sd_pair[::-1]
        self._intfs.append((intf, {}, (up, down)))
        self._intfs.sort(key=lambda t: -t[0].dim)

    def subdomains(self, return_data=False, dim=None):
        sel = [(g, d) for g, d in self._sds if dim is None or g.dim == dim]
        return sel if return_data else [g for g, _ in sel]

    def interfaces(self, return_data=False, dim=None):
        sel = [(g, d) for g, d, _ in self._intfs if dim is None or g.dim == dim]
        return sel if return_data else [g for g, _ in sel]

    def interface_to_subdomain_pair(self, intf):
        return next(pair for g, _, pair in self._intfs if g is intf)

    def subdomain_data(self, sd):
        return next(d for g, d in self._sds if g is sd)

    def interface_data(self, intf):
        return next(d for g, d, _ in self._intfs if g is intf)

    def num_subdomains(self):
        return len(self._sds)

    def num_interfaces(self):
        return len(self._intfs)

    def num_subdomain_cells(self, cond=None):
        return int(sum(g.num_cells for g, _ in self._sds if cond is None or cond(g)))

    def dim_max(self):
        return max(g.dim for g, _ in self._sds)

    def dim_min(self):
        return min(g.dim for g, _ in self._sds)


class _Dummy:
    def __init__(self, name="dummy"):
        self._name = name

    def __getattr__(self, item):
        return _Dummy(f"{self._name}.{item}")

    def __call__(self, *a, **k):
        raise NotImplementedError(f"porepy stand-in: {self._name} is not available")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Dummy(name)
