"""Mixed-dimensional grid container and mortar grid.

The reference's ``pg.MixedDimensionalGrid`` / ``pg.MortarGrid`` (``grids/md_grid.py:11-153``,
``grids/mortar_grid.py:19-208``) subclass PorePy's containers; PorePy is not part of this package,
so the subset of that container API which the operator layer reads
(``numerics/innerproducts.py:162-232``, ``numerics/differentials.py:144-192``,
``numerics/restrictions.py:11-72``) is provided here directly:

* ``subdomains(return_data=, dim=)`` / ``interfaces(return_data=, dim=)`` in PorePy's order
  (descending dimension, insertion order within a dimension),
* ``interface_to_subdomain_pair(intf) -> (sd_up, sd_down)``, ``subdomain_data`` / ``interface_data``,
* ``num_subdomains``, ``num_interfaces``, ``num_subdomain_cells / faces / ridges / peaks``,
  ``dim_max`` / ``dim_min``, ``compute_geometry`` and ``tag_leafs``.

A :class:`MortarGrid` is built for **conforming** interfaces: each mortar cell coincides with one
face of the higher-dimensional grid and one cell of the lower-dimensional one, which is what
PorePy's fracture meshing produces (two mortar cells per fracture cell, one per side).
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import scipy.sparse as sps


def _all(_sd) -> bool:
    return True


class MortarGrid:
    """Interface between a grid of dimension ``dim + 1`` (primary, "up") and one of dimension
    ``dim`` (secondary, "down").

    ``face_up[m]`` / ``cell_down[m]`` are the primary face and the secondary cell that mortar cell
    ``m`` coincides with.  The projections ``primary_to_mortar_int`` / ``secondary_to_mortar_int``
    are then 0/1 matrices (the conforming case of ``pp.MortarGrid``)."""

    def __init__(self, dim: int, face_up, cell_down, name: str = "mortar_grid") -> None:
        self.dim = int(dim)
        self.name = name
        self.face_up = np.asarray(face_up, dtype=np.int64)
        self.cell_down = np.asarray(cell_down, dtype=np.int64)
        if self.face_up.shape != self.cell_down.shape or self.face_up.ndim != 1:
            raise ValueError("face_up and cell_down must be 1-D arrays of the same length")
        self.num_cells = int(self.face_up.size)
        self.sd_pair = None
        self.tags: dict = {}

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    # ---- pp.MortarGrid subset ---------------------------------------------------------------
    def primary_to_mortar_int(self) -> sps.csc_array:
        nf = self.sd_pair[0].num_faces
        m = np.arange(self.num_cells)
        return sps.csc_array((np.ones(self.num_cells), (m, self.face_up)), shape=(self.num_cells, nf))

    def secondary_to_mortar_int(self) -> sps.csc_array:
        nc = self.sd_pair[1].num_cells
        m = np.arange(self.num_cells)
        return sps.csc_array((np.ones(self.num_cells), (m, self.cell_down)), shape=(self.num_cells, nc))

    def mortar_to_primary_int(self) -> sps.csc_array:
        return self.primary_to_mortar_int().T.tocsc()

    def mortar_to_secondary_int(self) -> sps.csc_array:
        return self.secondary_to_mortar_int().T.tocsc()

    # ---- pg.MortarGrid ----------------------------------------------------------------------
    def assign_sd_pair(self, sd_pair) -> None:
        self.sd_pair = tuple(sd_pair)

    def compute_geometry(self) -> None:
        """Mortar cell measures and centres (those of the matching secondary cells), the signed map
        to the primary faces and the jump operator ``cell_faces`` (``mortar_grid.py:27-44,168-208``)."""
        sd_up, sd_down = self.sd_pair
        self.cell_volumes = np.asarray(sd_down.cell_volumes)[self.cell_down].copy()
        self.cell_centers = np.asarray(sd_down.cell_centers)[:, self.cell_down].copy()
        # orientation of the primary face w.r.t. its only cell (fracture faces are boundary faces)
        cf = sd_up.cell_faces.tocsr()
        first = cf.indptr[self.face_up]
        if np.any(cf.indptr[self.face_up + 1] - first != 1):
            raise ValueError("a mortar cell must coincide with a primary face that has exactly one cell")
        sign = np.asarray(cf.data)[first]
        m = np.arange(self.num_cells)
        self.signed_mortar_to_primary = sps.csc_array(
            (sign, (self.face_up, m)), shape=(sd_up.num_faces, self.num_cells))
        self.cell_faces = sps.csc_array(-self.signed_mortar_to_primary @ self.secondary_to_mortar_int())
        if self.dim >= 2:
            self.compute_ridges()

    def compute_ridges(self) -> None:
        """Jump operators one and two codimensions down (``mortar_grid.py:46-166``): high-dimensional
        ridges matched to the low-dimensional faces (``face_ridges``) and high-dimensional peaks
        matched to low-dimensional ridges (``ridge_peaks``, 3D-2D only), entities paired through
        their coordinates cell by cell, orientations from the mortar normal."""
        sd_up, sd_down = self.sd_pair
        d = self.dim
        if d != 2:
            raise NotImplementedError("only 3D-2D interfaces are supported (this package has no 1D grids)")
        nr_up, nf_down = sd_up.num_ridges, sd_down.num_faces
        rows_fr, cols_fr, vals_fr = [], [], []
        rows_rp, cols_rp, vals_rp = [], [], []
        cf_up = sd_up.cell_faces.tocsr()
        out_sign = np.asarray(cf_up.data)[cf_up.indptr[self.face_up]]
        normal_up = np.asarray(sd_up.face_normals)[:, self.face_up] * out_sign  # outward w.r.t. sd_up
        cfd, fru = sd_down.cell_faces, sd_up.face_ridges
        nfc, nrf = 3, 3
        faces_down = cfd.indices.reshape(sd_down.num_cells, nfc)[self.cell_down]  # (m, d+1)
        ridges_up = fru.indices.reshape(sd_up.num_faces, nrf)[self.face_up]  # (m, nrf)
        fxyz = np.asarray(sd_down.face_centers)[:, faces_down]  # (3, m, k)
        rp = sd_up.ridge_peaks.indices.reshape(sd_up.num_ridges, 2)
        rxyz = np.asarray(sd_up.nodes)[:, rp[ridges_up]].mean(axis=3)
        # pair every low-dimensional face with the coinciding high-dimensional ridge
        dist = np.linalg.norm(fxyz[:, :, :, None] - rxyz[:, :, None, :], axis=0)  # (m, k, k)
        match = dist.argmin(axis=2)
        if np.any(np.take_along_axis(dist, match[:, :, None], axis=2) > 1e-8 * (1 + np.abs(fxyz).max())):
            raise ValueError("interface is not conforming: a face of the lower grid has no matching ridge")
        ridges_m = np.take_along_axis(ridges_up, match, axis=1)
        n_down = np.asarray(sd_down.face_normals)[:, faces_down]  # (3, m, k)
        tang = np.asarray(sd_up.edge_tangents)[:, ridges_m]  # (3, m, k)
        prod = np.cross(tang, normal_up[:, :, None], axisa=0, axisb=0, axisc=0)
        orient = np.einsum("imk,imk->mk", prod, n_down)
        # ridge_peaks: nodes of the high-dimensional face matched to the cell's nodes below
        Rd = np.asarray(sd_down.rotation_matrix)
        normal_plane = np.cross(Rd[0], Rd[1])
        cn = sd_down.cell_nodes().indices.reshape(sd_down.num_cells, 3)[self.cell_down]
        fn = sd_up.face_nodes.indices.reshape(sd_up.num_faces, 3)[self.face_up]
        pd = np.asarray(sd_down.nodes)[:, cn]
        pu = np.asarray(sd_up.nodes)[:, fn]
        dist2 = np.linalg.norm(pd[:, :, :, None] - pu[:, :, None, :], axis=0)
        pm = np.take_along_axis(fn, dist2.argmin(axis=2), axis=1)
        o_rp = -np.sign(normal_up.T @ normal_plane)  # (m,)
        rows_rp.append(pm.ravel())
        cols_rp.append(cn.ravel())
        vals_rp.append(np.repeat(o_rp, 3))
        rows_fr.append(ridges_m.ravel())
        cols_fr.append(faces_down.ravel())
        vals_fr.append(np.sign(orient).ravel())

        def build(rows, cols, vals, shape):
            if not rows:
                return sps.csc_array(shape, dtype=int)
            M = sps.csc_array((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=shape)
            M.sum_duplicates()
            M.data = np.sign(M.data).astype(int)  # both sides of a fracture hit the same pair: keep +-1, 0 at tips
            return M.astype(int)

        self.face_ridges = build(rows_fr, cols_fr, vals_fr, (nr_up, nf_down))
        self.ridge_peaks = build(rows_rp, cols_rp, vals_rp, (sd_up.num_peaks, sd_down.num_ridges))


class MixedDimensionalGrid:
    """Subdomains + interfaces with their data dictionaries."""

    def __init__(self) -> None:
        self._sd: list = []  # (grid, data)
        self._intf: list = []  # (mortar grid, data, (sd_up, sd_down))
        self.name = "Mixed-dimensional grid"

    # ---- construction ---------------------------------------------------------------------
    def add_subdomains(self, sds) -> None:
        for sd in (sds if isinstance(sds, (list, tuple)) else [sds]):
            if any(sd is s for s, _ in self._sd):
                raise ValueError("subdomain already present")
            self._sd.append((sd, {}))
        self._sd.sort(key=lambda t: -t[0].dim)  # stable: insertion order inside a dimension

    def add_interface(self, intf: MortarGrid, sd_pair, _primary_secondary_map=None) -> None:
        up, down = sd_pair
        if up.dim < down.dim:
            up, down = down, up
        intf.assign_sd_pair((up, down))
        self._intf.append((intf, {}, (up, down)))
        self._intf.sort(key=lambda t: -t[0].dim)

    # ---- access -------------------------------------------------------------------------------
    def subdomains(self, return_data: bool = False, dim: int | None = None) -> list:
        sel = [(sd, d) for sd, d in self._sd if dim is None or sd.dim == dim]
        return sel if return_data else [sd for sd, _ in sel]

    def interfaces(self, return_data: bool = False, dim: int | None = None) -> list:
        sel = [(g, d) for g, d, _ in self._intf if dim is None or g.dim == dim]
        return sel if return_data else [g for g, _ in sel]

    def interface_to_subdomain_pair(self, intf) -> tuple:
        for g, _, pair in self._intf:
            if g is intf:
                return pair
        raise KeyError("unknown interface")

    def subdomain_pair_to_interface(self, sd_pair):
        for g, _, pair in self._intf:
            if (pair[0] is sd_pair[0] and pair[1] is sd_pair[1]) or (pair[0] is sd_pair[1] and pair[1] is sd_pair[0]):
                return g
        raise KeyError("no interface between the subdomains")

    def subdomain_data(self, sd) -> dict:
        for s, d in self._sd:
            if s is sd:
                return d
        raise KeyError("unknown subdomain")

    def interface_data(self, intf) -> dict:
        for g, d, _ in self._intf:
            if g is intf:
                return d
        raise KeyError("unknown interface")

    def num_subdomains(self) -> int:
        return len(self._sd)

    def num_interfaces(self) -> int:
        return len(self._intf)

    def dim_max(self) -> int:
        return max(sd.dim for sd, _ in self._sd)

    def dim_min(self) -> int:
        return min(sd.dim for sd, _ in self._sd)

    def _count(self, attr: str, cond: Callable | None) -> int:
        cond = _all if cond is None else cond
        return int(sum(getattr(sd, attr) for sd, _ in self._sd if cond(sd)))

    def num_subdomain_cells(self, cond: Callable | None = None) -> int:
        return self._count("num_cells", cond)

    def num_subdomain_faces(self, cond: Callable | None = None) -> int:
        return self._count("num_faces", cond)

    def num_subdomain_ridges(self, cond: Callable | None = None) -> int:
        return self._count("num_ridges", cond)

    def num_subdomain_peaks(self, cond: Callable | None = None) -> int:
        return self._count("num_peaks", cond)

    def num_interface_cells(self, cond: Callable | None = None) -> int:
        cond = _all if cond is None else cond
        return int(sum(g.num_cells for g, _, _ in self._intf if cond(g)))

    # ---- geometry / tags (grids/md_grid.py:20-153) ----------------------------------------------
    def compute_geometry(self) -> None:
        for sd, _ in self._sd:
            sd.compute_geometry()
        for intf, _, pair in self._intf:
            intf.assign_sd_pair(pair)
            intf.compute_geometry()
        self.tag_leafs()

    def tag_leafs(self) -> None:
        """Faces / ridges / peaks of a subdomain that coincide with an entity of a lower-dimensional
        one (``md_grid.py:113-153``)."""
        for sd, _ in self._sd:
            zeros = lambda n: np.zeros(n, dtype=bool)  # noqa: E731
            tip = sd.tags.get("tip_faces", zeros(sd.num_faces))
            frac = sd.tags.get("fracture_faces", zeros(sd.num_faces))
            sd.tags["leaf_faces"] = np.logical_or(tip, frac)
            sd.tags["leaf_ridges"] = zeros(getattr(sd, "num_ridges", 0))
            sd.tags["leaf_peaks"] = zeros(getattr(sd, "num_peaks", 0))
        for codim, (up_tag, down_tag, op) in enumerate(
                (("leaf_ridges", "leaf_faces", "face_ridges"), ("leaf_peaks", "leaf_ridges", "ridge_peaks")), start=1):
            for intf, _, (sd_up, sd_down) in self._intf:
                if intf.dim >= codim:
                    hit = abs(getattr(intf, op)) @ sd_down.tags[down_tag].astype(int)
                    sd_up.tags[up_tag] = np.logical_or(sd_up.tags[up_tag], hit.astype(bool))

    def __repr__(self) -> str:
        return (f"Mixed-dimensional grid: {self.num_subdomains()} subdomains, {self.num_interfaces()} interfaces, "
                f"dimensions {self.dim_max()}..{self.dim_min()}")


def fractured_unit_cube(n: int, perturb: float = 0.0, seed: int = 0) -> MixedDimensionalGrid:
    """Unit cube of ``n x n x n`` Kuhn cubes (``n`` even) cut by the horizontal fracture ``z = 1/2``: the 3D
    grid is split along the fracture (nodes, ridges and faces of the plane exist once per side, as PorePy's
    fracture meshing produces them), the fracture is the matching ``n x n`` structured triangle grid in that
    plane, and one conforming mortar grid of ``2 * 2 n^2`` cells couples them.  A gmsh-free stand-in for
    ``pp.mdg_library.cube_with_orthogonal_fractures`` (tests/conftest.py:277-282 of the reference) that
    scales to any size; ``perturb`` moves the nodes off the fracture plane randomly."""
    from .grid import SimplexGrid, structured_grid

    if n % 2:
        raise ValueError("n must be even")
    h = n // 2
    bot = structured_grid(3, n, nz=h, perturb=perturb, seed=seed)
    top = structured_grid(3, n, nz=h, z0=h, perturb=perturb, seed=seed + 1)
    nn_b, nf_b, nc_b = bot.num_nodes, bot.num_faces, bot.num_cells
    nodes = np.hstack([bot.nodes, top.nodes])
    face_nodes = sps.block_diag([bot.face_nodes.astype(float), top.face_nodes.astype(float)], format="csc").astype(bool)
    cell_faces = sps.block_diag([bot.cell_faces, top.cell_faces], format="csc")
    sd3 = SimplexGrid(3, nodes, face_nodes, cell_faces, "fractured_cube_3d")
    sd2 = structured_grid(2, n)
    sd2.nodes[2] = 0.5
    sd2.name = "fracture_2d"
    # fracture faces: the top plane of the bottom half, the bottom plane of the top half; a face of the plane
    # is matched to the fracture cell with the same three lattice nodes
    m2 = (n + 1) ** 2
    fn3 = sd3.face_nodes.indices.reshape(sd3.num_faces, 3).astype(np.int64)
    plane_b = np.flatnonzero((fn3[:nf_b] >= h * m2).all(axis=1))  # bottom half: nodes of its last layer
    plane_t = nf_b + np.flatnonzero((fn3[nf_b:] - nn_b < m2).all(axis=1))  # top half: nodes of its first layer

    def key(tri):
        tri = np.sort(tri, axis=1)
        return (tri[:, 0] * m2 + tri[:, 1]) * m2 + tri[:, 2]

    cn2 = sd2.cell_node_table().astype(np.int64)
    order = np.argsort(key(cn2))
    sorted_keys = key(cn2)[order]
    face_up, cell_down = [], []
    for faces, shift in ((plane_b, h * m2), (plane_t, nn_b)):
        pos = np.searchsorted(sorted_keys, key(fn3[faces] - shift))
        if np.any(sorted_keys[np.minimum(pos, sorted_keys.size - 1)] != key(fn3[faces] - shift)):
            raise RuntimeError("fracture faces do not match the fracture grid")
        face_up.append(faces)
        cell_down.append(order[pos])
    face_up, cell_down = np.concatenate(face_up), np.concatenate(cell_down)
    frac = np.zeros(sd3.num_faces, dtype=bool)
    frac[face_up] = True
    sd3.tags["fracture_faces"] = frac
    mdg = MixedDimensionalGrid()
    mdg.add_subdomains([sd3, sd2])
    mdg.add_interface(MortarGrid(2, face_up, cell_down, "mortar_2d"), (sd3, sd2))
    return mdg


def as_mdg(obj) -> MixedDimensionalGrid:
    """A grid wrapped as a one-subdomain mixed-dimensional grid; a mixed-dimensional grid is returned
    unchanged (``filters/convert_from_pp.py:66-87``)."""
    if isinstance(obj, MixedDimensionalGrid):
        return obj
    if hasattr(obj, "cell_faces") and hasattr(obj, "num_cells") and hasattr(obj, "dim"):
        mdg = MixedDimensionalGrid()
        mdg.add_subdomains([obj])
        return mdg
    raise ValueError("expected a grid or a mixed-dimensional grid")
