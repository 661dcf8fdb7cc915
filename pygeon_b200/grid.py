"""Simplicial grid container and structured unit-square / unit-cube generators.

The discretizations in this package read a grid through the attribute contract
of ``pg.Grid`` (reference ``src/pygeon/grids/grid.py:18-370`` on top of
``pp.Grid``): ``dim, num_cells, num_faces, num_nodes, nodes, face_nodes,
cell_faces, face_ridges, ridge_peaks, num_ridges, num_edges, rotation_matrix,
tags`` plus the derived geometry.  A real ``pg.Grid`` satisfies it; so does
:class:`SimplexGrid`, which exists because PorePy (the provider of ``pp.Grid``
and of ``pp.StructuredTriangleGrid`` / ``pp.StructuredTetrahedralGrid``,
``src/pygeon/grids/create_grid.py:134-149``) is not part of this package.

Conventions:

* ``cell_faces[f, c] = +1`` iff ``face_normals[:, f]`` points out of ``c``.
* the normal is tied to the node order stored in ``face_nodes.indices``:
  2D ``n = (t_y, -t_x)`` with ``t = x1 - x0``; 3D ``n = (x1-x0) x (x2-x0) / 2``
  (right-hand rule, as assumed by ``grid.py:168-182``).
* 3D ridges are the unique sorted node pairs numbered lexicographically,
  ``ridge_peaks[:, e] = {lo: -1, hi: +1}`` (``grid.py:187-210``);
  ``face_ridges[e, f] = sign(next - prev)`` walking the face nodes cyclically.
* 2D: ridges are nodes, ``face_ridges = face_nodes`` with signs (-1, +1) times
  the orientation that makes the tangent the +90 degree rotation of the normal
  (``grid.py:132-147``); with the normal convention above that is always +1.

The structured generators number faces / edges as "sorted unique node tuples",
evaluated with integer lattice keys so they scale to 10^8 cells (on the GPU
when the tensors live there).  They are input providers, not part of the
assembly hot path.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import torch

AMBIENT_DIM = 3


class SimplexGrid:
    """Duck-typed stand-in for ``pg.Grid`` restricted to simplicial grids."""

    def __init__(self, dim, nodes, face_nodes, cell_faces, name="simplex_grid"):
        self.dim = int(dim)
        self.name = name
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.face_nodes = sps.csc_array(face_nodes)
        self.cell_faces = sps.csc_array(cell_faces)
        self.num_nodes = self.nodes.shape[1]
        self.num_faces = self.face_nodes.shape[1]
        self.num_cells = self.cell_faces.shape[1]
        self.tags: dict = {}
        self._geometry_done = False

    # pg.Grid objects are used as dictionary keys by the reference's caches
    # (hdiv.py:255,450); keep identity semantics.
    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    # ------------------------------------------------------------------ pp.Grid
    def cell_node_table(self) -> np.ndarray:
        """(nc, d+1) ascending node ids of every cell, without a sparse product.

        Every node of a simplex lies on exactly d of its d+1 faces, so the sorted
        multiset of the faces' nodes holds each cell node d times.
        """
        d = self.dim
        if np.any(np.diff(self.face_nodes.indptr) != d) or np.any(
            np.diff(self.cell_faces.indptr) != d + 1
        ):
            raise NotImplementedError("Grid is not simplicial")
        fn = self.face_nodes.indices.reshape(self.num_faces, d)
        cf = self.cell_faces.indices.reshape(self.num_cells, d + 1)
        alln = np.sort(fn[cf].reshape(self.num_cells, d * (d + 1)), axis=1)
        tab = alln[:, ::d]
        if not np.all(alln.reshape(self.num_cells, d + 1, d) == tab[:, :, None]):
            raise NotImplementedError("Grid is not simplicial")
        return tab

    def cell_nodes(self) -> sps.csc_array:
        tab = self.cell_node_table()
        d = self.dim
        indptr = np.arange(0, tab.size + 1, d + 1)
        return sps.csc_array(
            (np.ones(tab.size, dtype=bool), tab.ravel(), indptr),
            shape=(self.num_nodes, self.num_cells),
        )

    def num_cell_nodes(self) -> np.ndarray:
        return np.diff(self.cell_nodes().indptr)

    def compute_connectivity(self, device=None) -> None:
        """Ridges / peaks only (no metric quantities): what the assembly engine reads besides
        ``nodes``, ``cell_faces`` and ``face_nodes``.  Cheap enough for 10^8-cell grids.
        With a CUDA ``device`` the ridges of an unstructured 3D grid are numbered by
        ``pgb_ridges3d`` (:mod:`pygeon_b200.connectivity`) instead of numpy."""
        self.rotation_matrix = np.eye(self.dim, AMBIENT_DIM)
        self.tags.setdefault("domain_boundary_faces", np.bincount(self.cell_faces.indices, minlength=self.num_faces) == 1)
        self._compute_ridges(device)
        self.num_edges = self.num_faces if self.dim == 2 else self.num_ridges

    def compute_geometry(self, device=None) -> None:
        """Geometry + ridges/peaks + edge properties (``grid.py:39-64``).  With a CUDA ``device``
        every step runs in libpgb.so kernels (:mod:`pygeon_b200.connectivity`) and only the results
        come back to the host attributes; without one, numpy."""
        d = self.dim
        if d not in (2, 3):
            raise NotImplementedError("SimplexGrid supports dim 2 and 3")
        if device is not None:
            return self._compute_geometry_device(device)
        x = self.nodes
        if np.any(np.diff(self.face_nodes.indptr) != d):
            raise NotImplementedError("faces must be simplices")
        fn = self.face_nodes.indices.reshape(self.num_faces, d)
        self.face_centers = x[:, fn].mean(axis=2)
        cni = self.cell_node_table()
        self.cell_centers = x[:, cni].mean(axis=2)
        self.compute_rotation_matrix()
        if d == 2:
            # in-plane normal of the stored node order: t x n_plane (= (t_y, -t_x, 0) in the xy-plane)
            R = self.rotation_matrix
            t = x[:, fn[:, 1]] - x[:, fn[:, 0]]
            self.face_normals = np.cross(t, np.cross(R[0], R[1])[:, None], axis=0)
            a = x[:, cni[:, 1]] - x[:, cni[:, 0]]
            b = x[:, cni[:, 2]] - x[:, cni[:, 0]]
            self.cell_volumes = 0.5 * np.linalg.norm(np.cross(a, b, axis=0), axis=0)
        else:
            a = x[:, fn[:, 1]] - x[:, fn[:, 0]]
            b = x[:, fn[:, 2]] - x[:, fn[:, 0]]
            self.face_normals = 0.5 * np.cross(a, b, axis=0)
            p = x[:, cni]
            e1, e2, e3 = (p[:, :, k] - p[:, :, 0] for k in (1, 2, 3))
            self.cell_volumes = np.abs(np.einsum("ic,ic->c", e1, np.cross(e2, e3, axis=0))) / 6.0
        self.face_areas = np.linalg.norm(self.face_normals, axis=0)

        f, c, s = sps.find(self.cell_faces)
        out = np.einsum(
            "if,if->f", self.face_normals[:, f], self.face_centers[:, f] - self.cell_centers[:, c]
        )
        if not np.all(np.sign(out) == np.sign(s)):
            raise ValueError("cell_faces signs are inconsistent with the node-order normals")

        # faces tagged as fracture / tip faces by a mixed-dimensional construction stay tagged and are not
        # part of the domain boundary (PorePy's convention)
        self.tags.setdefault("tip_faces", np.zeros(self.num_faces, dtype=bool))
        self.tags.setdefault("fracture_faces", np.zeros(self.num_faces, dtype=bool))
        bnd = np.bincount(self.cell_faces.indices, minlength=self.num_faces) == 1
        bnd &= ~(self.tags["tip_faces"] | self.tags["fracture_faces"])
        self.tags["domain_boundary_faces"] = bnd
        bn = np.zeros(self.num_nodes, dtype=bool)
        bn[fn[bnd].ravel()] = True
        self.tags["domain_boundary_nodes"] = bn

        self._compute_ridges()
        self._compute_edge_properties()
        self._geometry_done = True

    def compute_rotation_matrix(self) -> None:
        """``rotation_matrix`` (dim x 3, orthonormal rows): maps the grid's plane to the xy-plane
        (``grids/grid.py:355-370``).  The reference takes it from ``pp.map_geometry.map_grid`` (PorePy,
        absent here): the rotation about ``n x e_z`` by the angle between the plane normal ``n`` and
        ``e_z`` (Rodrigues' formula), of which the first ``dim`` rows are kept.  Restated here; every
        matrix assembled from it only depends on the plane, not on the in-plane basis."""
        d = self.dim
        if d == 3 or not np.any(self.nodes[d:, :]):
            self.rotation_matrix = np.eye(d, AMBIENT_DIM)
            return
        if d != 2:
            raise NotImplementedError("only 2D grids embedded in 3D are supported")
        x = self.nodes
        fn = self.face_nodes.indices.reshape(self.num_faces, 2)
        cf = self.cell_faces.indices.reshape(self.num_cells, 3)
        # normal of the largest-area corner of the first cell; the sign convention follows that cell
        tri = np.unique(fn[cf[0]].ravel())
        nvec = np.cross(x[:, tri[1]] - x[:, tri[0]], x[:, tri[2]] - x[:, tri[0]])
        nvec = nvec / np.linalg.norm(nvec)
        if nvec[2] < 0:
            nvec = -nvec
        ez = np.array([0.0, 0.0, 1.0])
        axis = np.cross(nvec, ez)
        s, c = np.linalg.norm(axis), float(nvec @ ez)
        if s < 1e-15:
            R = np.eye(3)
        else:
            axis /= s
            W = np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]])
            R = c * np.eye(3) + s * W + (1.0 - c) * np.outer(axis, axis)
        planar = R @ (x - x[:, :1])
        if np.abs(planar[2]).max() > 1e-10 * max(np.abs(planar).max(), 1e-300):
            raise ValueError("the nodes of a 2D grid must lie in one plane")
        self.rotation_matrix = R[:2]

    def _compute_geometry_device(self, device) -> None:
        from . import connectivity, engine

        d, nc, nf = self.dim, self.num_cells, self.num_faces
        if np.any(self.nodes[d:, :]):
            raise NotImplementedError("device geometry supports axis-aligned grids; use compute_geometry() for tilted ones")
        self.rotation_matrix = np.eye(d, AMBIENT_DIM)
        raw = engine.RawGrid.upload(self, device, signs=True)
        pack = engine.GridPack.from_raw(raw)  # NotImplementedError for non-simplicial grids
        dev = raw.device
        with torch.cuda.device(dev):
            if raw.nodes_ready is not None:
                torch.cuda.current_stream().wait_event(raw.nodes_ready)
            cf, fn = raw.cf_idx.view(nc, d + 1), raw.fn_idx.view(nf, d)
            geo = connectivity.geometry(d, raw.nodes, cf, raw.cf_sgn.view(nc, d + 1), fn, pack.cell_nodes)
            bnd = torch.bincount(raw.cf_idx.long(), minlength=nf) == 1
            bn = torch.zeros(self.num_nodes, dtype=torch.bool, device=dev)
            bn[fn[bnd].reshape(-1).long()] = True
            for k, v in geo.items():
                setattr(self, k, v.cpu().numpy())
            self.tags["domain_boundary_faces"] = bnd.cpu().numpy()
            self.tags["domain_boundary_nodes"] = bn.cpu().numpy()
        self.tags["tip_faces"] = np.zeros(nf, dtype=bool)
        self.tags["fracture_faces"] = np.zeros(nf, dtype=bool)
        self._compute_ridges(dev)
        with torch.cuda.device(dev):
            if d == 2:
                en, es = raw.fn_idx.view(nf, 2), _i8(self.face_ridges.data, dev).view(nf, 2)
            else:
                dr = getattr(self, "_device_ridges", None)  # absent for the structured generators
                en = dr[2] if dr is not None else engine._to_dev(self.ridge_peaks.indices, dev, torch.int32).view(-1, 2)
                es = None
            t, ln = connectivity.edge_geometry(raw.nodes, en, es)
            self.edge_tangents, self.edge_lengths = t.cpu().numpy(), ln.cpu().numpy()
            self.num_edges = self.edge_lengths.size
            self.mesh_size = float(ln.mean().item()) if ln.numel() else 0.0
        self._geometry_done = True

    # ------------------------------------------------------------------ pg.Grid
    def _compute_ridges_device(self, device) -> None:
        """3D ridges through ``pgb_ridges3d``; keeps the device tensors for the assembly engine
        (``RawGrid.need_ridges`` then uploads nothing)."""
        from . import connectivity, engine

        dev = engine._device(device)
        if np.any(np.diff(self.face_nodes.indptr) != 3):
            raise NotImplementedError("faces must be simplices")
        with torch.cuda.device(dev):
            fn = engine._to_dev(self.face_nodes.indices, dev, torch.int32).view(self.num_faces, 3)
            fr, fs, rp = connectivity.ridges3d(fn, self.num_nodes)
            self._device_ridges = (fr, fs, rp)
            bnd = engine._to_dev(self.tags["domain_boundary_faces"], dev)
            br = torch.zeros(rp.shape[0], dtype=torch.bool, device=dev)
            br[fr[bnd].reshape(-1).long()] = True
            fr_h, fs_h, rp_h, br_h = fr.cpu().numpy(), fs.cpu().numpy(), rp.cpu().numpy(), br.cpu().numpy()
        self.num_peaks = self.num_nodes
        self.num_ridges = rp_h.shape[0]
        self.ridge_peaks = _csc_from_rows(rp_h, np.tile(np.array([-1, 1]), rp_h.shape[0]).reshape(-1, 2), self.num_nodes)
        self.face_ridges = _csc_from_rows(fr_h, fs_h.astype(int), self.num_ridges)
        self.tags["tip_peaks"] = np.zeros(self.num_peaks, dtype=bool)
        self.tags["tip_ridges"] = np.zeros(self.num_ridges, dtype=bool)
        self.tags["domain_boundary_ridges"] = br_h

    def _compute_ridges(self, device=None) -> None:
        d = self.dim
        if device is not None and d == 3 and not getattr(self, "_structured", None):
            return self._compute_ridges_device(device)
        if d == 2:
            self.num_peaks = 0
            self.num_ridges = self.num_nodes
            self.ridge_peaks = sps.csc_array((0, self.num_ridges), dtype=int)
            fr = self.face_nodes.copy().astype(int)
            fr.data[::2] = -1
            fr.data[1::2] = 1
            self.face_ridges = fr
        elif getattr(self, "_structured", None):
            st = self._structured
            self.num_peaks = self.num_nodes
            rp = st["ridge_peaks"].cpu().numpy()
            self.num_ridges = rp.shape[0]
            self.ridge_peaks = _csc_from_rows(rp, np.tile(np.array([-1, 1]), rp.shape[0]).reshape(-1, 2), self.num_nodes)
            self.face_ridges = _csc_from_rows(
                st["face_ridges"].cpu().numpy(), st["face_ridge_sign"].cpu().numpy().astype(int), self.num_ridges
            )
        else:
            self.num_peaks = self.num_nodes
            fn = self.face_nodes.indices.reshape(self.num_faces, 3).astype(np.int64)
            a = fn.ravel()
            b = np.roll(fn, -1, axis=1).ravel()
            orient = np.sign(b - a).astype(int)
            lo, hi = np.minimum(a, b), np.maximum(a, b)
            key = lo * self.num_nodes + hi
            ukey, inv = np.unique(key, return_inverse=True)
            self.num_ridges = ukey.size
            rp_idx = np.empty(2 * ukey.size, dtype=np.int64)
            rp_idx[0::2] = ukey // self.num_nodes
            rp_idx[1::2] = ukey % self.num_nodes
            rp_dat = np.tile(np.array([-1, 1]), ukey.size)
            self.ridge_peaks = sps.csc_array(
                (rp_dat, rp_idx, np.arange(0, 2 * ukey.size + 1, 2)),
                shape=(self.num_nodes, ukey.size),
            )
            self.face_ridges = sps.csc_array(
                (orient, inv, self.face_nodes.indptr.copy()),
                shape=(ukey.size, self.num_faces),
            )
        self.tags["tip_peaks"] = np.zeros(self.num_peaks, dtype=bool)
        self.tags["tip_ridges"] = np.zeros(self.num_ridges, dtype=bool)
        # astype() copies: abs()/sum_duplicates() would sort the indices in place
        frb = self.face_ridges.astype(bool).astype(np.int32)
        self.tags["domain_boundary_ridges"] = (
            frb @ self.tags["domain_boundary_faces"].astype(np.int32)
        ) > 0

    def _compute_edge_properties(self) -> None:
        edge_nodes = self.face_ridges if self.dim == 2 else self.ridge_peaks
        self.edge_tangents = self.nodes @ edge_nodes
        self.edge_lengths = np.sqrt(np.sum(self.edge_tangents**2, axis=0))
        self.num_edges = self.edge_lengths.size
        self.mesh_size = float(np.mean(self.edge_lengths))

    def compute_opposite_nodes(self, recompute: bool = False) -> sps.csc_array:
        """(face, cell) -> id of the cell node not on the face (``grid.py:279-309``)."""
        if recompute or not hasattr(self, "opposite_nodes"):
            d = self.dim
            cni = self.cell_node_table().astype(np.int64)
            cf = self.cell_faces.indices.reshape(self.num_cells, d + 1)
            fn = self.face_nodes.indices.reshape(self.num_faces, d).astype(np.int64)
            opp = cni.sum(axis=1)[:, None] - fn[cf].sum(axis=2)
            self.opposite_nodes = sps.csc_array(
                (opp.ravel(), self.cell_faces.indices.copy(), self.cell_faces.indptr.copy()),
                shape=self.cell_faces.shape,
            )
        return self.opposite_nodes


# ---------------------------------------------------------------------------
# structured generators (integer lattice, torch; CPU or CUDA)
# ---------------------------------------------------------------------------

_PERMS3 = ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0))


def _unique_sorted(keys: torch.Tensor):
    """Sorted unique + inverse.  torch.unique on CUDA sorts; on CPU too."""
    return torch.unique(keys, sorted=True, return_inverse=True)


def structured_connectivity(dim: int, n: int, device="cpu", nz: int | None = None) -> dict:
    """Connectivity of the 2-triangle / 6-Kuhn-tet split of an n^dim lattice (3D: n x n x nz
    cubes when ``nz`` is given; z is the slowest axis, so a z-slab of a taller box has the same
    numbering up to constant shifts).

    Returns int32 torch tensors on ``device`` (all row-major, one row per entity):

    ``cell_faces (nc, d+1)`` sorted face ids, ``cell_face_sign (nc, d+1)`` int8,
    ``face_nodes (nf, d)`` ascending node ids, and in 3D ``face_ridges (nf, 3)``,
    ``face_ridge_sign (nf, 3)`` int8, ``ridge_peaks (nr, 2)``; plus
    ``lattice (nn, dim)`` integer node coordinates (int32).
    """
    dev = torch.device(device)
    m = n + 1
    nz = n if nz is None else int(nz)
    i64 = torch.int64
    ar = torch.arange(n, device=dev, dtype=i64)
    if dim == 2:
        jj, ii = torch.meshgrid(ar, ar, indexing="ij")
        base = (ii + m * jj).reshape(-1)
        off = (1, m)  # offsets of e_x, e_y
        # triangles: b, b+e_x, b+e_x+e_y  and  b, b+e_y, b+e_x+e_y
        v = torch.stack(
            [
                torch.stack([base, base + off[0], base + off[0] + off[1]], dim=1),
                torch.stack([base, base + off[1], base + off[0] + off[1]], dim=1),
            ],
            dim=1,
        ).reshape(-1, 3)
        nbits, stride = 2, (1, m)
    elif dim == 3:
        kk, jj, ii = torch.meshgrid(torch.arange(nz, device=dev, dtype=i64), ar, ar, indexing="ij")
        base = (ii + m * jj + m * m * kk).reshape(-1)
        off = (1, m, m * m)
        tets = []
        for p in _PERMS3:
            v1 = base + off[p[0]]
            v2 = v1 + off[p[1]]
            v3 = v2 + off[p[2]]
            tets.append(torch.stack([base, v1, v2, v3], dim=1))
        v = torch.stack(tets, dim=1).reshape(-1, 4)
        nbits, stride = 3, (1, m, m * m)
    else:
        raise ValueError("dim must be 2 or 3")
    nc = v.shape[0]
    size = (m, m) if dim == 2 else (m, m, nz + 1)
    nn = int(np.prod(size))

    def step_code(delta: torch.Tensor) -> torch.Tensor:
        # delta = sx + sy*m (+ sz*m^2) with s in {0,1}; code = sx + 2 sy + 4 sz.
        code = torch.zeros_like(delta)
        rem = delta
        for axis in range(dim - 1, -1, -1):
            s = torch.div(rem, stride[axis], rounding_mode="floor")
            # guard: for axis>0 a delta of (stride-1)... cannot happen (s in {0,1})
            rem = rem - s * stride[axis]
            code = code + (s << axis)
        return code

    def code_offset(code: torch.Tensor) -> torch.Tensor:
        o = torch.zeros_like(code)
        for axis in range(dim):
            o = o + ((code >> axis) & 1) * stride[axis]
        return o

    sh = nbits
    # ---- faces: drop local node a  -> ascending (d)-tuple
    cols = list(range(dim + 1))
    fkeys = []
    for a in cols:
        keep = [c for c in cols if c != a]
        t = v[:, keep]
        key = t[:, 0]
        for q in range(1, dim):
            key = (key << sh) | step_code(t[:, q] - t[:, q - 1])
        fkeys.append(key)
    fkeys = torch.stack(fkeys, dim=1)  # (nc, d+1): face opposite local node a
    ufk, inv = _unique_sorted(fkeys.reshape(-1))
    nf = ufk.numel()
    face_of = inv.reshape(nc, dim + 1)
    # decode face nodes
    fn = [None] * dim
    k = ufk
    codes = []
    for q in range(dim - 1):
        codes.append(k & ((1 << sh) - 1))
        k = k >> sh
    codes = codes[::-1]
    fn[0] = k
    for q in range(1, dim):
        fn[q] = fn[q - 1] + code_offset(codes[q - 1])
    face_nodes = torch.stack(fn, dim=1)

    # lattice coordinates
    ids = torch.arange(nn, device=dev, dtype=i64)
    lat = torch.stack([(ids // stride[a]) % size[a] for a in range(dim)], dim=1)

    # ---- signs: +1 iff node-order normal points out of the cell (away from the opposite node)
    sign = torch.empty((nc, dim + 1), dtype=torch.int8, device=dev)
    for a in cols:
        f = face_of[:, a]
        x0 = lat[face_nodes[f, 0]]
        xo = lat[v[:, a]]
        if dim == 2:
            t = lat[face_nodes[f, 1]] - x0
            nrm = torch.stack([t[:, 1], -t[:, 0]], dim=1)
        else:
            e1 = lat[face_nodes[f, 1]] - x0
            e2 = lat[face_nodes[f, 2]] - x0
            nrm = torch.linalg.cross(e1, e2)
        s = ((x0 - xo) * nrm).sum(dim=1)
        sign[:, a] = torch.sign(s).to(torch.int8)
    # sort each cell's faces ascending (CSC sorted indices)
    face_sorted, order = torch.sort(face_of, dim=1)
    sign_sorted = torch.gather(sign, 1, order)

    out = {
        "dim": dim,
        "n": n,
        "num_cells": nc,
        "num_faces": nf,
        "num_nodes": nn,
        "cell_faces": face_sorted.to(torch.int32).contiguous(),
        "cell_face_sign": sign_sorted.contiguous(),
        "face_nodes": face_nodes.to(torch.int32).contiguous(),
        "lattice": lat.to(torch.int32).contiguous(),
        "cell_nodes": v,  # (nc, d+1) int64 ascending node ids
        "face_keys": ufk,  # sorted int64 lattice keys: lo node << (d-1)*key_bits | step codes
        "key_bits": sh,
    }
    if dim == 3:
        # ---- edges: all 6 local pairs
        ekeys = []
        for p in range(4):
            for q in range(p + 1, 4):
                ekeys.append((v[:, p] << sh) | step_code(v[:, q] - v[:, p]))
        uek = torch.unique(torch.stack(ekeys, dim=1).reshape(-1), sorted=True)
        nr = uek.numel()
        lo = uek >> sh
        hi = lo + code_offset(uek & ((1 << sh) - 1))
        out["ridge_peaks"] = torch.stack([lo, hi], dim=1).to(torch.int32).contiguous()
        out["num_ridges"] = nr
        out["edge_keys"] = uek  # sorted int64: lo node << key_bits | step code
        # face_ridges: ridge k joins face node k and k+1 (cyclic)
        fr = []
        fs = []
        for q in range(3):
            a = face_nodes[:, q]
            b = face_nodes[:, (q + 1) % 3]
            l, h = torch.minimum(a, b), torch.maximum(a, b)
            key = (l << sh) | step_code(h - l)
            fr.append(torch.searchsorted(uek, key))
            fs.append(torch.sign(b - a).to(torch.int8))
        out["face_ridges"] = torch.stack(fr, dim=1).to(torch.int32).contiguous()
        out["face_ridge_sign"] = torch.stack(fs, dim=1).contiguous()
    return out


def _i8(data: np.ndarray, dev) -> torch.Tensor:
    return torch.from_numpy(np.sign(np.asarray(data)).astype(np.int8)).to(dev)


def _csc_from_rows(idx: np.ndarray, data: np.ndarray, nrows: int) -> sps.csc_array:
    ncols, per = idx.shape
    indptr = np.arange(0, ncols * per + 1, per, dtype=idx.dtype if ncols * per < 2**31 else np.int64)
    return sps.csc_array((data.ravel(), idx.ravel(), indptr), shape=(nrows, ncols))


def structured_grid(dim: int, n: int, perturb: float = 0.0, seed: int = 0, device="cpu",
                    nz: int | None = None, z0: int = 0) -> SimplexGrid:
    """Unit square / cube, ``n`` cells per side, 2-triangle / 6-Kuhn-tet split.

    The analogue of ``pg.unit_grid(dim, 1/n, structured=True, as_mdg=False)``
    followed by ``compute_geometry()`` (``create_grid.py:115-156``).  ``perturb``
    moves interior nodes by ``U(-perturb, perturb)/n`` per coordinate (defeats the
    exact cancellations of the lattice in parity tests).  In 3D ``nz`` / ``z0`` give the slab of
    ``nz`` cube layers starting at layer ``z0`` of a taller box with the same mesh size ``1/n``
    (used by the cell-partitioned multi-GPU path); the lattice keys of its faces / edges are kept
    in ``sd._structured`` so that dofs can be matched across slabs.
    """
    con = structured_connectivity(dim, n, device=device, nz=nz)
    lat = con["lattice"].cpu().numpy().astype(np.float64)
    nodes = np.zeros((3, con["num_nodes"]))
    nodes[:dim] = lat.T / n
    if dim == 3 and z0:
        nodes[2] += z0 / n
    if perturb:
        rng = np.random.default_rng(seed)
        top = np.array([n] * dim) if nz is None or dim == 2 else np.array([n, n, nz])
        interior = np.all((lat > 0) & (lat < top), axis=1)
        nodes[:dim, interior] += rng.uniform(-perturb, perturb, size=(dim, int(interior.sum()))) / n
    cf = con["cell_faces"].cpu().numpy()
    sg = con["cell_face_sign"].cpu().numpy()
    fn = con["face_nodes"].cpu().numpy()
    sd = SimplexGrid.__new__(SimplexGrid)
    sd.dim = dim
    sd.name = f"structured_{dim}d_{n}"
    sd.nodes = nodes
    sd.num_nodes, sd.num_faces, sd.num_cells = con["num_nodes"], con["num_faces"], con["num_cells"]
    sd.face_nodes = _csc_from_rows(fn, np.ones(fn.size, dtype=bool), sd.num_nodes)
    sd.cell_faces = _csc_from_rows(cf, sg.astype(np.float64), sd.num_faces)
    sd.tags = {}
    sd._geometry_done = False
    keep = ("face_ridges", "face_ridge_sign", "ridge_peaks", "face_keys", "edge_keys", "key_bits", "cell_nodes")
    sd._structured = {k: con[k] for k in con if k in keep}
    sd._structured.update(n=n, nz=(n if nz is None else nz), z0=z0)
    return sd


def unit_grid(dim: int, mesh_size: float, as_mdg: bool = True, structured: bool = True, **kwargs):
    """``pg.unit_grid(dim, mesh_size, as_mdg=..., structured=True)`` (``grids/create_grid.py:115-156``):
    the structured simplicial unit square / cube with ``round(1 / mesh_size)`` cells per side.  The
    unstructured (gmsh) branch of the reference is not available."""
    if not structured:
        raise NotImplementedError("unstructured grids need gmsh; use structured=True")
    sd = structured_grid(dim, max(int(np.round(1.0 / mesh_size)), 1), **kwargs)
    if as_mdg:
        from .md_grid import as_mdg as _as_mdg

        return _as_mdg(sd)
    return sd


def reference_element(dim: int) -> SimplexGrid:
    """Reference triangle / tetrahedron with the face numbering and signs of
    ``pg.reference_element`` (``grids/create_grid.py:159-195``): face a is opposite node a,
    nodes are the origin and the unit vectors."""
    if dim not in (2, 3):
        raise ValueError("Dimension must be 2 or 3.")
    nodes = np.eye(AMBIENT_DIM, dim + 1, k=1)
    faces = [[j for j in range(dim + 1) if j != a] for a in range(dim + 1)]
    indices = np.array(faces).ravel()
    indptr = np.arange(0, indices.size + 1, dim)
    face_nodes = sps.csc_array((np.ones(indices.size), indices, indptr))
    signs = np.array([1.0, -1.0, 1.0, -1.0][: dim + 1])
    cell_faces = sps.csc_array(signs[:, None])
    name = "reference_triangle" if dim == 2 else "reference_tetrahedron"
    return SimplexGrid(dim, nodes, face_nodes, cell_faces, name)
