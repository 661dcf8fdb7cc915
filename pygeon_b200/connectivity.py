"""Grid connectivity and geometry on the device.

Device counterparts of ``pg.Grid._compute_ridges_3d`` (``grids/grid.py:149-216``),
``pp.Grid.compute_geometry`` for simplices (called from ``grids/grid.py:59``) and
``pg.Grid.compute_edge_properties`` (``grids/grid.py:311-337``).  They are the input-provider side
of the hot path: the assembly engine reads ``face_ridges`` / ``ridge_peaks`` (Nedelec0, Lagrange2),
and at 10^8 cells the numpy versions of these steps take longer than the assembly itself.

All functions take and return device tensors; :class:`pygeon_b200.grid.SimplexGrid` calls them when
``compute_geometry(device=...)`` / ``compute_connectivity(device=...)`` is given a CUDA device.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib
from .engine import _device, _ptr, _stream


def ridges3d(face_nodes: torch.Tensor, num_nodes: int):
    """``face_nodes`` int32 ``(nf, 3)`` on the device (``face_nodes.indices``) ->
    ``(face_ridges (nf, 3) int32, face_ridge_sign (nf, 3) int8, ridge_peaks (nr, 2) int32)``
    with the reference's lexicographic ridge numbering."""
    dev = _device(face_nodes.device)
    fn = face_nodes.to(torch.int32).contiguous()
    if fn.ndim != 2 or fn.shape[1] != 3:
        raise NotImplementedError("faces must be triangles")
    nf = fn.shape[0]
    with torch.cuda.device(dev):
        fr = torch.empty((nf, 3), dtype=torch.int32, device=dev)
        fs = torch.empty((nf, 3), dtype=torch.int8, device=dev)
        rp = torch.empty((3 * nf, 2), dtype=torch.int32, device=dev)
        nbytes = lib.pgb_ridges3d_workspace_bytes(nf)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        nr = C.c_int64(0)
        check(lib.pgb_ridges3d(nf, int(num_nodes), fn.data_ptr(), fr.data_ptr(), fs.data_ptr(), rp.data_ptr(),
                               C.byref(nr), ws.data_ptr(), nbytes, _stream()))
        _lib.count(10 + 3 * 4 * ((2 * max(int(num_nodes) - 1, 1).bit_length() + 7) // 8))
        rp = rp[: nr.value].clone()
    return fr, fs, rp


def geometry(dim: int, nodes: torch.Tensor, cf_idx: torch.Tensor, cf_sgn, fn_idx: torch.Tensor,
             cell_nodes: torch.Tensor) -> dict:
    """Metric quantities of a simplicial grid.  ``nodes`` f64 ``(3, nn)``; ``cf_idx`` int32
    ``(nc, d+1)``; ``cf_sgn`` int8 ``(nc, d+1)`` or None (no orientation check); ``fn_idx`` int32
    ``(nf, d)``; ``cell_nodes`` int32 ``(nc, d+1)`` (``GridPack.cell_nodes``).
    Raises ``ValueError`` when the ``cell_faces`` signs contradict the node-order normals."""
    dev = _device(nodes.device)
    nn, nc, nf = nodes.shape[1], cf_idx.shape[0], fn_idx.shape[0]
    f64 = dict(dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        out = {
            "face_normals": torch.empty((3, nf), **f64),
            "face_areas": torch.empty(nf, **f64),
            "face_centers": torch.empty((3, nf), **f64),
            "cell_centers": torch.empty((3, nc), **f64),
            "cell_volumes": torch.empty(nc, **f64),
        }
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.pgb_grid_geometry(int(dim), nc, nf, nn, nodes.data_ptr(), cf_idx.data_ptr(), _ptr(cf_sgn),
                                    fn_idx.data_ptr(), cell_nodes.data_ptr(), out["face_normals"].data_ptr(),
                                    out["face_areas"].data_ptr(), out["face_centers"].data_ptr(),
                                    out["cell_centers"].data_ptr(), out["cell_volumes"].data_ptr(),
                                    err.data_ptr(), _stream()))
        _lib.count(2)
        if cf_sgn is not None and int(err.item()):
            raise ValueError("cell_faces signs are inconsistent with the node-order normals")
    return out


def edge_geometry(nodes: torch.Tensor, edge_nodes: torch.Tensor, edge_signs=None):
    """``(edge_tangents (3, ne), edge_lengths (ne,))`` for ``edge_nodes`` int32 ``(ne, 2)``."""
    dev = _device(nodes.device)
    ne = edge_nodes.shape[0]
    with torch.cuda.device(dev):
        t = torch.empty((3, ne), dtype=torch.float64, device=dev)
        ln = torch.empty(ne, dtype=torch.float64, device=dev)
        check(lib.pgb_edge_geometry(ne, nodes.shape[1], nodes.data_ptr(), edge_nodes.data_ptr(), _ptr(edge_signs),
                                    t.data_ptr(), ln.data_ptr(), _stream()))
        _lib.count(1)
    return t, ln
