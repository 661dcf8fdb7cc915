"""Device-side assembly engine: grid packing, local matrices (K1), COO->CSR (K2).

torch is used for device memory, streams and copies only; every compute step is a call into
libpgb.so through :mod:`pygeon_b200._lib`.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sps
import torch

from . import _lib
from ._lib import KIND, check, lib

_I32 = torch.int32


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("pygeon_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _to_dev(arr: np.ndarray, dev, dtype=None) -> torch.Tensor:
    """Host numpy -> device tensor (async when the source is pinned), cast on the device."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    t = t.to(dev, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


_SIDE: dict = {}
# copy / compute overlap of the host-in / host-out path (PGB_OVERLAP=0 switches both off: debugging)
overlap_upload = os.environ.get("PGB_OVERLAP", "1") != "0"    # `nodes` goes up on a side stream (pinned sources only)
overlap_download = os.environ.get("PGB_OVERLAP", "1") != "0"  # the pattern comes down on a side stream while K1 / K2 run


def _side_stream(dev) -> "torch.cuda.Stream":
    """One extra stream per device for copies that may overlap the kernels of the current stream."""
    key = torch.device(dev).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(dev)
    return _SIDE[key]


def _to_dev_side(arr: np.ndarray, dev):
    """Upload ``arr`` on the side stream when it is pinned, so that kernels which do not read it
    need not wait for it.  Returns (tensor, event to wait for before the first read | None)."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if not overlap_upload or not t.is_pinned():
        return t.to(dev, non_blocking=True), None
    main, side = torch.cuda.current_stream(dev), _side_stream(dev)
    dst = torch.empty(t.shape, dtype=t.dtype, device=dev)
    side.wait_stream(main)  # the block may have been used by earlier work of the current stream
    with torch.cuda.stream(side):
        dst.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
    dst.record_stream(side)
    return dst, ev


class _PinnedPool:
    """Pinned host blocks for results.  A block is handed out as the memory of the numpy arrays
    returned to the caller (who owns them: the block is not reused while any array derived from it
    is alive) and comes back to the pool through a weakref finalizer.  Page-locking costs
    ~0.5 ms/MB, so recycling blocks is what makes device->host run at PCIe speed; a pageable
    ``.cpu()`` manages ~2 GB/s on this platform."""

    GRAIN = 1 << 22  # 4 MiB size classes

    def __init__(self):
        self.free: dict = {}

    def take(self, nbytes: int) -> torch.Tensor:
        cls = max((int(nbytes) + self.GRAIN - 1) // self.GRAIN, 1) * self.GRAIN
        lst = self.free.get(cls)
        if lst:
            return lst.pop()
        return torch.empty(cls, dtype=torch.uint8).pin_memory()

    def give(self, blk: torch.Tensor) -> None:
        lst = self.free.setdefault(blk.numel(), [])
        if len(lst) < 4:
            lst.append(blk)


_POOL = _PinnedPool()


def _d2h_start(tensors, stream=None):
    """Queue device -> pinned-host copies of ``tensors`` on ``stream`` (default: the current one)."""
    dev = tensors[0].device
    views, blocks = [], []
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream()
        with torch.cuda.stream(st):
            for t in tensors:
                nbytes = t.numel() * t.element_size()
                blk = _POOL.take(nbytes)
                view = blk[:nbytes].view(t.dtype).view(t.shape)
                view.copy_(t, non_blocking=True)
                if stream is not None:
                    t.record_stream(st)
                views.append(view)
                blocks.append(blk)
    return views, blocks, st


def _d2h_finish(pending) -> list:
    """Wait for the copies and hand the pinned blocks to the caller as numpy arrays."""
    import weakref

    views, blocks, st = pending
    st.synchronize()
    out = []
    for view, blk in zip(views, blocks):
        arr = view.numpy()
        # arr.base is the tensor object numpy holds on to; every array derived from `arr` keeps
        # `arr` (hence arr.base) alive, so the block returns to the pool only when all are gone
        weakref.finalize(arr.base, _POOL.give, blk)
        out.append(arr)
    return out


def _to_host(tensors) -> list:
    """Device tensors -> numpy arrays owned by the caller, backed by recycled pinned blocks."""
    return _d2h_finish(_d2h_start(tensors))


def _sign_i8(t: torch.Tensor) -> torch.Tensor:
    if t.dtype == torch.int8:
        return t
    return torch.where(t < 0, -1, 1).to(torch.int8)


# --------------------------------------------------------------------------------------
# raw grid on the device
# --------------------------------------------------------------------------------------
@dataclass
class RawGrid:
    """The grid arrays the engine reads, resident on the device, exactly as the grid object
    stores them: ``cell_faces`` (indices, signs), ``face_nodes.indices``, ``nodes`` (3, nn) and,
    for Nedelec0, ``face_ridges`` (indices, signs) / ``ridge_peaks.indices``."""

    dim: int
    nc: int
    nf: int
    nn: int
    ne: int
    device: torch.device
    R: np.ndarray
    cf_idx: torch.Tensor
    cf_sgn: torch.Tensor | None  # uploaded on demand (need_signs): Lagrange1 never reads it
    fn_idx: torch.Tensor
    nodes: torch.Tensor
    fr_idx: torch.Tensor | None = None
    fr_sgn: torch.Tensor | None = None
    rp_idx: torch.Tensor | None = None
    check: bool = True
    _sd: object = None
    h2d_bytes: int = 0
    nodes_ready: object = None  # event of the side-stream upload of `nodes` (only K1 reads the coordinates)

    @staticmethod
    def upload(sd, device=None, ridges: bool = False, signs: bool = False) -> "RawGrid":
        dev = _device(device)
        d = int(sd.dim)
        if d not in (2, 3):
            raise NotImplementedError("only 2D triangle and 3D tetrahedral grids are supported")
        nc, nf, nn = int(sd.num_cells), int(sd.num_faces), int(sd.num_nodes)
        cf, fn = sd.cell_faces, sd.face_nodes
        if cf.indices.size != nc * (d + 1) or fn.indices.size != nf * d:
            raise NotImplementedError("Grid is not simplicial; cannot compute opposite node.")
        if nc and (np.any(cf.indptr[:: max(nc // 64, 1)] % (d + 1)) or cf.indptr[-1] != nc * (d + 1)):
            raise NotImplementedError("Grid is not simplicial; cannot compute opposite node.")
        R = np.ascontiguousarray(getattr(sd, "rotation_matrix", np.eye(d, 3)), dtype=np.float64)
        nodes_h = np.asarray(sd.nodes, dtype=np.float64)
        with torch.cuda.device(dev):
            cf_idx, fn_idx = _to_dev(cf.indices, dev, _I32), _to_dev(fn.indices, dev, _I32)
            nodes_d, ev = _to_dev_side(nodes_h, dev)  # last, and off the main stream: pack / symbolic do not wait
            raw = RawGrid(
                d, nc, nf, nn, nf if d == 2 else int(getattr(sd, "num_ridges", 0)), dev, R,
                cf_idx, None, fn_idx, nodes_d, _sd=sd, nodes_ready=ev,
            )
        raw.h2d_bytes = cf.indices.nbytes + fn.indices.nbytes + nodes_h.nbytes
        if signs:
            raw.need_signs()
        if ridges:
            raw.need_ridges()
        return raw

    def need_signs(self) -> None:
        if self.cf_sgn is not None:
            return
        data = self._sd.cell_faces.data
        with torch.cuda.device(self.device):
            self.cf_sgn = _sign_i8(_to_dev(data, self.device))
        self.h2d_bytes += data.nbytes

    def need_ridges(self) -> None:
        if self.fr_idx is not None:
            return
        sd, dev, d = self._sd, self.device, self.dim
        cached = getattr(sd, "_device_ridges", None)  # left by SimplexGrid.compute_*(device=...)
        if d == 3 and cached is not None and cached[0].device == dev and cached[2].shape[0] == self.ne:
            self.fr_idx, self.fr_sgn, self.rp_idx = (t.reshape(-1) for t in cached)
            return
        fr = sd.face_ridges
        with torch.cuda.device(dev):
            if d == 2:
                if fr.indices.size != 2 * self.nf:
                    raise NotImplementedError("face_ridges must have 2 entries per face")
                self.fr_idx = _to_dev(fr.indices, dev, _I32)
                self.h2d_bytes += fr.indices.nbytes
            else:
                rp = sd.ridge_peaks
                if fr.indices.size != 3 * self.nf or rp.indices.size != 2 * self.ne:
                    raise NotImplementedError("Grid is not simplicial")
                self.fr_idx = _to_dev(fr.indices, dev, _I32)
                self.fr_sgn = _sign_i8(_to_dev(fr.data, dev))
                self.rp_idx = _to_dev(rp.indices, dev, _I32)
                self.h2d_bytes += fr.indices.nbytes + fr.data.nbytes + rp.indices.nbytes


# --------------------------------------------------------------------------------------
# packed grid
# --------------------------------------------------------------------------------------
@dataclass
class GridPack:
    """Device-resident per-cell tables of one grid (see include/pgb.h, "grid packing")."""

    dim: int
    nc: int
    nf: int
    nn: int
    ne: int
    device: torch.device
    R: np.ndarray
    _coords: torch.Tensor | None  # filled on first use (K1): see `coords`
    cf_idx: torch.Tensor
    fn_idx: torch.Tensor
    cell_nodes: torch.Tensor
    _sign_bits: torch.Tensor | None = None
    _raw: object = None
    _edges: tuple | None = None
    _bdm1: tuple | None = None
    _rt1: tuple | None = None
    _darcy: object = None
    _pwl: object = None
    _rt1cell: object = None
    _l2: object = None
    plans: dict = field(default_factory=dict)

    @staticmethod
    def from_grid(sd, device=None) -> "GridPack":
        """Upload the grid's raw arrays (H2D) and pack them."""
        return GridPack.from_raw(RawGrid.upload(sd, device))

    @staticmethod
    def from_raw(raw: "RawGrid") -> "GridPack":
        """Pack device-resident raw connectivity: coordinates + node-opposite-face tables."""
        dev, d, nc, nf, nn = raw.device, raw.dim, raw.nc, raw.nf, raw.nn
        with torch.cuda.device(dev):
            s = _stream()
            cell_nodes = torch.empty((nc, d + 1), dtype=_I32, device=dev)
            sign_bits = None
            if raw.cf_sgn is not None:
                sign_bits = torch.empty((nc,), dtype=torch.uint8, device=dev)
            err = torch.zeros((1,), dtype=_I32, device=dev)
            check(
                lib.pgb_pack_cells(
                    d, nc, raw.cf_idx.data_ptr(), _ptr(raw.cf_sgn), raw.fn_idx.data_ptr(),
                    cell_nodes.data_ptr(), _ptr(sign_bits), err.data_ptr(), s,
                )
            )
            _lib.count(1)
            if raw.check and int(err.item()):
                raise NotImplementedError("Grid is not simplicial; cannot compute opposite node.")
        return GridPack(d, nc, nf, nn, raw.ne, dev, raw.R, None, raw.cf_idx, raw.fn_idx, cell_nodes,
                        sign_bits, _raw=raw)

    @property
    def coords(self) -> torch.Tensor:
        """Rotated node coordinates (nn, 4 | 2) fp64.  Only K1 reads them, so they are prepared on
        first use: the upload of ``nodes`` (side stream) overlaps the packing and the symbolic phase."""
        if self._coords is None:
            raw, d = self._raw, self.dim
            with torch.cuda.device(self.device):
                if raw.nodes_ready is not None:
                    torch.cuda.current_stream().wait_event(raw.nodes_ready)
                coords = torch.empty((self.nn, 4 if d == 3 else 2), dtype=torch.float64, device=self.device)
                check(lib.pgb_prep_coords(d, self.nn, raw.nodes.data_ptr(), raw.R.ctypes.data, coords.data_ptr(),
                                          _stream()))
                _lib.count(1)
            self._coords = coords
        return self._coords

    @property
    def sign_bits(self) -> torch.Tensor:
        """bit a set <=> cell_faces sign of face slot a is -1 (uploads cell_faces.data on first use)."""
        if self._sign_bits is None:
            raw = self._raw
            raw.need_signs()
            with torch.cuda.device(self.device):
                bits = torch.empty((self.nc,), dtype=torch.uint8, device=self.device)
                check(lib.pgb_pack_signs(self.dim, self.nc, raw.cf_sgn.data_ptr(), bits.data_ptr(), _stream()))
                _lib.count(1)
            self._sign_bits = bits
        return self._sign_bits

    # -- Nedelec0 tables ----------------------------------------------------------------
    def edges(self):
        """(cell_edges (nc, E) int32, edge_bits (nc,) int32-as-uint32)."""
        if self._edges is None:
            raw, dev, d, nc = self._raw, self.device, self.dim, self.nc
            raw.need_ridges()
            with torch.cuda.device(dev):
                s = _stream()
                bits = torch.empty((nc,), dtype=_I32, device=dev)
                if d == 2:
                    check(lib.pgb_pack_edges2d(nc, self.cf_idx.data_ptr(), raw.fr_idx.data_ptr(),
                                               self.cell_nodes.data_ptr(), bits.data_ptr(), s))
                    self._edges = (self.cf_idx.view(nc, 3), bits)
                else:
                    ce = torch.empty((nc, 6), dtype=_I32, device=dev)
                    err = torch.zeros((1,), dtype=_I32, device=dev)
                    check(lib.pgb_pack_edges3d(nc, self.cf_idx.data_ptr(), raw.fr_idx.data_ptr(),
                                               raw.fr_sgn.data_ptr(), raw.rp_idx.data_ptr(),
                                               self.cell_nodes.data_ptr(), ce.data_ptr(),
                                               bits.data_ptr(), err.data_ptr(), s))
                    if raw.check and int(err.item()):
                        raise ValueError("face_ridges / ridge_peaks are inconsistent with the cells")
                    self._edges = (ce, bits)
                _lib.count(1)
        return self._edges

    def bdm1(self):
        """(cell_dofs (nc, d(d+1)) int32, jslot_bits (nc,))."""
        if self._bdm1 is None:
            d, nc, dev = self.dim, self.nc, self.device
            with torch.cuda.device(dev):
                dofs = torch.empty((nc, d * (d + 1)), dtype=_I32, device=dev)
                bits = torch.empty((nc,), dtype=_I32, device=dev)
                check(lib.pgb_pack_bdm1(d, nc, self.nf, self.cf_idx.data_ptr(), self.fn_idx.data_ptr(),
                                        self.cell_nodes.data_ptr(), dofs.data_ptr(), bits.data_ptr(),
                                        _stream()))
                _lib.count(1)
            self._bdm1 = (dofs, bits)
        return self._bdm1

    def rt1(self):
        """(cell_dofs (nc, d(d+2)) int32 in the reference's local order, rank / sign bits (nc,))."""
        if self._rt1 is None:
            d, nc, dev = self.dim, self.nc, self.device
            with torch.cuda.device(dev):
                dofs = torch.empty((nc, d * (d + 2)), dtype=_I32, device=dev)
                bits = torch.empty((nc,), dtype=_I32, device=dev)
                check(lib.pgb_pack_rt1(d, nc, self.nf, self.cf_idx.data_ptr(), self.cell_nodes.data_ptr(),
                                       self.sign_bits.data_ptr(), dofs.data_ptr(), bits.data_ptr(), _stream()))
                _lib.count(1)
            self._rt1 = (dofs, bits)
        return self._rt1

    # -- dof tables per space --------------------------------------------------------------
    def dofs(self, space: str):
        """(cell_dofs (nc, L) int32 tensor, ndof)."""
        d = self.dim
        if space == "lagrange1":
            return self.cell_nodes, self.nn
        if space == "rt0":
            return self.cf_idx.view(self.nc, d + 1), self.nf
        if space == "nedelec0":
            return self.edges()[0], self.ne
        if space == "bdm1":
            return self.bdm1()[0], d * self.nf
        if space == "rt1":
            return self.rt1()[0], d * (self.nf + self.nc)
        if space == "lagrange2":  # nodes in slot order, then num_nodes + edge (p, q), pairs in lexicographic order
            if self._l2 is None:
                ce = self.cf_idx if d == 2 else self.edges()[0]
                with torch.cuda.device(self.device):
                    L = (d + 1) + d * (d + 1) // 2
                    out = torch.empty((self.nc, L), dtype=_I32, device=self.device)
                    check(lib.pgb_pack_l2_dofs(d, self.nc, self.nn, self.cell_nodes.data_ptr(), ce.data_ptr(),
                                               out.data_ptr(), _stream()))
                    _lib.count(1)
                self._l2 = out
            return self._l2, self.nn + self.ne
        if space == "rt1cell":  # the d cell functions of RT1, numbered (k, c) -> k * nc + c inside their block
            if self._rt1cell is None:
                cells = torch.arange(self.nc, dtype=_I32, device=self.device)
                self._rt1cell = (cells[:, None] + self.nc * torch.arange(d, dtype=_I32, device=self.device)[None, :]).contiguous()
            return self._rt1cell, d * self.nc
        if space == "pwlinears":  # broken P1: dof (slot k, cell c) = k * num_cells + c (fem/l2.py:38-49)
            if self._pwl is None:
                cells = torch.arange(self.nc, dtype=_I32, device=self.device)
                self._pwl = (cells[:, None] + self.nc * torch.arange(d + 1, dtype=_I32, device=self.device)[None, :]).contiguous()
            return self._pwl, (d + 1) * self.nc
        if space == "darcy":  # RT0 x P0: faces of the cell, then num_faces + cell
            if self._darcy is None:
                cells = torch.arange(self.nc, dtype=_I32, device=self.device) + self.nf
                self._darcy = torch.cat([self.cf_idx.view(self.nc, d + 1), cells[:, None]], dim=1).contiguous()
            return self._darcy, self.nf + self.nc
        raise KeyError(space)


# --------------------------------------------------------------------------------------
# symbolic plan + device CSR
# --------------------------------------------------------------------------------------
@dataclass
class CsrPlan:
    """Result of the symbolic phase, reusable by every operator of a space on the grid.

    mode 0 (two level): K1 writes the local rows to ``inc_pos`` (row-sorted order) and the
    streaming numeric kernel reduces them through ``row_start`` / ``slots``;
    mode 1 (full sort): ``perm, seg`` drive the segmented-reduce numeric kernel."""

    n_rows: int
    n_coo: int
    nnz: int
    L: int
    mode: int
    indptr: torch.Tensor
    indices: torch.Tensor
    perm: torch.Tensor | None = None  # mode 1: uint32 stored as int32
    seg: torch.Tensor | None = None
    inc_pos: torch.Tensor | None = None  # mode 0: row-sorted position of every local row (for K1)
    row_start: torch.Tensor | None = None
    slots: torch.Tensor | None = None
    max_row_nnz: int = 0


@dataclass
class DeviceCSR:
    """A CSR matrix resident on the GPU (int32 indices; indptr int32 or int64; fp64 data)."""

    shape: tuple
    indptr: torch.Tensor
    indices: torch.Tensor
    data: torch.Tensor
    # True: the matrix equals its transpose (symmetric coefficient tensor), so the arrays may be read
    # as CSR or CSC alike.  False: the arrays are the CSR arrays of the matrix (what SpMV / CG /
    # MINRES / reduce_system consume); a CSC host copy then needs a real transposition.
    symmetric: bool = True

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())

    @property
    def idx64(self) -> int:
        return 1 if self.indptr.dtype == torch.int64 else 0

    def to_scipy(self, fmt: str = "csc"):
        """Host copy owned by the caller.  ``fmt='csc'`` reinterprets the CSR arrays as CSC (valid
        because the engine builds the transpose-consistent matrix, see pgb.h K1 note).

        Device->host goes through a persistent pinned staging buffer (a pageable ``.cpu()`` runs at
        ~2 GB/s on this platform, pinned at ~55 GB/s); the fresh numpy arrays are then filled by a
        multi-threaded host copy."""
        cls = sps.csc_array if fmt == "csc" else sps.csr_array
        if self.shape[0] == 0:
            return cls(self.shape)
        ip, ix, da = _to_host([self.indptr, self.indices, self.data])
        if fmt == "csc" and not self.symmetric:  # arrays are CSR of a non-symmetric matrix
            return sps.csr_array((da, ix, ip), shape=self.shape).tocsc()
        return cls((da, ix, ip), shape=self.shape)

    def diagonal(self) -> torch.Tensor:
        """Diagonal entries (``pgb_csr_diagonal``; 0 for a row without one)."""
        n = self.shape[0]
        dev = self.data.device
        out = torch.empty(n, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib.pgb_csr_diagonal(n, self.indptr.data_ptr(), self.idx64, self.indices.data_ptr(),
                                       self.data.data_ptr(), out.data_ptr(), _stream()))
            _lib.count(1)
        return out

    def matvec(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """``A @ x`` on the device (``pgb_spmv``)."""
        n = self.shape[0]
        y = torch.empty(n, dtype=torch.float64, device=self.data.device) if out is None else out
        with torch.cuda.device(self.data.device):
            check(lib.pgb_spmv(n, self.indptr.data_ptr(), self.idx64, self.indices.data_ptr(), self.data.data_ptr(),
                               x.data_ptr(), y.data_ptr(), 0, _stream()))
            _lib.count(1)
        return y

    def extract(self, rows: torch.Tensor | None, col_map: torch.Tensor | None, n_cols: int) -> "DeviceCSR":
        """Rows ``rows`` (int32 ids, None = all) with the columns renumbered through ``col_map`` (int32,
        -1 = dropped, None = unchanged): count -> scan -> fill kernels, O(rows) scratch."""
        rows_out = self.shape[0] if rows is None else int(rows.numel())
        ip, ix, da = extract_rows(self.indptr, self.indices, self.data, rows, col_map)
        return DeviceCSR((rows_out, n_cols), ip, ix, da, symmetric=self.symmetric)

    def transpose(self) -> "DeviceCSR":
        """CSR arrays of the transpose (``pgb_csr_transpose``: stable radix sort by column, deterministic)."""
        m, n = self.shape
        if self.symmetric and m == n:
            return self
        dev = self.data.device
        nnz = self.nnz
        it = torch.int64 if (nnz >= 2**31 or force_idx64) else _I32
        ip = torch.empty(n + 1, dtype=it, device=dev)
        ix = torch.empty(nnz, dtype=_I32, device=dev)
        da = torch.empty(nnz, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ws_bytes = int(lib.pgb_csr_transpose_workspace_bytes(nnz))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib.pgb_csr_transpose(m, n, nnz, self.indptr.data_ptr(), self.idx64, self.indices.data_ptr(),
                                        self.data.data_ptr(), ip.data_ptr(), 1 if it == torch.int64 else 0,
                                        ix.data_ptr(), da.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
            _lib.count(2 + 5 * (((max(n, 2) - 1).bit_length() + 7) // 8))  # per radix pass: histogram, 3 scan kernels, scatter
        return DeviceCSR((n, m), ip, ix, da, symmetric=False)

    def add_diagonal(self, rows: torch.Tensor, vals: torch.Tensor) -> None:
        """In place ``A[rows[i], rows[i]] += vals[i]`` for distinct rows whose diagonal entry is stored
        (``pgb_csr_add_diagonal``)."""
        dev = self.data.device
        rows = rows.to(device=dev, dtype=_I32).contiguous()
        vals = vals.to(device=dev, dtype=torch.float64).contiguous()
        with torch.cuda.device(dev):
            flag = torch.empty(1, dtype=_I32, device=dev)
            check(lib.pgb_csr_add_diagonal(int(rows.numel()), rows.data_ptr(), vals.data_ptr(), self.indptr.data_ptr(),
                                           self.idx64, self.indices.data_ptr(), self.data.data_ptr(), flag.data_ptr(),
                                           _stream()))
            _lib.count(1)

    @staticmethod
    def from_scipy(A, device=None) -> "DeviceCSR":
        dev = _device(device)
        A = sps.csr_array(A)
        A.sum_duplicates()
        A.sort_indices()
        it = torch.int64 if A.nnz >= 2**31 else _I32
        return DeviceCSR(
            tuple(A.shape),
            _to_dev(A.indptr, dev, it),
            _to_dev(A.indices, dev, _I32),
            _to_dev(A.data.astype(np.float64, copy=False), dev),
            symmetric=False,  # unknown: the arrays are CSR, a CSC copy is made by transposition
        )


def device_bmat(blocks, device=None, symmetric: bool = False) -> DeviceCSR:
    """``sps.block_array(blocks).tocsr()`` in HBM (``utils/bmat.py``; the block operators of
    ``numerics/innerproducts.py:162-232`` / ``differentials.py:144-192``): ``blocks`` is a 2-D nested list
    whose entries are ``DeviceCSR``, ``(DeviceCSR, scale)`` or ``None``.  Count -> scan -> place kernels
    (``pgb_csr_block_lengths`` / ``pgb_scan_i64`` / ``pgb_csr_block_place``), O(rows) scratch."""
    grid = [[(b, 1.0) if isinstance(b, DeviceCSR) else b for b in row] for row in blocks]
    nbr, nbc = len(grid), len(grid[0])
    if any(len(row) != nbc for row in grid):
        raise ValueError("device_bmat: ragged block grid")
    heights, widths = [None] * nbr, [None] * nbc
    dev = None
    for i, row in enumerate(grid):
        for j, b in enumerate(row):
            if b is None:
                continue
            A = b[0]
            dev = A.data.device if dev is None else dev
            if A.data.device != dev:
                raise ValueError("device_bmat: blocks on different devices")
            for sizes, k, v in ((heights, i, A.shape[0]), (widths, j, A.shape[1])):
                if sizes[k] is not None and sizes[k] != v:
                    raise ValueError(f"device_bmat: block ({i}, {j}) has shape {A.shape}, expected {heights[i]} x {widths[j]}")
                sizes[k] = v
    if dev is None or None in heights or None in widths:
        raise ValueError("device_bmat: a block row or column holds no block (its size is unknown)")
    row0 = np.concatenate([[0], np.cumsum(heights)]).astype(np.int64)
    col0 = np.concatenate([[0], np.cumsum(widths)]).astype(np.int64)
    n, m = int(row0[-1]), int(col0[-1])
    if n >= 2**31 or m >= 2**31:
        raise ValueError("device_bmat: dimensions must be < 2^31")
    with torch.cuda.device(dev):
        s = _stream()
        lens = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        nl = 0
        for i, row in enumerate(grid):
            for b in row:
                if b is not None and b[0].shape[0]:
                    A = b[0]
                    check(lib.pgb_csr_block_lengths(A.shape[0], A.indptr.data_ptr(), A.idx64,
                                                    lens.data_ptr() + 8 * int(row0[i]), s))
                    nl += 1
        ws_bytes = int(lib.pgb_system_workspace_bytes(0))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ip = torch.empty(n + 1, dtype=torch.int64, device=dev)
        tot = C.c_int64(0)
        check(lib.pgb_scan_i64(n, lens.data_ptr(), ip.data_ptr(), C.byref(tot), ws.data_ptr(), ws_bytes, s))
        nnz = int(tot.value)
        ix = torch.empty(nnz, dtype=_I32, device=dev)
        da = torch.empty(nnz, dtype=torch.float64, device=dev)
        cursor = lens  # reuse: the cursors start at the row starts
        cursor.copy_(ip)
        for i, row in enumerate(grid):
            for j, b in enumerate(row):  # left to right: the columns of a row stay sorted
                if b is not None and b[0].shape[0]:
                    A, scale = b
                    check(lib.pgb_csr_block_place(A.shape[0], A.indptr.data_ptr(), A.idx64, A.indices.data_ptr(),
                                                  A.data.data_ptr(), int(col0[j]), float(scale),
                                                  cursor.data_ptr() + 8 * int(row0[i]), ix.data_ptr(), da.data_ptr(), s))
        _lib.count(3 * nl + 3)
        if nnz < 2**31 and not force_idx64:
            ip = ip.to(_I32)
    return DeviceCSR((n, m), ip, ix, da, symmetric=symmetric)


def mask_to_map(mask: torch.Tensor, want_newid: bool = True):
    """(newid int32 (n,) with -1 for dropped entries | None, kept ids int32 (count,)) of a bool mask
    (``pgb_mask_to_map``)."""
    dev = mask.device
    n = int(mask.numel())
    m8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else (mask != 0).view(torch.uint8)
    with torch.cuda.device(dev):
        s = _stream()
        ws_bytes = int(lib.pgb_system_workspace_bytes(n))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        cnt = C.c_int64(0)
        check(lib.pgb_mask_to_map(n, m8.data_ptr(), 0, 0, C.byref(cnt), ws.data_ptr(), ws_bytes, s))
        newid = torch.empty(n, dtype=_I32, device=dev) if want_newid else None
        rows = torch.empty(int(cnt.value), dtype=_I32, device=dev)
        check(lib.pgb_mask_to_map(n, m8.data_ptr(), _ptr(newid), rows.data_ptr(), None, ws.data_ptr(), ws_bytes, s))
        _lib.count(7)
    return newid, rows


def extract_rows(indptr, indices, data, rows, col_map):
    """``pgb_csr_extract_symbolic`` + ``pgb_csr_extract_fill``: (indptr, indices, data) of the selected rows."""
    dev = indices.device
    n_out = int(indptr.numel()) - 1 if rows is None else int(rows.numel())
    idx64 = 1 if indptr.dtype == torch.int64 else 0
    with torch.cuda.device(dev):
        s = _stream()
        ws_bytes = int(lib.pgb_system_workspace_bytes(n_out))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ip = torch.empty(n_out + 1, dtype=torch.int64, device=dev)
        nnz = C.c_int64(0)
        check(lib.pgb_csr_extract_symbolic(n_out, _ptr(rows), indptr.data_ptr(), idx64, indices.data_ptr(),
                                           _ptr(col_map), ip.data_ptr(), C.byref(nnz), ws.data_ptr(), ws_bytes, s))
        del ws
        nnz = int(nnz.value)
        ix = torch.empty(nnz, dtype=_I32, device=dev)
        da = torch.empty(nnz, dtype=torch.float64, device=dev)
        check(lib.pgb_csr_extract_fill(n_out, _ptr(rows), indptr.data_ptr(), idx64, indices.data_ptr(),
                                       data.data_ptr(), _ptr(col_map), ip.data_ptr(), ix.data_ptr(), da.data_ptr(), s))
        _lib.count(5)
        if nnz < 2**31 and not force_idx64:
            ip = ip.to(_I32)
    return ip, ix, da


# K2 numeric variant (pgb.h): 0 = one thread per row (the fastest for every space measured, profiles/r2_numeric_ab.txt),
# bit 1: bulk-async staging of the block's local rows, bit 2: lane groups per row
numeric_flags = int(os.environ.get("PGB_NUMERIC_FLAGS", "0"))
symbolic_mode = 0  # 0 = two-level (default), 1 = full (row,col) radix sort; see include/pgb.h
force_idx64 = False  # tests: emit int64 indptr even when nnz < 2^31


_TOTAL_MEM: dict = {}


def _total_memory(dev) -> int:
    key = torch.device(dev).index or 0
    if key not in _TOTAL_MEM:
        _TOTAL_MEM[key] = int(torch.cuda.get_device_properties(key).total_memory)
    return _TOTAL_MEM[key]


def symbolic(pack: GridPack, space: str, mode: int | None = None, overlap=None) -> CsrPlan:
    """Sort-by-(row,col) of the cell->dof incidences: CSR pattern + the gather plan.

    ``overlap(partial_plan)``: called (two-level modes, plan not cached) once ``inc_pos`` is complete in stream
    order but before the host waits for ``nnz``: what it queues - K1, which needs ``inc_pos`` only - keeps the GPU
    busy while the host allocates ``indptr`` / ``indices`` and launches the compaction.  The partial plan has
    ``mode == 0`` and no pattern yet; when the call has to fall back to the full sort afterwards, the caller
    sees ``plan.mode == 1`` and must discard what the hook produced."""
    mode = symbolic_mode if mode is None else mode
    key = ("plan", space)
    if key in pack.plans:
        return pack.plans[key]
    dofs, n_rows = pack.dofs(space)
    nc, L = int(dofs.shape[0]), int(dofs.shape[1])
    n_coo, n_inc = nc * L * L, nc * L
    dev = pack.device
    with torch.cuda.device(dev):
        s = _stream()
        dofs = dofs.contiguous()
        while True:
            ws_bytes = int(lib.pgb_csr_workspace_bytes(nc, L, n_rows, mode))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            if mode == 1:
                first = torch.empty((max(n_coo, 1),), dtype=_I32, device=dev)  # perm
                slots = row_start = None
            else:
                first = torch.empty((max(n_inc, 1),), dtype=_I32, device=dev)  # inc_pos
                slots = torch.zeros((max(n_inc * ((L + 3) & ~3), 1),), dtype=torch.uint8, device=dev)
                row_start = torch.empty((n_rows + 1,), dtype=_I32, device=dev)
            nnz, mx = C.c_int64(0), C.c_int(0)
            if overlap is not None and mode != 1:
                rc = lib.pgb_csr_symbolic_sort(n_rows, nc, L, dofs.data_ptr(), mode, ws.data_ptr(), ws_bytes,
                                               first.data_ptr(), _ptr(slots), _ptr(row_start), None, None, s)
                if rc == 0:
                    # the hook's output (the local rows) then lives beside the workspace: only when both are small
                    # against the device memory (no cudaMemGetInfo here: it costs more than the overlap gains)
                    if 4 * (ws_bytes + n_inc * ((L + 3) & ~3) * 8) < _total_memory(dev):
                        overlap(CsrPlan(n_rows, n_coo, -1, L, 0, None, None, inc_pos=first, row_start=row_start,
                                        slots=slots, max_row_nnz=0))
                    rc = lib.pgb_csr_symbolic_result(C.byref(nnz), C.byref(mx))
            else:
                rc = lib.pgb_csr_symbolic_sort(n_rows, nc, L, dofs.data_ptr(), mode, ws.data_ptr(), ws_bytes,
                                               first.data_ptr(), _ptr(slots), _ptr(row_start), C.byref(nnz),
                                               C.byref(mx), s)
            if rc == 3 and mode != 1:  # a dof shared by too many cells / a row too long for the register sort
                mode = 1
                del ws, first, slots, row_start
                continue
            check(rc)
            break
        nnz = int(nnz.value)
        it = torch.int64 if (nnz >= 2**31 or force_idx64) else _I32
        indptr = torch.empty((n_rows + 1,), dtype=it, device=dev)
        indices = torch.empty((nnz,), dtype=_I32, device=dev)
        seg = torch.empty((nnz + 1,), dtype=_I32, device=dev) if mode == 1 else None
        check(lib.pgb_csr_symbolic_finish(n_rows, nc, L, mode, nnz, ws.data_ptr(), _ptr(row_start),
                                          indptr.data_ptr(), 1 if it == torch.int64 else 0, indices.data_ptr(),
                                          _ptr(seg), s))
        bits = max(1, int(n_rows - 1).bit_length()) if n_rows > 1 else 1
        if mode == 1:
            _lib.count(((2 * bits + 7) // 8) * 5 + 3)
        elif mode == 4:
            _lib.count(((bits + 7) // 8) * 5 + 7)
        else:
            _lib.count(12)  # count, 3 scan, max, place, (canon,) row sort|hash, 3 scan, compact (+memsets)
        del ws
    if mode == 1:
        plan = CsrPlan(n_rows, n_coo, nnz, L, 1, indptr, indices, perm=first, seg=seg)
    else:
        plan = CsrPlan(n_rows, n_coo, nnz, L, 0, indptr, indices, inc_pos=first, row_start=row_start,
                       slots=slots, max_row_nnz=int(mx.value))
    pack.plans[key] = plan
    return plan


_FORMS = {
    ("lagrange1", "mass"): "l1_mass",
    ("lagrange1", "stiff"): "l1_stiff",
    ("lagrange1", "stiff_rot"): "l1_stiff_rot",
    ("rt0", "mass"): "rt0_mass",
    ("bdm1", "mass"): "bdm1_mass",
    ("nedelec0", "mass"): "n0_mass",
    ("nedelec0", "stiff"): "n0_curlcurl",
    ("nedelec0", "stiff2d"): "rt0_divdiv",
    ("rt0", "stiff"): "rt0_divdiv",
    ("bdm1", "stiff"): "bdm1_divdiv",
    ("darcy", "saddle"): "darcy_saddle",
    ("lagrange1", "lumped"): "l1_lumped",
    ("lagrange2", "mass"): "l2_mass",
    ("lagrange2", "stiff"): "l2_stiff",
    ("rt0", "lumped"): "rt0_lumped",
    ("bdm1", "lumped"): "bdm1_lumped",
    ("rt1", "mass"): "rt1_mass",
    ("rt1", "stiff"): "rt1_divdiv",
    ("rt1cell", "lumped"): "rt1_lumped_cell",
    ("pwlinears", "mass"): "l1_mass",
    ("pwlinears", "lumped"): "l1_lumped",
}


def local_matrices(pack: GridPack, kind: str, w=None, K=None, out=None, plan: "CsrPlan | None" = None) -> torch.Tensor:
    """K1: vals (nc, L*L) on the device.  ``w`` (nc,) / ``K`` (3,3,nc) are device tensors or None.
    With a mode-0 ``plan`` the local rows are written in the plan's row-sorted order (what
    :func:`numeric` expects for that plan); without, in plain [cell][a][b] order."""
    d, nc, dev = pack.dim, pack.nc, pack.device
    L = int(lib.pgb_local_size(KIND[kind], d))
    aux = None
    needs_sign = kind in ("rt0_mass", "bdm1_mass", "bdm1_lumped", "n0_curlcurl", "rt0_divdiv", "bdm1_divdiv",
                          "darcy_saddle")
    sign_bits = pack.sign_bits if needs_sign else None
    if kind in ("n0_mass",) or (kind == "n0_curlcurl" and d == 3):
        aux = pack.edges()[1]
    elif kind in ("bdm1_mass", "bdm1_lumped"):
        aux = pack.bdm1()[1]
    elif kind in ("rt1_mass", "rt1_divdiv", "rt1_lumped_cell"):
        aux = pack.rt1()[1]
    with torch.cuda.device(dev):
        rowmajor = plan is not None and plan.mode == 0
        pitch = ((L + 3) & ~3) if rowmajor else L  # row-sorted layout: rows padded to whole 32-byte sectors
        vals = out if out is not None else torch.empty((nc, L * pitch), dtype=torch.float64, device=dev)
        check(
            lib.pgb_local_matrices(
                KIND[kind], d, nc, pack.cell_nodes.data_ptr(), _ptr(sign_bits), _ptr(aux),
                pack.coords.data_ptr(), _ptr(K), pack.R.ctypes.data, _ptr(w),
                _ptr(plan.inc_pos) if (plan is not None and plan.mode == 0) else 0, vals.data_ptr(), _stream(),
            )
        )
        _lib.count(1)
    return vals


def numeric(plan: CsrPlan, vals: torch.Tensor, out=None) -> torch.Tensor:
    """K2 numeric phase: local values -> CSR data through the plan (fixed summation order)."""
    dev = vals.device
    with torch.cuda.device(dev):
        data = out if out is not None else torch.empty((plan.nnz,), dtype=torch.float64, device=dev)
        if plan.mode == 1:
            check(lib.pgb_csr_numeric(plan.nnz, plan.perm.data_ptr(), plan.seg.data_ptr(), vals.data_ptr(),
                                      data.data_ptr(), _stream()))
        else:
            # coalesced output through shared memory when the matrix has about as many entries as
            # COO candidates (RT0, BDM1, N0, Darcy); not for P1 in 3D (6 candidates per entry)
            staged = (1 if 3 * plan.nnz > plan.n_coo else 0) | (numeric_flags & (2 | 64))
            if numeric_flags & 4:  # G lanes per row from the average number of local rows per CSR row
                avg = plan.n_coo / plan.L / max(plan.n_rows, 1)
                staged |= (8 if avg >= 12 else 4 if avg >= 6 else 2 if avg >= 3 else 0) << 2
            check(lib.pgb_csr_numeric_rows(plan.n_rows, plan.L, plan.max_row_nnz, plan.row_start.data_ptr(),
                                           plan.slots.data_ptr(), plan.indptr.data_ptr(),
                                           1 if plan.indptr.dtype == torch.int64 else 0,
                                           vals.data_ptr(), data.data_ptr(), staged, _stream()))
        _lib.count(1)
    return data


_CACHE_ATTR = "_pgb200_pack"
cache_enabled = True


def get_pack(sd, device=None) -> GridPack:
    """Packed grid, cached on the grid object (the reference caches per grid too:
    ``sd.opposite_nodes`` grid.py:291-307 and ``functools.cache`` hdiv.py:255,450)."""
    dev = _device(device)
    pack = getattr(sd, _CACHE_ATTR, None) if cache_enabled else None
    if pack is None or pack.device != dev:
        pack = GridPack.from_grid(sd, dev)
        if cache_enabled:
            try:
                setattr(sd, _CACHE_ATTR, pack)
            except AttributeError:
                pass
    return pack


def clear_cache(sd) -> None:
    if hasattr(sd, _CACHE_ATTR):
        delattr(sd, _CACHE_ATTR)


def assemble_on_pack(pack: GridPack, space: str, kind: str, wd=None, Kd=None):
    """symbolic (cached per pack) -> K1 (writing the local rows where the plan wants them) -> K2."""
    early = []  # cold plan: K1 is queued as soon as inc_pos exists, ahead of the host's wait for nnz
    plan = symbolic(pack, space, overlap=lambda partial: early.append(local_matrices(pack, kind, wd, Kd, plan=partial)))
    vals = early[0] if (early and plan.mode == 0) else local_matrices(pack, kind, wd, Kd, plan=plan)
    return plan, numeric(plan, vals)


def tensor_is_symmetric(K) -> bool:
    """True when every (3, 3) block of ``K`` (3, 3, nc) equals its transpose."""
    return K is None or all(np.array_equal(K[i, j], K[j, i]) for i, j in ((0, 1), (0, 2), (1, 2)))


def assemble_device(space: str, form: str, sd, w=None, K=None, device=None) -> DeviceCSR:
    """Full assembly on the device: pack (cached) -> K1 -> K2.  ``w``/``K`` are host numpy
    arrays in the reference's layout or None.

    The result is consumed as **CSR** on the device (SpMV, CG, MINRES, ``reduce_system``).  K1 builds
    ``A[a][b] = u_a^T (R K R^T) u_b`` whose CSR arrays are the CSC arrays of the reference's matrix
    (its contraction yields ``R K^T R^T``, fem/vec_l2.py:113-114; pgb.h K1 note), so a
    non-symmetric tensor goes in transposed here: the arrays are then the CSR arrays of the
    reference's matrix itself."""
    pack = get_pack(sd, device)
    dev = pack.device
    kind = _FORMS[(space, form)]
    sym = tensor_is_symmetric(K)
    if not sym:
        K = np.ascontiguousarray(np.transpose(K, (1, 0, 2)))
    wd = None if w is None else _to_dev(w, dev)
    Kd = None if K is None else _to_dev(K, dev)
    plan, data = assemble_on_pack(pack, space, kind, wd, Kd)
    n = plan.n_rows
    return DeviceCSR((n, n), plan.indptr, plan.indices, data, symmetric=sym)


def assemble_host(space: str, form: str, sd, w=None, K=None, device=None, fmt: str = "csc"):
    """The public host-in / host-out call: upload, assemble, download, with the copies overlapped
    where the data flow allows it: ``nodes`` goes up on a side stream while the grid is packed and
    the pattern is built (only K1 reads coordinates); indptr / indices come down on the side stream
    while K1 and the numeric phase run; the values follow on the main stream."""
    pack = get_pack(sd, device)
    dev = pack.device
    kind = _FORMS[(space, form)]
    cls = sps.csc_array if fmt == "csc" else sps.csr_array
    if fmt != "csc" and not tensor_is_symmetric(K):  # CSR of the reference matrix: see assemble_device
        K = np.ascontiguousarray(np.transpose(K, (1, 0, 2)))
    with torch.cuda.device(dev):
        wd = None if w is None else _to_dev(w, dev)
        Kd = None if K is None else _to_dev(K, dev)
        plan = symbolic(pack, space)  # ends with a stream synchronisation: the pattern is complete
        n = plan.n_rows
        if n == 0 or plan.nnz == 0:
            return DeviceCSR((n, n), plan.indptr, plan.indices, numeric(plan, local_matrices(pack, kind, wd, Kd, plan=plan))).to_scipy(fmt)
        side = _side_stream(dev) if overlap_download else None
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
        pat = _d2h_start([plan.indptr, plan.indices], side)
        data = numeric(plan, local_matrices(pack, kind, wd, Kd, plan=plan))
        val = _d2h_start([data])
        ip, ix = _d2h_finish(pat)
        (da,) = _d2h_finish(val)
    return cls((da, ix, ip), shape=(n, n))


def cell_volumes_inverse_weighted(sd, w=None, device=None) -> torch.Tensor:
    """diag of the P0 mass matrix, w/|c| (fem/l2.py:335-353)."""
    pack = get_pack(sd, device)
    wd = None if w is None else _to_dev(w, pack.device)
    return local_matrices(pack, "p0_mass", wd, None).view(-1)
