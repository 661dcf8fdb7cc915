"""Cell-partitioned multi-GPU path.

One process per GPU.  The grid is partitioned **by cells** into contiguous z-slabs; a dof (node,
face, edge) is owned by the lowest rank that has a cell touching it.  Every rank assembles all
rows it owns by also evaluating the single layer of halo cells that touch owned dofs - redundant
compute, **no communication during assembly**.  The result is a row-block-distributed CSR whose
columns are renumbered ``[owned | ghost]``; the distributed solve needs one halo exchange of
ghost vector entries per SpMV (grouped point-to-point over NCCL / NVLink, at most two neighbours
for slabs) and one all-reduce of a scalar per dot product.

The reference has no distributed path at all (no MPI / NCCL anywhere); the
single-GPU operators are the oracle: tests check that the distributed rows equal the global
matrix's rows and that the distributed CG reproduces the single-GPU iterates.

Everything in this module except the ``pgb_*`` kernel calls is device agnostic, so the
partition / exchange logic is covered by world_size-2 ``gloo`` tests on CPU.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

from .grid import SimplexGrid, structured_grid


def bind_to_gpu_numa(device_index: int) -> list[int] | None:
    """Pin the calling process to the CPUs NVML reports as local to the GPU, so that the pinned host
    buffers it allocates afterwards (grid arrays, result blocks) live on the GPU's NUMA node.  With
    one process per GPU this keeps every rank's PCIe traffic off the inter-socket link.  Returns the
    CPU list, or None when NVML / the affinity call is unavailable (nothing is changed then)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


# --------------------------------------------------------------------------------------
# structured slab partition
# --------------------------------------------------------------------------------------
@dataclass
class SlabLayout:
    n: int  # cubes per side in x and y
    nz_total: int  # cube layers of the global box
    world: int
    rank: int
    z0: int  # first owned layer
    z1: int  # one past the last owned layer

    @property
    def nz_owned(self) -> int:
        return self.z1 - self.z0

    @property
    def halo_top(self) -> bool:
        return self.rank < self.world - 1

    @property
    def nz_local(self) -> int:
        return self.nz_owned + (1 if self.halo_top else 0)


def slab_layout(n: int, nz_total: int, world: int, rank: int) -> SlabLayout:
    """Contiguous split of the ``nz_total`` cube layers (cells are numbered layer by layer)."""
    if nz_total < world:
        raise ValueError("need at least one cube layer per rank")
    base, rem = divmod(nz_total, world)
    z0 = rank * base + min(rank, rem)
    z1 = z0 + base + (1 if rank < rem else 0)
    return SlabLayout(n, nz_total, world, rank, z0, z1)


def local_slab_grid(layout: SlabLayout, device="cpu", geometry: bool = False) -> SimplexGrid:
    """Owned cube layers plus the halo layer above them, as a local grid."""
    sd = structured_grid(3, layout.n, device=device, nz=layout.nz_local, z0=layout.z0)
    if geometry:
        sd.compute_geometry()
    else:
        sd.compute_connectivity()
    sd._layout = layout
    return sd


def _dof_z_range(sd: SimplexGrid, space: str):
    """(zmin, zmax) local lattice layer of the nodes of every local dof, as int64 tensors."""
    st = sd._structured
    n = st["n"]
    m2 = (n + 1) ** 2
    if space == "lagrange1":
        z = torch.arange(sd.num_nodes, dtype=torch.int64) // m2
        return z, z
    if space == "rt0":
        fn = torch.from_numpy(sd.face_nodes.indices.reshape(sd.num_faces, 3).astype(np.int64)) // m2
        return fn.min(dim=1).values, fn.max(dim=1).values
    if space == "nedelec0":
        rp = torch.from_numpy(sd.ridge_peaks.indices.reshape(sd.num_ridges, 2).astype(np.int64)) // m2
        return rp.min(dim=1).values, rp.max(dim=1).values
    if space == "p0":
        z = torch.arange(sd.num_cells, dtype=torch.int64) // (6 * n * n)
        return z, z + 1
    if space == "darcy":  # RT0 x P0: face dofs then cell dofs
        a, b = _dof_z_range(sd, "rt0"), _dof_z_range(sd, "p0")
        return torch.cat([a[0], b[0]]), torch.cat([a[1], b[1]])
    raise KeyError(space)


def owned_mask(sd: SimplexGrid, space: str) -> torch.Tensor:
    """True for the local dofs this rank owns (lowest rank with a cell touching the dof)."""
    if getattr(sd, "_part", None) is not None:
        return _part_owned_mask(sd, space)
    lay: SlabLayout = sd._layout
    zmin, zmax = _dof_z_range(sd, space)
    in_plane = zmin == zmax
    # cells touching the dof live in layers zmin-1 and zmin (in-plane dof) or zmin (spanning dof)
    lowest_layer = torch.where(in_plane, torch.clamp(zmin - 1, min=0), zmin)
    mask = lowest_layer < lay.nz_owned
    if lay.rank > 0:
        mask &= ~(in_plane & (zmin == 0))  # bottom plane belongs to the rank below
    return mask


def global_keys(sd: SimplexGrid, space: str) -> torch.Tensor:
    """A partition-independent int64 key per local dof (lattice key shifted by the slab offset; the global
    dof id for a grid cut out of a global grid by :func:`partition_local_grid`)."""
    if getattr(sd, "_part", None) is not None:
        return _part_global_keys(sd, space)
    st = sd._structured
    lay: SlabLayout = sd._layout
    n, sh = st["n"], st["key_bits"]
    node_off = lay.z0 * (n + 1) ** 2
    if space == "lagrange1":
        return torch.arange(sd.num_nodes, dtype=torch.int64) + node_off
    if space == "rt0":
        return st["face_keys"].cpu() + (node_off << (2 * sh))
    if space == "nedelec0":
        return st["edge_keys"].cpu() + (node_off << sh)
    if space == "p0":
        return torch.arange(sd.num_cells, dtype=torch.int64) + lay.z0 * 6 * n * n
    if space == "darcy":
        return torch.cat([global_keys(sd, "rt0"), global_keys(sd, "p0") + (1 << 60)])
    raise KeyError(space)


# --------------------------------------------------------------------------------------
# arbitrary cell partition of a global grid ("any cell partition + owner = lowest-rank cell
# touching the dof")
# --------------------------------------------------------------------------------------
def partition_local_grid(sd: SimplexGrid, cell_owner: np.ndarray, rank: int) -> SimplexGrid:
    """Local grid of ``rank`` for the cell partition ``cell_owner`` (one rank id per global cell): the
    rank's own cells plus every cell that shares a node with one of them (all cells touching any dof
    the rank can own), with nodes / faces / cells renumbered in ascending global order.  The returned
    grid carries ``_part``: the global ids of its nodes / faces / cells / ridges and the owner of every
    local cell, from which :func:`owned_mask` / :func:`global_keys` work for every space."""
    import scipy.sparse as sps

    d = sd.dim
    cell_owner = np.asarray(cell_owner)
    nc, nf, nn = sd.num_cells, sd.num_faces, sd.num_nodes
    cf = sd.cell_faces.indices.reshape(nc, d + 1)
    fn = sd.face_nodes.indices.reshape(nf, d)
    cn = sd.cell_node_table()
    mine = cell_owner == rank
    node_hit = np.zeros(nn, dtype=bool)
    node_hit[cn[mine].ravel()] = True
    cells = np.flatnonzero(node_hit[cn].any(axis=1))  # own cells + one node-layer of halo cells
    faces = np.unique(cf[cells].ravel())
    nodes = np.unique(fn[faces].ravel())
    f_new = -np.ones(nf, dtype=np.int64)
    f_new[faces] = np.arange(faces.size)
    n_new = -np.ones(nn, dtype=np.int64)
    n_new[nodes] = np.arange(nodes.size)
    lcf = f_new[cf[cells]]  # ascending inside a cell: the renumbering is monotone
    lsg = np.asarray(sd.cell_faces.data).reshape(nc, d + 1)[cells]
    lfn = n_new[fn[faces]]
    cell_faces = sps.csc_array((lsg.ravel(), lcf.ravel(), np.arange(0, lcf.size + 1, d + 1)),
                               shape=(faces.size, cells.size))
    face_nodes = sps.csc_array((np.ones(lfn.size, dtype=bool), lfn.ravel(), np.arange(0, lfn.size + 1, d)),
                               shape=(nodes.size, faces.size))
    loc = SimplexGrid(d, np.asarray(sd.nodes)[:, nodes], face_nodes, cell_faces, f"{sd.name}_part{rank}")
    loc.compute_connectivity()
    part = {"rank": rank, "cells": cells, "faces": faces, "nodes": nodes, "cell_owner": cell_owner[cells],
            "num_global": {"lagrange1": nn, "rt0": nf, "p0": nc}}
    if d == 3:  # global ridge ids of the local ridges through their global (lo, hi) node pairs
        rp = loc.ridge_peaks.indices.reshape(loc.num_ridges, 2)
        g = nodes[rp]
        grp = sd.ridge_peaks.indices.reshape(sd.num_ridges, 2).astype(np.int64)
        gkey = grp.min(axis=1) * nn + grp.max(axis=1)
        order = np.argsort(gkey)
        want = g.min(axis=1) * nn + g.max(axis=1)
        pos = np.searchsorted(gkey[order], want)
        if np.any(gkey[order][np.minimum(pos, gkey.size - 1)] != want):
            raise ValueError("a local ridge is missing from the global grid")
        part["ridges"] = order[pos]
        part["num_global"]["nedelec0"] = sd.num_ridges
    else:
        part["ridges"] = faces
        part["num_global"]["nedelec0"] = nf
    loc._part = part
    return loc


def _part_cell_dofs(sd: SimplexGrid, space: str) -> np.ndarray:
    """(nc_local, k) local dof ids touched by every local cell."""
    d, nc = sd.dim, sd.num_cells
    if space == "lagrange1":
        return sd.cell_node_table()
    if space == "rt0":
        return sd.cell_faces.indices.reshape(nc, d + 1)
    if space == "p0":
        return np.arange(nc)[:, None]
    if space == "nedelec0":
        if d == 2:
            return sd.cell_faces.indices.reshape(nc, 3)
        fr = sd.face_ridges.indices.reshape(sd.num_faces, 3)
        return fr[sd.cell_faces.indices.reshape(nc, 4)].reshape(nc, 12)
    raise KeyError(space)


def _part_owned_mask(sd: SimplexGrid, space: str) -> torch.Tensor:
    if space == "darcy":
        return torch.cat([_part_owned_mask(sd, "rt0"), _part_owned_mask(sd, "p0")])
    part = sd._part
    dofs = _part_cell_dofs(sd, space)
    ndof = {"lagrange1": sd.num_nodes, "rt0": sd.num_faces, "p0": sd.num_cells, "nedelec0": sd.num_edges}[space]
    owner = np.full(ndof, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(owner, dofs.ravel(), np.repeat(part["cell_owner"].astype(np.int64), dofs.shape[1]))
    # a dof is decided only if one of this rank's OWN cells touches it (then all its cells are local)
    touched = np.zeros(ndof, dtype=bool)
    touched[dofs[part["cell_owner"] == part["rank"]].ravel()] = True
    return torch.from_numpy(touched & (owner == part["rank"]))


def _part_global_keys(sd: SimplexGrid, space: str) -> torch.Tensor:
    part = sd._part
    if space == "darcy":
        return torch.cat([_part_global_keys(sd, "rt0"), _part_global_keys(sd, "p0") + (1 << 60)])
    ids = {"lagrange1": part["nodes"], "rt0": part["faces"], "p0": part["cells"], "nedelec0": part["ridges"]}[space]
    return torch.from_numpy(np.asarray(ids, dtype=np.int64))


# --------------------------------------------------------------------------------------
# distributed CSR
# --------------------------------------------------------------------------------------
@dataclass
class DistCSR:
    """Rows of the owned dofs; columns renumbered [owned (n_owned) | ghost (n_ghost)]."""

    n_owned: int
    n_ghost: int
    indptr: torch.Tensor
    indices: torch.Tensor
    data: torch.Tensor
    owned_local: torch.Tensor  # local grid ids of the owned dofs (row order)
    ghost_local: torch.Tensor  # local grid ids of the ghost dofs (ghost-block order)
    owned_keys: torch.Tensor
    send_idx: dict = field(default_factory=dict)  # peer -> indices into the owned block
    recv_range: dict = field(default_factory=dict)  # peer -> (start, count) inside the ghost block
    group: object = None
    send_dst: dict = field(default_factory=dict)  # peer -> first entry of this rank's block in the peer's local vector
    n_local_max: int = 0  # largest local vector length over the ranks (sizes the peer-memory arena)
    _peer: object = None  # PeerComm + flattened halo description, built on first use
    _interior: tuple | None = None

    @property
    def ghost_start(self) -> int:
        """First ghost entry of a local vector: the owned block padded to a multiple of 16 doubles, so that
        no 128-byte line holds owned and ghost entries together (peers store into the ghost block while the
        interior rows of the SpMV are already reading owned entries)."""
        return (self.n_owned + 15) // 16 * 16

    @property
    def n_local(self) -> int:
        return self.ghost_start + self.n_ghost

    def column_ids(self) -> torch.Tensor:
        """Local grid id of every local column (length ``n_local``; -1 in the padding)."""
        ids = torch.full((self.n_local,), -1, dtype=torch.int64, device=self.owned_local.device)
        ids[: self.n_owned] = self.owned_local
        ids[self.ghost_start:] = self.ghost_local
        return ids

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())


def _all_gather_var(t: torch.Tensor, group=None) -> list:
    """all_gather of 1-D int64 tensors of different lengths (pads to the max length)."""
    world = dist.get_world_size(group)
    dev = t.device
    size = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=t.dtype, device=dev)
    buf[: t.numel()] = t
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def build_dist_csr(indptr, indices, data, owned: torch.Tensor, keys: torch.Tensor, group=None,
                   comm_device=None) -> DistCSR:
    """Extract the owned rows of a locally assembled CSR (local ids) and set up the halo exchange.

    ``indptr / indices / data``: CSR over all local dofs (torch tensors, any device);
    ``owned``: bool mask over local dofs; ``keys``: partition-independent int64 key per local dof.
    ``comm_device``: device of the tensors handed to the collectives (cuda for nccl, cpu for gloo).
    On a CUDA device the rows are selected and renumbered by the extraction kernels
    (``pgb_mask_to_map`` / ``pgb_csr_extract_*``: O(rows) scratch); the torch version below serves the
    CPU (gloo) tests of the partition logic."""
    dev = indices.device
    comm_device = dev if comm_device is None else torch.device(comm_device)
    owned = owned.to(dev)
    keys = keys.to(dev)
    n_local = owned.numel()
    cuda = dev.type == "cuda"
    if cuda:
        from . import engine

        _, own32 = engine.mask_to_map(owned, want_newid=False)
        own = own32.long()
        n_owned = int(own.numel())
        # referenced columns of the owned rows: one marking pass over those rows (no per-entry temporary)
        from ._lib import check, lib

        ref = torch.zeros(n_local, dtype=torch.bool, device=dev)
        with torch.cuda.device(dev):
            check(lib.pgb_csr_mark_columns(n_owned, own32.data_ptr(), indptr.data_ptr(),
                                           1 if indptr.dtype == torch.int64 else 0, indices.data_ptr(),
                                           ref.data_ptr(), torch.cuda.current_stream().cuda_stream))
    else:
        own = torch.nonzero(owned).flatten()
        n_owned = int(own.numel())
        ip = indptr.to(torch.int64)
        lens = ip[own + 1] - ip[own]
        new_ip = torch.zeros(n_owned + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens, 0, out=new_ip[1:])
        nnz = int(new_ip[-1].item())
        row_of = torch.repeat_interleave(torch.arange(n_owned, device=dev), lens)
        pos = ip[own][row_of] + (torch.arange(nnz, device=dev) - new_ip[:-1][row_of])
        cols = indices[pos].to(torch.int64)
        vals = data[pos]
        ref = torch.zeros(n_local, dtype=torch.bool, device=dev)
        ref[cols] = True
    ghost = torch.nonzero(ref & ~owned).flatten()
    gkeys = keys[ghost]

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    okeys = keys[own]
    send_idx, recv_range = {}, {}
    if world > 1:
        okeys_sorted, oorder = torch.sort(okeys)
        all_g = _all_gather_var(gkeys.to(comm_device), group)
        masks = []
        hits = {}
        for q, gq in enumerate(all_g):
            gq = gq.to(dev)
            if q == rank or gq.numel() == 0 or n_owned == 0:
                m = torch.zeros(gq.numel(), dtype=torch.bool, device=dev)
            else:
                p = torch.searchsorted(okeys_sorted, gq).clamp(max=n_owned - 1)
                m = okeys_sorted[p] == gq
                hits[q] = oorder[p[m]]
            masks.append(m)
        for q, h in hits.items():
            if h.numel():
                send_idx[q] = h
        allm = _all_gather_var(torch.cat(masks).to(torch.int64).to(comm_device), group)
        off = sum(g.numel() for g in all_g[:rank])
        ng = int(ghost.numel())
        owner = torch.full((ng,), -1, dtype=torch.int64, device=dev)
        for q, mq in enumerate(allm):
            mine = mq[off: off + ng].to(dev).bool()
            if (owner[mine] >= 0).any():
                raise RuntimeError("a ghost dof is claimed by two owners")
            owner[mine] = q
        if ng and (owner < 0).any():
            raise RuntimeError("a ghost dof has no owner")
        # order the ghost block by (owner, position in the original ghost list) so that every peer's
        # values arrive contiguously and in the order the peer sends them
        order = torch.argsort(owner * max(ng, 1) + torch.arange(ng, device=dev))
        ghost = ghost[order]
        owner = owner[order]
        for q in range(world):
            cnt = int((owner == q).sum().item())
            if cnt:
                start = int(torch.nonzero(owner == q)[0].item())
                recv_range[q] = (start, cnt)
    elif ghost.numel():
        raise RuntimeError("ghost dofs in a single-rank run")
    # renumber the columns: [owned | pad to a multiple of 16 | ghosts]
    gstart = (n_owned + 15) // 16 * 16
    newid = torch.full((n_local,), -1, dtype=torch.int32, device=dev)
    newid[own] = torch.arange(n_owned, device=dev, dtype=torch.int32)
    newid[ghost] = gstart + torch.arange(ghost.numel(), device=dev, dtype=torch.int32)
    if cuda:
        new_ip, ncols, vals = engine.extract_rows(indptr, indices, data, own32, newid)
    else:
        ncols = newid[cols].to(torch.int32)
        it = torch.int64 if nnz >= 2**31 else torch.int32
        new_ip = new_ip.to(it)
    A = DistCSR(n_owned, int(ghost.numel()), new_ip, ncols, vals, own, ghost, okeys, send_idx, recv_range, group)
    # where this rank's values land in each peer's local vector, and the arena size all ranks agree on
    A.n_local_max = A.n_local
    if world > 1:
        info = [None] * world
        dist.all_gather_object(info, {"n_local": A.n_local, "gstart": gstart, "recv": recv_range}, group=group)
        A.n_local_max = max(i["n_local"] for i in info)
        for q in send_idx:
            start, cnt = info[q]["recv"][rank]
            if cnt != int(send_idx[q].numel()):
                raise RuntimeError("halo plan mismatch between sender and receiver")
            A.send_dst[q] = info[q]["gstart"] + start
    return A


def interior_range(A: DistCSR) -> tuple[int, int]:
    """Longest run ``[a, b)`` of consecutive owned rows without a ghost column.

    In the slab numbering the rows that read ghost values sit at both ends of the owned block (the
    planes shared with the lower / upper neighbour), so ``[a, b)`` holds nearly all rows: an SpMV over
    it needs no halo data and can run while ``halo_exchange`` is in flight; the two interface ranges
    ``[0, a)`` and ``[b, n_owned)`` follow.  (Pointer offsets are enough for a row-range SpMV: indptr
    holds absolute positions.)  The peer-memory SpMV (``kd_spmv``) multiplies ``[a, b)`` while the halo is in flight."""
    n = A.n_owned
    if n == 0:
        return 0, 0
    ip = A.indptr
    if ip.device.type == "cuda":  # one pass over the rows (pgb_csr_rows_touching), O(rows) scratch
        from ._lib import check, lib

        touches = torch.zeros(n, dtype=torch.bool, device=ip.device)
        with torch.cuda.device(ip.device):
            check(lib.pgb_csr_rows_touching(n, ip.data_ptr(), 1 if ip.dtype == torch.int64 else 0, A.indices.data_ptr(),
                                            A.ghost_start, touches.data_ptr(), torch.cuda.current_stream().cuda_stream))
    else:
        ip = ip.to(torch.int64)
        lens = ip[1:] - ip[:-1]
        row_of = torch.repeat_interleave(torch.arange(n, device=ip.device), lens)
        touches = torch.zeros(n, dtype=torch.bool, device=ip.device)
        touches[row_of[A.indices.to(torch.int64) >= A.ghost_start]] = True
    if not bool(touches.any()):
        return 0, n
    # run lengths of the ghost-free rows: boundaries at the interface rows
    bad = torch.nonzero(touches).flatten()
    edges = torch.cat([bad.new_tensor([-1]), bad, bad.new_tensor([n])])
    gaps = edges[1:] - edges[:-1] - 1
    k = int(torch.argmax(gaps).item())
    a = int(edges[k].item()) + 1
    return a, a + int(gaps[k].item())


def halo_exchange(A: DistCSR, x: torch.Tensor, comm_device=None) -> None:
    """Fill the ghost block ``x[n_owned:]`` with the owners' values (grouped isend/irecv)."""
    if not A.send_idx and not A.recv_range:
        return
    cdev = x.device if comm_device is None else torch.device(comm_device)
    ops, recvs, keep = [], [], []
    for q, (start, cnt) in sorted(A.recv_range.items()):
        view = x[A.ghost_start + start: A.ghost_start + start + cnt]
        if view.device == cdev:
            buf = view
        else:
            buf = torch.empty(cnt, dtype=x.dtype, device=cdev)
            recvs.append((view, buf))
        ops.append(dist.P2POp(dist.irecv, buf, q, group=A.group))
    for q, idx in sorted(A.send_idx.items()):
        buf = x[idx].to(cdev)
        keep.append(buf)
        ops.append(dist.P2POp(dist.isend, buf, q, group=A.group))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    for view, buf in recvs:
        view.copy_(buf)


# --------------------------------------------------------------------------------------
# distributed assembly + solve (GPU)
# --------------------------------------------------------------------------------------
def assemble_slab(space: str, form: str, sd: SimplexGrid, data=None, keyword="unitary_data", group=None) -> DistCSR:
    """Assemble the owned rows of ``space``/``form`` on this rank's slab grid (owned + halo cells)."""
    from . import engine
    from .constants import SECOND_ORDER_TENSOR, WEIGHT
    from .params import coefficient

    w = coefficient(sd, data, keyword, WEIGHT)
    K = coefficient(sd, data, keyword, SECOND_ORDER_TENSOR)
    if space == "lagrange1":
        K = None if form == "mass" else K
    A = engine.assemble_device(space, form, sd, w, K)
    return build_dist_csr(A.indptr, A.indices, A.data, owned_mask(sd, space), global_keys(sd, space), group)


def dist_spmv(A: DistCSR, x: torch.Tensor, y: torch.Tensor, tpr: int = 0) -> None:
    """y[:n_owned] = A @ x after refreshing the ghost block of x."""
    from ._lib import check, lib
    from .engine import _stream

    halo_exchange(A, x)
    check(lib.pgb_spmv(A.n_owned, A.indptr.data_ptr(), 1 if A.indptr.dtype == torch.int64 else 0,
                       A.indices.data_ptr(), A.data.data_ptr(), x.data_ptr(), y.data_ptr(), tpr, _stream()))


class _Dot:
    def __init__(self, device, group):
        from ._lib import NPART

        self.parts = torch.zeros(NPART, dtype=torch.float64, device=device)
        self.out = torch.zeros(1, dtype=torch.float64, device=device)
        self.group = group
        self.npart = NPART

    def __call__(self, n, x, y) -> float:
        from ._lib import check, lib
        from .engine import _stream

        check(lib.pgb_dot_partials(n, x.data_ptr(), y.data_ptr(), self.parts.data_ptr(), _stream()))
        check(lib.pgb_reduce_partials(self.parts.data_ptr(), self.npart, self.out.data_ptr(), _stream()))
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.out, group=self.group)
        return float(self.out.item())


def _tpr(A: DistCSR) -> int:
    if not A.nnz or not A.n_owned:
        return 2
    avg = A.nnz / A.n_owned
    return 2 if avg <= 9 else 4 if avg <= 18 else 8 if avg <= 40 else 16 if avg <= 100 else 32


def dist_cg(A: DistCSR, b: torch.Tensor, *, rtol=1e-5, atol=0.0, maxiter=1000, check_every=8):
    """CG on the row-distributed operator with the **same fused device kernels** as the single-GPU
    driver (``pgb_cg_step_*``): scalars never leave the device; per iteration one grouped halo
    exchange (before the SpMV) and two tiny all-reduces of partial sums (p.q; then r.r and r.z).
    The host only polls the ``done`` flag every ``check_every`` iterations.  Recurrence and stopping
    rule are those of ``scipy.sparse.linalg.cg``.  ``b`` holds the owned entries.
    Returns (x_owned, info, iterations, residual norms)."""
    from ._lib import NPART, check, lib
    from .engine import _stream

    dev = b.device
    n = A.n_owned
    multi = dist.is_initialized() and dist.get_world_size(A.group) > 1
    tpr = _tpr(A)
    idx64 = 1 if A.indptr.dtype == torch.int64 else 0
    scratch = torch.zeros(int(lib.pgb_cg_scratch_doubles(int(maxiter))), dtype=torch.float64, device=dev)
    sp = scratch.data_ptr()
    red = scratch[4 * NPART: 4 * NPART + 4]
    flags = scratch[4 * NPART + 8: 4 * NPART + 9].view(torch.int32)
    scal_i = scratch[4 * NPART + 16 + 5: 4 * NPART + 16 + 6].view(torch.int32)  # (done, iters)
    hist_d = scratch[4 * NPART + 32:]
    x = torch.zeros(A.n_local, dtype=torch.float64, device=dev)
    p = torch.zeros(A.n_local, dtype=torch.float64, device=dev)
    r = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    q = torch.zeros(max(n, 1), dtype=torch.float64, device=dev)

    def allreduce(lo, hi):
        mask = sum(1 << i for i in range(lo, hi))
        check(lib.pgb_cg_step_reduce(sp, mask, 0, _stream()))
        if multi:
            dist.all_reduce(red[lo:hi], group=A.group)
        check(lib.pgb_cg_step_reduce(sp, mask, 1, _stream()))

    check(lib.pgb_cg_step_init(n, b.data_ptr(), q.data_ptr(), 0, r.data_ptr(), float(rtol), float(atol), sp, _stream()))
    allreduce(0, 3)
    if float(red[2].item()) == 0.0:  # scipy: b == 0 -> x = 0
        return x[:n], 0, 0, np.zeros(0)
    launched = 0
    for k in range(int(maxiter)):
        check(lib.pgb_cg_step_p(n, k, r.data_ptr(), 0, p.data_ptr(), sp, _stream()))
        halo_exchange(A, p)
        check(lib.pgb_cg_step_spmv(n, A.indptr.data_ptr(), idx64, A.indices.data_ptr(), A.data.data_ptr(),
                                   p.data_ptr(), q.data_ptr(), tpr, sp, _stream()))
        allreduce(3, 4)
        check(lib.pgb_cg_step_xr(n, p.data_ptr(), q.data_ptr(), 0, x.data_ptr(), r.data_ptr(), sp, _stream()))
        allreduce(0, 2)
        launched += 1
        if check_every and (k + 1) % check_every == 0 and int(flags[0].item()):
            break
    done, iters = (int(v) for v in scal_i.cpu())
    if done:
        hist = hist_d[: iters + 1].cpu().numpy()
        return x[:n], 0, iters, hist
    hist = np.concatenate([hist_d[:iters].cpu().numpy(), [float(np.sqrt(max(float(red[0].item()), 0.0)))]])
    return x[:n], int(maxiter), iters, hist


def dist_minres(A: DistCSR, b: torch.Tensor, *, rtol=1e-5, shift=0.0, maxiter=1000, check_every=8):
    """MINRES (Paige-Saunders, the recurrence and stopping tests of ``scipy.sparse.linalg.minres``)
    on the row-distributed operator with the fused device kernels of the single-GPU driver
    (``pgb_mr_step_*``): per iteration one halo exchange of ``v`` and three one-scalar all-reduces
    (v.y, r2.y, x.x); all Lanczos / rotation scalars stay on the device.
    Returns (x_owned, info, iterations, phibar history)."""
    import ctypes as C

    from ._lib import NPART, check, lib
    from .engine import _stream

    dev = b.device
    n = A.n_owned
    multi = dist.is_initialized() and dist.get_world_size(A.group) > 1
    tpr = _tpr(A)
    idx64 = 1 if A.indptr.dtype == torch.int64 else 0
    scratch = torch.zeros(int(lib.pgb_mr_scratch_doubles(int(maxiter))), dtype=torch.float64, device=dev)
    sp = scratch.data_ptr()
    red = scratch[5 * NPART: 5 * NPART + 5]
    flags = scratch[5 * NPART + 8: 5 * NPART + 9].view(torch.int32)
    out2 = scratch[5 * NPART + 80: 5 * NPART + 82]
    hist_d = scratch[5 * NPART + 82:]

    def allreduce(lo, hi):
        mask = sum(1 << i for i in range(lo, hi))
        check(lib.pgb_mr_step_reduce(sp, mask, 0, _stream()))
        if multi:
            dist.all_reduce(red[lo:hi], group=A.group)
        check(lib.pgb_mr_step_reduce(sp, mask, 1, _stream()))

    f64 = dict(dtype=torch.float64, device=dev)
    x = torch.zeros(max(n, 1), **f64)
    v = torch.zeros(A.n_local, **f64)
    rb = [torch.zeros(max(n, 1), **f64) for _ in range(3)]
    wb = [torch.zeros(max(n, 1), **f64) for _ in range(2)]
    check(lib.pgb_mr_step_init(n, b.data_ptr(), rb[2].data_ptr(), rb[0].data_ptr(), sp, _stream()))
    allreduce(3, 5)
    check(lib.pgb_mr_step_start(n, rb[0].data_ptr(), v.data_ptr(), wb[0].data_ptr(), wb[1].data_ptr(),
                                float(rtol), float(shift), int(maxiter), sp, _stream()))
    b1sq, bb = (float(t) for t in out2.cpu())
    if b1sq == 0.0 or bb == 0.0:  # x0 = 0 solves the system / b == 0
        return x[:n], 0, 0, np.zeros(0)
    r1 = r2 = rb[0]
    y, spare = rb[1], rb[2]
    w_new, w_old = wb[0], wb[1]
    itn = 0
    while itn < maxiter:
        if check_every and itn and itn % check_every == 0 and int(flags[0].item()):
            break
        itn += 1
        halo_exchange(A, v)
        check(lib.pgb_mr_step_spmv(n, A.indptr.data_ptr(), idx64, A.indices.data_ptr(), A.data.data_ptr(),
                                   v.data_ptr(), y.data_ptr(), r1.data_ptr(), itn, tpr, sp, _stream()))
        allreduce(0, 1)
        check(lib.pgb_mr_step_c(n, y.data_ptr(), r2.data_ptr(), sp, _stream()))
        allreduce(1, 2)
        old_r1 = r1
        r1, r2 = r2, y
        y = spare if old_r1 is r1 else old_r1
        check(lib.pgb_mr_step_scalars(itn, 0, sp, _stream()))
        check(lib.pgb_mr_step_d(n, v.data_ptr(), w_old.data_ptr(), w_new.data_ptr(), r2.data_ptr(), x.data_ptr(),
                                sp, _stream()))
        allreduce(2, 3)
        w_old, w_new = w_new, w_old
    check(lib.pgb_mr_step_scalars(itn + 1, 1, sp, _stream()))
    it_h, st_h = C.c_int(0), C.c_int(0)
    check(lib.pgb_mr_state(sp, C.byref(it_h), C.byref(st_h), _stream()))
    k = it_h.value
    hist = hist_d[1: k + 1].cpu().numpy()
    return x[:n], (int(maxiter) if st_h.value == 6 else 0), k, hist


# --------------------------------------------------------------------------------------
# NVLink peer-memory transport (pgb_dist.cu): halo stores and global sums inside the Krylov kernels
# --------------------------------------------------------------------------------------
class PeerComm:
    """Arena + flag header of this rank, exported to / mapped from the other ranks of the node through
    CUDA IPC handles (exchanged with one ``all_gather``).  ``arena_bytes`` must agree on all ranks."""

    def __init__(self, arena_bytes: int, group=None, device=None) -> None:
        import ctypes as C

        from ._lib import check, lib

        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.pgb_comm_create(self.rank, self.world, int(arena_bytes), C.byref(h)))
            self.handle = h
            if self.world > 1:
                nb = int(lib.pgb_comm_handle_bytes())
                buf = (C.c_ubyte * nb)()
                check(lib.pgb_comm_get_handle(h, buf))
                mine = torch.tensor(list(buf), dtype=torch.uint8, device=self.device)
                allh = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(allh, mine, group=group)
                flat = torch.cat(allh).cpu().numpy().tobytes()
                check(lib.pgb_comm_connect(h, flat))
                dist.barrier(group=group)  # every rank has mapped every arena before anyone stores into one
        self.arena_bytes = int(lib.pgb_comm_arena_bytes(h))

    def close(self) -> None:
        from ._lib import lib

        if self.handle is not None:
            torch.cuda.synchronize(self.device)
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)  # nobody is still storing into a peer's arena
            lib.pgb_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):  # best effort: no collective here
        try:
            from ._lib import lib

            if getattr(self, "handle", None) is not None:
                lib.pgb_comm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _PeerPlan:
    """Flattened halo description in the form pgb_dist_cg / pgb_dist_minres take."""

    def __init__(self, A: DistCSR) -> None:
        import ctypes as C

        dev = A.data.device
        speers = sorted(A.send_idx)
        rpeers = sorted(A.recv_range)
        offs = [0]
        for q in speers:
            offs.append(offs[-1] + int(A.send_idx[q].numel()))
        self.nsend, self.nrecv = len(speers), len(rpeers)
        self.send_peer = (C.c_int * max(self.nsend, 1))(*speers)
        self.send_off = (C.c_int64 * (self.nsend + 1))(*offs)
        self.send_dst = (C.c_int64 * max(self.nsend, 1))(*[int(A.send_dst[q]) for q in speers])
        self.recv_peer = (C.c_int * max(self.nrecv, 1))(*rpeers)
        self.send_idx = (torch.cat([A.send_idx[q] for q in speers]).to(torch.int32).contiguous() if speers
                         else torch.zeros(1, dtype=torch.int32, device=dev))
        self.halo_bytes = 8 * offs[-1]


def peer_setup(A: DistCSR) -> None:
    """Communicator (sized for the double-buffered CG direction), halo plan and interior range of ``A``."""
    if A._peer is None:
        nbytes = 2 * ((A.n_local_max * 8 + 255) // 256 * 256)
        A._peer = (PeerComm(nbytes, A.group, A.data.device), _PeerPlan(A))
        A._interior = interior_range(A)


def peer_release(A: DistCSR) -> None:
    if A._peer is not None:
        try:
            A._peer[0].close()
        finally:
            A._peer = None


def _peer_solve(fn_name: str, A: DistCSR, b: torch.Tensor, dinv, work_vectors: int, extra: tuple, maxiter: int):
    import ctypes as C

    from . import _lib
    from ._lib import check, lib
    from .engine import _ptr, _stream
    from .solvers import RESID_CAP, _maxiter

    peer_setup(A)
    comm, plan = A._peer
    dev = b.device
    n = A.n_owned
    maxiter = _maxiter(maxiter, 10 * max(n, 1))
    lo, hi = A._interior
    x = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    work = torch.empty(work_vectors * max(n, 1), dtype=torch.float64, device=dev)
    resid = np.zeros(min(maxiter + 1, RESID_CAP))
    it, info = C.c_int(0), C.c_int(0)
    idx64 = 1 if A.indptr.dtype == torch.int64 else 0
    with torch.cuda.device(dev):
        if comm.world > 1:
            dist.barrier(group=A.group)  # ranks enter the solve together (the device waits are bounded)
        check(getattr(lib, fn_name)(
            comm.handle, n, A.ghost_start, A.n_ghost, A.indptr.data_ptr(), idx64, A.indices.data_ptr(),
            A.data.data_ptr(), b.data_ptr(), _ptr(dinv), x.data_ptr(), work.data_ptr(), plan.nsend, plan.send_peer,
            plan.send_off, plan.send_dst, plan.send_idx.data_ptr(), plan.nrecv, plan.recv_peer, lo, hi, _tpr(A),
            *extra, int(maxiter), C.byref(it), C.byref(info), resid.ctypes.data, _stream()))
        _lib.count(2 + 3 * it.value)
    return x[:n], info.value, it.value, resid


def dist_cg_peer(A: DistCSR, b: torch.Tensor, *, rtol=1e-5, atol=0.0, maxiter=1000, dinv=None):
    """CG on the row-distributed operator with the halo exchange and the global sums done by the Krylov
    kernels over NVLink peer memory (``pgb_dist_cg``): 3 kernel launches per iteration, no NCCL call, the
    halo in flight under the interior rows of the SpMV.  Same recurrence and stopping rule as
    ``scipy.sparse.linalg.cg`` with ``x0 = 0``.  Returns (x_owned, info, iterations, residual norms)."""
    x, info, k, resid = _peer_solve("pgb_dist_cg", A, b, dinv, 2, (float(rtol), float(atol)), maxiter)
    return x, info, k, resid[: min(k + 1, resid.size)].copy()


def dist_minres_peer(A: DistCSR, b: torch.Tensor, *, rtol=1e-5, shift=0.0, maxiter=1000, dinv=None):
    """MINRES (``pgb_dist_minres``), the peer-memory counterpart of :func:`dist_minres`.
    Returns (x_owned, info, iterations, phibar history)."""
    x, info, k, resid = _peer_solve("pgb_dist_minres", A, b, dinv, 5, (float(shift), float(rtol)), maxiter)
    return x, info, k, resid[: min(k, resid.size)].copy()


def self_check(n: int = 6, nz: int = 10, device=None, group=None) -> dict:
    """Distributed parity on a small box, run by every rank (tests, ``bench.py --gpus N``):

    * the rows this rank owns, assembled on its slab, equal the same rows of the single-GPU assembly of
      the whole box (P1 stiffness, RT0 mass, Nedelec0 mass, Darcy saddle);
    * distributed CG (P1 stiffness + mass) and MINRES (Darcy saddle) over NVLink peer memory and over NCCL
      reproduce the single-GPU drivers' iteration counts and iterates.

    Returns {"ok": bool, "max_row_err": ..., "cg": {...}, "minres": {...}} (identical on all ranks)."""
    import scipy.sparse as sps

    from . import darcy, discretization as D_, engine
    from .solvers import cg_device, minres_device

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    full = structured_grid(3, n, nz=nz, perturb=0.1)
    full.compute_geometry()
    lay = slab_layout(n, nz, world, rank)
    sd = local_slab_grid(lay, geometry=False)
    # same perturbed coordinates as the box: local nodes are a contiguous z-range of the global ones
    m2 = (n + 1) ** 2
    sd.nodes = full.nodes[:, lay.z0 * m2: lay.z0 * m2 + sd.num_nodes].copy()
    sd.compute_connectivity()
    out = {"world": world, "grid": f"{n}x{n}x{nz}", "max_row_err": 0.0}

    def gids(space, keys):
        keys = keys.cpu()
        if space == "lagrange1":
            return keys.numpy()
        if space == "darcy":
            nf = full.num_faces
            f = keys < (1 << 60)
            g = np.empty(keys.numel(), dtype=np.int64)
            g[f.numpy()] = torch.searchsorted(full._structured["face_keys"].cpu(), keys[f]).numpy()
            g[~f.numpy()] = (keys[~f] - (1 << 60)).numpy() + nf
            return g
        table = full._structured["face_keys" if space == "rt0" else "edge_keys"].cpu()
        return torch.searchsorted(table, keys).numpy()

    def rows_match(space, Dm, Ag):
        go = gids(space, Dm.owned_keys)
        keys = global_keys(sd, space)
        cid = Dm.column_ids().cpu()
        valid = (cid >= 0).numpy()
        gcol = np.zeros(Dm.n_local, dtype=np.int64)
        gcol[valid] = gids(space, keys[cid[cid >= 0]])
        loc = sps.csr_array((Dm.data.cpu().numpy(), Dm.indices.cpu().numpy(), Dm.indptr.cpu().numpy()),
                            shape=(Dm.n_owned, Dm.n_local))
        P = sps.csr_array((valid.astype(float), (np.arange(Dm.n_local), gcol)), shape=(Dm.n_local, Ag.shape[1]))
        return go, float(abs(loc @ P - Ag[go]).max() / abs(Ag).max())

    systems = {}
    for space, form, cls in (("lagrange1", "stiff", D_.Lagrange1), ("rt0", "mass", D_.RT0),
                             ("nedelec0", "mass", D_.Nedelec0), ("darcy", "saddle", None)):
        if space == "darcy":
            Dm = assemble_darcy_slab(sd, group=group)
            Ag = darcy.darcy_saddle_device(full).to_scipy("csr")
        else:
            Dm = assemble_slab(space, form, sd, group=group)
            Ad = cls().assemble_stiff_matrix_device(full) if form == "stiff" else cls().assemble_mass_matrix_device(full)
            Ag = Ad.to_scipy("csr")
        go, err = rows_match(space, Dm, Ag)
        out["max_row_err"] = max(out["max_row_err"], err)
        systems[space] = (Dm, Ag, go)

    def agree(xd, go, xs):
        return float(np.abs(xd.cpu().numpy() - xs[go]).max() / max(np.abs(xs).max(), 1e-300))

    def hist_err(h, href, m=30):
        m = min(m, len(h), len(href))
        return float(np.abs(np.asarray(h[:m]) - np.asarray(href[:m])).max() / max(abs(href[0]), 1e-300)) if m else 0.0

    def run(fn, Dm, bo, go, xs, href, **kw):
        try:
            r = fn(Dm, bo, **kw)
        except Exception as e:  # a transport that cannot run here (e.g. no CUDA IPC) fails the check, not the caller
            return {"iters": -1, "info": -1, "x_err": float("inf"), "hist_err": float("inf"), "error": repr(e)[:300]}
        return {"iters": int(r[2]), "info": int(r[1]), "x_err": agree(r[0], go, xs), "hist_err": hist_err(r[3], href)}

    # CG on S + M (SPD), right-hand side from a global seeded vector
    Dm, S, go = systems["lagrange1"]
    Mloc = assemble_slab("lagrange1", "mass", sd, group=group)
    Dm.data += Mloc.data  # same pattern, same extraction order
    Ag = S + D_.Lagrange1().assemble_mass_matrix_device(full).to_scipy("csr")
    bg = np.random.default_rng(7).normal(size=Ag.shape[0])
    xs, si = cg_device(engine.DeviceCSR.from_scipy(Ag, dev), torch.from_numpy(bg).to(dev), rtol=1e-10, maxiter=2000)
    xs = xs.cpu().numpy()
    bo = torch.from_numpy(bg[go]).to(dev)
    out["cg"] = {"single_gpu_iters": int(si.iterations),
                 "peer": run(dist_cg_peer, Dm, bo, go, xs, si.residuals, rtol=1e-10, maxiter=2000),
                 "nccl": run(dist_cg, Dm, bo, go, xs, si.residuals, rtol=1e-10, maxiter=2000)}
    peer_release(Dm)
    # MINRES on the symmetrised saddle system: block-Jacobi preconditioned over peer memory (iterates
    # compared), plain over NCCL (badly conditioned: the first 30 residual estimates are compared)
    Dm, Ag, go = systems["darcy"]
    Afull = engine.DeviceCSR.from_scipy(Ag, dev)
    dinv = darcy.darcy_block_jacobi(Afull, full.num_faces)
    bg = np.random.default_rng(8).normal(size=Ag.shape[0])
    bd = torch.from_numpy(bg).to(dev)
    bo = torch.from_numpy(bg[go]).to(dev)
    xs, si = minres_device(Afull, bd, rtol=1e-10, maxiter=20000, dinv=dinv)
    dio = dinv[torch.from_numpy(go).to(dev)].contiguous()
    out["minres"] = {"single_gpu_iters": int(si.iterations),
                     "peer": run(dist_minres_peer, Dm, bo, go, xs.cpu().numpy(), si.residuals, rtol=1e-10,
                                 maxiter=20000, dinv=dio)}
    xs2, si2 = minres_device(Afull, bd, rtol=1e-8, maxiter=20000)
    out["minres"]["single_gpu_iters_plain"] = int(si2.iterations)
    out["minres"]["nccl"] = run(dist_minres, Dm, bo, go, xs2.cpu().numpy(), si2.residuals, rtol=1e-8, maxiter=20000)
    peer_release(Dm)
    ok = out["max_row_err"] <= 1e-13
    for k, t, ref_it, xtol in (("cg", "peer", "single_gpu_iters", 1e-7), ("cg", "nccl", "single_gpu_iters", 1e-7),
                               ("minres", "peer", "single_gpu_iters", 1e-6),
                               ("minres", "nccl", "single_gpu_iters_plain", None)):
        r, its = out[k][t], out[k][ref_it]
        ok = ok and r["info"] == 0 and r["hist_err"] <= 1e-8 and abs(r["iters"] - its) <= max(3, its // 20)
        if xtol is not None:
            ok = ok and r["x_err"] <= xtol
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        ok = bool(flag.item())
    out["ok"] = ok
    return out


def assemble_darcy_slab(sd: SimplexGrid, data=None, keyword="unitary_data", group=None) -> DistCSR:
    """Owned rows of the symmetrised mixed-Darcy system [[M, -B^T], [-B, 0]] on this rank's slab."""
    from . import darcy

    A = darcy.darcy_saddle_device(sd, data, keyword)
    return build_dist_csr(A.indptr, A.indices, A.data, owned_mask(sd, "darcy"), global_keys(sd, "darcy"), group)


def dist_cg_host(A: DistCSR, b: torch.Tensor, *, rtol=1e-5, atol=0.0, maxiter=1000):
    """Host-driven variant (generic kernels ``pgb_spmv / pgb_dot_partials / pgb_axpby``, two host
    syncs per iteration); kept as a cross-check of :func:`dist_cg`."""
    from ._lib import check, lib
    from .engine import _stream

    dev = b.device
    n = A.n_owned
    dot = _Dot(dev, A.group)
    tpr = _tpr(A)
    x = torch.zeros(A.n_local, dtype=torch.float64, device=dev)
    p = torch.zeros(A.n_local, dtype=torch.float64, device=dev)
    r = b.clone()
    q = torch.empty(n, dtype=torch.float64, device=dev)
    bnorm = np.sqrt(dot(n, b, b))
    atol = max(float(atol), float(rtol) * bnorm)
    hist = []
    if bnorm == 0.0:
        return x[:n], 0, 0, np.zeros(0)
    rho_prev = 1.0
    s = _stream
    for it in range(maxiter):
        rho = dot(n, r, r)
        rn = np.sqrt(rho)
        hist.append(rn)
        if rn < atol:
            return x[:n], 0, it, np.array(hist)
        beta = rho / rho_prev if it > 0 else 0.0
        check(lib.pgb_axpby(n, 1.0, r.data_ptr(), beta, p.data_ptr(), s()))  # p = r + beta p
        dist_spmv(A, p, q, tpr)
        alpha = rho / dot(n, p, q)
        check(lib.pgb_axpby(n, alpha, p.data_ptr(), 1.0, x.data_ptr(), s()))  # x += alpha p
        check(lib.pgb_axpby(n, -alpha, q.data_ptr(), 1.0, r.data_ptr(), s()))  # r -= alpha q
        rho_prev = rho
    hist.append(np.sqrt(dot(n, r, r)))
    return x[:n], maxiter, maxiter, np.array(hist)
