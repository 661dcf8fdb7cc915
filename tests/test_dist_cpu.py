"""world_size-2/3 gloo tests (CPU) of the cell-partitioned multi-GPU host logic: slab layout,
dof ownership, owned-row extraction, ghost renumbering and the halo exchange.  The local matrices
come from the CPU oracle here; on the GPU box tests/test_dist_gpu.py does the same with the CUDA
assembly."""

import os
import socket

import numpy as np
import pytest
import scipy.sparse as sps
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygeon_b200 import dist as pd
from pygeon_b200.grid import structured_grid

N, NZ = 3, 5
CASES = [("lagrange1", "stiff"), ("rt0", "mass"), ("nedelec0", "mass")]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_ids(full, space, keys):
    st = full._structured
    if space == "lagrange1":
        return keys.numpy()
    table = st["face_keys"] if space == "rt0" else st["edge_keys"]
    pos = torch.searchsorted(table, keys)
    assert torch.equal(table[pos], keys)
    return pos.numpy()


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import fem_oracle as fo

        full = structured_grid(3, N, nz=NZ)
        full.compute_geometry()
        lay = pd.slab_layout(N, NZ, world, rank)
        sd = pd.local_slab_grid(lay, geometry=True)
        assert sd.num_cells == 6 * N * N * lay.nz_local
        rng = np.random.default_rng(0)
        for space, form in CASES:
            kw = {"rot2d": False} if space == "lagrange1" else {}
            Aloc = fo.closed_form(space, form, sd, **kw)
            Aglob = fo.closed_form(space, form, full, **kw).tocsr()
            own = pd.owned_mask(sd, space)
            keys = pd.global_keys(sd, space)
            D = pd.build_dist_csr(torch.from_numpy(Aloc.indptr.astype(np.int64)), torch.from_numpy(Aloc.indices),
                                  torch.from_numpy(Aloc.data), own, keys)
            # ownership partitions the global dofs
            gid_owned = _global_ids(full, space, D.owned_keys)
            allo = pd._all_gather_var(torch.from_numpy(gid_owned.astype(np.int64)))
            cat = np.sort(torch.cat(allo).numpy())
            assert np.array_equal(cat, np.arange(Aglob.shape[0])), space
            # rows equal the global rows
            loc = sps.csr_array((D.data.numpy(), D.indices.numpy(), D.indptr.numpy()), shape=(D.n_owned, D.n_local))
            cid = D.column_ids()
            valid = (cid >= 0).numpy()  # the owned block is padded to a multiple of 16 entries
            assert D.ghost_start % 16 == 0 and D.n_owned <= D.ghost_start < D.n_owned + 16
            assert valid.sum() == D.n_owned + D.n_ghost and not valid[D.n_owned: D.ghost_start].any()
            gcol = np.zeros(D.n_local, dtype=np.int64)
            gcol[valid] = _global_ids(full, space, keys[cid[cid >= 0]])
            P = sps.csr_array((valid.astype(float), (np.arange(D.n_local), gcol)), shape=(D.n_local, Aglob.shape[1]))
            assert valid[np.unique(loc.indices)].all()  # no entry points into the padding
            rows_glob = Aglob[gid_owned]
            diff = abs(loc @ P - rows_glob).max()
            assert diff <= 1e-13 * abs(Aglob).max(), (space, diff)
            ones = loc.copy()
            ones.data = np.ones_like(ones.data)
            gones = rows_glob.copy()
            gones.data = np.ones_like(gones.data)
            assert abs(ones @ P - gones).max() == 0  # identical structural pattern
            # halo exchange + local matvec == global matvec
            xg = rng.normal(size=Aglob.shape[1])
            x = torch.zeros(D.n_local, dtype=torch.float64)
            x[: D.n_owned] = torch.from_numpy(xg[gid_owned])
            pd.halo_exchange(D, x)
            assert np.array_equal(x.numpy()[valid], xg[gcol[valid]]), space
            # the plan for direct stores into the peers' vectors (NVLink path): where my values land
            lands = [None] * world
            dist.all_gather_object(lands, {q: (D.send_dst[q], keys[D.owned_local][D.send_idx[q]]) for q in D.send_idx})
            for src, plan in enumerate(lands):
                if rank in plan:
                    dst, ks = plan[rank]
                    assert torch.equal(keys[cid[dst: dst + ks.numel()]], ks), (space, src)
            assert D.n_local_max >= D.n_local
            y = loc @ x.numpy()
            assert np.abs(y - (Aglob @ xg)[gid_owned]).max() <= 1e-12 * np.abs(Aglob @ xg).max()
            # interior range: no ghost column inside, every interface row outside, and it is the bulk
            a, b = pd.interior_range(D)
            ghost_rows = np.unique(loc.tocoo().row[loc.tocoo().col >= D.ghost_start])
            assert 0 <= a <= b <= D.n_owned
            assert not np.any((ghost_rows >= a) & (ghost_rows < b)), space
            if ghost_rows.size:
                assert a == 0 or (a - 1) in ghost_rows
                assert b == D.n_owned or b in ghost_rows
            # at most two neighbours for a slab partition
            assert set(D.send_idx) <= {rank - 1, rank + 1} and set(D.recv_range) <= {rank - 1, rank + 1}
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
        raise


@pytest.mark.parametrize("world", [2, 3])
def test_slab_partition_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_slab_layout_covers_all_layers():
    for world in (1, 2, 3, 8):
        lays = [pd.slab_layout(4, 19, world, r) for r in range(world)]
        assert lays[0].z0 == 0 and lays[-1].z1 == 19
        assert all(a.z1 == b.z0 for a, b in zip(lays, lays[1:]))
        assert max(l.nz_owned for l in lays) - min(l.nz_owned for l in lays) <= 1
        assert not lays[-1].halo_top and all(l.halo_top for l in lays[:-1])


# ---- arbitrary cell partition of an unstructured grid ------------------------------------
def _worker_unstructured(rank, world, port, q, dim):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import sys

        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from oracle import fem_oracle as fo
        from pgb_helpers import delaunay_grid

        full = delaunay_grid(dim, 250 if dim == 3 else 400, seed=3)
        full.compute_geometry()
        # a partition with ragged, non-convex interfaces: by angle around the centre
        cc = full.cell_centers
        ang = np.arctan2(cc[1] - 0.5, cc[0] - 0.5)
        owner = np.floor((ang + np.pi) / (2 * np.pi) * world).astype(int).clip(0, world - 1)
        sd = pd.partition_local_grid(full, owner, rank)
        assert set(np.flatnonzero(owner == rank)) <= set(sd._part["cells"])
        rng = np.random.default_rng(1)
        for space, form in CASES:
            kw = {"rot2d": False} if space == "lagrange1" else {}
            Aloc = fo.closed_form(space, form, sd, **kw)
            Aglob = fo.closed_form(space, form, full, **kw).tocsr()
            D = pd.build_dist_csr(torch.from_numpy(Aloc.indptr.astype(np.int64)), torch.from_numpy(Aloc.indices),
                                  torch.from_numpy(Aloc.data), pd.owned_mask(sd, space), pd.global_keys(sd, space))
            keys = pd.global_keys(sd, space)
            gid_owned = D.owned_keys.numpy()
            allo = pd._all_gather_var(D.owned_keys)
            assert np.array_equal(np.sort(torch.cat(allo).numpy()), np.arange(Aglob.shape[0])), space
            cid = D.column_ids()
            valid = (cid >= 0).numpy()
            gcol = np.zeros(D.n_local, dtype=np.int64)
            gcol[valid] = keys[cid[cid >= 0]].numpy()
            loc = sps.csr_array((D.data.numpy(), D.indices.numpy(), D.indptr.numpy()), shape=(D.n_owned, D.n_local))
            P = sps.csr_array((valid.astype(float), (np.arange(D.n_local), gcol)), shape=(D.n_local, Aglob.shape[1]))
            assert abs(loc @ P - Aglob[gid_owned]).max() <= 1e-13 * abs(Aglob).max(), space
            xg = rng.normal(size=Aglob.shape[1])
            x = torch.zeros(D.n_local, dtype=torch.float64)
            x[: D.n_owned] = torch.from_numpy(xg[gid_owned])
            pd.halo_exchange(D, x)
            assert np.array_equal(x.numpy()[valid], xg[gcol[valid]]), space
            assert np.abs(loc @ x.numpy() - (Aglob @ xg)[gid_owned]).max() <= 1e-12 * np.abs(Aglob @ xg).max()
            a, b = pd.interior_range(D)
            ghost_rows = np.unique(loc.tocoo().row[loc.tocoo().col >= D.ghost_start])
            assert not np.any((ghost_rows >= a) & (ghost_rows < b))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
        raise


@pytest.mark.parametrize("dim,world", [(2, 3), (3, 2)])
def test_arbitrary_partition_gloo(dim, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_unstructured, args=(r, world, port, q, dim)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
