"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_check.py [n] [nz_total]

Every rank assembles its slab with the CUDA path; rank-local rows are compared with the rows of the
single-GPU assembly of the whole box, and the distributed CG with the single-GPU CG.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sps
import torch
import torch.distributed as dist

import pygeon_b200 as pg
from pygeon_b200 import dist as pd
from pygeon_b200 import engine
from pygeon_b200.solvers import cg_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

full = pg.structured_grid(3, n, nz=nz)
full.compute_geometry()
lay = pd.slab_layout(n, nz, world, rank)
sd = pd.local_slab_grid(lay, device=dev, geometry=True)


def gids(space, keys):
    if space == "lagrange1":
        return keys.cpu().numpy()
    table = full._structured["face_keys" if space == "rt0" else "edge_keys"]
    return torch.searchsorted(table, keys.cpu()).numpy()


for space, form, cls in (("lagrange1", "stiff", pg.Lagrange1), ("rt0", "mass", pg.RT0), ("nedelec0", "mass", pg.Nedelec0)):
    D = pd.assemble_slab(space, form, sd)
    Ag = sps.csr_array((cls().assemble_stiff_matrix(full) if form == "stiff" else cls().assemble_mass_matrix(full)).T)
    go = gids(space, D.owned_keys)
    keys = pd.global_keys(sd, space)
    gcol = gids(space, torch.cat([keys[D.owned_local.cpu()], keys[D.ghost_local.cpu()]]))
    loc = sps.csr_array((D.data.cpu().numpy(), D.indices.cpu().numpy(), D.indptr.cpu().numpy()), shape=(D.n_owned, D.n_local))
    P = sps.csr_array((np.ones(D.n_local), (np.arange(D.n_local), gcol)), shape=(D.n_local, Ag.shape[1]))
    err = abs(loc @ P - Ag[go]).max() / abs(Ag).max()
    assert err <= 1e-13, (space, err)
    cnt = torch.tensor([D.n_owned], device=dev)
    dist.all_reduce(cnt)
    assert int(cnt.item()) == Ag.shape[0]
    if rank == 0:
        print(f"{space} {form}: distributed rows == single-GPU rows (rel err {err:.1e}), {Ag.shape[0]} dofs over {world} ranks")

# distributed CG on S + M (SPD) vs single-GPU CG
DS = pd.assemble_slab("lagrange1", "stiff", sd)
DM = pd.assemble_slab("lagrange1", "mass", sd)
DS.data += DM.data
S = pg.Lagrange1().assemble_stiff_matrix_device(full)
M = pg.Lagrange1().assemble_mass_matrix_device(full)
A1 = engine.DeviceCSR(S.shape, S.indptr, S.indices, S.data + M.data)
bg = torch.from_numpy(np.random.default_rng(0).normal(size=S.shape[0])).to(dev)
x1, info1 = cg_device(A1, bg, rtol=1e-10, maxiter=500)
go = torch.from_numpy(gids("lagrange1", DS.owned_keys)).to(dev)
xd, info, k, hist = pd.dist_cg(DS, bg[go].contiguous(), rtol=1e-10, maxiter=500)
if rank == 0:
    m_ = min(len(hist), len(info1.residuals))
    print("iters", k, info1.iterations, "max hist diff", np.abs(hist[:m_] - info1.residuals[:m_]).max(), "hist0", hist[0])
    d_ = np.abs(hist[:m_] - info1.residuals[:m_]) / hist[0]
    print("rel diff at its 0,5,10,20,40,80:", [float(d_[i]) for i in (0, 5, 10, 20, 40, 80) if i < m_], "argmax", int(d_.argmax()))
assert info == 0 and info1.info == 0
assert abs(k - info1.iterations) <= 1, (k, info1.iterations)
m = min(len(hist), len(info1.residuals))
# rounding differences of the dot products (different partition => different summation order)
# are amplified along the Lanczos recurrence; the early history must agree to 1e-10
assert np.abs(hist[:15] - info1.residuals[:15]).max() <= 1e-10 * hist[0]
assert np.abs(hist[:m] - info1.residuals[:m]).max() <= 1e-5 * hist[0]
err = float((xd - x1[go]).abs().max() / x1.abs().max())
assert err <= 1e-9, err
xh, infoh, kh, histh = pd.dist_cg_host(DS, bg[go].contiguous(), rtol=1e-10, maxiter=500)
assert infoh == 0 and abs(kh - k) <= 1 and float((xh - xd).abs().max() / xd.abs().max()) <= 1e-9
xm, infom, km, histm = pd.dist_cg(DS, bg[go].contiguous(), rtol=1e-14, maxiter=7)
assert infom == 7 and km == 7 and len(histm) == 8
if rank == 0:
    print(f"dist_cg: {k} iterations (single GPU {info1.iterations}), max rel diff {err:.1e}")
# distributed MINRES on the symmetrised Darcy saddle system vs the single-GPU MINRES
from pygeon_b200.solvers import minres_device

DD = pd.assemble_darcy_slab(sd)
A1 = pg.darcy_saddle_device(full)
nf, ncg = full.num_faces, full.num_cells
bg = torch.from_numpy(np.random.default_rng(1).normal(size=nf + ncg)).to(dev)
keysd = DD.owned_keys.cpu()
isf = keysd < (1 << 60)
god = torch.empty(keysd.numel(), dtype=torch.int64)
god[isf] = torch.searchsorted(full._structured["face_keys"], keysd[isf])
god[~isf] = nf + (keysd[~isf] - (1 << 60))
god = god.to(dev)
cnt = torch.tensor([DD.n_owned], device=dev)
dist.all_reduce(cnt)
assert int(cnt.item()) == nf + ncg
x1, i1 = minres_device(A1, bg, rtol=1e-9, maxiter=4000)
xd, infod, kd, histd = pd.dist_minres(DD, bg[god].contiguous(), rtol=1e-9, maxiter=4000)
assert infod == i1.info == 0, (infod, i1.info)
assert abs(kd - i1.iterations) <= max(3, i1.iterations // 50), (kd, i1.iterations)
mm = min(30, kd, i1.iterations)
assert np.abs(histd[:mm] - i1.residuals[:mm]).max() <= 1e-9 * i1.residuals[0]
# residual of the distributed solution measured with the single-GPU operator
xfull = torch.zeros(nf + ncg, dtype=torch.float64, device=dev)
xfull[god] = xd
dist.all_reduce(xfull)
A1s = sps.csr_array((A1.data.cpu().numpy(), A1.indices.cpu().numpy(), A1.indptr.cpu().numpy()), shape=A1.shape)
res_d = np.linalg.norm(bg.cpu().numpy() - A1s @ xfull.cpu().numpy())
res_1 = np.linalg.norm(bg.cpu().numpy() - A1s @ x1.cpu().numpy())
assert res_d <= 10 * max(res_1, 1e-9 * float(bg.norm())), (res_d, res_1)
if rank == 0:
    print(f"dist_minres: {kd} iterations (single GPU {i1.iterations}), residual {res_d:.2e} vs {res_1:.2e}")
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST CHECK OK")
