"""One GPU: the single-GPU Krylov drivers beside the peer-memory (multi-GPU) kernels on a one-rank
communicator, on a system whose vectors do not fit L2 (RT0 mass or Darcy saddle on n^3 tets).  Run under
`ncu --metrics gpu__time_duration.sum` for a per-kernel comparison.
usage: python scripts/peer_vs_single.py [n] [iters] [space: rt0|darcy|lagrange1]"""
import sys

sys.path.insert(0, "/root/repo")
import torch

import pygeon_b200 as pg
from pygeon_b200 import dist as pd
from pygeon_b200 import engine
from pygeon_b200.solvers import cg_device, minres_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
space = sys.argv[3] if len(sys.argv) > 3 else "darcy"
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
if space == "darcy":
    A = pg.darcy_saddle_device(sd)
elif space == "rt0":
    A = pg.RT0().assemble_mass_matrix_device(sd)
else:
    A = pg.Lagrange1().assemble_stiff_matrix_device(sd)
    A.data += pg.Lagrange1().assemble_mass_matrix_device(sd).data
N = A.shape[0]
D = pd.build_dist_csr(A.indptr, A.indices, A.data, torch.ones(N, dtype=torch.bool), torch.arange(N))
b = torch.randn(N, dtype=torch.float64, device=dev)


def timed(fn):
    fn(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(iters)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(f"{space} n={n}: rows {N}, nnz {A.nnz}")
print("minres single %.4f ms/it   peer(1 rank) %.4f ms/it" % (
    timed(lambda k: minres_device(A, b, rtol=0.0, maxiter=k)), timed(lambda k: pd.dist_minres_peer(D, b, rtol=0.0, maxiter=k))))
print("cg     single %.4f ms/it   peer(1 rank) %.4f ms/it" % (
    timed(lambda k: cg_device(A, b, rtol=0.0, atol=0.0, maxiter=k)), timed(lambda k: pd.dist_cg_peer(D, b, rtol=0.0, atol=0.0, maxiter=k))))
pd.peer_release(D)
