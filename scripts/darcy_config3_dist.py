"""BASELINE config #3 across N GPUs (torchrun, one rank per GPU): mixed Darcy RT0/P0 saddle system on
the n x n x nz tet box, cell-partitioned in z-slabs; one-pass assembly per rank (owned + halo layer,
no communication) and distributed MINRES (NCCL halo exchange + 3 scalar all-reduces / iteration).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/darcy_config3_dist.py [n] [nz_total] [minres_iters]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pygeon_b200 import dist as pd
from pygeon_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nz = int(sys.argv[2]) if len(sys.argv) > 2 else n
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
peak = 6525.9

lay = pd.slab_layout(n, nz, world, rank)
sd = pd.local_slab_grid(lay, device=dev)
raw = engine.RawGrid.upload(sd, dev)
raw.check = False


def ev():
    return torch.cuda.Event(enable_timing=True)


def barrier():
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()


best = None
for rep in range(3):
    barrier()
    e = [ev() for _ in range(3)]
    e[0].record()
    pack = engine.GridPack.from_raw(raw)
    plan = engine.symbolic(pack, "darcy")
    e[1].record()
    data = engine.numeric(plan, engine.local_matrices(pack, "darcy_saddle", plan=plan))
    e[2].record()
    barrier()
    t = torch.tensor([e[0].elapsed_time(e[2]), e[1].elapsed_time(e[2])], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = t.tolist()
cold_ms, warm_ms = best
D = pd.build_dist_csr(plan.indptr, plan.indices, data, pd.owned_mask(sd, "darcy"), pd.global_keys(sd, "darcy"))
del pack, data
b = torch.randn(D.n_owned, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
pd.dist_minres(D, b, rtol=0.0, maxiter=3)
barrier()
c0, c1 = ev(), ev()
c0.record()
x, info, k, hist = pd.dist_minres(D, b, rtol=0.0, maxiter=iters, check_every=0)
c1.record()
barrier()
t = torch.tensor([c0.elapsed_time(c1) / max(k, 1)], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms_it = float(t.item())
ipb = 8 if D.indptr.dtype == torch.int64 else 4
tot = torch.tensor([12.0 * D.nnz + ipb * (D.n_owned + 1) + 112.0 * D.n_owned, float(D.n_owned), float(D.nnz),
                    float(sum(int(v.numel()) for v in D.send_idx.values()) * 8)], dtype=torch.float64, device=dev)
dist.all_reduce(tot)
cells = 6 * n * n * nz
if rank == 0:
    byt = float(tot[0])
    print(json.dumps({
        "workload": f"mixed Darcy RT0/P0 saddle on {n}x{n}x{nz} cubes ({cells} tets), {world} GPUs, z-slabs",
        "n_gpus": world, "cells": cells, "unknowns": int(tot[1]), "nnz": int(tot[2]),
        "assembly_cold_ms": cold_ms, "assembly_warm_ms": warm_ms,
        "cold_cells_per_s": cells / (cold_ms * 1e-3), "warm_cells_per_s": cells / (warm_ms * 1e-3),
        "minres_iters": k, "minres_ms_per_iter": ms_it, "minres_gbs_aggregate": byt / ms_it / 1e6,
        "minres_frac_of_measured_peak": byt / ms_it / 1e6 / (peak * world),
        "halo_bytes_sent_per_iter_all_ranks": float(tot[3])}))
dist.destroy_process_group()
