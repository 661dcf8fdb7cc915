"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_check.py [n] [nz_total]

Runs pygeon_b200.dist.self_check: every rank assembles its slab with the CUDA path, the rows it owns are
compared with the rows of the single-GPU assembly of the whole box, and the distributed CG / MINRES
(NVLink peer-memory kernels and the NCCL variant) with the single-GPU drivers.  Also works without
torchrun (one rank).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pygeon_b200 import dist as pd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 10
world = int(os.environ.get("WORLD_SIZE", "1"))
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = pd.self_check(n, nz)
if rank == 0:
    print(json.dumps(res))
    print("DIST CHECK OK" if res["ok"] else "DIST CHECK FAILED")
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if res["ok"] else 1)
