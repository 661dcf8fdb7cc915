"""A/B of the K2 numeric kernels on one GPU: one thread per row (flags 0), bulk-async staging (flags 2),
lane groups per row (flags 4); time of the numeric phase alone and max difference to the flags-0 values.
usage: python scripts/numeric_ab.py [n] [spaces...]"""
import sys

sys.path.insert(0, "/root/repo")
import torch

import pygeon_b200 as pg
from pygeon_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
sd = pg.structured_grid(3, n, device=dev)
sd.compute_connectivity()
for space, kind in (("lagrange1", "l1_stiff"), ("rt0", "rt0_mass"), ("nedelec0", "n0_mass"), ("bdm1", "bdm1_mass"),
                    ("darcy", "darcy_saddle"), ("lagrange2", "l2_stiff")):
    if len(sys.argv) > 2 and space not in sys.argv[2:]:
        continue
    raw = engine.RawGrid.upload(sd, dev, ridges=(space in ("nedelec0", "lagrange2")), signs=True)
    raw.check = False
    pack = engine.GridPack.from_raw(raw)
    plan = engine.symbolic(pack, space)
    vals = engine.local_matrices(pack, kind, plan=plan)
    ref = None
    line = f"{space:10s} rows {plan.n_rows:>10d} nnz {plan.nnz:>11d}"
    for flags in (0, 2, 4, 64):
        engine.numeric_flags = flags
        out = engine.numeric(plan, vals)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            engine.numeric(plan, vals, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if ref is None:
            ref = out.clone()
            diff = 0.0
        else:
            diff = float((out - ref).abs().max() / ref.abs().max())
        line += f" | flags {flags}: {ms:7.3f} ms (rel diff {diff:.1e})"
    print(line, flush=True)
    del pack, plan, vals, ref, out, raw
    torch.cuda.empty_cache()
