#!/usr/bin/env python
"""Benchmark of the FEM assembly-and-solve hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3] [--ncube NCUBE]

Workloads
  --config 2 (default at N = 1): BASELINE configs[1] - pg.Lagrange1 stiffness on the 128^3 structured
      tetrahedral unit cube (12 582 912 tets) per GPU, CG iterations on the assembled operator
      (weak scaling when N > 1: 128^3 per GPU).
  --config 3 (default at N > 1): BASELINE configs[2], the north-star scaling case - mixed Darcy RT0/P0
      saddle-point system on the 256^3 tetrahedral cube (100 663 296 tets, 302 M unknowns), cell
      partitioned in z-slabs over the N GPUs (STRONG scaling: the same 100 M tets at every N), one-pass
      assembly per rank + distributed MINRES over NVLink peer memory.  `--config 3 --gpus 1` runs the same
      workload on one GPU (93.5 GiB).

A *step* is one complete (cold) assembly into CSR with NOTHING cached: grid packing, the symbolic phase
(incidence transpose + per-row sort = the sort by (row, col); pattern), the per-cell local matrices and
the ordered reduce.  `value` = cells/s with the grid's raw arrays already resident in HBM; `e2e` = the
same through the public call (`pg.Lagrange1().assemble_stiff_matrix(sd)` / `pg.darcy_saddle_matrix(sd)`)
with host (pinned) grid arrays in and a host scipy matrix out, host<->device copies inside the timed
region.  `--impl reference` / `cpu_baseline` time the UNMODIFIED reference package staged in
baseline/_ref (kind "reference"; the oracle's port of its algorithm, kind "port", only when the staged
copy is absent) on the host cores, on a bounded sample of the workload that the line names.

Timing: CUDA events on the launching stream, W >= 3 warm-up steps, inputs far larger than L2 (the working
set of a step is > 5 GB), max over ranks, clocks sampled during the timed region.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NOMINAL_HBM_GBS = 8000.0  # north_star's nominal HBM3e figure; the measured copy peak is what `frac` uses
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/: config #2, L1 stiffness on 128^3)
TRAFFIC = {"symbolic_rows": 0.908e9, "local_matrices": 2.027e9, "csr_numeric": 2.069e9}

METRIC = "cells assembled/s into CSR"
UNIT = "cells/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms 20` in the background for the duration of the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index=0):
        self.index, self.proc, self.rows = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)  # first samples before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.1)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        self.rows = [[c.strip() for c in line.split(",")] for line in out.strip().splitlines() if line.strip()]

    def summary(self):
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        sm = sorted(int(r[0]) for r in rows)
        mx = [int(r[1]) for r in rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].startswith("Active")})
        pw = [float(r[6]) for r in rows if len(r) > 6 and r[6].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def count_launches(fn) -> int | None:
    """Kernel launches of one call of ``fn``, counted by CUPTI (torch.profiler sees every kernel of the
    process, including libpgb.so's).  None when the profiler is unavailable."""
    try:
        import torch
        from torch.profiler import ProfilerActivity, profile

        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        n = 0
        for e in prof.events():
            if e.device_type is not None and "cuda" in str(e.device_type).lower() and "memcpy" not in e.name.lower() \
                    and "memset" not in e.name.lower():
                n += 1
        return n or None
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ----------------------------------------------------------------------------------------------
def load_reference():
    """(pg_ref, pp_standin, kind): the unmodified reference from baseline/_ref (staged by
    scripts/stage_reference.py; travels to the GPU box) with the porepy stand-in, else (None, None, 'port')."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "pygeon")):
        return None, None, "port"
    for p in (ref_dir, os.path.join(ROOT, "oracle", "_standins")):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import porepy as pp
        import pygeon as rpg

        return rpg, pp, "reference"
    except Exception as e:  # pragma: no cover
        print(f"bench: staged reference failed to import ({e!r}); using the oracle port", file=sys.stderr)
        return None, None, "port"


def cpu_workload(config: int, n_sample: int):
    """(step function -> matrix, cells, description, kind) of the CPU arm on an n_sample^3 Kuhn-tet cube."""
    import scipy.sparse as sps

    import pygeon_b200 as pg

    rpg, pp, kind = load_reference()
    sd = pg.structured_grid(3, n_sample)
    if kind == "reference":
        g = rpg.Grid(3, sd.nodes.copy(), sd.face_nodes.copy(), sd.cell_faces.copy(), "bench")
        g.compute_geometry()
        if config == 2:
            step = lambda: rpg.Lagrange1().assemble_stiff_matrix(g)  # noqa: E731  discretization.py:154-183
        elif config == 4:
            step = lambda: rpg.Nedelec0().assemble_stiff_matrix(g) + rpg.Nedelec0().assemble_mass_matrix(g)  # noqa: E731
        else:
            def step():  # darcy.ipynb cell 10 / test_laplacian_mixed.py:50
                M = rpg.RT0().assemble_mass_matrix(g)
                B = rpg.PwConstants().assemble_mass_matrix(g) @ rpg.RT0().assemble_diff_matrix(g)
                return sps.block_array([[M, -B.T], [B, None]]).tocsc()
        what = "unmodified reference (baseline/_ref, porepy stand-in)"
    else:
        from oracle import fem_oracle as fo

        sd.compute_geometry()
        if config == 2:
            step = lambda: fo.port_stiff("lagrange1", sd)  # noqa: E731
        elif config == 4:
            step = lambda: fo.port_stiff("nedelec0", sd) + fo.port_mass("nedelec0", sd)  # noqa: E731
        else:
            def step():
                M = fo.port_mass("rt0", sd)
                B = sps.diags_array(1.0 / sd.cell_volumes) @ sd.cell_faces.T
                return sps.block_array([[M, -B.T], [B, None]]).tocsc()
        what = "oracle port of the reference algorithm (baseline/_ref absent)"
    name = {2: "pg.Lagrange1().assemble_stiff_matrix", 3: "RT0 mass + P0 mass @ div + block_array (Darcy saddle)",
            4: "pg.Nedelec0().assemble_stiff_matrix + assemble_mass_matrix"}[config]
    return step, sd.num_cells, f"{name} on {n_sample}^3 structured tet cube ({sd.num_cells} tets); {what}", kind


def time_cpu(step, max_steps: int, budget_s: float):
    times = []
    t_start = time.perf_counter()
    while len(times) < max(max_steps, 1):
        t = time.perf_counter()
        step()
        times.append(time.perf_counter() - t)
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    step, cells, what, kind = cpu_workload(cfg, args.cpu_n)
    if args.warmup > 0:
        step()
    times = time_cpu(step, args.steps, budget_s=150.0)
    dt = sum(times) / len(times)
    v = cells / dt
    full = workload_config(cfg, {2: args.n, 3: args.n3, 4: args.n4}[cfg], args.gpus)
    conf = {"workload": what + " per step; scipy sparse kernels are single-threaded",
            "n_per_side": args.cpu_n, "cells": cells, "extrapolated_to": full["workload"],
            "note": "CPU throughput is flat or falling with size (SURVEY section 6), so cells/s measured on the "
                    "sample is an upper bound for the full workload; 5 kB/tet of host RAM rules out the full size"}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak" if cfg == 2 else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": what, "host_cpus": os.cpu_count()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(cfg, n, gpus):
    if cfg == 2:
        return {"workload": f"pg.Lagrange1 stiffness (K=I, w=1) on {n}^3 structured Kuhn-tet unit cube "
                            f"({6 * n ** 3} tets, {(n + 1) ** 3} nodes) per GPU; BASELINE configs[1]",
                "n_per_side": n, "cells_per_gpu": 6 * n ** 3, "parallelism": f"cell-partitioned x{gpus}",
                "l2": "inputs larger than L2 (step working set > 5 GB)"}
    if cfg == 4:
        return {"workload": f"pg.Nedelec0 curl-curl + mass (Maxwell type) on {n}^3 structured Kuhn-tet cube "
                            f"({6 * n ** 3} tets, {7 * n ** 3 + 9 * n ** 2 + 3 * n} edges), one symbolic plan, CG; "
                            f"BASELINE configs[3], the same cells at every GPU count",
                "n_per_side": n, "cells_total": 6 * n ** 3, "parallelism": f"cell-partitioned z-slabs x{gpus}",
                "l2": "inputs larger than L2 (step working set > 10 GB per GPU)"}
    return {"workload": f"mixed Darcy RT0/P0 saddle system [[M,-B^T],[-B,0]] on {n}^3 structured Kuhn-tet cube "
                        f"({6 * n ** 3} tets, {12 * n ** 3 + 6 * n ** 2 + 6 * n ** 3} unknowns), one-pass assembly + MINRES; "
                        f"BASELINE configs[2], the same cells at every GPU count",
            "n_per_side": n, "cells_total": 6 * n ** 3, "parallelism": f"cell-partitioned z-slabs x{gpus}",
            "l2": "inputs larger than L2 (step working set > 10 GB per GPU)"}


# ----------------------------------------------------------------------------------------------
# helpers of our arm
# ----------------------------------------------------------------------------------------------
def pinned_view(arr):
    import numpy as np
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t.numpy(), t


def pin_grid(sd):
    keep = []
    for mat in (sd.cell_faces, sd.face_nodes):
        mat.indices, t1 = pinned_view(mat.indices)
        mat.data, t2 = pinned_view(mat.data)
        keep += [t1, t2]
    sd.nodes, t3 = pinned_view(sd.nodes)
    keep.append(t3)
    return keep


class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from pygeon_b200 import dist as pgd

        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa_cpus = pgd.bind_to_gpu_numa(self.local) if os.environ.get("PGB_BIND_NUMA", "1") != "0" else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.hbm_peak, self.peak_src = peaks()
        self.torch, self.dist = torch, dist

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def fractions(gbs, hbm_peak, world=1):
    return {"gbs": gbs, "frac": gbs / (hbm_peak * world), "frac_nominal_8tbs": gbs / (NOMINAL_HBM_GBS * world)}


def cpu_baseline_leg(args, cfg):
    step, cells, what, kind = cpu_workload(cfg, args.cpu_n)
    times = time_cpu(step, 1, budget_s=60.0)
    dt = times[0]
    return {"value": cells / dt, "unit": UNIT, "cores": 1, "kind": kind, "host_cpus": os.cpu_count(),
            "sample": f"one call of {what}: {dt:.1f} s"}


def verify_config2(A, sd):
    """Full CSR of the benchmarked matrix against the oracle's closed form (bit-exact pattern, 1e-12 values)."""
    import numpy as np

    from oracle import fem_oracle as fo

    t = time.perf_counter()
    ref = fo.closed_form("lagrange1", "stiff", sd, fast=True).tocsc()
    ref.sort_indices()
    ok = (A.shape == ref.shape and A.nnz == ref.nnz and np.array_equal(A.indptr, ref.indptr)
          and np.array_equal(A.indices, ref.indices))
    err = float(np.abs(A.data - ref.data).max() / np.abs(ref.data).max()) if ok else None
    ok = ok and err <= 1e-12
    return {"parity_at_config": "ok" if ok else "FAILED", "max_rel_err": err, "nnz": int(ref.nnz),
            "checked": "indptr, indices bit-exact; values 1e-12 * max|A| against oracle/fem_oracle.closed_form",
            "oracle_s": time.perf_counter() - t}


# ----------------------------------------------------------------------------------------------
# config #2: P1 stiffness, 128^3 per GPU
# ----------------------------------------------------------------------------------------------
def run_config2(args, cx: Ctx):
    import ctypes as C

    import numpy as np

    import pygeon_b200 as pg
    from pygeon_b200 import _lib, engine
    from pygeon_b200 import dist as pgd
    from pygeon_b200.solvers import cg_device

    torch, dist, dev, world, rank = cx.torch, cx.dist, cx.dev, cx.world, cx.rank
    hbm_peak = cx.hbm_peak
    n = args.n
    dist_parity = pgd.self_check(6, 4 * world + 2) if world > 1 else None

    # N>1: rank r owns n cube layers of the n x n x (n*N) box and also evaluates the halo layer above them
    lay = pgd.slab_layout(n, n * world, world, rank)
    sd = pgd.local_slab_grid(lay, device=dev)
    keep = pin_grid(sd)
    nc, nn = sd.num_cells, sd.num_nodes
    nc_owned = 6 * n * n * lay.nz_owned
    disc = pg.Lagrange1()
    raw = engine.RawGrid.upload(sd, dev)
    raw.check = False  # no device->host validity read inside the device-resident step
    torch.cuda.synchronize()
    ev = cx.ev
    phase_names = ["pack", "symbolic", "local", "numeric"]

    def device_step(record=None):
        if record is None:
            pack = engine.GridPack.from_raw(raw)
            return engine.assemble_on_pack(pack, "lagrange1", "l1_stiff")
        e = [ev() for _ in range(5)]
        e[0].record()
        pack = engine.GridPack.from_raw(raw)
        e[1].record()
        plan = engine.symbolic(pack, "lagrange1")
        e[2].record()
        vals = engine.local_matrices(pack, "l1_stiff", plan=plan)
        e[3].record()
        data = engine.numeric(plan, vals)
        e[4].record()
        record.append(e)
        return plan, data

    for _ in range(max(args.warmup, 3)):
        plan, data = device_step()
    cx.barrier()
    launches_per_step = count_launches(device_step)
    rec = []
    start, stop = ev(), ev()
    with ClockSampler(cx.local) as clocks:
        cx.barrier()
        start.record()
        for _ in range(args.steps):
            plan, data = device_step()
        stop.record()
        cx.barrier()
        ms_total = start.elapsed_time(stop)
        for _ in range(min(args.steps, 5)):  # serialised steps, only for the per-phase attribution
            device_step(rec)
        torch.cuda.synchronize()
        # ---- warm steps (plan reused) ------------------------------------------------------
        pack = engine.GridPack.from_raw(raw)
        plan = engine.symbolic(pack, "lagrange1")
        vals = torch.empty((nc, 16), dtype=torch.float64, device=dev)
        out = torch.empty((plan.nnz,), dtype=torch.float64, device=dev)
        for _ in range(3):
            engine.numeric(plan, engine.local_matrices(pack, "l1_stiff", out=vals, plan=plan), out=out)
        w0, w1, w2 = ev(), ev(), ev()
        wl, wn = 0.0, 0.0
        torch.cuda.synchronize()
        for _ in range(args.steps):
            w0.record()
            engine.local_matrices(pack, "l1_stiff", out=vals, plan=plan)
            w1.record()
            engine.numeric(plan, vals, out=out)
            w2.record()
            torch.cuda.synchronize()
            wl += w0.elapsed_time(w1)
            wn += w1.elapsed_time(w2)
        wl /= args.steps
        wn /= args.steps
        _lib.lib.pgb_set_profile(1)
        ph = np.zeros(4)
        for _ in range(3):
            engine.symbolic(engine.GridPack.from_raw(raw), "lagrange1")
            buf = (C.c_double * 4)()
            _lib.lib.pgb_get_phase_ms(buf, 4)
            ph += np.array(list(buf)) / 3
        _lib.lib.pgb_set_profile(0)
    ms_step = cx.max_over_ranks(ms_total / args.steps)
    phases = {k: 0.0 for k in phase_names}
    for e in rec:
        for i, k in enumerate(phase_names):
            phases[k] += e[i].elapsed_time(e[i + 1]) / len(rec)
    nnz, n_coo = plan.nnz, plan.n_coo

    # ---- roofline bookkeeping (algorithmic bytes) --------------------------
    n_inc = nc * 4
    b_local = nc * (16 + 16 + 128) + 32 * nn
    b_numeric = n_coo * 9 + 8 * nn + nnz * 8
    b_transpose = n_inc * (4 + 4 + 1 + 1 + 4 + 4 + 4 + 4) + 12 * nn
    b_rowsort = n_inc * 4 + 8 * nn + nc * 16 + n_inc * 4 + n_coo * 1 + nnz * 4 + 4 * nn
    b_step_algo = nc * 16 + 32 * nn + nnz * 12 + 4 * (nn + 1)
    kern = {
        "local_matrices": {"ms": wl, "bytes": b_local, **fractions(b_local / wl / 1e6, hbm_peak)},
        "csr_numeric": {"ms": wn, "bytes": b_numeric, **fractions(b_numeric / wn / 1e6, hbm_peak)},
        "symbolic_total": {"ms": phases["symbolic"], "bytes": b_transpose + b_rowsort + 16 * nnz,
                           "gbs": (b_transpose + b_rowsort + 16 * nnz) / phases["symbolic"] / 1e6},
        "pack": {"ms": phases["pack"]},
        "symbolic_count_scan": {"ms": float(ph[0])},
        "symbolic_place_rowcanon": {"ms": float(ph[1])},
        "symbolic_rowsort": {"ms": float(ph[2]), "bytes": b_rowsort, "gbs": b_rowsort / max(ph[2], 1e-9) / 1e6,
                             "kernel": _lib.lib.pgb_csr_last_row_kernel().decode()},
        "symbolic_scan": {"ms": float(ph[3])},
    }

    # ---- CG on the assembled operator (S + M, SPD, same pattern) --------------------------------
    S = engine.DeviceCSR((nn, nn), plan.indptr, plan.indices, out.clone())
    Mv = engine.numeric(plan, engine.local_matrices(pack, "l1_mass", out=vals, plan=plan))
    S.data += Mv
    cg_iters = args.cg_iters
    if world == 1:
        b = torch.ones(nn, dtype=torch.float64, device=dev)
        cg_device(S, b, rtol=0.0, atol=0.0, maxiter=5)
        torch.cuda.synchronize()
        c0, c1 = ev(), ev()
        c0.record()
        x, info = cg_device(S, b, rtol=0.0, atol=0.0, maxiter=cg_iters)
        c1.record()
        torch.cuda.synchronize()
        cg_ms = c0.elapsed_time(c1) / max(info.iterations, 1)
        b_cg = 12 * nnz + 4 * (nn + 1) + 88 * nn
        solve = {"solver": "cg (device-resident recurrence, pgb_cg)", "iters": info.iterations, "ms_per_iter": cg_ms,
                 "bytes_per_iter": b_cg, **fractions(b_cg / cg_ms / 1e6, hbm_peak), "n": nn, "nnz": nnz,
                 "note": "vectors (17 MB each) are L2 resident at this size"}
        # the multi-GPU kernels (pgb_dist_cg) on a one-rank communicator: their intrinsic cost without a peer
        D1 = pgd.build_dist_csr(S.indptr, S.indices, S.data, torch.ones(nn, dtype=torch.bool), torch.arange(nn))
        pgd.dist_cg_peer(D1, b, rtol=0.0, atol=0.0, maxiter=5)
        torch.cuda.synchronize()
        c0.record()
        _, _, its1, _ = pgd.dist_cg_peer(D1, b, rtol=0.0, atol=0.0, maxiter=cg_iters)
        c1.record()
        torch.cuda.synchronize()
        solve["peer_kernels_one_rank_ms_per_iter"] = c0.elapsed_time(c1) / max(its1, 1)
        pgd.peer_release(D1)
        del D1
    else:
        D = pgd.build_dist_csr(S.indptr, S.indices, S.data, pgd.owned_mask(sd, "lagrange1"),
                               pgd.global_keys(sd, "lagrange1"))
        bo = torch.ones(D.n_owned, dtype=torch.float64, device=dev)
        solve = {}
        for name, fn in (("peer", pgd.dist_cg_peer), ("nccl", pgd.dist_cg)):
            try:
                fn(D, bo, rtol=0.0, atol=0.0, maxiter=3)
                cx.barrier()
                c0, c1 = ev(), ev()
                c0.record()
                _, _, its, _ = fn(D, bo, rtol=0.0, atol=0.0, maxiter=cg_iters)
                c1.record()
                cx.barrier()
            except Exception as e:  # e.g. CUDA IPC unavailable: the NCCL variant then carries the line
                if name != "peer":
                    raise
                solve["peer_error"] = repr(e)[:300]
                continue
            ms = cx.max_over_ranks(c0.elapsed_time(c1) / max(its, 1))
            b_cg = cx.sum_over_ranks(12.0 * D.nnz + 4 * (D.n_owned + 1) + 88 * D.n_owned)
            solve[name] = {"iters": its, "ms_per_iter": ms, "bytes_per_iter": b_cg,
                           **fractions(b_cg / ms / 1e6, hbm_peak, world)}
        halo = sum(int(v.numel()) for v in D.send_idx.values()) * 8
        solve.update({"solver": "cg, row-distributed: 'peer' = halo stores + global sums inside the Krylov kernels over "
                                "NVLink peer memory (pgb_dist_cg, 3 launches / iteration); 'nccl' = grouped send/recv + "
                                "2 all-reduces per iteration",
                      "halo_bytes_sent_per_iter_rank0": halo, "n_owned_rank0": D.n_owned,
                      **{k: (solve.get("peer") or solve["nccl"])[k] for k in ("ms_per_iter", "gbs", "frac", "frac_nominal_8tbs")}})
        try:
            pgd.peer_release(D)
        except Exception:
            pass

    # ---- end to end through the public API (host arrays in, scipy matrix out) -------------------
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(3):
        engine.clear_cache(sd)
        A = disc.assemble_stiff_matrix(sd)
    engine.clear_cache(sd)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        engine.clear_cache(sd)
        A = disc.assemble_stiff_matrix(sd)
    torch.cuda.synchronize()
    e2e_s = cx.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    h2d = raw.h2d_bytes
    d2h = A.indptr.nbytes + A.indices.nbytes + A.data.nbytes
    assert A.shape == (nn, nn) and A.nnz == nnz
    verify = None
    if rank == 0 and world == 1 and args.verify:
        verify = verify_config2(A, sd)
    cells_note = (f"{nc} cells evaluated per rank of which {nc_owned} owned "
                  f"(halo layer is redundant work and is not counted)")
    if rank != 0:
        return None

    cpu = cpu_baseline_leg(args, 2) if (world == 1 and not args.no_cpu) else None
    dom = max(("symbolic", "local", "numeric", "pack"), key=lambda k: phases[k])
    # the dominant single kernel of the cold step
    rows_label = kern["symbolic_rowsort"]["kernel"] + " (per-row dedupe + sort of the symbolic phase)"
    cand = {rows_label: (kern["symbolic_rowsort"]["ms"], kern["symbolic_rowsort"]["gbs"], "symbolic_rows"),
            "k_local<L1_STIFF,3> (K1)": (wl, kern["local_matrices"]["gbs"], "local_matrices"),
            "k_numeric_t<4,128> (K2 numeric)": (wn, kern["csr_numeric"]["gbs"], "csr_numeric")}
    kname = max(cand, key=lambda k: cand[k][0])
    roof = {"bound": "hbm", "kernel": kname, "achieved": cand[kname][1], "peak": hbm_peak, "unit": "GB/s",
            "frac": cand[kname][1] / hbm_peak, "frac_nominal_8tbs": cand[kname][1] / NOMINAL_HBM_GBS,
            "traffic": TRAFFIC.get(cand[kname][2]), "peak_source": cx.peak_src,
            "dominant_phase": dom,
            "kernels": {"K1_local_matrices": kern["local_matrices"], "K2_csr_numeric": kern["csr_numeric"],
                        "K3_cg_iteration": {k: solve[k] for k in ("ms_per_iter", "gbs", "frac", "frac_nominal_8tbs")},
                        "symbolic_row_kernel": kern["symbolic_rowsort"]},
            "note": "achieved = algorithmic bytes / CUDA-event time of the kernel; K1 / K2 / K3 are the HBM-bound "
                    "kernels (fractions against the measured copy peak and the nominal 8 TB/s)"}
    value = world * nc_owned / (ms_step * 1e-3)
    conf = workload_config(2, n, world)
    if verify:
        conf["parity_at_config"] = verify["parity_at_config"]
    if dist_parity:
        conf["dist_parity"] = "ok" if dist_parity["ok"] else "FAILED"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf,
        "roofline": roof,
        "step_algorithmic": {"bytes": b_step_algo, **fractions(b_step_algo / ms_step / 1e6, hbm_peak),
                             "note": "read grid once + write CSR once, over the whole cold step"},
        "phases_ms": phases, "kernels": kern, "cells": cells_note,
        "warm": {"value": world * nc_owned / ((wl + wn) * 1e-3), "unit": UNIT, "ms_per_step": wl + wn,
                 "note": "symbolic plan reused (K1 + K2 numeric only)"},
        "solve": solve, "verify": verify, "dist_parity": dist_parity,
        "e2e": {"value": world * nc_owned / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "host_cpus_bound": len(cx.numa_cpus) if cx.numa_cpus else None},
        "cpu_baseline": cpu,
        "gpu_launches": (launches_per_step * args.steps) if launches_per_step else None,
        "gpu_launches_note": "kernel launches of one cold step counted by CUPTI (torch.profiler) x steps",
        "clocks": clocks.summary(),
    }
    del keep
    return line


# ----------------------------------------------------------------------------------------------
# configs #3 / #4: one box cut into z-slabs over the N GPUs (strong scaling), assembly + distributed solve
# ----------------------------------------------------------------------------------------------
STRONG = {
    3: {"space": "darcy", "kinds": ["darcy_saddle"], "solver": "minres", "vec_bytes": 112.0, "ridges": False,
        "solver_name": "distributed MINRES iteration: kd_spmv<MINRES> + kd_mr_c + kd_mr_d (pgb_dist_minres"},
    4: {"space": "nedelec0", "kinds": ["n0_curlcurl", "n0_mass"], "solver": "cg", "vec_bytes": 88.0, "ridges": True,
        "solver_name": "distributed CG iteration: kd_cg_p + kd_spmv<CG> + kd_cg_xr (pgb_dist_cg"},
}


def single_gpu_base(args, cx: Ctx, cfg: int):
    """The same workload on ONE GPU, measured live by rank 0 of a multi-GPU run before the distributed part (the
    other ranks wait): the base of the strong-scaling series, so that every multi-GPU line carries its own
    single-GPU reference.  A short version of run_strong's device part: 3 cold steps, 10 solver iterations."""
    import pygeon_b200 as pg  # noqa: F401
    from pygeon_b200 import engine
    from pygeon_b200 import dist as pgd
    from pygeon_b200.solvers import cg_device, minres_device

    torch, dev = cx.torch, cx.dev
    W = STRONG[cfg]
    n = args.n3 if cfg == 3 else args.n4
    out = None
    try:
        sd = pgd.local_slab_grid(pgd.slab_layout(n, n, 1, 0), device=dev)
        raw = engine.RawGrid.upload(sd, dev, signs=True, ridges=W["ridges"])
        raw.check = False

        def step():
            pack = engine.GridPack.from_raw(raw)
            plan = engine.symbolic(pack, W["space"])
            data = None
            for kind in W["kinds"]:
                d = engine.numeric(plan, engine.local_matrices(pack, kind, plan=plan))
                data = d if data is None else data.add_(d)
            return plan, data

        for _ in range(2):
            plan, data = step()
            del plan, data
        torch.cuda.synchronize()
        e0, e1 = cx.ev(), cx.ev()
        e0.record()
        for _ in range(3):
            plan, data = step()
            if _ < 2:
                del plan, data
        e1.record()
        torch.cuda.synchronize()
        cold_ms = e0.elapsed_time(e1) / 3
        A = engine.DeviceCSR((plan.n_rows, plan.n_rows), plan.indptr, plan.indices, data)
        b = torch.randn(plan.n_rows, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
        solve = (lambda k: minres_device(A, b, rtol=0.0, maxiter=k)) if W["solver"] == "minres" else (
            lambda k: cg_device(A, b, rtol=0.0, atol=0.0, maxiter=k))
        solve(3)
        torch.cuda.synchronize()
        e0.record()
        _, si = solve(10)
        e1.record()
        torch.cuda.synchronize()
        out = {"n_gpus": 1, "measured": "live by rank 0 of this run, before the distributed part",
               "cold_ms": cold_ms, "value": 6 * n ** 3 / (cold_ms * 1e-3), "unit": UNIT,
               "solver_ms_per_iter": e0.elapsed_time(e1) / max(si.iterations, 1), "unknowns": plan.n_rows, "nnz": plan.nnz}
        del A, b, plan, data, raw, sd
    except Exception as e:  # e.g. out of memory on a smaller GPU: the series then has no live base
        out = {"n_gpus": 1, "skipped": repr(e)[:200]}
    torch.cuda.empty_cache()
    return out


def run_strong(args, cx: Ctx, cfg: int):
    import pygeon_b200 as pg
    from pygeon_b200 import engine
    from pygeon_b200 import dist as pgd
    from pygeon_b200.solvers import cg_device, minres_device

    torch, dev, world, rank = cx.torch, cx.dev, cx.world, cx.rank
    hbm_peak = cx.hbm_peak
    W = STRONG[cfg]
    space, kinds = W["space"], W["kinds"]
    n = args.n3 if cfg == 3 else args.n4
    dist_parity = pgd.self_check(6, 4 * world + 2)
    base = None
    if world > 1 and not args.no_base:
        if rank == 0:
            base = single_gpu_base(args, cx, cfg)
        cx.barrier()

    lay = pgd.slab_layout(n, n, world, rank)
    sd = pgd.local_slab_grid(lay, device=dev)
    nc = sd.num_cells
    cells_total = 6 * n ** 3
    raw = engine.RawGrid.upload(sd, dev, signs=True, ridges=W["ridges"])
    raw.check = False
    torch.cuda.synchronize()
    ev = cx.ev

    def device_step(split=None):
        """pack + symbolic (one plan for all operators of the space) + K1 / K2 per operator, summed"""
        pack = engine.GridPack.from_raw(raw)
        plan = engine.symbolic(pack, space)
        if split is not None:
            split.record()
        data = None
        for kind in kinds:
            d = engine.numeric(plan, engine.local_matrices(pack, kind, plan=plan))
            data = d if data is None else data.add_(d)
        return pack, plan, data

    launches_per_step = count_launches(lambda: device_step())
    torch.cuda.empty_cache()  # the warm-up steps below leave the caching allocator with exactly the blocks a step needs
    for _ in range(max(args.warmup, 3)):
        pack, plan, data = device_step()
        del pack, plan, data
    cx.barrier()
    start, stop = ev(), ev()
    with ClockSampler(cx.local) as clocks:
        cx.barrier()
        start.record()
        for _ in range(args.steps):
            pack, plan, data = device_step()
            del pack, plan, data  # nothing of a step survives into the next one
        stop.record()
        cx.barrier()
        ms_step = cx.max_over_ranks(start.elapsed_time(stop) / args.steps)
        mid, end = ev(), ev()
        pack, plan, data = device_step(mid)
        end.record()
        torch.cuda.synchronize()
        warm_ms = cx.max_over_ranks(mid.elapsed_time(end))
        # ---- distributed solve on the assembled system ----------------------------------------
        D = pgd.build_dist_csr(plan.indptr, plan.indices, data, pgd.owned_mask(sd, space), pgd.global_keys(sd, space))
        del pack, plan, data
        torch.cuda.empty_cache()
        gen = torch.Generator(device=dev).manual_seed(rank)
        b = torch.randn(D.n_owned, dtype=torch.float64, device=dev, generator=gen)
        iters = args.solve_iters
        if W["solver"] == "minres":
            variants = [("peer", lambda **kw: pgd.dist_minres_peer(D, b, rtol=0.0, **kw)),
                        ("nccl", lambda **kw: pgd.dist_minres(D, b, rtol=0.0, **kw))]
        else:
            variants = [("peer", lambda **kw: pgd.dist_cg_peer(D, b, rtol=0.0, atol=0.0, **kw)),
                        ("nccl", lambda **kw: pgd.dist_cg(D, b, rtol=0.0, atol=0.0, **kw))]
        solve = {}
        for name, fn in variants[: (2 if world > 1 else 1)]:
            try:
                fn(maxiter=3)
                cx.barrier()
                c0, c1 = ev(), ev()
                c0.record()
                _, _, k, _ = fn(maxiter=iters)
                c1.record()
                cx.barrier()
            except Exception as e:  # e.g. CUDA IPC unavailable: the NCCL variant then carries the line
                if name != "peer" or world == 1:
                    raise
                solve["peer_error"] = repr(e)[:300]
                continue
            ms = cx.max_over_ranks(c0.elapsed_time(c1) / max(k, 1))
            ipb = 8 if D.indptr.dtype == torch.int64 else 4
            byt = cx.sum_over_ranks(12.0 * D.nnz + ipb * (D.n_owned + 1) + W["vec_bytes"] * D.n_owned)
            solve[name] = {"iters": k, "ms_per_iter": ms, "bytes_per_iter": byt, **fractions(byt / ms / 1e6, hbm_peak, world)}
        unknowns = int(cx.sum_over_ranks(D.n_owned))
        nnz_total = int(cx.sum_over_ranks(D.nnz))
        halo = cx.sum_over_ranks(sum(int(v.numel()) for v in D.send_idx.values()) * 8)
        if world == 1:  # the single-GPU driver beside the one-rank peer path
            A1 = engine.DeviceCSR((D.n_owned, D.n_owned), D.indptr, D.indices, D.data)
            one = (lambda **kw: minres_device(A1, b, rtol=0.0, **kw)) if W["solver"] == "minres" else (
                lambda **kw: cg_device(A1, b, rtol=0.0, atol=0.0, **kw))
            one(maxiter=3)
            c0, c1 = ev(), ev()
            c0.record()
            _, si = one(maxiter=iters)
            c1.record()
            torch.cuda.synchronize()
            ms = c0.elapsed_time(c1) / max(si.iterations, 1)
            solve["single_gpu_driver"] = {"iters": si.iterations, "ms_per_iter": ms,
                                          **fractions(solve["peer"]["bytes_per_iter"] / ms / 1e6, hbm_peak)}
        try:
            pgd.peer_release(D)
        except Exception:
            pass
        del D
        torch.cuda.empty_cache()

    # ---- end to end: public host-in / host-out calls on this rank's slab ------------------------
    def public_call():
        engine.clear_cache(sd)
        if cfg == 3:
            return [pg.darcy_saddle_matrix(sd)]
        return [pg.Nedelec0().assemble_stiff_matrix(sd), pg.Nedelec0().assemble_mass_matrix(sd)]

    e2e = None
    try:
        keep = pin_grid(sd)
        for _ in range(2):
            mats = public_call()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            mats = public_call()
        torch.cuda.synchronize()
        e2e_s = cx.max_over_ranks((time.perf_counter() - t0) / 2)
        h2d = sd.cell_faces.indices.nbytes + sd.cell_faces.data.nbytes + sd.face_nodes.indices.nbytes + sd.nodes.nbytes
        if W["ridges"]:
            h2d += sd.face_ridges.indices.nbytes + sd.face_ridges.data.nbytes + sd.ridge_peaks.indices.nbytes
        d2h = sum(A.indptr.nbytes + A.indices.nbytes + A.data.nbytes for A in mats)
        e2e = {"value": cells_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(cx.sum_over_ranks(h2d)),
               "d2h_bytes_per_step": int(cx.sum_over_ranks(d2h)), "ms_per_step": e2e_s * 1e3, "steps": 2,
               "call": ("pg.darcy_saddle_matrix(sd)" if cfg == 3 else
                        "pg.Nedelec0().assemble_stiff_matrix(sd) + assemble_mass_matrix(sd)")
                       + " on every rank's slab (host grid arrays in, scipy csc out)"}
        del mats, keep
    except (MemoryError, RuntimeError) as e:  # host / device memory at --gpus 1
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": repr(e)[:200]}
        cx.barrier()
    if rank != 0:
        return None
    cpu = cpu_baseline_leg(args, cfg) if (world == 1 and not args.no_cpu) else None
    sv = solve.get("peer") or solve["nccl"]
    roof = {"bound": "hbm", "kernel": W["solver_name"] + (", halo + global sums over NVLink peer memory)" if "peer" in solve
                                                          else "; NCCL transport: the peer-memory one failed)"),
            "achieved": sv["gbs"] / world, "peak": hbm_peak, "unit": "GB/s", "frac": sv["frac"],
            "frac_nominal_8tbs": sv["frac_nominal_8tbs"], "traffic": None, "peak_source": cx.peak_src,
            "aggregate_gbs": sv["gbs"], "ms_per_iter": sv["ms_per_iter"], "bytes_per_iter": sv["bytes_per_iter"],
            "nccl_variant": solve.get("nccl"), "single_gpu_driver": solve.get("single_gpu_driver"),
            "assembly": {"cold_ms": ms_step, "warm_ms": warm_ms, "cold_cells_per_s": cells_total / (ms_step * 1e-3),
                         "warm_cells_per_s": cells_total / (warm_ms * 1e-3)},
            "note": "achieved = per-GPU share of the algorithmic bytes of one solver iteration (12 nnz + 4|8 (n+1) + "
                    f"{W['vec_bytes']:.0f} n) / CUDA-event time, max over ranks"}
    conf = workload_config(cfg, n, world)
    conf["dist_parity"] = "ok" if dist_parity["ok"] else "FAILED"
    conf["unknowns"], conf["nnz"] = unknowns, nnz_total
    if base is not None:
        conf["single_gpu_base"] = base  # raw values only: ratios are the reader's to compute
    return {
        "metric": METRIC, "value": cells_total / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf, "roofline": roof,
        "solve": solve, "dist_parity": dist_parity, "halo_bytes_sent_per_iter_all_ranks": halo,
        "cells": f"{nc} cells evaluated on rank 0 (owned slab + one halo layer), {cells_total} in total",
        "e2e": e2e, "cpu_baseline": cpu,
        "gpu_launches": (launches_per_step * args.steps) if launches_per_step else None,
        "gpu_launches_note": "kernel launches of one cold step counted by CUPTI (torch.profiler) x steps",
        "clocks": clocks.summary(),
    }


def run_ours(args):
    cx = Ctx(args)
    try:
        line = run_config2(args, cx) if args.config == 2 else run_strong(args, cx, args.config)
        if line is not None:
            print(json.dumps(line))
    finally:
        cx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4],
                    help="BASELINE config number (1-based): 2 = P1 stiffness 128^3 per GPU, 3 = Darcy saddle 256^3 strong "
                         "scaling, 4 = Nedelec0 curl-curl + mass 192^3 strong scaling + CG; default 2 at --gpus 1, 3 otherwise")
    ap.add_argument("--ncube", dest="n", type=int, default=128, help="config 2: cubes per side (per GPU)")
    ap.add_argument("--ncube3", dest="n3", type=int, default=256, help="config 3: cubes per side of the whole box")
    ap.add_argument("--cpu-n", type=int, default=0,
                    help="cubes per side of the CPU-arm sample (default 64 for config 2, 48 for the heavier 3 / 4)")
    ap.add_argument("--cg-iters", type=int, default=300)
    ap.add_argument("--ncube4", dest="n4", type=int, default=192, help="config 4: cubes per side of the whole box")
    ap.add_argument("--solve-iters", "--minres-iters", dest="solve_iters", type=int, default=50)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-base", action="store_true",
                    help="multi-GPU strong-scaling runs: skip the live single-GPU measurement of the same workload")
    ap.add_argument("--no-verify", dest="verify", action="store_false",
                    help="skip the full-CSR oracle comparison of config 2 (about one minute of host time)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if args.config == 0:
        args.config = 2 if max(world, args.gpus) == 1 else 3
    if not args.cpu_n:
        args.cpu_n = 64 if args.config == 2 else 48
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
