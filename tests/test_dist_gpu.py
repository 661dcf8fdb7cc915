"""Distributed path on the GPU.

* one GPU (always runs): the NVLink peer-memory drivers ``pgb_dist_cg`` / ``pgb_dist_minres`` with a
  one-rank communicator - the same kernels (ticket reductions, flag waits, interior-first SpMV) as on N
  ranks, against the single-GPU drivers; plus their edge cases (zero right-hand side, maxiter);
* two GPUs (skipped on a one-GPU box): ``scripts/dist_check.py`` under torchrun, one rank per GPU."""

import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import pygeon_b200 as pg
from pygeon_b200 import dist as pd

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_self_check_one_rank():
    res = pd.self_check(6, 10)
    assert res["ok"], res
    assert res["cg"]["peer"]["iters"] == res["cg"]["single_gpu_iters"]
    assert res["minres"]["peer"]["iters"] == res["minres"]["single_gpu_iters"]


def _one_rank_system(n=9):
    sd = pg.structured_grid(3, n, perturb=0.1)
    sd.compute_geometry()
    A = pg.Lagrange1().assemble_stiff_matrix_device(sd)
    A.data += pg.Lagrange1().assemble_mass_matrix_device(sd).data
    owned = torch.ones(sd.num_nodes, dtype=torch.bool)
    D = pd.build_dist_csr(A.indptr, A.indices, A.data, owned, torch.arange(sd.num_nodes))
    return sd, A, D


def test_peer_cg_matches_single_gpu_driver_bitwise_history():
    sd, A, D = _one_rank_system()
    b = torch.from_numpy(np.random.default_rng(0).normal(size=sd.num_nodes)).cuda()
    xs, si = pg.cg_device(A, b, rtol=1e-11, maxiter=1000)
    x, info, k, hist = pd.dist_cg_peer(D, b, rtol=1e-11, maxiter=1000)
    assert info == 0 and k == si.iterations
    # one rank: same partial sums; the interior-first row order of the SpMV changes the order of the
    # p.q partials only, so the histories agree to rounding
    assert np.abs(hist - si.residuals).max() <= 1e-10 * si.residuals[0]
    assert float((x - xs).abs().max()) <= 1e-10 * float(xs.abs().max())
    # Jacobi preconditioner
    dinv = 1.0 / A.diagonal()
    xs, si = pg.cg_device(A, b, rtol=1e-11, maxiter=1000, dinv=dinv)
    x, info, k, hist = pd.dist_cg_peer(D, b, rtol=1e-11, maxiter=1000, dinv=dinv)
    assert info == 0 and k == si.iterations and float((x - xs).abs().max()) <= 1e-10 * float(xs.abs().max())
    pd.peer_release(D)


def test_peer_edge_cases():
    sd, A, D = _one_rank_system(5)
    z = torch.zeros(sd.num_nodes, dtype=torch.float64, device="cuda")
    x, info, k, hist = pd.dist_cg_peer(D, z)
    assert info == 0 and k == 0 and not bool(x.any())
    x, info, k, hist = pd.dist_minres_peer(D, z)
    assert info == 0 and k == 0 and not bool(x.any())
    b = torch.ones(sd.num_nodes, dtype=torch.float64, device="cuda")
    xs, si = pg.cg_device(A, b, rtol=1e-30, maxiter=7)
    x, info, k, hist = pd.dist_cg_peer(D, b, rtol=1e-30, maxiter=7)
    assert info == 7 and k == 7 and hist.size == 8
    assert np.abs(hist - si.residuals).max() <= 1e-10 * si.residuals[0]
    xs, si = pg.minres_device(A, b, rtol=1e-30, maxiter=7)
    x, info, k, hist = pd.dist_minres_peer(D, b, rtol=1e-30, maxiter=7)
    assert info == 7 and k == 7 and float((x - xs).abs().max()) <= 1e-9 * float(xs.abs().max())
    # two solves on one communicator (sequence numbers carry on)
    x2, info2, k2, _ = pd.dist_minres_peer(D, b, rtol=1e-10, maxiter=500)
    xs, si = pg.minres_device(A, b, rtol=1e-10, maxiter=500)
    assert info2 == 0 and k2 == si.iterations and float((x2 - xs).abs().max()) <= 1e-8 * float(xs.abs().max())
    pd.peer_release(D)


@pytest.mark.parametrize("world", [2])
def test_distributed_assembly_and_solves(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "dist_check.py"), "6", "10"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
