"""Shared fixtures.  GPU tests are marked ``gpu``; everything else runs on CPU."""

from __future__ import annotations

import glob
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class GoldenCase:
    """A grid + coefficients + the reference's matrices, loaded from tests/golden/<name>.npz."""

    def __init__(self, path):
        from pygeon_b200.grid import SimplexGrid
        from pygeon_b200 import SecondOrderTensor, PARAMETERS, SECOND_ORDER_TENSOR, WEIGHT

        self.name = os.path.splitext(os.path.basename(path))[0]
        z = np.load(path)
        self.z = z
        d = int(z["dim"])
        nn = z["nodes"].shape[1]
        nf = z["fn_indptr"].size - 1
        nc = z["cf_indptr"].size - 1
        fn = sps.csc_array((np.ones(z["fn_indices"].size, dtype=bool), z["fn_indices"], z["fn_indptr"]), shape=(nn, nf))
        cf = sps.csc_array((z["cf_data"], z["cf_indices"], z["cf_indptr"]), shape=(nf, nc))
        sd = SimplexGrid(d, z["nodes"], fn, cf, self.name)
        sd.compute_geometry()
        # use the reference's own face_ridges / ridge_peaks storage (index order matters in 2D)
        nr = int(z["num_ridges"])
        sd.face_ridges = sps.csc_array((z["fr_data"], z["fr_indices"], z["fr_indptr"]), shape=(nr, nf))
        if d == 3:
            sd.ridge_peaks = sps.csc_array((z["rp_data"], z["rp_indices"], z["rp_indptr"]), shape=(nn, nr))
        assert sd.num_edges == int(z["num_edges"])
        self.sd = sd
        self.data = None
        if "K" in z.files:
            sot = SecondOrderTensor(np.ones(nc))
            sot.values = z["K"]
            self.data = {PARAMETERS: {"key": {SECOND_ORDER_TENSOR: sot, WEIGHT: z["w"]}}}

    def matrix(self, space, form):
        k = f"{space}_{form}"
        z = self.z
        n = z[k + "_indptr"].size - 1
        return sps.csr_array((z[k + "_data"], z[k + "_indices"], z[k + "_indptr"]), shape=(n, n))


GOLDEN_NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


@pytest.fixture(scope="session", params=GOLDEN_NAMES)
def golden(request):
    return GoldenCase(os.path.join(GOLDEN_DIR, request.param + ".npz"))
