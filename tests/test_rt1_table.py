"""The generated RT1 table (pygeon_b200/csrc/pgb_rt1_table.cuh): the committed header is what
oracle/gen_rt1_table.py produces, and evaluating the table the way the CUDA kernel does reproduces the
closed-form oracle (itself pinned against the reference and its known answers).  CPU only."""

import numpy as np
import pytest

from oracle import fem_oracle as fo
from oracle import gen_rt1_table
from pgb_helpers import delaunay_grid
from pygeon_b200.grid import reference_element, structured_grid


def test_header_is_up_to_date():
    assert open(gen_rt1_table.OUT).read() == gen_rt1_table.render()


def _table_from_header(d):
    import re

    text = open(gen_rt1_table.OUT).read()
    body = re.search(r"double T%d\[(\d+)\] = \{(.*?)\};" % d, text, flags=re.S)
    vals = np.array([float(v) for v in body.group(2).replace("\n", " ").split(",") if v.strip()])
    L = d * (d + 2)
    assert vals.size == int(body.group(1)) == L * L * d * d
    return vals.reshape(L, L, d, d)


@pytest.mark.parametrize("case", ["ref2", "ref3", "kuhn2", "kuhn3", "delaunay3"])
@pytest.mark.parametrize("sym", [True, False])
def test_kernel_formula_matches_closed_form(case, sym):
    """A^T[p, q] = s_p s_q / (d^2 |c|) sum_nm T[p, q, n, m] (K y_n) . y_m, the statement k_local evaluates."""
    sd = {"ref2": lambda: reference_element(2), "ref3": lambda: reference_element(3),
          "kuhn2": lambda: structured_grid(2, 3, perturb=0.2), "kuhn3": lambda: structured_grid(3, 2, perturb=0.2),
          "delaunay3": lambda: delaunay_grid(3, 40, seed=2)}[case]()
    sd.compute_geometry()
    d, nc = sd.dim, sd.num_cells
    rng = np.random.default_rng(3)
    B = rng.normal(size=(3, 3, nc))
    K = np.einsum("iac,jac->ijc", B, B) + np.eye(3)[:, :, None]
    if not sym:
        K = K + 0.3 * rng.normal(size=(3, 3, nc))
    tabs = fo.cell_tables(sd)
    Aref, _ = fo.local_rt1_mass(sd, tabs, None, K)
    nodes, _, signs = fo.rt1_rank_tables(sd, tabs)
    X = sd.nodes[:d].T[nodes]
    Y = X[:, 1:, :] - X[:, :1, :]
    vol = np.abs(np.linalg.det(Y)) / (2.0 if d == 2 else 6.0)
    Q = np.einsum("xyc,cny,cmx->cnm", K[:d, :d, :], Y, Y)  # (K y_n) . y_m
    T = _table_from_header(d)
    q = np.arange(d * (d + 1))
    s = np.ones((nc, d * (d + 2)))
    s[:, : d * (d + 1)] = signs[:, (d * (d + 1) - 1 - q) % (d + 1)]
    At = np.einsum("pqnm,cnm->cpq", T, Q) * (s[:, :, None] * s[:, None, :]) / (d * d * vol)[:, None, None]
    assert np.abs(At - np.transpose(Aref, (0, 2, 1))).max() <= 1e-13 * np.abs(Aref).max()


def test_dof_table_is_a_valid_numbering():
    sd = structured_grid(3, 2)
    sd.compute_geometry()
    dofs = fo.rt1_dofs(sd, fo.cell_tables(sd))
    n = 3 * (sd.num_faces + sd.num_cells)
    assert dofs.min() == 0 and dofs.max() == n - 1 and np.unique(dofs).size == n
    counts = np.bincount(dofs.ravel(), minlength=n)
    assert counts[: 3 * sd.num_faces].max() == 2 and np.all(counts[3 * sd.num_faces:] == 1)
