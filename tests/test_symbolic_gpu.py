"""COO -> CSR on random cell->dof tables: every row kernel of the two-level symbolic phase
(k_rowthread, k_rowins with its overflow list, k_rowhash with 8/16/32 lanes per row, k_rowsort, k_row_canon) and the full-sort
fallback against a scipy oracle, through engine.symbolic / engine.numeric (ctypes -> C ABI).

The tables have no geometry: `L` distinct dofs per cell drawn from a sliding window (locality as in
a mesh) plus one hub dof shared by exactly `hub` cells, which sets the largest incidence count and
therefore the kernel that pgb_csr_symbolic_sort picks."""

import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu


class _Table:
    """Stand-in for engine.GridPack: just the cell->dof table of one 'space'."""

    def __init__(self, dofs, n_rows, device):
        self._dofs, self._n, self.device, self.plans = dofs, n_rows, device, {}

    def dofs(self, space):
        return self._dofs, self._n


def _random_table(L, nc, n_rows, hub, seed):
    rng = np.random.default_rng(seed)
    W = L + 6
    off = np.argsort(rng.random((nc, W)), axis=1)[:, :L] + 1  # L distinct offsets in [1, W]
    start = rng.integers(1, n_rows - W - 1, size=nc)
    dofs = (start[:, None] + off).astype(np.int32)  # never dof 0
    cells = rng.choice(nc, size=hub, replace=False)
    # the hub: dof 0 in exactly `hub` cells, whose other dofs come from a small set (a row may have at
    # most 255 distinct columns in the two-level path)
    dofs[cells] = (rng.integers(1, 40, size=hub)[:, None] + off[cells]).astype(np.int32)
    dofs[cells, rng.integers(0, L, size=hub)] = 0
    return dofs


def _oracle(dofs, n_rows, vals):
    nc, L = dofs.shape
    rows = np.repeat(dofs, L, axis=1).ravel()
    cols = np.tile(dofs, (1, L)).ravel()
    A = sps.coo_array((vals.ravel(), (rows, cols)), shape=(n_rows, n_rows)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


CASES = [
    # L, nc, n_rows, hub, row kernel expected in mode 0 (prefix of pgb_csr_last_row_kernel())
    (3, 4000, 9000, 1, "k_rowthread<"),       # k_rowthread MI<=4
    (3, 4000, 3000, 8, "k_rowthread<8,3,32"),
    (3, 4001, 3000, 8, "k_rowthread<8,3,32"),   # nc*L not a multiple of 4: scalar tails of k_inc_count / k_inc_place
    (5, 2999, 3100, 8, "k_rowthread<8,5,64"),
    (3, 130, 131, 3, "k_rowthread<"),           # a single block, ragged last group       # k_rowthread MI=8
    (3, 6000, 2500, 16, "k_rowthread<16,3,64"),      # k_rowthread MI=16, 48 keys
    (3, 6000, 2500, 30, "k_rowins<3,"),      # k_rowins; the hub row (> 24 distinct columns) comes back through k_rowhash_list
    (3, 6000, 2500, 80, "k_rowhash<32,4,3,512>"),      # k_rowhash 32 lanes x 4, table 512
    (3, 6000, 2500, 150, "k_rowsort<32,16,3"),     # k_row_canon + k_rowsort (450 candidates)
    (3, 6000, 2500, 200, "k_rowlong<3>"),  # > 512 candidates: one block per row
    (3, 6000, 2500, 300, "k_rowlong<3>"),  # > 255 incidences: saturated ranks take the overflow counter
    (4, 9000, 2500, 700, "k_rowlong<4>"),
    (3, 9000, 2500, 2100, ""),             # > 2048 incidences: full-sort fallback
    (4, 3000, 9000, 2, "k_rowthread<"),       # k_rowthread MI=2
    (4, 3000, 4000, 4, "k_rowthread<"),
    (4, 5000, 2500, 16, "k_rowthread<16,4,64"),      # k_rowthread MI=16, 64 keys
    (4, 5000, 2500, 24, "k_rowins<4,24>"),      # k_rowins + k_rowhash_list (table 128)
    (4, 5000, 2500, 32, "k_rowins<4,32>"),      # k_rowins + k_rowhash_list (table 512)
    (4, 5000, 2500, 60, "k_rowhash<32,4,4,512>"),      # 32 lanes x 4
    (4, 5000, 2500, 100, "k_rowsort<32,16,4"),     # k_rowsort
    (5, 3000, 9000, 2, "k_rowthread<"),
    (5, 3000, 3000, 8, "k_rowthread<8,5,64"),       # k_rowthread MI=8, 40 of 64 keys
    (5, 4000, 2500, 12, "k_rowhash<16,1,5,128>"),      # k_rowhash 16 lanes, table 128
    (5, 4000, 2500, 30, "k_rowhash<32,1,5,512>"),      # 32 lanes, table 512
    (5, 4000, 2500, 50, "k_rowhash<32,4,5,512>"),      # 32 lanes x 4
    (6, 3000, 9000, 2, "k_rowthread<"),
    (6, 3000, 3500, 8, "k_rowthread<8,6,64"),       # k_rowthread MI=8, 48 of 64 keys
    (6, 4000, 3000, 16, "k_rowhash<16,1,6,128>"),      # k_rowhash 16 lanes, table 128
    (6, 4000, 3000, 40, "k_rowhash<32,4,6,512>"),      # 32 lanes x 4, table 512
    (10, 2000, 9000, 2, "k_rowthread<"),        # Lagrange2 in 3D: edge rows ...
    (10, 3000, 6000, 24, "k_rowhash<32,1,10,512>"),  # ... and node rows (240 candidates)
    (12, 2000, 9000, 1, "k_rowthread<"),
    (12, 2000, 9000, 2, "k_rowthread<"),      # k_rowthread MI=2 (BDM1 in 3D)
    (12, 2000, 9000, 4, "k_rowthread<4,12,64"),      # 48 of 64 keys
    (12, 3000, 9000, 7, "k_rowhash<8,1,12,128>"),      # k_rowhash 8 lanes, table 128
    (12, 3000, 6000, 14, "k_rowhash<16,1,12,256>"),     # 16 lanes, table 256
    (12, 3000, 6000, 20, "k_rowhash<32,1,12,512>"),     # 32 lanes, table 512
]


@pytest.mark.parametrize("L,nc,n_rows,hub,kernel", CASES)
def test_random_tables_all_modes(L, nc, n_rows, hub, kernel):
    import torch

    from pygeon_b200 import _lib, engine

    dev = torch.device("cuda", 0)
    dofs = _random_table(L, nc, n_rows, hub, seed=L * 1000 + hub)
    counts = np.bincount(dofs.ravel(), minlength=n_rows)
    assert counts[0] == hub
    rng = np.random.default_rng(7)
    vals = rng.uniform(0.5, 1.5, size=(nc, L, L))
    ref = _oracle(dofs, n_rows, vals)
    dd = torch.from_numpy(dofs).to(dev)
    vd = torch.from_numpy(vals).to(dev)
    LP = (L + 3) & ~3
    first = None
    for it, mode in enumerate((0, 0, 3, 4, 2, 1)):
        plan = engine.symbolic(_Table(dd, n_rows, dev), "t", mode=mode)
        used = _lib.lib.pgb_csr_last_row_kernel().decode().split("+")  # one name per kind of row range
        if it == 0 and counts.max() == hub:  # the hub's chunk of rows gets this kernel
            assert (kernel and any(u.startswith(kernel) for u in used)) or (kernel == "" and plan.mode == 1), used
        if mode == 3 and plan.mode == 0:
            assert all(u.startswith("k_rowsort<") for u in used), used
        assert np.array_equal(plan.indptr.cpu().numpy(), ref.indptr), (mode, plan.mode)
        assert np.array_equal(plan.indices.cpu().numpy(), ref.indices), (mode, plan.mode)
        if plan.mode == 0:
            # the local rows where K1 would put them: row (c, a) at inc_pos[c*L + a]
            vt = torch.zeros((nc * L, LP), dtype=torch.float64, device=dev)
            vt[plan.inc_pos.long(), :L] = vd.reshape(nc * L, L)
            data = engine.numeric(plan, vt.reshape(nc, L * LP))
            state = (plan.inc_pos.clone(), plan.slots.clone(), plan.row_start.clone())
            if first is None:
                first = state
            else:  # bit-identical plans whatever the kernel / the order of the integer atomics
                assert all(torch.equal(a, b) for a, b in zip(first, state)), mode
            inc = np.argsort(plan.inc_pos.cpu().numpy().astype(np.int64), kind="stable")
            rows_sorted = dofs.ravel()[inc]
            assert np.all(np.diff(rows_sorted) >= 0)  # (dof, cell) order ...
            same = rows_sorted[1:] == rows_sorted[:-1]
            assert np.all(np.diff(inc)[same] > 0)  # ... ascending incidence inside a row
        else:
            data = engine.numeric(plan, vd.reshape(nc, L * L).contiguous())
        err = np.abs(data.cpu().numpy() - ref.data).max()
        assert err <= 1e-13 * np.abs(ref.data).max(), (mode, err)
    if hub <= 2048 and np.diff(ref.indptr).max() <= 255:
        assert first is not None  # the two-level path did run


def test_overlap_hook_and_late_fallback():
    """engine.symbolic(overlap=...): the hook runs once inc_pos is complete and before the host waits for nnz
    (pgb_csr_symbolic_sort with nnz_host = NULL + pgb_csr_symbolic_result); the plan is the same as without the
    hook; when a row turns out to have more than 255 entries - known only at the end - the call still falls
    back to the full sort and reports mode 1."""
    import torch

    from pygeon_b200 import engine

    dev = torch.device("cuda", 0)
    dofs = _random_table(4, 5000, 2500, 24, seed=3)
    dd = torch.from_numpy(dofs).to(dev)
    seen = []
    plan = engine.symbolic(_Table(dd, 2500, dev), "t", overlap=lambda p: seen.append((p.mode, p.nnz, p.inc_pos.clone())))
    ref = engine.symbolic(_Table(dd, 2500, dev), "t")
    assert len(seen) == 1 and seen[0][0] == 0 and seen[0][1] == -1 and plan.mode == 0
    for a, b in ((plan.indptr, ref.indptr), (plan.indices, ref.indices), (plan.inc_pos, ref.inc_pos),
                 (plan.slots, ref.slots), (plan.row_start, ref.row_start), (seen[0][2], ref.inc_pos)):
        assert torch.equal(a, b)
    # dof 0 in 150 cells whose other dofs are all different: 451 distinct columns in row 0
    wide = np.arange(1, 1 + 150 * 3, dtype=np.int32).reshape(150, 3)
    table = np.concatenate([np.zeros((150, 1), np.int32), wide], axis=1)
    filler = (np.arange(2000)[:, None] % 1500 + np.arange(4)[None, :] + 460).astype(np.int32)
    dofs = np.concatenate([table, filler])
    dd = torch.from_numpy(dofs).to(dev)
    seen = []
    plan = engine.symbolic(_Table(dd, 2500, dev), "t", overlap=lambda p: seen.append(p.mode))
    assert plan.mode == 1 and seen == [0]
    vals = np.random.default_rng(0).uniform(0.5, 1.5, size=(dofs.shape[0], 4, 4))
    ref = _oracle(dofs, 2500, vals)
    assert np.array_equal(plan.indptr.cpu().numpy(), ref.indptr) and np.array_equal(plan.indices.cpu().numpy(), ref.indices)
