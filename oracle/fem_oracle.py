"""CPU oracle for the FEM assembly hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
``pygeon_b200`` never does.

Two independent restatements of the reference (compgeo-mox/pygeon, paths below
relative to ``/root/reference/src/pygeon``):

**Part A - "port"**: the reference's own algorithm, i.e. project the conforming
space onto broken piecewise linears and form the sparse triple product
``Pi.T @ M @ Pi`` with scipy (``discretizations/discretization.py:112-136``),
stiffness as ``B.T @ A @ B`` (``:154-183``).  This is what the CPU baseline
times: it has the reference's cost structure (COO->CSC conversions + SpGEMMs).

**Part B - "closed form"**: per-cell local matrices in closed form (SURVEY
section 8(a)) scattered through scipy's COO->CSR.  This is the spec the CUDA
kernels implement.

Pinning: ``tests/test_oracle_vs_reference.py`` checks A and B against the
UNMODIFIED reference imported from ``/root/reference`` (when present), and
``tests/test_oracle_golden.py`` checks them against the reference's own
known-answer matrices (``tests/discretizations/fem/**``) and the committed
fixtures in ``tests/golden/`` generated from the reference by
``oracle/gen_golden.py``.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps

PARAMETERS = "parameters"  # pp.PARAMETERS
SECOND_ORDER_TENSOR = "second_order_tensor"  # utils/common_constants.py:37
WEIGHT = "weight"  # utils/common_constants.py:44


# ---------------------------------------------------------------------------
# coefficient fetch  (params/data.py:61-118)
# ---------------------------------------------------------------------------
def cell_data(sd, data, keyword):
    """Return ``(w[nc], K[3,3,nc])`` with the reference's defaults (w=1, K=I)."""
    params = {}
    if data and PARAMETERS in data and keyword in data[PARAMETERS]:
        params = data[PARAMETERS][keyword]
    w = params.get(WEIGHT, 1)
    if np.isscalar(w):
        w = np.full(sd.num_cells, float(w))
    sot = params.get(SECOND_ORDER_TENSOR, 1)
    if np.isscalar(sot):
        K = np.zeros((3, 3, sd.num_cells))
        K[0, 0] = K[1, 1] = K[2, 2] = float(sot)
    else:
        K = np.asarray(sot.values, dtype=float)
    return np.asarray(w, dtype=float), K


# ---------------------------------------------------------------------------
# per-cell connectivity tables shared by both parts
# ---------------------------------------------------------------------------
def cell_tables(sd):
    """Per-cell tables in the slot order of ``cell_faces`` columns.

    Returns dict with ``faces (nc,d+1)``, ``signs (nc,d+1)``, ``opp (nc,d+1)``
    (node opposite each face slot, ``grids/grid.py:279-309``), ``fnodes
    (nc,d+1,d)`` (nodes of each face in ``face_nodes.indices`` order) and
    ``cnodes (nc,d+1)`` ascending node ids (slot order of the broken P1 space,
    ``fem/l2.py:598-612``).
    """
    d, nc, nf = sd.dim, sd.num_cells, sd.num_faces
    if np.any(np.diff(sd.cell_faces.indptr) != d + 1) or np.any(np.diff(sd.face_nodes.indptr) != d):
        raise NotImplementedError("Grid is not simplicial; cannot compute opposite node.")
    faces = sd.cell_faces.indices.reshape(nc, d + 1).astype(np.int64)
    signs = np.asarray(sd.cell_faces.data, dtype=float).reshape(nc, d + 1)
    fn = sd.face_nodes.indices.reshape(nf, d).astype(np.int64)
    fnodes = fn[faces]  # (nc, d+1, d)
    alln = np.sort(fnodes.reshape(nc, d * (d + 1)), axis=1)
    cnodes = alln[:, ::d]
    if not np.all(alln.reshape(nc, d + 1, d) == cnodes[:, :, None]):
        raise NotImplementedError("Grid is not simplicial; cannot compute opposite node.")
    opp = cnodes.sum(axis=1)[:, None] - fnodes.sum(axis=2)
    return dict(faces=faces, signs=signs, opp=opp, fnodes=fnodes, cnodes=cnodes)


def _slot_of(cnodes, nodes):
    """Slot (0..d) of ``nodes[c, ...]`` inside the ascending ``cnodes[c]``."""
    shp = nodes.shape
    flat = nodes.reshape(shp[0], -1)
    return (flat[:, :, None] > cnodes[:, None, :]).sum(axis=2).reshape(shp)


def _geometry(sd, tabs):
    """|c| and rotated node coordinates; intrinsic to the d+1 node positions."""
    d = sd.dim
    R = np.asarray(getattr(sd, "rotation_matrix", np.eye(d, 3)), dtype=float)
    x = R @ sd.nodes  # (d, nn)
    return x, np.asarray(sd.cell_volumes, dtype=float), R


# ===========================================================================
# Part A: the reference's projection route
# ===========================================================================
def pwlinears_mass(sd, w):
    """Broken-P1 mass ``kron(Mhat, diag(|c| w))`` (fem/l2.py:63-86, 457-470)."""
    d = sd.dim
    mhat = (np.ones((d + 1, d + 1)) + np.eye(d + 1)) / ((d + 1) * (d + 2))
    return sps.kron(mhat, sps.diags_array(sd.cell_volumes * w)).tocsc()


def vecpwlinears_mass(sd, w, K):
    """``kron(I_d, M_P1) @ W(K)`` (vec_discretization.py:235-246, fem/vec_l2.py:66-126)."""
    d = sd.dim
    R = np.asarray(sd.rotation_matrix, dtype=float)
    m0 = sps.kron(sps.eye_array(d), pwlinears_mass(sd, w)).tocsc()
    # same contraction as fem/vec_l2.py:113-114; note it yields (R K^T R^T)[i, j] at [i, j, c]
    Kr = np.tensordot(R.T, np.tensordot(R, K, (1, 0)), (0, 1))
    tiled = np.tile(Kr, d + 1)
    blocks = [[sps.diags_array(tiled[i, j]) for j in range(d)] for i in range(d)]
    return m0 @ sps.block_array(blocks).tocsc()


def proj_lagrange1(sd, tabs):
    """L1 -> broken P1: ones at (slot*nc + c, node) (fem/h1.py:218-237)."""
    d, nc = sd.dim, sd.num_cells
    rows = (np.arange(d + 1)[None, :] * nc + np.arange(nc)[:, None]).ravel()
    cols = tabs["cnodes"].ravel()
    return sps.csc_array((np.ones(rows.size), (rows, cols)), shape=((d + 1) * nc, sd.num_nodes))


def proj_bdm1(sd, tabs):
    """BDM1 -> vector broken P1 (fem/hdiv.py:450-509).

    One entry per (cell, face, face node, component): the dof (f, pos) has the
    nodal value (x_node - x_opp) / (sign * |c| * d) at that node of that cell.
    """
    d, nc, nf = sd.dim, sd.num_cells, sd.num_faces
    x, vol, _ = _geometry(sd, tabs)
    faces, signs, opp, fnodes, cnodes = (tabs[k] for k in ("faces", "signs", "opp", "fnodes", "cnodes"))
    tang = x[:, fnodes] - x[:, opp][:, :, :, None]  # (d, nc, d+1, d)
    vals = tang / (signs[None, :, :, None] * vol[None, :, None, None] * d)
    slot = _slot_of(cnodes, fnodes)  # (nc, d+1, d)
    p1 = slot * nc + np.arange(nc)[:, None, None]
    rows = p1[None] + (d + 1) * nc * np.arange(d)[:, None, None, None]
    cols = faces[:, :, None] + nf * np.arange(d)[None, None, :]
    cols = np.broadcast_to(cols[None], rows.shape)
    return sps.csc_array(
        (vals.ravel(), (rows.ravel(), cols.ravel())), shape=(d * (d + 1) * nc, d * nf)
    )


def proj_rt0(sd, tabs):
    """RT0 -> vector broken P1 = Pi_BDM1 @ vstack([I]*d) (fem/hdiv.py:255-274, 338-351)."""
    up = sps.vstack([sps.eye_array(sd.num_faces)] * sd.dim).tocsc()
    return proj_bdm1(sd, tabs) @ up


def _edge_tables(sd, tabs):
    """(cell, edge) incidences with the edge's end nodes in ``edge_nodes.indices`` order."""
    d, nc = sd.dim, sd.num_cells
    if d == 2:
        cell_edges = tabs["faces"]  # (nc, 3)
        en = sd.face_ridges.indices.reshape(sd.num_faces, 2).astype(np.int64)
    else:
        # NB: astype() copies; abs() on the live matrix would sort the grid's indices in place
        ce = (sd.face_ridges.astype(bool).astype(np.int32) @ sd.cell_faces.astype(bool).astype(np.int32)).tocsc()
        ce.sort_indices()
        cell_edges = ce.indices.reshape(nc, 6).astype(np.int64)
        en = sd.ridge_peaks.indices.reshape(sd.num_ridges, 2).astype(np.int64)
    return cell_edges, en


def proj_nedelec1(sd, tabs):
    """Nedelec1 -> vector broken P1 (fem/hcurl.py:215-291).

    The dof (e, i) is lambda_{n_i} grad(lambda_{partner}); its nodal value at
    n_i is the gradient of the partner's barycentric function, i.e.
    -normal(f)/(sign * |c| * d) with f the face opposite the partner node.
    """
    d, nc = sd.dim, sd.num_cells
    _, vol, R = _geometry(sd, tabs)
    normals = R @ sd.face_normals
    cell_edges, en = _edge_tables(sd, tabs)
    ne = en.shape[0]
    nodes = en[cell_edges]  # (nc, E, 2)
    partner = nodes[:, :, ::-1]
    # face opposite the partner node, per cell
    match = partner[:, :, :, None] == tabs["opp"][:, None, None, :]  # (nc,E,2,d+1)
    fslot = match.argmax(axis=3)
    cidx = np.arange(nc)[:, None, None]
    faces = tabs["faces"][cidx, fslot]
    orien = tabs["signs"][cidx, fslot]
    vals = -normals[:, faces] / (orien[None] * vol[None, :, None, None] * d)  # (d,nc,E,2)
    slot = _slot_of(tabs["cnodes"], nodes)
    p1 = slot * nc + cidx
    rows = p1[None] + (d + 1) * nc * np.arange(d)[:, None, None, None]
    cols = cell_edges[:, :, None] + ne * np.arange(2)[None, None, :]
    cols = np.broadcast_to(cols[None], rows.shape)
    return sps.csc_array(
        (vals.ravel(), (rows.ravel(), cols.ravel())), shape=(d * (d + 1) * nc, 2 * ne)
    )


def proj_nedelec0(sd, tabs):
    """Nedelec0 -> vector broken P1 = Pi_Ne1 @ [I; -I] (fem/hcurl.py:44-61, 166-180)."""
    ne = sd.num_edges
    to_ne1 = sps.vstack([sps.eye_array(ne), -sps.eye_array(ne)]).tocsc()
    return proj_nedelec1(sd, tabs) @ to_ne1


_PROJ = {"lagrange1": proj_lagrange1, "rt0": proj_rt0, "bdm1": proj_bdm1, "nedelec0": proj_nedelec0}


def port_mass(space: str, sd, data=None, keyword="unitary_data", tabs=None):
    """``(Pi.T @ M @ Pi).tocsc()`` (discretization.py:112-136)."""
    tabs = tabs or cell_tables(sd)
    w, K = cell_data(sd, data, keyword)
    Pi = _PROJ[space](sd, tabs)
    M = pwlinears_mass(sd, w) if space == "lagrange1" else vecpwlinears_mass(sd, w, K)
    return (Pi.T @ M @ Pi).tocsc()


def diff_matrix(space: str, sd):
    """Exterior derivative of the space (h1.py:152-179, hcurl.py:82-106, hdiv.py:177-192, 353-371)."""
    d = sd.dim
    if space == "lagrange1":
        return (sd.ridge_peaks.T if d == 3 else sd.face_ridges.T).tocsc()
    if space == "nedelec0":
        return (sd.face_ridges.T if d == 3 else sd.cell_faces.T).tocsc()
    if space == "rt0":
        return sps.csc_array(sd.cell_faces.T)
    if space == "bdm1":
        return (sps.csc_array(sd.cell_faces.T) @ (sps.hstack([sps.eye_array(sd.num_faces)] * d) / d)).tocsc()
    raise KeyError(space)


def p0_mass(sd, data=None, keyword="unitary_data"):
    """``diag(w / |c|)`` (fem/l2.py:335-353)."""
    w, _ = cell_data(sd, data, keyword)
    return sps.diags_array(w / sd.cell_volumes).tocsc()


def darcy_saddle(sd, data=None, keyword="unitary_data", tabs=None, route="closed"):
    """Symmetrised mixed-Darcy system [[M, -B^T], [-B, 0]] built the reference's way
    (tutorials/fluid_flow/darcy.ipynb cell 10 with the second block row negated): M = RT0 mass,
    B = P0 mass @ div."""
    M = port_mass("rt0", sd, data, keyword, tabs) if route == "port" else closed_form("rt0", "mass", sd, data, keyword, tabs)
    B = p0_mass(sd, data, keyword) @ diff_matrix("rt0", sd)
    return sps.block_array([[M, -B.T], [-B, None]]).tocsr()


def port_stiff(space: str, sd, data=None, keyword="unitary_data", tabs=None):
    """``(B.T @ A @ B).tocsc()`` with A the range-space mass (discretization.py:154-183)."""
    d = sd.dim
    rng = {
        ("lagrange1", 3): "nedelec0",
        ("lagrange1", 2): "rt0",
        ("nedelec0", 3): "rt0",
        ("nedelec0", 2): "p0",
        ("rt0", 2): "p0",
        ("rt0", 3): "p0",
        ("bdm1", 2): "p0",
        ("bdm1", 3): "p0",
    }[(space, d)]
    B = diff_matrix(space, sd)
    A = p0_mass(sd, data, keyword) if rng == "p0" else port_mass(rng, sd, data, keyword, tabs)
    return (B.T @ A @ B).tocsc()


# ===========================================================================
# Part B: closed-form local matrices (the spec of the CUDA kernels)
# ===========================================================================
def local_geometry(sd, tabs):
    """Slot-ordered node coordinates, |c|, and barycentric gradients per cell.

    Local slot a <-> the node opposite face slot a.  ``X (nc, d+1, d)``,
    ``vol (nc,)``, ``G (nc, d+1, d)`` with ``G[c, a] = grad(lambda_a)``; all
    recomputed from the node coordinates (SURVEY 8(a) "implementation freedom").
    """
    d = sd.dim
    R = np.asarray(getattr(sd, "rotation_matrix", np.eye(d, 3)), dtype=float)
    x = (R @ sd.nodes).T  # (nn, d)
    X = x[tabs["opp"]]  # (nc, d+1, d)
    E = X[:, 1:, :] - X[:, :1, :]  # (nc, d, d) rows = edge vectors from slot 0
    det = np.linalg.det(E)
    fact = 2.0 if d == 2 else 6.0
    vol = np.abs(det) / fact
    # grad lambda_a for a>=1 are the columns of E^{-1}; grad lambda_0 = -sum
    Einv = np.linalg.inv(E)  # (nc, d, d): Einv[:, :, k] = grad lambda_{k+1}
    G = np.empty_like(X)
    G[:, 1:, :] = np.swapaxes(Einv, 1, 2)
    G[:, 0, :] = -G[:, 1:, :].sum(axis=1)
    return X, vol, G


def rotated_tensor(sd, K):
    d = sd.dim
    R = np.asarray(getattr(sd, "rotation_matrix", np.eye(d, 3)), dtype=float)
    # (nc, d, d) = R K^T R^T, the tensor the reference's contraction produces
    # (fem/vec_l2.py:113-114); identical to R K R^T for the symmetric tensors
    # pp.SecondOrderTensor builds.
    return np.einsum("ia,bac,jb->cij", R, K, R)


def local_lagrange1_mass(sd, tabs, w, K):
    d = sd.dim
    _, vol, _ = local_geometry(sd, tabs)
    mhat = (np.ones((d + 1, d + 1)) + np.eye(d + 1)) / ((d + 1) * (d + 2))
    return (w * vol)[:, None, None] * mhat[None], tabs["opp"]


def local_lagrange1_stiff(sd, tabs, w, K, rot2d=True):
    """w|c| g_i^T K g_j; in 2D with ``rot2d`` the reference's B^T M_RT0 B = (K rot g_i, rot g_j)."""
    d = sd.dim
    _, vol, G = local_geometry(sd, tabs)
    Kr = rotated_tensor(sd, K)
    if d == 2 and rot2d:
        G = np.stack([G[:, :, 1], -G[:, :, 0]], axis=2)
    A = np.einsum("cai,cij,cbj->cab", G, Kr, G) * (w * vol)[:, None, None]
    return A, tabs["opp"]


def local_rt0_mass(sd, tabs, w, K):
    """A_ab = w s_a s_b / (d^2 |c|) [ T/((d+1)(d+2)) + y_a^T K y_b ],  y = x - centroid."""
    d = sd.dim
    X, vol, _ = local_geometry(sd, tabs)
    Kr = rotated_tensor(sd, K)
    Y = X - X.mean(axis=1, keepdims=True)
    YKY = np.einsum("cai,cij,cbj->cab", Y, Kr, Y)
    T = np.einsum("caa->c", YKY)
    s = tabs["signs"]
    A = (YKY + (T / ((d + 1) * (d + 2)))[:, None, None]) * (s[:, :, None] * s[:, None, :])
    A *= (w / (d * d * vol))[:, None, None]
    return A, tabs["faces"]


def local_bdm1_mass(sd, tabs, w, K):
    """dof (a, j!=a): lambda_j * t_aj, t_aj = (x_j - x_a)/(s_a d |c|); A = w|c| Mhat_jl t_aj^T K t_bl."""
    d, nc, nf = sd.dim, sd.num_cells, sd.num_faces
    X, vol, _ = local_geometry(sd, tabs)
    Kr = rotated_tensor(sd, K)
    s = tabs["signs"]
    mhat = (np.ones((d + 1, d + 1)) + np.eye(d + 1)) / ((d + 1) * (d + 2))
    # local dof list: for face slot a, the d nodes of the face in face_nodes order
    fnodes = tabs["fnodes"]  # (nc, d+1, d) global node ids
    jslot = (fnodes[:, :, :, None] == tabs["opp"][:, None, None, :]).argmax(axis=3)  # local slot of node
    cidx = np.arange(nc)[:, None, None]
    aslot = np.broadcast_to(np.arange(d + 1)[None, :, None], jslot.shape)
    t = (X[cidx, jslot] - X[cidx, aslot]) / (s[:, :, None, None] * d * vol[:, None, None, None])
    L = d * (d + 1)
    t = t.reshape(nc, L, d)
    js = jslot.reshape(nc, L)
    tKt = np.einsum("cpi,cij,cqj->cpq", t, Kr, t)
    A = tKt * mhat[js[:, :, None], js[:, None, :]] * (w * vol)[:, None, None]
    dofs = (tabs["faces"][:, :, None] + nf * np.arange(d)[None, None, :]).reshape(nc, L)
    return A, dofs


_PAIRS3 = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
_PAIRS2 = ((1, 2), (0, 2), (0, 1))  # 2D: edge (face) slot a joins the two other slots


def local_edges(sd, tabs):
    """Per cell: global edge ids ``(nc, E)`` and the local slots (p, q) of the edge's
    end nodes in ``edge_nodes.indices`` order (Whitney form lambda_p grad lambda_q -
    lambda_q grad lambda_p, fem/hcurl.py:166-180, 215-291)."""
    d, nc = sd.dim, sd.num_cells
    opp = tabs["opp"]
    if d == 2:
        eid = tabs["faces"]
        en = sd.face_ridges.indices.reshape(sd.num_faces, 2).astype(np.int64)
    else:
        en = sd.ridge_peaks.indices.reshape(sd.num_ridges, 2).astype(np.int64)
        # edge id of local pair (p, q): the ridge of a face containing both nodes whose
        # end nodes (ridge_peaks) are that pair; independent of the storage order of
        # face_ridges.indices within a column.
        fr = sd.face_ridges.indices.reshape(sd.num_faces, 3).astype(np.int64)
        eid = np.empty((nc, 6), dtype=np.int64)
        for k, (p, q) in enumerate(_PAIRS3):
            a = min(set(range(4)) - {p, q})
            cand = fr[tabs["faces"][:, a]]  # (nc, 3) ridge ids of that face
            ends = np.sort(en[cand], axis=2)  # (nc, 3, 2)
            lo = np.minimum(opp[:, p], opp[:, q])[:, None]
            hi = np.maximum(opp[:, p], opp[:, q])[:, None]
            hit = (ends[:, :, 0] == lo) & (ends[:, :, 1] == hi)
            assert np.all(hit.sum(axis=1) == 1)
            eid[:, k] = cand[np.arange(nc), hit.argmax(axis=1)]
    ends = en[eid]  # (nc, E, 2) global nodes in edge_nodes order
    slots = (ends[:, :, :, None] == opp[:, None, None, :]).argmax(axis=3)  # (nc, E, 2)
    return eid, slots


def local_nedelec0_mass(sd, tabs, w, K):
    """A_ee' = w|c| [Mh_pr G_qs - Mh_ps G_qr - Mh_qr G_ps + Mh_qs G_pr],  G_ij = g_i^T K g_j."""
    d = sd.dim
    _, vol, G = local_geometry(sd, tabs)
    Kr = rotated_tensor(sd, K)
    GKG = np.einsum("cai,cij,cbj->cab", G, Kr, G)
    mhat = (np.ones((d + 1, d + 1)) + np.eye(d + 1)) / ((d + 1) * (d + 2))
    eid, sl = local_edges(sd, tabs)
    p, q = sl[:, :, 0], sl[:, :, 1]
    c = np.arange(sd.num_cells)[:, None, None]
    P, Q = p[:, :, None], q[:, :, None]
    Rr, S = p[:, None, :], q[:, None, :]
    A = (
        mhat[P, Rr] * GKG[c, Q, S]
        - mhat[P, S] * GKG[c, Q, Rr]
        - mhat[Q, Rr] * GKG[c, P, S]
        + mhat[Q, S] * GKG[c, P, Rr]
    )
    return A * (w * vol)[:, None, None], eid


def local_nedelec0_curlcurl(sd, tabs, w, K):
    """3D: C_loc^T A_RT0 C_loc with C_loc the (face slot, edge) signs of face_ridges;
    2D: s_a s_b w/|c| (cell_faces diag(w/|c|) cell_faces^T)."""
    d, nc = sd.dim, sd.num_cells
    if d == 2:
        _, vol, _ = local_geometry(sd, tabs)
        s = tabs["signs"]
        return s[:, :, None] * s[:, None, :] * (w / vol)[:, None, None], tabs["faces"]
    A_rt0, _ = local_rt0_mass(sd, tabs, w, K)
    eid, _ = local_edges(sd, tabs)
    fr_idx = sd.face_ridges.indices.reshape(sd.num_faces, 3).astype(np.int64)
    fr_sgn = np.asarray(sd.face_ridges.data, dtype=float).reshape(sd.num_faces, 3)
    C = np.zeros((nc, 4, 6))
    for a in range(4):
        f = tabs["faces"][:, a]
        for k in range(3):
            hit = fr_idx[f, k][:, None] == eid  # (nc, 6)
            C[:, a, :] += hit * fr_sgn[f, k][:, None]
    return np.einsum("cae,cab,cbf->cef", C, A_rt0, C), eid


def local_rt0_divdiv(sd, tabs, w, K):
    """RT0 stiffness B^T diag(w/|c|) B with B = cell_faces^T: s_a s_b w/|c| (hdiv.py:177-192,
    l2.py:335-353 through discretization.py:154-183)."""
    _, vol, _ = local_geometry(sd, tabs)
    s = tabs["signs"]
    return s[:, :, None] * s[:, None, :] * (w / vol)[:, None, None], tabs["faces"]


def local_bdm1_divdiv(sd, tabs, w, K):
    """BDM1 stiffness: div = div_RT0 @ hstack([I]*d)/d (hdiv.py:322-336, 353-371)."""
    d, nc, nf = sd.dim, sd.num_cells, sd.num_faces
    _, vol, _ = local_geometry(sd, tabs)
    s = np.repeat(tabs["signs"], d, axis=1)
    A = s[:, :, None] * s[:, None, :] * (w / (vol * d * d))[:, None, None]
    dofs = (tabs["faces"][:, :, None] + nf * np.arange(d)[None, None, :]).reshape(nc, d * (d + 1))
    return A, dofs


def local_lagrange1_lumped(sd, tabs, w, K):
    """Pi^T kron(eye/(d+1), diag(|c| w)) Pi (discretization.py:91-110, l2.py:472-483): diag w|c|/(d+1)."""
    d = sd.dim
    _, vol, _ = local_geometry(sd, tabs)
    return (w * vol)[:, None, None] * (np.eye(d + 1) / (d + 1))[None], tabs["opp"]


def local_rt0_lumped(sd, tabs, w, K):
    """RT0.assemble_lumped_matrix (hdiv.py:140-175): L_ff = sum_c (dist^T K dist)/|dist| / |f|,
    dist = face centre - cell centre, K = the stored 3x3 tensor (unrotated, as in the reference)."""
    d, nc = sd.dim, sd.num_cells
    x = sd.nodes.T  # (nn, 3) ambient coordinates, as the reference uses face_centers / cell_centers
    X = x[tabs["opp"]]  # (nc, d+1, 3)
    cc = X.mean(axis=1)
    A = np.zeros((nc, d + 1, d + 1))
    for a in range(d + 1):
        others = [t for t in range(d + 1) if t != a]
        P = X[:, others, :]
        dist = P.mean(axis=1) - cc
        if d == 3:
            area = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
        else:
            area = np.linalg.norm(P[:, 1] - P[:, 0], axis=1)
        q = np.einsum("ci,ijc,cj->c", dist, K, dist)
        nd = np.linalg.norm(dist, axis=1)
        A[:, a, a] = np.where(nd > 0, q / np.where(nd > 0, nd, 1.0), 0.0) / area
    return A, tabs["faces"]


def local_bdm1_lumped(sd, tabs, w, K):
    """BDM1 lumped mass: Pi^T (kron(I_d, lumped P1) W(K)) Pi: w|c| delta_jl/(d+1) t_aj^T K t_bl."""
    d = sd.dim
    A, dofs = local_bdm1_mass(sd, tabs, w, K)
    fnodes = tabs["fnodes"]
    js = (fnodes[:, :, :, None] == tabs["opp"][:, None, None, :]).argmax(axis=3).reshape(sd.num_cells, -1)
    same = js[:, :, None] == js[:, None, :]
    # undo Mhat (2 on the diagonal blocks j == l, 1 elsewhere), apply delta_jl/(d+1)
    return np.where(same, A / 2.0 * (d + 2), 0.0), dofs


def port_lumped(space: str, sd, data=None, keyword="unitary_data", tabs=None):
    """The reference's route: Pi.T @ M_lumped @ Pi (discretization.py:91-136); RT0 / Nedelec0 overrides."""
    tabs = tabs or cell_tables(sd)
    w, K = cell_data(sd, data, keyword)
    d = sd.dim
    if space == "rt0":
        A, dofs = local_rt0_lumped(sd, tabs, w, K)
        diag = np.zeros(sd.num_faces)
        np.add.at(diag, dofs.ravel(), np.einsum("caa->ca", A).ravel())
        return sps.diags_array(diag).tocsc()
    if space == "nedelec0":
        return sps.diags_array(np.asarray(port_mass(space, sd, data, keyword, tabs).sum(axis=0)).ravel()).tocsc()
    lump = sps.kron(np.eye(d + 1) / (d + 1), sps.diags_array(sd.cell_volumes * w)).tocsc()
    Pi = _PROJ[space](sd, tabs)
    if space == "lagrange1":
        M = lump
    else:
        R = np.asarray(sd.rotation_matrix, dtype=float)
        Kr = np.tensordot(R.T, np.tensordot(R, K, (1, 0)), (0, 1))
        tiled = np.tile(Kr, d + 1)
        W = sps.block_array([[sps.diags_array(tiled[i, j]) for j in range(d)] for i in range(d)]).tocsc()
        M = sps.kron(sps.eye_array(d), lump).tocsc() @ W
    return (Pi.T @ M @ Pi).tocsc()


# ---- Lagrange2 (the reference's per-cell Python loop, fem/h1.py:545-599) ----
def lagrange2_pairs(d):
    """Local edge -> (p, q) slot pairs, ``Lagrange2.get_local_edge_nodes`` (fem/h1.py:449-469)."""
    return [(p, q) for p in range(d + 1) for q in range(p + 1, d + 1)]


def lagrange2_dofs(sd, tabs):
    """Cell -> dof table (nc, (d+1) + d(d+1)/2): the nodes in slot order, then ``num_nodes`` + the
    edge the reference attaches to the local pair (p, q) of ``lagrange2_pairs`` order
    (``get_edge_dof_indices`` fem/h1.py:507-543): in 2D the face opposite the third slot
    (``faces[::-1]``); in 3D a purely index-based rule - the cell's six ridge ids sorted ascending
    and permuted by [5, 4, 2, 3, 1, 0] ("experimentally, we always find the following numbering").
    On grids whose ridges are numbered lexicographically by their end nodes and whose opposite
    nodes come in descending order (all structured grids) that is the ridge joining the two
    nodes; elsewhere (e.g. ``reference_element(3)``) it is not, and parity means following the rule."""
    d, opp = sd.dim, tabs["opp"]
    if d == 2:
        edges = tabs["faces"][:, ::-1]
    else:
        eid, _ = local_edges(sd, tabs)
        edges = np.sort(eid, axis=1)[:, [5, 4, 2, 3, 1, 0]]
    return np.hstack([opp, sd.num_nodes + edges]).astype(np.int64)


def lagrange2_local_mass(d):
    """(phi_i, phi_j) on a d-simplex of measure 1, phi = lambda_i(2 lambda_i - 1) | 4 lambda_p lambda_q
    (``Lagrange2.assemble_local_mass`` fem/h1.py:352-392), from the monomial integrals
    d! prod(alpha_k!) / (d + |alpha|)!  (fem/h1.py:418-436)."""
    from math import factorial

    pairs = lagrange2_pairs(d)
    n, ne = d + 1, len(pairs)
    # basis functions as {exponent tuple: coefficient}
    def mono(*idx):
        a = [0] * n
        for i in idx:
            a[i] += 1
        return tuple(a)

    basis = [{mono(i, i): 2.0, mono(i): -1.0} for i in range(n)] + [{mono(p, q): 4.0} for p, q in pairs]

    def integ(a):
        return factorial(d) * np.prod([factorial(k) for k in a]) / factorial(d + sum(a))

    M = np.zeros((n + ne, n + ne))
    for i, bi in enumerate(basis):
        for j, bj in enumerate(basis):
            M[i, j] = sum(ci * cj * integ(tuple(x + y for x, y in zip(ai, aj))) for ai, ci in bi.items() for aj, cj in bj.items())
    return M


def local_lagrange2_mass(sd, tabs, w, K):
    _, vol, _ = local_geometry(sd, tabs)
    return (w * vol)[:, None, None] * lagrange2_local_mass(sd.dim)[None], lagrange2_dofs(sd, tabs)


def local_lagrange2_stiff(sd, tabs, w, K):
    """|c| sum_{j,l} Mhat_jl grad phi_i(x_j)^T K grad phi_k(x_l) with the P1 mass Mhat of the nodal
    values of the (linear) gradients; grad phi_i(x_j) = (4 delta_ij - 1) g_i for a node,
    4 (delta_jp g_q + delta_jq g_p) for the edge (p, q) (fem/h1.py:471-505, 545-599).  The
    reference does not read pg.WEIGHT here."""
    d = sd.dim
    _, vol, G = local_geometry(sd, tabs)
    # entry (i, k) = Psi_i^T K Psi_k with K itself (h1.py:585-587), unlike the projection-based
    # spaces whose contraction produces K^T (rotated_tensor)
    R = np.asarray(getattr(sd, "rotation_matrix", np.eye(d, 3)), dtype=float)
    Kr = np.einsum("ia,abc,jb->cij", R, K, R)
    pairs = lagrange2_pairs(d)
    n, ne = d + 1, len(pairs)
    nc = sd.num_cells
    # nodal values of the gradients: Psi (nc, n+ne, n nodes, d)
    Psi = np.zeros((nc, n + ne, n, d))
    for i in range(n):
        for j in range(n):
            Psi[:, i, j, :] = (4.0 * (i == j) - 1.0) * G[:, i, :]
    for e, (p_, q_) in enumerate(pairs):
        Psi[:, n + e, p_, :] = 4.0 * G[:, q_, :]
        Psi[:, n + e, q_, :] = 4.0 * G[:, p_, :]
    mhat = (np.ones((n, n)) + np.eye(n)) / ((d + 1) * (d + 2))
    A = np.einsum("jl,cijx,cxy,ckly->cik", mhat, Psi, Kr, Psi) * vol[:, None, None]
    return A, lagrange2_dofs(sd, tabs)


# ---- RT1 (the reference's per-cell Python loop, fem/hdiv.py:555-760) --------
def rt1_point_coefficients(d):
    """Nodal values of the RT1 basis of one cell as combinations of the cell's node positions:
    ``c[p, j, n]`` = coefficient of ``x_n`` in the value of local basis function ``p`` at P2 point
    ``j`` (the d+1 nodes, then the edges in ``lagrange2_pairs`` order), before the factor
    ``s_p / (d |c|)``.  Local nodes are the cell's nodes in ascending id ("rank"), local dof
    ``q < d(d+1)`` sits at node ``i = q // d`` of the face opposite node ``j = (d(d+1)-1-q) % (d+1)``
    and the last d dofs are the cell functions of nodes 0..d-1 (``RT1.eval_basis_functions``,
    fem/hdiv.py:676-744).  Closed form of the reference's construction: with t = x_i - x_j the face
    function is t at node i, t/2 + (x_o - x_i)/4 on the edges {i, o}, o != j, -(x_o - x_j)/4 on the
    edges {j, o}, o != i, and 0 elsewhere (in particular on the edge {i, j}); the cell function of node
    m is (x_o - x_m)/4 on every edge {m, o}."""
    n = d + 1
    pairs = lagrange2_pairs(d)
    P, L = n + len(pairs), d * (d + 2)
    c = np.zeros((L, P, n))
    for q in range(d * n):
        i, j = q // d, (d * n - 1 - q) % n
        c[q, i, i] += 1.0
        c[q, i, j] -= 1.0
        for e, (a, b) in enumerate(pairs):
            if i in (a, b) and j not in (a, b):
                o = b if a == i else a
                c[q, n + e, i] += 0.5 - 0.25
                c[q, n + e, j] -= 0.5
                c[q, n + e, o] += 0.25
            elif j in (a, b) and i not in (a, b):
                o = b if a == j else a
                c[q, n + e, o] -= 0.25
                c[q, n + e, j] += 0.25
    for m in range(d):
        for e, (a, b) in enumerate(pairs):
            if m in (a, b):
                o = b if a == m else a
                c[d * n + m, n + e, o] += 0.25
                c[d * n + m, n + e, m] -= 0.25
    return c


def rt1_gram_table(d):
    """``C[p, q, n, n']`` (n, n' = 1..d): A_pq = s_p s_q / (d^2 |c|) * sum C[p,q,n,n'] Q[n,n'] with
    Q[n,n'] = (K y_n) . y_n', y_n = x_n - x_0.  The table the CUDA kernel is generated from
    (``oracle/gen_rt1_table.py``); the n = 0 terms drop out because every value is a combination of
    differences of node positions."""
    c = rt1_point_coefficients(d)
    assert np.abs(c.sum(axis=2)).max() < 1e-15
    M2 = lagrange2_local_mass(d)
    return np.einsum("jl,pjn,qlm->pqnm", M2, c, c)[:, :, 1:, 1:]


def rt1_rank_tables(sd, tabs):
    """Per cell, in rank order (ascending node id): node ids ``(nc, d+1)``, and for the face opposite
    each node its id and sign (``RT1.reorder_faces`` fem/hdiv.py:629-674 lists the faces by
    DESCENDING opposite node; here index r = the face opposite the node of rank r)."""
    order = np.argsort(tabs["opp"], axis=1, kind="stable")
    take = lambda a: np.take_along_axis(a, order, axis=1)  # noqa: E731
    return take(tabs["opp"]), take(tabs["faces"]), take(tabs["signs"])


def rt1_dofs(sd, tabs):
    """``RT1.local_dofs_of_cell`` (fem/hdiv.py:534-552): dof q < d(d+1) = face opposite rank
    d - q % (d+1), copy q // (d+1); then d*num_faces + k*num_cells + c."""
    d, nc, nf = sd.dim, sd.num_cells, sd.num_faces
    _, faces, _ = rt1_rank_tables(sd, tabs)
    q = np.arange(d * (d + 1))
    face_dofs = faces[:, d - q % (d + 1)] + (q // (d + 1)) * nf
    cell_dofs = d * nf + np.arange(d)[None, :] * nc + np.arange(nc)[:, None]
    return np.hstack([face_dofs, cell_dofs]).astype(np.int64)


def local_rt1_mass(sd, tabs, w, K):
    """|c| Psi (M2 x K) Psi^T with the P2 local mass M2 (fem/hdiv.py:555-621); entry (p, q) =
    sum_jl M2_jl Psi_pj^T K Psi_ql with K itself; pg.WEIGHT is not read."""
    d, nc = sd.dim, sd.num_cells
    R = np.asarray(getattr(sd, "rotation_matrix", np.eye(d, 3)), dtype=float)
    Kr = np.einsum("ia,abc,jb->cij", R, K, R)
    nodes, _, signs = rt1_rank_tables(sd, tabs)
    X = (R @ sd.nodes).T[nodes]  # (nc, d+1, d) in rank order
    E = X[:, 1:, :] - X[:, :1, :]
    vol = np.abs(np.linalg.det(E)) / (2.0 if d == 2 else 6.0)
    c = rt1_point_coefficients(d)
    Psi = np.einsum("pjn,cnx->cpjx", c, X)  # (nc, L, P, d)
    q = np.arange(d * (d + 1))
    s = np.ones((nc, d * (d + 2)))
    s[:, : d * (d + 1)] = signs[:, (d * (d + 1) - 1 - q) % (d + 1)]
    Psi *= (s / (d * vol)[:, None])[:, :, None, None]
    M2 = lagrange2_local_mass(d)
    A = np.einsum("jl,cpjx,cxy,cqly->cpq", M2, Psi, Kr, Psi) * vol[:, None, None]
    return A, rt1_dofs(sd, tabs)


_LOCAL = {
    ("rt1", "mass"): local_rt1_mass,
    ("lagrange2", "mass"): local_lagrange2_mass,
    ("lagrange2", "stiff"): local_lagrange2_stiff,
    ("lagrange1", "lumped"): local_lagrange1_lumped,
    ("rt0", "lumped"): local_rt0_lumped,
    ("bdm1", "lumped"): local_bdm1_lumped,
    ("rt0", "stiff"): local_rt0_divdiv,
    ("bdm1", "stiff"): local_bdm1_divdiv,
    ("lagrange1", "mass"): local_lagrange1_mass,
    ("lagrange1", "stiff"): local_lagrange1_stiff,
    ("rt0", "mass"): local_rt0_mass,
    ("bdm1", "mass"): local_bdm1_mass,
    ("nedelec0", "mass"): local_nedelec0_mass,
    ("nedelec0", "stiff"): local_nedelec0_curlcurl,
}

NDOF = {
    "lagrange1": lambda sd: sd.num_nodes,
    "rt0": lambda sd: sd.num_faces,
    "bdm1": lambda sd: sd.dim * sd.num_faces,
    "nedelec0": lambda sd: sd.num_edges,
    "lagrange2": lambda sd: sd.num_nodes + sd.num_edges,
    "rt1": lambda sd: sd.dim * (sd.num_faces + sd.num_cells),
}


def local_matrices(space, form, sd, data=None, keyword="unitary_data", tabs=None, **kw):
    tabs = tabs or cell_tables(sd)
    w, K = cell_data(sd, data, keyword)
    return _LOCAL[(space, form)](sd, tabs, w, K, **kw)


def closed_form(space, form, sd, data=None, keyword="unitary_data", tabs=None, fast=False, **kw) -> sps.csr_array:
    """Structural-pattern CSR (sorted, int32, explicit zeros kept) from local matrices.

    Duplicates are summed in COO order = ascending cell id (the order the CUDA
    segmented reduce uses), so values agree to the last few ulps.  ``fast=True`` lets scipy's
    COO -> CSR constructor do the duplicate summation (same pattern, explicit zeros kept, summation
    order unspecified: values agree to rounding) - for grids with 10^7 cells and more.
    """
    A, dofs = local_matrices(space, form, sd, data, keyword, tabs, **kw)
    n = NDOF[space](sd)
    L = dofs.shape[1]
    rows = np.broadcast_to(dofs[:, :, None], A.shape).ravel()
    cols = np.broadcast_to(dofs[:, None, :], A.shape).ravel()
    if fast:
        M = sps.coo_array((A.ravel(), (rows, cols)), shape=(n, n)).tocsr()
        M.sort_indices()
        return M
    return coo_to_csr(rows, cols, A.ravel(), n)


def coo_to_csr(rows, cols, vals, n) -> sps.csr_array:
    """Stable sort by (row, col) + sequential segmented sum; keeps explicit zeros."""
    key = rows.astype(np.int64) * n + cols.astype(np.int64)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    head = np.ones(ks.size, dtype=bool)
    head[1:] = ks[1:] != ks[:-1]
    seg = np.flatnonzero(head)
    data = np.add.reduceat(vals[order], seg) if ks.size else np.zeros(0)
    ukey = ks[seg]
    r, c = ukey // n, ukey % n
    indptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=n))]).astype(np.int64)
    it = np.int32 if ukey.size < 2**31 else np.int64
    return sps.csr_array((data, c.astype(np.int32), indptr.astype(it)), shape=(n, n))


def canonicalise(ref, pattern: sps.csr_array) -> sps.csr_array:
    """Put a reference matrix on the structural pattern: ``ref + 0*S``.

    Asserts pattern(ref) is a subset of the structural pattern.
    """
    ref = sps.csr_array(ref)
    ref.sum_duplicates()
    S = pattern.copy()
    S.data = np.ones_like(S.data)
    R1 = ref.copy()
    R1.data = np.ones_like(R1.data)
    outside = (R1 - R1.multiply(S)).count_nonzero() if ref.nnz else 0
    if outside:
        raise AssertionError(f"{outside} reference entries fall outside the structural pattern")
    n = pattern.shape[0]
    key_s = np.repeat(np.arange(n, dtype=np.int64), np.diff(pattern.indptr)) * n + pattern.indices
    ref.sort_indices()
    key_r = np.repeat(np.arange(n, dtype=np.int64), np.diff(ref.indptr)) * n + ref.indices
    pos = np.searchsorted(key_s, key_r)
    data = np.zeros(pattern.nnz)
    data[pos] = ref.data
    return sps.csr_array((data, pattern.indices.copy(), pattern.indptr.copy()), shape=pattern.shape)
