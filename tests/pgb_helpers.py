"""Helpers shared by the test modules."""

import numpy as np


def random_data(nc, seed=0, sym=True):
    from pygeon_b200 import SecondOrderTensor, PARAMETERS, SECOND_ORDER_TENSOR, WEIGHT

    rng = np.random.default_rng(seed)
    A = rng.normal(size=(3, 3, nc))
    K = np.einsum("iac,jac->ijc", A, A) + np.eye(3)[:, :, None]
    if not sym:
        K = K + 0.3 * rng.normal(size=(3, 3, nc))
    sot = SecondOrderTensor(np.ones(nc))
    sot.values = K
    return {PARAMETERS: {"key": {SECOND_ORDER_TENSOR: sot, WEIGHT: rng.uniform(0.5, 2.0, nc)}}}


def delaunay_grid(dim, npts, seed=0, min_rel_vol=1e-4):
    """An unstructured simplicial grid: Delaunay triangulation (scipy / Qhull) of random points in the
    unit square / cube plus its corners, slivers dropped.  Faces are numbered by their sorted node
    tuples, ``face_nodes`` lists the nodes of a face in ascending order and ``cell_faces`` carries
    +1 where the normal implied by that node order points out of the cell (the convention
    ``SimplexGrid.compute_geometry`` checks).  Node valences vary widely (unlike the Kuhn meshes)."""
    import itertools

    import scipy.sparse as sps
    from scipy.spatial import Delaunay

    from pygeon_b200.grid import SimplexGrid

    rng = np.random.default_rng(seed)
    corners = np.array(list(itertools.product((0.0, 1.0), repeat=dim)))
    pts = np.vstack([corners, rng.random((npts, dim))])
    cells = np.sort(Delaunay(pts).simplices.astype(np.int64), axis=1)
    X = pts[cells]  # (nc, d+1, d)
    vol = np.abs(np.linalg.det(X[:, 1:, :] - X[:, :1, :])) / (2.0 if dim == 2 else 6.0)
    cells = cells[vol > min_rel_vol * vol.mean()]
    used = np.unique(cells)
    renum = -np.ones(pts.shape[0], dtype=np.int64)
    renum[used] = np.arange(used.size)
    cells, pts = renum[cells], pts[used]
    nc, nn = cells.shape[0], pts.shape[0]
    # faces: drop one node of the cell at a time
    loc = np.array([[j for j in range(dim + 1) if j != a] for a in range(dim + 1)])  # (d+1, d)
    fnodes = cells[:, loc].reshape(nc * (dim + 1), dim)  # sorted because cells are
    ufaces, inv = np.unique(fnodes, axis=0, return_inverse=True)
    inv = inv.reshape(nc, dim + 1)
    nf = ufaces.shape[0]
    x = np.zeros((3, nn))
    x[:dim] = pts.T
    if dim == 2:
        t = x[:, ufaces[:, 1]] - x[:, ufaces[:, 0]]
        normals = np.vstack([t[1], -t[0], np.zeros(nf)])
    else:
        normals = 0.5 * np.cross(x[:, ufaces[:, 1]] - x[:, ufaces[:, 0]], x[:, ufaces[:, 2]] - x[:, ufaces[:, 0]], axis=0)
    fc = x[:, ufaces].mean(axis=2)
    cc = x[:, cells].mean(axis=2)
    sign = np.sign(np.einsum("ifc,ifc->fc", normals[:, inv.T], fc[:, inv.T] - cc[:, None, :]))  # (d+1, nc)
    face_nodes = sps.csc_array((np.ones(nf * dim, dtype=bool), ufaces.ravel(), np.arange(0, nf * dim + 1, dim)),
                               shape=(nn, nf))
    cell_faces = sps.csc_array((sign.T.ravel(), (inv.ravel(), np.repeat(np.arange(nc), dim + 1))), shape=(nf, nc))
    cell_faces.sort_indices()
    return SimplexGrid(dim, x, face_nodes, cell_faces, f"delaunay{dim}d")
