"""Import the UNMODIFIED reference (compgeo-mox/pygeon) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has no
/root/reference); callers must check :func:`available` first.  porepy and
matplotlib are replaced by the stand-ins under ``oracle/_standins``.
"""

from __future__ import annotations

import os
import sys

REF_SRC = "/root/reference/src"
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_standins")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "pygeon"))


def load():
    """Return ``(pg, pp)``: the reference package and the porepy stand-in."""
    if not available():
        raise RuntimeError("reference source not present (expected in the build container only)")
    for p in (REF_SRC, _STANDINS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import porepy as pp  # the stand-in
    import pygeon as pg  # the reference

    return pg, pp


def ref_grid(sd):
    """Reference ``pg.Grid`` with the same nodes / face_nodes / cell_faces as ``sd``."""
    pg, _ = load()
    g = pg.Grid(sd.dim, sd.nodes.copy(), sd.face_nodes.copy(), sd.cell_faces.copy(), "ref")
    g.compute_geometry()
    return g
