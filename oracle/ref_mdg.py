"""The UNMODIFIED reference's mixed-dimensional objects built on the same raw arrays as a
pygeon_b200.MixedDimensionalGrid.  TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).

The reference's pg.MixedDimensionalGrid / pg.MortarGrid subclass PorePy's containers, for which
oracle/_standins/porepy holds minimal restatements (conforming mortar grids, container bookkeeping); all the
arithmetic under test - signed mortar maps, the jump operators cell_faces / face_ridges / ridge_peaks
(grids/mortar_grid.py:46-208), tip / leaf tags (grids/md_grid.py:113-153), block operators
(numerics/differentials.py:144-192, innerproducts.py:162-232) - is the reference's own code."""

from __future__ import annotations

import numpy as np

from oracle import ref_loader


def ref_mdg(mdg):
    """Reference MDG with the same subdomains / interface as ``mdg``; returns (ref_mdg, {id(sd): ref_sd})."""
    rpg, pp = ref_loader.load()
    out = rpg.MixedDimensionalGrid()
    table = {}
    for sd in mdg.subdomains():
        g = rpg.Grid(sd.dim, sd.nodes.copy(), sd.face_nodes.copy(), sd.cell_faces.copy(), sd.name)
        for tag in ("fracture_faces", "tip_faces", "domain_boundary_faces"):
            if tag in sd.tags:
                g.tags[tag] = np.asarray(sd.tags[tag]).copy()
        table[id(sd)] = g
    out.add_subdomains(list(table.values()))
    for intf in mdg.interfaces():
        up, down = mdg.interface_to_subdomain_pair(intf)
        gu, gd = table[id(up)], table[id(down)]
        gd.compute_geometry()  # the stand-in's mortar grid takes the secondary cell measures
        m = rpg.MortarGrid(intf.dim, intf.face_up, intf.cell_down, gu.num_faces, gd.num_cells, gd.cell_volumes)
        out.add_interface(m, (gu, gd))
    out.compute_geometry()
    return out, table
