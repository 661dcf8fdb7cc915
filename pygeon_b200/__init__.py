"""B200-native FEM assembly-and-solve hot path behind PyGeoN's operator API.

``import pygeon_b200 as pg`` gives the reference's names for the path this package covers:
``pg.Lagrange1, pg.Lagrange2, pg.RT0, pg.BDM1, pg.RT1, pg.Nedelec0, pg.PwConstants, pg.LinearSystem`` and the parameter
constants, plus the GPU ``solver`` callables ``pg.cg_solver`` / ``pg.minres_solver``.

Importing the package never touches CUDA; the first assemble/solve call loads ``libpgb.so``
(built by ``python -m pygeon_b200._build``) and fails loudly if it or the GPU is missing.
"""

from .constants import *  # noqa: F401,F403
from .grid import SimplexGrid, structured_grid
from .params import SecondOrderTensor, get_cell_data, initialize_data


def __getattr__(name):
    # lazy: these pull in libpgb.so
    if name in ("Discretization", "Lagrange1", "Lagrange2", "RT0", "BDM1", "RT1", "Nedelec0", "PwConstants"):
        from . import discretization as _d

        return getattr(_d, name)
    if name in ("LinearSystem", "create_restriction"):
        from . import linear_system as _l

        return getattr(_l, name)
    if name in ("cg", "minres", "cg_solver", "minres_solver", "cg_device", "minres_device"):
        from . import solvers as _s

        return getattr(_s, name)
    if name in ("darcy_saddle_device", "darcy_saddle_matrix", "darcy_block_jacobi"):
        from . import darcy as _dz

        return getattr(_dz, name)
    if name in ("DeviceLinearSystem",):
        from . import device_system as _ds

        return getattr(_ds, name)
    if name in ("DeviceCSR",):
        from . import engine as _e

        return getattr(_e, name)
    raise AttributeError(name)
