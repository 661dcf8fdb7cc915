"""B200-native FEM assembly-and-solve hot path behind PyGeoN's operator API.

``import pygeon_b200 as pg`` gives the reference's names for the path this package covers:
``pg.Lagrange1, pg.Lagrange2, pg.RT0, pg.BDM1, pg.RT1, pg.Nedelec0, pg.PwConstants, pg.LinearSystem`` and the parameter
constants, plus the GPU ``solver`` callables ``pg.cg_solver`` / ``pg.minres_solver``.

Importing the package never touches CUDA; the first assemble/solve call loads ``libpgb.so``
(built by ``python -m pygeon_b200._build``) and fails loudly if it or the GPU is missing.
"""

from .constants import *  # noqa: F401,F403
from . import bmat, numerics  # noqa: F401
from .grid import SimplexGrid, reference_element, structured_grid, unit_grid  # noqa: F401
from .md_grid import MixedDimensionalGrid, MortarGrid, as_mdg, fractured_unit_cube  # noqa: F401
from .numerics.differentials import curl, div, grad  # noqa: F401
from .numerics.innerproducts import (cell_mass, face_mass, lumped_cell_mass, lumped_face_mass,  # noqa: F401
                                     lumped_peak_mass, lumped_ridge_mass, peak_mass, ridge_mass)
from .numerics.projections import eval_at_cell_centers  # noqa: F401
from .numerics.restrictions import remove_tip_dofs  # noqa: F401
from .numerics.stiffness import cell_stiff, face_stiff, peak_stiff, ridge_stiff  # noqa: F401
from .params import SecondOrderTensor, get_cell_data, initialize_data  # noqa: F401

Grid = SimplexGrid  # the grid class of this package under the reference's name


def __getattr__(name):
    # lazy: these pull in libpgb.so
    if name in ("Discretization", "Lagrange1", "Lagrange2", "RT0", "BDM1", "RT1", "Nedelec0", "PwConstants"):
        from . import discretization as _d

        return getattr(_d, name)
    if name in ("PwPolynomials", "PwLinears", "VecPwPolynomials", "VecPwConstants", "VecPwLinears", "get_PwPolynomials",
                "proj_to_PwPolynomials"):
        from . import broken as _b

        return getattr(_b, name)
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
    if name in ("DeviceCSR", "device_bmat"):
        from . import engine as _e

        return getattr(_e, name)
    if name in ("face_mass_device", "cell_div_device", "mdg_darcy_saddle_device"):
        from . import md_system as _m

        return getattr(_m, "darcy_saddle_device" if name == "mdg_darcy_saddle_device" else name)
    raise AttributeError(name)
