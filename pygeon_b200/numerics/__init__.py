"""Operator layer above the discretizations (reference ``src/pygeon/numerics``): exterior
derivatives, mass / lumped-mass / stiffness dispatchers over the subdomains of a (mixed-dimensional)
grid, tip-dof restrictions, cell-centre evaluation and the linear-system front end."""

from . import differentials, innerproducts, projections, restrictions, stiffness  # noqa: F401
from .. import linear_system  # noqa: F401
