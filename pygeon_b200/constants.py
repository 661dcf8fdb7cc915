"""Names shared with the reference (``utils/common_constants.py:5-47``; ``pp.PARAMETERS``)."""

AMBIENT_DIM = 3
UNITARY_DATA = "unitary_data"
SCALAR, VECTOR, MATRIX = 0, 1, 2
SECOND_ORDER_TENSOR = "second_order_tensor"
VECTOR_FIELD = "vector_field"
WEIGHT = "weight"
NORMAL_DIFFUSIVITY = "normal_diffusivity"
PARAMETERS = "parameters"  # porepy's pp.PARAMETERS
DISCRETIZATION_MATRICES = "discretization_matrices"
