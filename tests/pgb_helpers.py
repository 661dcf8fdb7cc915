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
