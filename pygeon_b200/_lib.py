"""ctypes binding of libpgb.so (the C ABI in include/pgb.h).

There is no CPU fallback: importing this module without the built library raises.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgb.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m pygeon_b200._build` "
        "(nvcc, sm_100a).  pygeon_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_p, _i, _l, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double

_SIGS = {
    "pgb_version": (_i, []),
    "pgb_last_error": (C.c_char_p, []),
    "pgb_csr_last_row_kernel": (C.c_char_p, []),
    "pgb_pack_l2_dofs": (_i, [_i, _l, _l, _p, _p, _p, _p]),
    "pgb_set_l2_fetch_granularity": (_i, [_i]),
    "pgb_get_l2_fetch_granularity": (_i, []),
    "pgb_pack_cells": (_i, [_i, _l, _p, _p, _p, _p, _p, _p, _p]),
    "pgb_pack_signs": (_i, [_i, _l, _p, _p, _p]),
    "pgb_pack_bdm1": (_i, [_i, _l, _l, _p, _p, _p, _p, _p, _p]),
    "pgb_pack_edges3d": (_i, [_l, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pgb_pack_edges2d": (_i, [_l, _p, _p, _p, _p, _p]),
    "pgb_pack_rt1": (_i, [_i, _l, _l, _p, _p, _p, _p, _p, _p]),
    "pgb_prep_coords": (_i, [_i, _l, _p, _p, _p, _p]),
    "pgb_ridges3d_workspace_bytes": (_l, [_l]),
    "pgb_ridges3d": (_i, [_l, _l, _p, _p, _p, _p, C.POINTER(C.c_int64), _p, _l, _p]),
    "pgb_grid_geometry": (_i, [_i, _l, _l, _l, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pgb_edge_geometry": (_i, [_l, _l, _p, _p, _p, _p, _p, _p]),
    "pgb_local_size": (_i, [_i, _i]),
    "pgb_local_matrices": (_i, [_i, _i, _l, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pgb_csr_workspace_bytes": (_l, [_l, _i, _l, _i]),
    "pgb_set_profile": (None, [_i]),
    "pgb_get_phase_ms": (None, [_p, _i]),
    "pgb_csr_symbolic_sort": (_i, [_l, _l, _i, _p, _i, _p, _l, _p, _p, _p, C.POINTER(C.c_int64), C.POINTER(_i), _p]),
    "pgb_csr_symbolic_result": (_i, [C.POINTER(C.c_int64), C.POINTER(_i)]),
    "pgb_csr_symbolic_finish": (_i, [_l, _l, _i, _i, _l, _p, _p, _p, _i, _p, _p, _p]),
    "pgb_csr_numeric_rows": (_i, [_l, _i, _i, _p, _p, _p, _i, _p, _p, _i, _p]),
    "pgb_csr_numeric": (_i, [_l, _p, _p, _p, _p, _p]),
    "pgb_system_workspace_bytes": (_l, [_l]),
    "pgb_mask_to_map": (_i, [_l, _p, _p, _p, C.POINTER(C.c_int64), _p, _l, _p]),
    "pgb_csr_extract_symbolic": (_i, [_l, _p, _p, _i, _p, _p, _p, C.POINTER(C.c_int64), _p, _l, _p]),
    "pgb_csr_extract_fill": (_i, [_l, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "pgb_csr_mark_columns": (_i, [_l, _p, _p, _i, _p, _p, _p]),
    "pgb_csr_rows_touching": (_i, [_l, _p, _i, _p, _i, _p, _p]),
    "pgb_scan_i64": (_i, [_l, _p, _p, C.POINTER(C.c_int64), _p, _l, _p]),
    "pgb_csr_block_lengths": (_i, [_l, _p, _i, _p, _p]),
    "pgb_csr_block_place": (_i, [_l, _p, _i, _p, _p, _i, _d, _p, _p, _p, _p]),
    "pgb_csr_add_diagonal": (_i, [_l, _p, _p, _p, _i, _p, _p, _p, _p]),
    "pgb_csr_transpose_workspace_bytes": (_l, [_l]),
    "pgb_csr_transpose": (_i, [_l, _l, _l, _p, _i, _p, _p, _p, _i, _p, _p, _p, _l, _p]),
    "pgb_csr_diagonal": (_i, [_l, _p, _i, _p, _p, _p, _p]),
    "pgb_darcy_block_jacobi": (_i, [_l, _l, _p, _i, _p, _p, _p, _p, _p]),
    "pgb_gather": (_i, [_l, _p, _p, _p, _d, _p, _p]),
    "pgb_scatter_add": (_i, [_l, _p, _p, _p, _p]),
    "pgb_spmv": (_i, [_l, _p, _i, _p, _p, _p, _p, _i, _p]),
    "pgb_cg": (_i, [_l, _p, _i, _p, _p, _p, _p, _p, _d, _d, _i, _p, C.POINTER(_i), C.POINTER(_i), _p, _p]),
    "pgb_minres": (_i, [_l, _p, _i, _p, _p, _p, _p, _p, _d, _d, _i, _p, C.POINTER(_i), C.POINTER(_i), _p, _p]),
    "pgb_comm_create": (_i, [_i, _i, _l, C.POINTER(C.c_void_p)]),
    "pgb_comm_handle_bytes": (_i, []),
    "pgb_comm_get_handle": (_i, [_p, _p]),
    "pgb_comm_connect": (_i, [_p, _p]),
    "pgb_comm_arena": (C.c_void_p, [_p]),
    "pgb_comm_arena_bytes": (_l, [_p]),
    "pgb_comm_barrier": (_i, [_p, _p]),
    "pgb_comm_destroy": (_i, [_p]),
    "pgb_dist_cg": (_i, [_p, _l, _l, _l, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _l, _l, _i,
                         _d, _d, _i, C.POINTER(_i), C.POINTER(_i), _p, _p]),
    "pgb_dist_minres": (_i, [_p, _l, _l, _l, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _l, _l, _i,
                             _d, _d, _i, C.POINTER(_i), C.POINTER(_i), _p, _p]),
    "pgb_cg_scratch_doubles": (_l, [_i]),
    "pgb_cg_step_init": (_i, [_l, _p, _p, _p, _p, _d, _d, _p, _p]),
    "pgb_cg_step_reduce": (_i, [_p, _i, _i, _p]),
    "pgb_cg_step_p": (_i, [_l, _i, _p, _p, _p, _p, _p]),
    "pgb_cg_step_spmv": (_i, [_l, _p, _i, _p, _p, _p, _p, _i, _p, _p]),
    "pgb_cg_step_xr": (_i, [_l, _p, _p, _p, _p, _p, _p, _p]),
    "pgb_mr_scratch_doubles": (_l, [_i]),
    "pgb_mr_step_reduce": (_i, [_p, _i, _i, _p]),
    "pgb_mr_step_init": (_i, [_l, _p, _p, _p, _p, _p]),
    "pgb_mr_step_start": (_i, [_l, _p, _p, _p, _p, _d, _d, _i, _p, _p]),
    "pgb_mr_step_spmv": (_i, [_l, _p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _p]),
    "pgb_mr_step_c": (_i, [_l, _p, _p, _p, _p]),
    "pgb_mr_step_scalars": (_i, [_i, _i, _p, _p]),
    "pgb_mr_step_d": (_i, [_l, _p, _p, _p, _p, _p, _p, _p]),
    "pgb_mr_state": (_i, [_p, C.POINTER(_i), C.POINTER(_i), _p]),
    "pgb_dot_partials": (_i, [_l, _p, _p, _p, _p]),
    "pgb_reduce_partials": (_i, [_p, _i, _p, _p]),
    "pgb_axpby": (_i, [_l, _d, _p, _d, _p, _p]),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)  # AttributeError here = header and library disagree
    _f.restype = _res
    _f.argtypes = _args

EXPORTS = tuple(_SIGS)

KIND = {
    "l1_mass": 0,
    "l1_stiff": 1,
    "l1_stiff_rot": 2,
    "rt0_mass": 3,
    "bdm1_mass": 4,
    "n0_mass": 5,
    "n0_curlcurl": 6,
    "p0_mass": 7,
    "rt0_divdiv": 8,
    "bdm1_divdiv": 9,
    "darcy_saddle": 10,
    "l1_lumped": 11,
    "rt0_lumped": 12,
    "bdm1_lumped": 13,
    "l2_mass": 14,
    "l2_stiff": 15,
    "rt1_mass": 16,
    "rt1_divdiv": 17,
    "rt1_lumped_cell": 18,
}
NPART = 1184


class PgbError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise PgbError(lib.pgb_last_error().decode())


def launches() -> int:
    return _launch_count[0]


_launch_count = [0]


def count(n: int) -> None:
    _launch_count[0] += n
