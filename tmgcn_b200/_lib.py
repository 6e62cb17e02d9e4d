"""ctypes binding of libtmgcn_b200.so (the C ABI in include/tmgcn.h).

The product path has no CPU fallback: if the shared library cannot be loaded
(and cannot be built because nvcc is missing) every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtmgcn_b200.so")

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_z = C.c_size_t

# name -> (restype, argtypes); must list every symbol include/tmgcn.h declares
PROTOTYPES = {
    "tmgcn_last_error": (C.c_char_p, []),
    "tmgcn_abi_version": (_i, []),
    "tmgcn_launch_count": (_l, []),
    "tmgcn_scan_ws_bytes": (_z, [_l]),
    "tmgcn_exclusive_scan_i64": (_i, [_p, _p, _l, _p, _p]),
    "tmgcn_rowptr_from_sorted_rows": (_i, [_p, _l, _l, _p, _p]),
    "tmgcn_mtransform_sparse_plan": (_i, [_p, _p, _i, _i, _l, _p, _i, _p, _p]),
    "tmgcn_mtransform_sparse_run": (_i, [_p, _p, _p, _i, _i, _l, _p, _i, _p, _p, _p, _i, _p]),
    "tmgcn_mtransform_sparse_ws_bytes": (_z, [_i, _i, _l, _i, _l]),
    "tmgcn_mtransform_sparse_plan_ws": (_i, [_p, _p, _i, _i, _l, _p, _i, _p, _p, _z, _p]),
    "tmgcn_mtransform_sparse_run_ws": (_i, [_p, _p, _p, _i, _i, _l, _p, _i, _p, _p, _p, _i, _p, _z, _p]),
    "tmgcn_csr_transpose_ws_bytes": (_z, [_l, _l, _i]),
    "tmgcn_csr_transpose_plan": (_i, [_p, _p, _i, _l, _p, _p]),
    "tmgcn_csr_transpose_run": (_i, [_p, _p, _p, _i, _l, _p, _p, _p, _i, _p, _p]),
    "tmgcn_csr_axpby_plan": (_i, [_p, _p, _p, _p, _l, _p, _p]),
    "tmgcn_csr_axpby_run": (_i, [_p, _p, _p, _p, _p, _p, C.c_double, C.c_double, _l, _p, _p, _p, _i, _p]),
    "tmgcn_csr_row_sums": (_i, [_p, _p, _l, _p, _i, _p]),
    "tmgcn_csr_scale_sym": (_i, [_p, _p, _p, _i, _l, _p, _i, _p]),
    "tmgcn_mtransform_dense_fwd": (_i, [_p, _p, _i, _i, _l, _p, _i, _p]),
    "tmgcn_mtransform_dense_bwd": (_i, [_p, _p, _i, _i, _l, _p, _i, _p]),
    "tmgcn_mtransform_dense_solve_fwd": (_i, [_p, _p, _i, _l, _p, _i, _p]),
    "tmgcn_mtransform_dense_solve_bwd": (_i, [_p, _p, _i, _l, _p, _i, _p]),
    "tmgcn_mtransform_dense_solve_part": (_i, [_p, _p, _p, _i, _i, _l, _l, _l, _p, _i, _i, _p]),
    "tmgcn_mtransform_dense_fwd_split": (_i, [_p, _p, _p, _i, _i, _l, _p, _i, _i, _p]),
    "tmgcn_mtransform_dense_bwd_range": (_i, [_p, _p, _i, _i, _l, _p, _i, _i, _i, _i, _p]),
    "tmgcn_spmm_fwd": (_i, [_p, _p, _p, _p, _p, _i, _l, _i, _i, _p]),
    "tmgcn_gemm_xw_fwd": (_i, [_p, _p, _p, _l, _i, _i, _i, _p]),
    "tmgcn_gemm_xw_bias_fwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p]),
    "tmgcn_gemm_dw_ws_bytes": (_z, [_i, _i]),
    "tmgcn_gemm_dw_dx_bwd": (_i, [_p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _p, _p]),
    "tmgcn_gemm_xw_sliced_fwd": (_i, [_p, _p, _p, _i, _l, _i, _i, _i, _p]),
    "tmgcn_gemm_sliced_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _l, _i, _i, _i, _p]),
    "tmgcn_flat_edge_ids": (_i, [_p, _l, _l, _l, _p, _p, _p]),
    "tmgcn_edge_gather_fwd": (_i, [_p, _p, _p, _p, _l, _i, _p]),
    "tmgcn_edge_readout_fwd": (_i, [_p, _p, _p, _p, _p, _l, _i, _i, _p]),
    "tmgcn_edge_gather_bwd": (_i, [_p, _p, _p, _p, _l, _i, _p]),
    "tmgcn_edge_readout_bwd_ws_bytes": (_z, [_l, _i, _i]),
    "tmgcn_edge_factor_ws_bytes": (_z, [_i, _i]),
    "tmgcn_edge_class_sums": (_i, [_p, _p, _p, _p, _l, _i, _p]),
    "tmgcn_edge_factor_apply": (_i, [_p, _p, _p, _p, _p, _l, _i, _i, _p, _p]),
    "tmgcn_edge_readout_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _p, _p]),
    "tmgcn_act_fwd": (_i, [_p, _p, _l, _i, _p]),
    "tmgcn_act_bwd": (_i, [_p, _p, _p, _l, _i, _p]),
}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the library (building it first when it is absent and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing (run `python -m tmgcn_b200.build`)")
        from .build import build
        build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.tmgcn_abi_version() != 2:
        raise RuntimeError("libtmgcn_b200.so: ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().tmgcn_last_error().decode()


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("tmgcn: " + last_error())


def launch_count() -> int:
    return int(load().tmgcn_launch_count())
