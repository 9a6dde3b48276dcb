"""ctypes binding of the C ABI declared in ``include/peppa_b200.h``.

The shared library is built in-tree (``peppa_b200/csrc/libpeppa_b200.so``) by
``peppa_b200.build``.  There is no fallback: if the library is missing or a symbol is
absent, importing/using the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpeppa_b200.so")
MEASURE_LIB_PATH = os.path.join(_HERE, "csrc", "libpeppa_b200_measure.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "peppa_b200.h")

PB2_BF16, PB2_F16, PB2_F32, PB2_U8, PB2_I8_PLANES = 0, 1, 2, 3, 4

_p = C.c_void_p
_i64 = C.c_int64
_i = C.c_int
_f = C.c_float

# name -> argument ctypes (all return int unless listed in _RESTYPE)
SIGNATURES = {
    "pb2_version": [],
    "pb2_last_error": [],
    "pb2_launch_count": [],
    "pb2_sim_grid": [],
    "pb2_triplet_score": [_p, _p, _p, _p, _p, _p, _i64, _i, _i64, _i, _i, _p, _p],
    "pb2_row_norms": [_p, _i, _i64, _i, _i64, _p, _p, _p],
    "pb2_split_f16": [_p, _p, _i64, _i, _i64, _i, _p, _i64, _p, _p],
    "pb2_pair_dot": [_p, _p, _i, _p, _p, _p, _p, _i64, _i, _i64, _i64, _p, _p, _p, _p],
    "pb2_sim_diag": [_p, _p, _p, _p, _i64, _i, _i, _i64, _i64, _p, _p, _p, _p],
    "pb2_sim_matrix": [_p, _p, _p, _p, _i64, _i64, _i, _i, _i64, _i64, _f, _p, _i64, _p],
    "pb2_sim_rank": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _i, _i64, _i64, _p, _p],
    "pb2_subset_rank": [_p, _i64, _p, _i, _i, _p, _p],
    "pb2_sim_hinge": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i, _i, _i64, _i64, _f, _p, _i, _p, _p, _p, _i, _i64, _p, _p, _p],
    "pb2_rows_quant_i8": [_p, _i, _p, _i64, _i, _i64, _p, _i64, _p],
    "pb2_sim_lse_parts": [_i64],
    "pb2_sim_lse_rows": [_p, _p, _p, _p, _i64, _i64, _i, _i, _i64, _i64, _f, _p, _p, _p],
    "pb2_lse_merge": [_p, _p, _i, _i64, _p, _i, _p],
    "pb2_sim_lse_col_parts": [_i64],
    "pb2_sim_lse_both": [_p, _p, _p, _p, _i64, _i64, _i, _i, _i64, _i64, _f, _f, _p, _p, _p],
    "pb2_sim_lse_both_rank": [_p, _p, _p, _p, _i64, _i64, _i, _i, _i64, _i64, _f, _f, _p, _p, _p, _p, _p, _i64, _i64, _p, _p],
    "pb2_lse_merge_const": [_p, _i, _i64, _f, _p, _i, _p],
    "pb2_lse_combine": [_p, _i, _i64, _p, _p],
    "pb2_sim_lse_grad": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i, _i, _i64, _i64, _f, _p, _i64, _p],
    "pb2_grad_gemm": [_p, _i, _i64, _i64, _i64, _i, _p, _i, _i, _i64, _f, _i, _p, _i64, _p],
    "pb2_grad_gemm_workspace": [],
    "pb2_grad_gemm_ws": [_p, _i, _i64, _i64, _i64, _i, _p, _i, _i, _i64, _f, _i, _p, _i64, _p, _i64, _p],
    "pb2_grad_gemm_dual": [_p, _i, _i64, _i64, _i64, _p, _p, _i, _i, _i64, _i64, _f, _p, _p, _i64, _i64, _p],
    "pb2_hinge_finish": [_p, _i64, _p, _p, _i, _p, _p, _p, _p, _i64, _i, _i64, _i64, _f, _p, _p, _i64, _p],
    "pb2_hinge_prep": [_p, _p, _i, _i64, _i, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p],
    "pb2_hinge_finish2": [_p, _p, _p, _p, _i, _i64, _i, _i64, _i64, _p, _p, _p, _p, _p, _p, _i, _f, _f, _p, _p, _p, _i, _p],
    "pb2_hinge_step_workspace": [_i64, _i, _i],
    "pb2_hinge_step": [_p, _p, _i, _i64, _i, _i64, _i64, _f, _p, _i64, _p, _p, _p, _i, _p, _p, _p],
    "pb2_hinge_forward_workspace": [_i64, _i, _i],
    "pb2_hinge_state_bytes": [_i64, _i],
    "pb2_hinge_forward": [_p, _p, _i, _i64, _i, _i64, _i64, _f, _p, _i64, _p, _i64, _p, _p, _p, _p],
    "pb2_hinge_backward": [_p, _i64, _p, _p, _i, _i64, _i, _i64, _i64, _p, _p, _p, _i, _p],
    "pb2_rows_scale_f16": [_p, _i, _p, _i64, _i, _i64, _p, _i64, _p],
    "pb2_scale_pair": [_p, _p, _i64, _i, _p, _p, _p, _p],
    "pb2_milnce_finish": [_p, _i64, _p, _i, _i64, _i, _i64, _f, _p, _p, _i64, _p],
    "pb2_milnce_finish_k": [_p, _i64, _p, _i, _p, _i64, _i, _i, _i, _i64, _f, _p, _p, _i64, _p],
    "pb2_project_normalize": [_p, _p, _p, _i64, _i, _i, _i64, _i64, _f, _p, _i64, _p, _p, _p],
    "pb2_sum_partials": [_p, _i, _f, _p, _p],
    "pb2_hinge_loss_terms": [_p, _i, _p, _p, _i64, _f, _f, _p, _i, _p],
    "pb2_milnce_loss": [_p, _p, _p, _i64, _p, _p, _p],
    "pb2_ipc_export": [_p, _p, _p],
    "pb2_ipc_open": [_p, _i64, _p, _p],
    "pb2_ipc_close": [_p],
    "pb2_peer_reduce": [_p, _i, _i64, _p, _p],
    "pb2_nccl_available": [],
    "pb2_nccl_gallery_allgather": [_p, _p, _i64, _i64, _p, _p],
    "pb2_nccl_colstat_merge": [_p, _p, _i64, _p, _p, _i, _p],
    "pb2_nccl_dv_reduce_scatter": [_p, _p, _i64, _i, _p, _p],
    "pb2_contrastive_matrix": [_p, _i64, _i64, _f, _p, _i, _p, _i64, _f, _p, _p],
    "pb2_host_random_doubles": [_p, _i64, _p],
    "pb2_host_sample_pairs": [_p, _p, _p, _i64, _i64, _p, _p],
}
_RESTYPE = {"pb2_last_error": C.c_char_p, "pb2_launch_count": C.c_longlong, "pb2_hinge_step_workspace": C.c_int64, "pb2_hinge_forward_workspace": C.c_int64, "pb2_hinge_state_bytes": C.c_int64,
            "pb2_grad_gemm_workspace": C.c_int64}
# selectors / knock-outs of the MEASUREMENT build only (libpeppa_b200_measure.so, -DPB2_MEASURE): not in the public
# header and not exported by the product library
_DEBUG = {"pb2_debug_force_bn": [_i], "pb2_debug_gg_pair": [_i], "pb2_debug_proj_variant": [_i], "pb2_debug_gg_units": [_i], "pb2_debug_sim_pair": [_i], "pb2_debug_step_stages": [_i], "pb2_debug_set_mn_desc": [C.c_uint32, C.c_uint32, C.c_uint32]}

_lib = None


def header_symbols():
    """Every function name declared in include/peppa_b200.h."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pb2_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"peppa_b200: native library not built ({LIB_PATH} missing). Run `python -m peppa_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    _lib = _bind(C.CDLL(LIB_PATH), SIGNATURES)
    return _lib


def _bind(handle, table, optional=()):
    for name, args in table.items():
        if name in optional and not hasattr(handle, name):      # an older variant build under tools/ab/ (A/B runs)
            continue
        fn = getattr(handle, name)  # AttributeError if the export is missing
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    return handle


FAST_PATH = os.path.join(_HERE, "csrc", "_pb2_fast.so")
_fast = None


def fast():
    """The C++ autograd glue over the same C ABI (csrc/torch_fast.cpp) for the launch-bound training step, or None:
    not built (older checkouts), the measurement build is active (tools / variant tests route through ctypes), or
    PEPPA_B200_NO_FAST=1.  Both paths end in pb2_hinge_forward / pb2_hinge_backward of the product library."""
    global _fast
    if _fast is None:
        _fast = False
        if os.environ.get("PEPPA_B200_NO_FAST") != "1" and os.path.exists(FAST_PATH):
            import importlib.util
            lib()                           # the product library first: the extension's DT_NEEDED resolves to the same file
            try:
                spec = importlib.util.spec_from_file_location("_pb2_fast", FAST_PATH)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _fast = mod
            except (ImportError, OSError) as e:     # e.g. built against another torch: the ctypes path does the same work
                import warnings
                warnings.warn(f"peppa_b200: {FAST_PATH} does not load ({e}); TripletLoss keeps the ctypes path "
                              "(rebuild with `python -m peppa_b200.build`)")
    if _fast is False or (_measure is not None and _lib is _measure):
        return None
    return _fast


_measure = None


class measurement_library:
    """``with measurement_library() as lib:`` routes every ``lib()`` call of this process through the measurement
    build (same kernels plus the ``pb2_debug_*`` selectors) for the duration of the block.  For tools/ and the
    variant tests only; the product path never enters it."""

    def __enter__(self):
        global _lib, _measure
        if _measure is None:
            if not os.path.exists(MEASURE_LIB_PATH):
                raise RuntimeError(f"peppa_b200: measurement library not built ({MEASURE_LIB_PATH}); run `python -m peppa_b200.build`")
            _measure = _bind(C.CDLL(MEASURE_LIB_PATH), {**SIGNATURES, **_DEBUG}, optional=_DEBUG)
        self._saved = lib()
        _lib = _measure
        return _measure

    def __exit__(self, *a):
        global _lib
        _lib = self._saved


class Pb2Error(RuntimeError):
    pass


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().pb2_last_error()
        raise Pb2Error(f"peppa_b200 {what} failed (status {status}): {msg.decode() if msg else ''}")


def use_measurement_library():
    """Switch this process to the measurement build for good (tools/ scripts call this first)."""
    return measurement_library().__enter__()
