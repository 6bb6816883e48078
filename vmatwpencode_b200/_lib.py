"""ctypes binding of libvmat_b200.so (C ABI declared in include/vmat_b200.h).

There is no CPU fallback: if the library is missing or a call fails, an exception
is raised.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C vmatwpencode_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VMAT_B200_LIB overrides the library path (A/B comparisons of kernel builds only)
LIB_PATH = os.environ.get("VMAT_B200_LIB") or os.path.join(_HERE, "libvmat_b200.so")

VMAT_F32, VMAT_BF16, VMAT_F64 = 0, 1, 2
VMAT_HOST, VMAT_DEVICE = 0, 1
VMAT_PRICE_SAMPLING, VMAT_PRICE_CLASSIC, VMAT_PRICE_SHORT = 0, 1, 2
DTYPE_CODES = {"f32": VMAT_F32, "bf16": VMAT_BF16, "f64": VMAT_F64}


class VmatError(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [("store_dtype", C.c_int32), ("acc_dtype", C.c_int32), ("tile_voxels", C.c_int32),
                ("k1_threads", C.c_int32), ("k2_threads", C.c_int32), ("k2_tasks", C.c_int32),
                ("use_graph", C.c_int32), ("k1_stages", C.c_int32), ("k2_stages", C.c_int32), ("chunk_cap", C.c_int32),
                ("no_pdl", C.c_int32), ("host_path", C.c_int32), ("index_bytes", C.c_int32), ("claim_lead", C.c_int32),
                ("k2_chunk_cap", C.c_int32), ("eval_kernels", C.c_int32), ("k2_pool_pct", C.c_int32),
                ("consume_batch", C.c_int32), ("reserved", C.c_int32 * 2)]


class PriceRow(C.Structure):
    _fields_ = [("l_first", C.c_double), ("n_l", C.c_int32), ("vb_min", C.c_int32),
                ("r_lo", C.c_double), ("r_hi", C.c_double), ("r_mid", C.c_double),
                ("vb_mask", C.c_uint64), ("grad_offset", C.c_int32), ("n_r_mid", C.c_int32)]


class PriceParams(C.Structure):
    _fields_ = [("C", C.c_double), ("C2", C.c_double), ("C3", C.c_double), ("b", C.c_double),
                ("M", C.c_int32), ("N", C.c_int32), ("variant", C.c_int32), ("reserved", C.c_int32 * 5)]


# every symbol include/vmat_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "vmat_version": (C.c_int, []),
    "vmat_last_error": (C.c_char_p, []),
    "vmat_default_options": (None, [C.POINTER(Options)]),
    "vmat_plan_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int64, C.c_int32, _P,
                                   _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, C.POINTER(Options)]),
    "vmat_plan_destroy": (C.c_int, [_P]),
    "vmat_plan_info": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vmat_set_aperture": (C.c_int, [_P, C.c_int32, _P, _P]),
    "vmat_eval": (C.c_int, [_P, _P, C.POINTER(C.c_double), _P]),
    "vmat_set_intensities": (C.c_int, [_P, _P]),
    "vmat_eval_async": (C.c_int, [_P, _P, C.c_int]),
    "vmat_fetch_result": (C.c_int, [_P, C.POINTER(C.c_double), _P]),
    "vmat_sync": (C.c_int, [_P]),
    "vmat_set_timing": (C.c_int, [_P, C.c_int]),
    "vmat_get_timing": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "vmat_trace_eval": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vmat_guard_status": (C.c_int, [_P, _P, C.c_int32]),
    "vmat_reduce_buffer": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "vmat_comm_handle": (C.c_int, [_P, _P]),
    "vmat_comm_attach": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "vmat_comm_attach_local": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "vmat_comm_set_timeout": (C.c_int, [_P, C.c_double]),
    "vmat_intensities_buffer": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "vmat_result_buffer": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "vmat_get_dose": (C.c_int, [_P, _P]),
    "vmat_get_voxelgrad": (C.c_int, [_P, _P]),
    "vmat_get_beamlet_grad": (C.c_int, [_P, _P]),
    "vmat_rmp_begin": (C.c_int, [_P, C.c_int32, _P]),
    "vmat_rmp_eval": (C.c_int, [_P, _P, C.POINTER(C.c_double), _P]),
    "vmat_rmp_end": (C.c_int, [_P]),
    "vmat_dose_max": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "vmat_dvh_histogram": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, _P]),
    "vmat_price": (C.c_int, [_P, C.c_int32, _P, C.POINTER(PriceParams), C.POINTER(PriceRow), _P, _P, _P]),
    "vmat_price_ex": (C.c_int, [_P, C.c_int32, _P, C.POINTER(PriceParams), C.POINTER(PriceRow), _P, _P, _P, _P]),
    "vmat_kmedoid_costs": (C.c_int, [C.c_int, C.c_int64, C.c_int32, _P, _P, _P]),
    "vmat_kmedoid_costs_shard": (C.c_int, [C.c_int, C.c_int64, _P, C.c_int64, C.c_int32, _P, _P, _P]),
    "vmat_kmedoid_open": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int32, _P, _P, C.c_int64, _P]),
    "vmat_kmedoid_insert": (C.c_int, [_P, _P]),
    "vmat_kmedoid_costs_resident": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "vmat_kmedoid_close": (C.c_int, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library once; raise VmatError (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise VmatError(f"{LIB_PATH} not built: run __graft_entry__.build() "
                            "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().vmat_last_error().decode(errors="replace")
        raise VmatError(f"{what} failed ({rc}): {msg}")
