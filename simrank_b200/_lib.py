"""ctypes binding of libsimrank_b200.so (the C ABI in include/simrank_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails this module
raises.  PyTorch is used by the callers only to own device memory and streams; nothing
here takes a torch type.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libsimrank_b200.so")
ABI_VERSION = 12

SRK_X2_MID, SRK_X2_FINAL, SRK_X2_COUNTS = 0, 1, 2
SRK_X2_DIRECT, SRK_X2_SYMMETRIC, SRK_X2_TRANSPOSED = 0, 1, 2


class EngineError(RuntimeError):
    """A libsimrank_b200 call returned a non-zero status."""


class Epilogue(C.Structure):
    _fields_ = [("coef", C.c_double),
                ("evidence", C.c_void_p), ("ld_evidence", C.c_int64),
                ("prior", C.c_void_p), ("ld_prior", C.c_int64),
                ("lambda_", C.c_double),
                ("s_old", C.c_void_p), ("ld_s_old", C.c_int64),
                ("maxdiff", C.c_void_p), ("maxoff", C.c_void_p),
                ("diag_offset", C.c_int64)]


class RowBound(C.Structure):
    """bound(r) = vec[r]*mul + add   (vec == NULL: add)."""
    _fields_ = [("vec", C.c_void_p), ("mul", C.c_double), ("add", C.c_double)]

    @classmethod
    def of(cls, vec_ptr, mul, add):
        rb = cls()
        rb.vec, rb.mul, rb.add = vec_ptr, float(mul), float(add)
        return rb


SRK_ELEM_F64, SRK_ELEM_U16 = 0, 1
SRK_CSR_FIRST, SRK_CSR_FINAL, SRK_CSR_ACCUM, SRK_CSR_FINISH, SRK_CSR_FINISH_FIRST = 0, 1, 2, 3, 4


class CsrArgs(C.Structure):
    _fields_ = [("elem", C.c_int), ("mode", C.c_int), ("symmetric", C.c_int),
                ("indptr", C.c_void_p), ("indices", C.c_void_p), ("g", C.c_void_p),
                ("M", C.c_int64), ("row_begin", C.c_int64), ("row_end", C.c_int64),
                ("X", C.c_void_p), ("ldx", C.c_int64), ("L", C.c_int64), ("K", C.c_int64),
                ("OUT", C.c_void_p), ("ldo", C.c_int64),
                ("in_unit", RowBound), ("out_bound", RowBound), ("qmax", C.c_double),
                ("g_col", C.c_void_p),
                ("counts", C.c_void_p), ("ld_counts", C.c_int64),
                ("counts_bits", C.c_int), ("add_counts", C.c_int), ("use_evidence", C.c_int),
                ("epi", Epilogue),
                ("row_lo", C.c_void_p), ("row_hi", C.c_void_p),
                ("accum", C.c_void_p), ("ld_accum", C.c_int64), ("accum_slot", C.c_void_p)]


class X2Args(C.Structure):
    _fields_ = [("mode", C.c_int), ("ns", C.c_int), ("layout", C.c_int),
                ("M", C.c_int64), ("R", C.c_int64), ("K", C.c_int64),
                ("A8", C.c_void_p), ("lda", C.c_int64),
                ("in_planes", C.c_void_p), ("ld_in", C.c_int64), ("in_plane_stride", C.c_int64),
                ("in_kblock", C.c_int64), ("in_kblock_stride", C.c_int64),
                ("in_rowbound", RowBound),
                ("out_planes", C.c_void_p), ("ld_outp", C.c_int64), ("out_plane_stride", C.c_int64),
                ("out_rowbound", RowBound),
                ("g_a", C.c_void_p), ("g_v", C.c_void_p),
                ("counts", C.c_void_p), ("ld_counts", C.c_int64),
                ("add_counts", C.c_int), ("use_evidence", C.c_int), ("counts_bits", C.c_int),
                ("out_f64", C.c_void_p), ("ld_out", C.c_int64), ("diag_offset", C.c_int64),
                ("mirror_out", C.c_void_p), ("ld_mirror", C.c_int64), ("mirror_col0", C.c_int64),
                ("rowmax_hi", C.c_void_p),
                ("epi", Epilogue),
                ("out_counts", C.c_void_p), ("ld_out_counts", C.c_int64),
                ("sync_ws", C.c_void_p), ("sync_ws_bytes", C.c_int64),
                ("mirror_rowmax_hi", C.c_void_p)]


_P, _I64, _INT, _DBL = C.c_void_p, C.c_int64, C.c_int, C.c_double
# name -> (restype, argtypes); must list every symbol include/simrank_b200.h declares
SYMBOLS = {
    "srk_abi_version": (_INT, []),
    "srk_last_error": (C.c_char_p, []),
    "srk_device_cc": (_INT, []),
    "srk_csr_half_f64": (_INT, [_P, _P, _P, _I64, _I64, _I64, _P, _I64, _I64, _P, _I64, C.POINTER(Epilogue), _P]),
    "srk_csr_half": (_INT, [C.POINTER(CsrArgs), _P]),
    "srk_quantize_rows_u16": (_INT, [_P, _I64, _I64, _I64, _I64, _P, _I64, _P, _DBL, _INT, _P]),
    "srk_edges_to_csr_workspace": (C.c_size_t, [_I64, _I64]),
    "srk_edges_to_csr": (_INT, [_P, _P, _I64, _I64, _I64, _P, _P, _P, _P, C.c_size_t, _P]),
    "srk_csr_evidence_counts": (_INT, [_P, _P, _P, _I64, _I64, _I64, _P, _I64, _P]),
    "srk_csr_row_spread": (_INT, [_P, _P, _P, _I64, _P, _P]),
    "srk_csr_to_dense_u8": (_INT, [_P, _P, _I64, _I64, _I64, _P, _I64, _P]),
    "srk_slice_rows_f64": (_INT, [_P, _I64, _I64, _I64, C.POINTER(RowBound), _I64, _INT, _P, _I64, _I64, _P]),
    "srk_x2_half": (_INT, [C.POINTER(X2Args), _P]),
    "srk_slice_rows_max_f64": (_INT, [_P, _I64, _I64, _I64, _I64, _INT, _P, _I64, _I64, _P, _P]),
    "srk_slice_rows_key_f64": (_INT, [_P, _I64, _I64, _I64, _I64, _INT, _P, _P, _I64, _I64, _P, _P]),
    "srk_i8_supported": (_INT, []),
    "srk_topk_rows": (_INT, [_P, _I64, _I64, _I64, _INT, _P, _P, _P]),
    "srk_set_identity_f64": (_INT, [_P, _I64, _I64, _I64, _I64, _P]),
}

_lib = None


def load():
    """Load the shared library once; raise if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m simrank_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.srk_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.srk_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


LAUNCHES = 0      # successful library calls so far; every one of them enqueues at least one kernel


def check(rc: int, what: str = ""):
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        msg = load().srk_last_error().decode("utf-8", "replace")
        raise EngineError(f"{what or 'libsimrank_b200'} failed ({rc}): {msg}")
