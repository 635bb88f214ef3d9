"""ctypes binding of include/tvidz_b200.h.  There is no CPU fallback: if the CUDA
library is missing or fails to load, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TVZ_LIB: tuning hook (scripts/sweep_fragment.py loads kernel-shape variants of the same library)
LIB_PATH = os.environ.get("TVZ_LIB") or os.path.join(_HERE, "libtvidz_b200.so")

# every symbol include/tvidz_b200.h declares: (name, restype, argtypes)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
SYMBOLS = [
    ("tvz_last_error", C.c_char_p, []),
    ("tvz_abi_version", _i, []),
    ("tvz_sad_luma_u8", _i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _vp, _vp]),
    ("tvz_sad_luma_u16", _i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _vp, _vp]),
    ("tvz_sad_luma_u8_path", _i, [_vp, _i, _i, _i64, _i64, _i64]),
    ("tvz_scene_select", _i, [_vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _vp]),
    ("tvz_scene_score_host", _i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _i, _d, _i, _vp, _vp, _vp]),
    ("tvz_nvdec_available", _i, []),
    ("tvz_nvdec_library", C.c_char_p, []),
    ("tvz_nvdec_caps", _i, [_i, _i, _vp]),
    ("tvz_decoder_create", _i, [_i, _i64, C.POINTER(_vp)]),
    ("tvz_decoder_destroy", None, [_vp]),
    ("tvz_decoder_feed", _i, [_vp, _vp, _i64, _i64, _i, C.POINTER(_i64)]),
    ("tvz_decoder_info", _i, [_vp, _vp]),
    ("tvz_decoder_ring", _vp, [_vp]),
    ("tvz_decoder_pts", _i, [_vp, _i64, _i64, _vp]),
    ("tvz_catalog_create", _i, [_vp, _vp, _vp, _i64, C.POINTER(_vp)]),
    ("tvz_catalog_create_mutable", _i, [_vp, _vp, _vp, _i64, _i64, C.POINTER(_vp)]),
    ("tvz_catalog_destroy", None, [_vp]),
    ("tvz_catalog_rows", _i64, [_vp]),
    ("tvz_catalog_values", _i64, [_vp]),
    ("tvz_catalog_algo_bytes", _i64, [_vp]),
    ("tvz_catalog_tiles", _i, [_vp]),
    ("tvz_catalog_upsert", _i, [_vp, C.c_int32, _vp, _i]),
    ("tvz_catalog_tail_info", _i, [_vp, _vp]),
    ("tvz_match_ws_create", _i, [_vp, _i64, C.POINTER(_vp)]),
    ("tvz_match_ws_destroy", None, [_vp]),
    ("tvz_catalog_match", _i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    ("tvz_catalog_batch_limit", _i, []),
    ("tvz_catalog_batch_size", _i, []),
    ("tvz_catalog_match_batch", _i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i64, C.POINTER(_i64),
                                     C.POINTER(_i64)]),
    ("tvz_catalog_match_async", _i, [_vp, _vp, _vp, _i, _i, _vp, _i64, _vp]),
    ("tvz_catalog_match_batch_async", _i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i64, _vp]),
    ("tvz_catalog_match_gather_async", _i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i64, _i64, C.c_uint32, _vp]),
    ("tvz_catalog_match_batch_gather_async", _i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i64, _i64, C.c_uint32,
                                                  _vp]),
    ("tvz_copy_records_to_host", _i, [_vp, _vp, _i, _i64, _i64, _i64, _i, _vp]),
    ("tvz_match_ws_hits", _vp, [_vp]),
    ("tvz_match_ws_nhits", _vp, [_vp]),
    ("tvz_fragcat_create", _i, [_vp, _vp, _vp, _i64, _d, _i64, C.POINTER(_vp)]),
    ("tvz_fragcat_destroy", None, [_vp]),
    ("tvz_fragcat_rows", _i64, [_vp]),
    ("tvz_fragcat_values", _i64, [_vp]),
    ("tvz_fragcat_match", _i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    ("tvz_fragcat_match_async", _i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    ("tvz_fragcat_match_gather_async", _i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i64, C.c_uint32, _vp]),
]
# debug hooks outside the public header
_DEBUG_SYMBOLS = [("tvz_debug_sad_tuning", _i, [_i, _i, _i, _i]),
                  ("tvz_debug_flush_l2", _i, [_vp, _i64, _vp]),
                  ("tvz_debug_match_timing", _i, [_vp, _i]),
                  ("tvz_debug_match_count_ms", _i, [_vp, C.POINTER(C.c_float)]),
                  ("tvz_debug_tile_trace", _i, [_vp, _vp]),
                  ("tvz_debug_arrange_fingerprints", _i, [_vp, _i64, _vp, _vp]),
                  ("tvz_debug_build_tiles", _i, [_vp, _i64, _i, _vp, _i, _vp])]

TVZ_ERR_OVERFLOW = -4
_lib = None


class TvzError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tvidz_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load the CUDA library; raise loudly (no fallback) when it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m tvidz_b200.build` "
                "(nvcc, sm_100a). tvidz_b200 has no CPU path.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS + _DEBUG_SYMBOLS:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise TvzError(rc, (lib().tvz_last_error() or b"").decode(errors="replace"))
