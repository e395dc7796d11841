"""ctypes binding of the C-ABI in include/nans_clip.h (libnans_clip.so, built in-tree).

There is no fallback: if the shared library is missing and cannot be built, importing the product
path raises.  Every non-zero return code becomes a NansError carrying nans_last_error().
"""
from __future__ import annotations

import ctypes
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libnans_clip.so"

NANS_F32, NANS_F16, NANS_BF16 = 0, 1, 2
NANS_LOSS_WITH_ACC = 1
NANS_LOSS_STRIP_IMG = 2
NANS_LOSS_STRIP_TXT = 4

ERR_NAMES = {-1: "NANS_ERR_ARG", -2: "NANS_ERR_DEVICE", -3: "NANS_ERR_CUDA", -4: "NANS_ERR_WORKSPACE"}


class NansError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


_SIGNATURES = {
    # name: (restype, argtypes)
    "nans_version": (c_int, []),
    "nans_last_error": (c_char_p, []),
    "nans_device_check": (c_int, []),
    "nans_l2norm_cast": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int,
                                 c_void_p, c_void_p, c_int, c_void_p]),
    "nans_l2norm_bwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                c_void_p, c_void_p]),
    "nans_clip_loss_fwd_phase_slots": (c_int64, [c_int64, c_int64, c_int64]),
    "nans_clip_loss_fwd_phase_slots_flags": (c_int64, [c_int64, c_int64, c_int64, c_int]),
    "nans_clip_loss_fwd_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "nans_clip_loss_fwd_phase": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                         c_int, c_int64, c_int64, c_int64, c_int64, c_int64,
                                         c_int64, c_int64,
                                         c_void_p, c_int, c_void_p, c_size_t, c_int64, c_void_p]),
    "nans_clip_loss_fwd_phase_rows": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                              c_int, c_int64, c_int64, c_int64, c_int64, c_int64,
                                              c_int64, c_int64, c_int64, c_int64,
                                              c_void_p, c_int, c_void_p, c_size_t, c_int64, c_void_p]),
    "nans_clip_loss_fwd_finalize_rows": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p,
                                                 c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nans_clip_loss_fwd_finalize": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p,
                                            c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nans_clip_loss_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int,
                                   c_int64, c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nans_clip_loss_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "nans_clip_loss_bwd_plan": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_int]),
    "nans_clip_loss_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int,
                                   c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_float, c_int64, c_int64, c_void_p, c_void_p, c_int,
                                   c_void_p, c_size_t, c_void_p]),
    "nans_clip_loss_bwd_minmax": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int,
                                          c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_float, c_int64, c_int64, c_void_p, c_void_p, c_int,
                                          c_void_p, c_size_t, c_void_p]),
    "nans_clip_loss_exchange_finish": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                               c_void_p, c_void_p]),
    "nans_peer_alloc": (c_int, [c_size_t, c_void_p, c_void_p]),
    "nans_peer_free": (c_int, [c_void_p]),
    "nans_peer_open": (c_int, [c_void_p, c_void_p]),
    "nans_peer_close": (c_int, [c_void_p]),
    "nans_peer_zero": (c_int, [c_void_p, c_size_t, c_void_p]),
    "nans_xchg_layout": (c_int, [c_void_p, c_int64, c_int64]),
    "nans_xchg_cast_local": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p]),
    "nans_xchg_push": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p]),
    "nans_xchg_cast_local_dma": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p,
                                         c_int, c_void_p]),
    "nans_xchg_push_dma": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "nans_xchg_push_peers": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "nans_xchg_push_dma_peers": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "nans_xchg_cast_push": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p]),
    "nans_clip_loss_fwd_xchg_slots": (c_int64, [c_int64, c_int64, c_int64]),
    "nans_clip_loss_fwd_xchg": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t,
                                        c_void_p]),
    "nans_clip_loss_fwd_finalize_push": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_size_t, c_void_p,
                                                 c_void_p, c_void_p, c_void_p]),
    "nans_clip_loss_exchange_finish_xchg": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                                    c_void_p]),
    "nans_clip_loss_bwd_xchg": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_float, c_int64, c_int64, c_void_p, c_void_p, c_int,
                                        c_void_p, c_size_t, c_void_p]),
    "nans_label_smooth_stats": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "nans_label_smooth_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                      c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p]),
    "nans_jsonl_scan": (c_int, [c_void_p, c_int64, c_char_p, c_void_p, c_void_p]),
    "nans_jsonl_parse": (c_int, [c_void_p, c_int64, c_char_p, c_int64, c_int64, c_void_p, c_void_p, c_int]),
    "nans_jsonl_parse_rows": (c_int, [c_void_p, c_int64, c_char_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                      c_void_p, c_int]),
    "nans_topk_ip_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "nans_topk_ip": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64,
                             c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p]),
    "nans_topk_merge": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p,
                                c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)

_lock = threading.Lock()
_lib = None


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building first if the .so is absent) and type the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            if not build_if_missing:
                raise FileNotFoundError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build`")
            import importlib.util

            spec = importlib.util.spec_from_file_location("_nans_build", PKG / "build.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().nans_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise NansError(rc, last_error())
