"""ctypes binding of libnlspn_b200.so (C ABI: include/nlspn_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Build it with ``python -m nlspn_eccv20_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
import threading

from .build import LIB_PATH

_c = ctypes
_fp = _c.c_void_p  # every tensor argument is passed as a raw device address

_lock = threading.Lock()
_lib = None

# name -> (restype, argtypes); mirrors include/nlspn_b200.h one to one
SIGNATURES = {
    "nlspn_abi_version": (_c.c_int, []),
    "nlspn_last_error": (_c.c_char_p, []),
    "nlspn_launch_count": (_c.c_ulonglong, []),
    "nlspn_set_option": (_c.c_int, [_c.c_char_p, _c.c_int]),
    "nlspn_get_option": (_c.c_int, [_c.c_char_p, _c.POINTER(_c.c_int)]),
    "nlspn_reset_options": (_c.c_int, []),
    "nlspn_profile_enable": (_c.c_int, [_c.c_int]),
    "nlspn_profile_classes": (_c.c_int, []),
    "nlspn_profile_class_name": (_c.c_char_p, [_c.c_int]),
    "nlspn_profile_read": (_c.c_int, [_c.POINTER(_c.c_double), _c.POINTER(_c.c_longlong), _c.c_int]),
    "nlspn_device_info": (_c.c_int, [_c.c_int, _c.POINTER(_c.c_int), _c.POINTER(_c.c_int)]),
    "nlspn_prologue_fwd": (_c.c_int, [_fp, _fp, _fp, _fp, _fp, _c.c_int, _c.c_uint,
                                      _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                      _fp, _fp, _fp, _fp, _fp]),
    "nlspn_propagate_fwd": (_c.c_int, [_fp, _fp, _fp, _fp, _c.c_uint,
                                       _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                       _fp, _c.c_int, _fp, _fp]),
    "nlspn_forward": (_c.c_int, [_fp, _fp, _fp, _fp, _fp, _c.c_int, _c.c_uint,
                                 _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                 _fp, _fp, _fp, _fp, _c.c_int, _fp, _fp]),
    "nlspn_backward_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    "nlspn_backward_workspace_bytes_ex": (_c.c_size_t, [_c.c_int] * 5 + [_c.c_uint]),
    "nlspn_heads_packed_floats": (_c.c_size_t, [_c.c_int]),
    "nlspn_heads_pack": (_c.c_int, [_c.c_void_p] * 3 + [_c.c_int, _c.c_void_p, _c.c_void_p]),
    "nlspn_heads_fwd": (_c.c_int, [_c.c_void_p] * 6 + [_c.c_int] * 4 + [_c.c_void_p] * 4),
    "nlspn_heads_prologue_supported": (_c.c_int, [_c.c_int, _c.c_int]),
    "nlspn_heads_wgrad_supported": (_c.c_int, [_c.c_int, _c.c_int]),
    "nlspn_heads_grad_prep": (_c.c_int, [_c.c_void_p] * 5 + [_c.c_int] * 4 + [_c.c_void_p] * 3),
    "nlspn_heads_wgrad": (_c.c_int, [_c.c_void_p] * 5 + [_c.c_int] * 4 + [_c.c_void_p] * 2),
    "nlspn_heads_dgrad_one": (_c.c_int, [_c.c_void_p] * 3 + [_c.c_int] * 4 + [_c.c_void_p] * 3),
    "nlspn_heads_dgrad_packed_floats": (_c.c_size_t, [_c.c_int]),
    "nlspn_heads_dgrad_supported": (_c.c_int, [_c.c_int, _c.c_int]),
    "nlspn_heads_dgrad_pack": (_c.c_int, [_c.c_void_p] * 3 + [_c.c_int, _c.c_void_p, _c.c_void_p]),
    "nlspn_heads_dgrad_wide": (_c.c_int, [_c.c_void_p] * 2 + [_c.c_int] * 4 + [_c.c_void_p] * 3),
    "nlspn_heads_prologue_fwd": (_c.c_int, [_c.c_void_p] * 8 + [_c.c_int, _c.c_uint] + [_c.c_int] * 4 + [_c.c_void_p] * 8),
    "nlspn_backward": (_c.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _c.c_int, _fp,
                                  _c.POINTER(_c.c_void_p), _fp, _fp, _fp, _c.c_int, _c.c_uint,
                                  _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                  _fp, _fp, _fp, _fp, _fp, _c.c_size_t, _fp]),
    "nlspn_dcn_forward": (_c.c_int, [_fp] * 5 + [_c.c_int] * 15 + [_fp, _fp]),
    "nlspn_dcn_backward": (_c.c_int, [_fp] * 6 + [_c.c_int] * 15 + [_fp] * 6),
    "nlspn_dcn_backward_workspace_bytes": (_c.c_size_t, [_c.c_int] * 4),
    "nlspn_dcn_backward_ws": (_c.c_int, [_fp] * 6 + [_c.c_int] * 15 + [_fp] * 5 + [_fp, _c.c_size_t, _fp]),
    "nlspn_step_fwd": (_c.c_int, [_fp] * 5 + [_c.c_uint] + [_c.c_int] * 4 + [_fp, _fp, _fp]),
    "nlspn_step_bwd_workspace_bytes": (_c.c_size_t, [_c.c_int] * 4 + [_c.c_uint]),
    "nlspn_step_bwd": (_c.c_int, [_fp] * 8 + [_c.c_uint] + [_c.c_int] * 4 + [_fp] * 4 + [_fp, _c.c_size_t, _fp]),
    "nlspn_dcn_forward_f64": (_c.c_int, [_fp] * 5 + [_c.c_int] * 15 + [_fp, _fp]),
    "nlspn_dcn_backward_f64": (_c.c_int, [_fp] * 6 + [_c.c_int] * 15 + [_fp] * 6),
    "nlspn_debug_indices": (_c.c_int, [_fp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _fp, _fp]),
}

AFFINITY = {"AS": 0, "ASS": 1, "TC": 2, "TGASS": 3}
FLAG_PRESERVE_INPUT = 1
FLAG_ALWAYS_CLIP = 2
FLAG_NO_OFFSET = 4
FLAG_BLEND_PRE = 8
FLAG_CONF_SAMPLED = 16
FLAG_LEGACY = 32
FLAG_BWD_PER_ITERATION = 0x100
FLAG_DETERMINISTIC = 0x200


def lib_path() -> str:
    return LIB_PATH


def load():
    """Load the shared library (once).  Raises RuntimeError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "nlspn_eccv20_b200: %s is missing -- the CUDA extension is REQUIRED (there is "
                    "no CPU/PyTorch fallback). Build it: python -m nlspn_eccv20_b200.build" % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the header and the .so disagree
                fn.restype = res
                fn.argtypes = args
            if lib.nlspn_abi_version() != 1:
                raise RuntimeError("libnlspn_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().nlspn_last_error().decode("utf-8", "replace")
        kind = "validation error" if rc < 0 else "CUDA error"
        raise RuntimeError("%s failed: %s %d: %s" % (what, kind, rc, msg))


def set_option(name: str, value: int):
    """Tuning knob of the library; -1 = auto where the default depends on the shape.  The
    NLSPN_<NAME> environment variables only seed the defaults when the library is loaded."""
    check(load().nlspn_set_option(name.encode(), int(value)), "nlspn_set_option(%s)" % name)


def get_option(name: str) -> int:
    v = _c.c_int(0)
    check(load().nlspn_get_option(name.encode(), _c.byref(v)), "nlspn_get_option(%s)" % name)
    return v.value


class options:
    """``with _lib.options(state_gather=1, gather_compact=0): ...`` -- set, then restore."""

    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = get_option(k)
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)
        return False


def profile_read():
    """-> {kernel class name: (total ms, launches)} for the launches recorded since profile_enable(1)."""
    lib = load()
    n = lib.nlspn_profile_classes()
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_longlong * n)()
    check(lib.nlspn_profile_read(ms, cnt, n), "nlspn_profile_read")
    return {lib.nlspn_profile_class_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i]}
