"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy binding of oracle/nlspn_oracle.c.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg.  The product package (nlspn_eccv20_b200) never imports this module.

Every function takes / returns C-contiguous numpy arrays of dtype float32 or
float64 (chosen by the dtype of the first array argument) in the reference's
NCHW layouts.  See nlspn_oracle.c for the reference file:line each entry follows.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

AFFINITY = {"AS": 0, "ASS": 1, "TC": 2, "TGASS": 3}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Returns the .so path."""
    src = os.path.join(_HERE, "nlspn_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _suf(a):
    if a.dtype == np.float32:
        return "f32", ctypes.c_float
    if a.dtype == np.float64:
        return "f64", ctypes.c_double
    raise TypeError("oracle supports float32/float64, got %s" % a.dtype)


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt=None):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dt)


def _call(name, suf, *args):
    fn = getattr(lib(), "%s_%s" % (name, suf))
    fn.restype = None
    fn(*args)


def dcn_step_fwd(x, off, msk, wgt=None, bias=None):
    x = _c(x); dt = x.dtype
    suf, _ = _suf(x)
    off, msk, wgt, bias = _c(off, dt), _c(msk, dt), _c(wgt, dt), _c(bias, dt)
    B, _, H, W = x.shape
    K = int(round(msk.shape[1] ** 0.5))
    out = np.empty_like(x)
    _call("dcn_step_fwd", suf, _p(x), _p(off), _p(msk), _p(wgt), _p(bias), B, H, W, K, _p(out))
    return out


def dcn_step_bwd(x, off, msk, gout, wgt=None, want_wb=True):
    x = _c(x); dt = x.dtype
    suf, _ = _suf(x)
    off, msk, gout, wgt = _c(off, dt), _c(msk, dt), _c(gout, dt), _c(wgt, dt)
    B, _, H, W = x.shape
    K = int(round(msk.shape[1] ** 0.5))
    gin = np.empty_like(x)
    goff = np.zeros_like(off)
    gmsk = np.zeros_like(msk)
    gw = np.zeros((1, 1, K, K), dt) if want_wb else None
    gb = np.zeros((1,), dt) if want_wb else None
    _call("dcn_step_bwd", suf, _p(x), _p(off), _p(msk), _p(wgt), _p(gout), B, H, W, K, 0,
          _p(gin), _p(goff), _p(gmsk), _p(gw), _p(gb))
    return gin, goff, gmsk, gw, gb


def dcn_debug_indices(off, K):
    off = _c(off)
    suf, _ = _suf(off)
    B, _, H, W = off.shape
    idx = np.empty((B, K * K, 3, H, W), np.int32)
    _call("dcn_debug_indices", suf, _p(off), B, H, W, K, _p(idx))
    return idx


def prologue_fwd(guidance, conf, dep, gamma, K, affinity="TGASS", preserve=True):
    guidance = _c(guidance); dt = guidance.dtype
    suf, cf = _suf(guidance)
    conf, dep = _c(conf, dt), _c(dep, dt)
    B, _, H, W = guidance.shape
    KK = K * K
    offset = np.empty((B, 2 * KK, H, W), dt)
    aff = np.empty((B, KK, H, W), dt)
    conf_out = np.empty((B, 1, H, W), dt) if conf is not None else None
    _call("nlspn_prologue_fwd", suf, _p(guidance), _p(conf), _p(dep), cf(gamma),
          AFFINITY[affinity], int(preserve), B, H, W, K, _p(offset), _p(aff), _p(conf_out))
    return offset, aff, conf_out


def propagate_fwd(feat_init, offset, aff, conf, dep, T, preserve=True, always_clip=False):
    feat_init = _c(feat_init); dt = feat_init.dtype
    suf, _ = _suf(feat_init)
    offset, aff, conf, dep = _c(offset, dt), _c(aff, dt), _c(conf, dt), _c(dep, dt)
    B, _, H, W = feat_init.shape
    K = int(round(aff.shape[1] ** 0.5))
    list_feat = np.empty((T, B, 1, H, W), dt)
    scratch = np.empty((2 * B * H * W,), dt)
    _call("nlspn_propagate_fwd", suf, _p(feat_init), _p(offset), _p(aff), _p(conf), _p(dep),
          int(preserve and dep is not None), int(always_clip), B, H, W, K, T,
          _p(list_feat), _p(scratch))
    return list_feat


def propagate_bwd(feat_init, offset, aff, conf, dep, list_feat, g_list, preserve=True):
    feat_init = _c(feat_init); dt = feat_init.dtype
    suf, _ = _suf(feat_init)
    offset, aff, conf, dep = _c(offset, dt), _c(aff, dt), _c(conf, dt), _c(dep, dt)
    list_feat, g_list = _c(list_feat, dt), _c(g_list, dt)
    T, B, _, H, W = list_feat.shape
    K = int(round(aff.shape[1] ** 0.5))
    g_init = np.empty_like(feat_init)
    g_off = np.empty_like(offset)
    g_aff = np.empty_like(aff)
    g_conf = np.empty_like(feat_init) if conf is not None else None
    scratch = np.empty((3 * B * H * W,), dt)
    _call("nlspn_propagate_bwd", suf, _p(feat_init), _p(offset), _p(aff), _p(conf), _p(dep),
          int(preserve and dep is not None), _p(list_feat), _p(g_list), B, H, W, K, T,
          _p(g_init), _p(g_off), _p(g_aff), _p(g_conf), _p(scratch))
    return g_init, g_off, g_aff, g_conf


def prologue_bwd(guidance, dep, gamma, K, g_offset, g_aff, g_conf_fixed,
                 affinity="TGASS", preserve=True):
    guidance = _c(guidance); dt = guidance.dtype
    suf, cf = _suf(guidance)
    dep, g_offset, g_aff, g_conf_fixed = _c(dep, dt), _c(g_offset, dt), _c(g_aff, dt), _c(g_conf_fixed, dt)
    B, _, H, W = guidance.shape
    g_guid = np.empty_like(guidance)
    g_conf = np.empty((B, 1, H, W), dt) if g_conf_fixed is not None else None
    gg = ctypes.c_double(0.0)
    _call("nlspn_prologue_bwd", suf, _p(guidance), _p(dep), cf(gamma), AFFINITY[affinity],
          int(preserve and dep is not None), _p(g_offset), _p(g_aff), _p(g_conf_fixed),
          B, H, W, K, _p(g_guid), _p(g_conf), ctypes.byref(gg))
    return g_guid, g_conf, gg.value


# ------------------------------------------------------------------------------------
# Module-level convenience: the whole path with the north-star NLSPN signature.
# ------------------------------------------------------------------------------------
def nlspn_forward(feat_init, guidance, confidence, feat_fix, gamma, K, T,
                  affinity="TGASS", preserve=True, always_clip=False):
    """-> dict(feat_result, list_feat, offset, aff, confidence)  (nlspnmodel.py:323-381)."""
    offset, aff, conf = prologue_fwd(guidance, confidence, feat_fix, gamma, K, affinity,
                                     preserve and feat_fix is not None)
    lf = propagate_fwd(feat_init, offset, aff, conf, feat_fix, T,
                       preserve and feat_fix is not None, always_clip)
    return dict(feat_result=lf[-1], list_feat=lf, offset=offset, aff=aff, confidence=conf)


def nlspn_backward(feat_init, guidance, confidence, feat_fix, gamma, K, fwd, g_list,
                   affinity="TGASS", preserve=True, g_offset_out=None, g_aff_out=None):
    """Gradients wrt (feat_init, guidance, confidence, gamma) given g_list [T,B,1,H,W]."""
    pres = preserve and feat_fix is not None
    g_init, g_off, g_aff, g_conf_fixed = propagate_bwd(
        feat_init, fwd["offset"], fwd["aff"], fwd["confidence"], feat_fix, fwd["list_feat"],
        g_list, pres)
    if g_offset_out is not None:
        g_off = g_off + g_offset_out
    if g_aff_out is not None:
        g_aff = g_aff + g_aff_out
    g_guid, g_conf, g_gamma = prologue_bwd(guidance, feat_fix, gamma, K, g_off, g_aff,
                                           g_conf_fixed, affinity, pres)
    return g_init, g_guid, g_conf, g_gamma
