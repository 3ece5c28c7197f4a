"""On-disk formats on either side of the path, host-side and dependency-light.

* KITTI depth PNG, 16 bit: value = depth[m] * 256, 0 = no measurement.  Reader: reference
  src/data/kittidc.py:71-82 (``read_depth``); writer: src/summary/nlspnsummary.py:176-182
  (``clamp(pred, 0) * 256 -> uint16``; truncation, not rounding, exactly as ``astype`` does).
* ``offset.npy`` / ``aff.npy`` / ``gamma.npy`` dumps of the propagation geometry
  (src/summary/nlspnsummary.py:265-268): plain ``numpy.save`` of the first batch element.
"""
from __future__ import annotations

import os

import numpy as np


def write_depth_png16(path, depth):
    """depth: [H,W] array-like in metres (torch tensor or numpy).  Negative values clamp to 0."""
    from PIL import Image
    if hasattr(depth, "detach"):
        depth = depth.detach().float().cpu().numpy()
    d = np.asarray(depth, dtype=np.float32)
    if d.ndim != 2:
        raise ValueError("write_depth_png16 expects a [H,W] map, got shape %s" % (d.shape,))
    d = np.clip(d, 0.0, None) * 256.0
    if float(d.max(initial=0.0)) > 65535.0:
        raise ValueError("depth * 256 exceeds the uint16 range")
    Image.fromarray(d.astype(np.uint16)).save(path)


def read_depth_png16(path):
    """-> float32 [H,W] in metres (kittidc.py:71-82, including its sanity assert)."""
    from PIL import Image
    if not os.path.exists(path):
        raise FileNotFoundError("file not found: {}".format(path))
    img = np.array(Image.open(path))
    if not (img.max() == 0 or img.max() > 255):
        raise ValueError("np.max(depth_png)={}, path={}".format(img.max(), path))
    return img.astype(np.float32) / 256.0


def dump_geometry(dirname, offset, aff, gamma):
    """offset.npy (skipped when None), aff.npy, gamma.npy of the FIRST image, as the reference writes
    them (nlspnsummary.py:185-191,265-268)."""
    os.makedirs(dirname, exist_ok=True)

    def first(t):
        if hasattr(t, "detach"):
            t = t.detach().float().cpu().numpy()
        return np.asarray(t)

    if offset is not None:
        np.save(os.path.join(dirname, "offset.npy"), first(offset)[0])
    np.save(os.path.join(dirname, "aff.npy"), first(aff)[0])
    np.save(os.path.join(dirname, "gamma.npy"), first(gamma))
