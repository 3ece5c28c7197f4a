"""CPU: the 16-bit depth PNG / .npy formats downstream of the path."""
import numpy as np
import pytest


def test_depth_png16_round_trip(tmp_path):
    from nlspn_eccv20_b200.io import read_depth_png16, write_depth_png16
    rng = np.random.default_rng(0)
    d = rng.uniform(0, 90, size=(37, 53)).astype(np.float32)
    d[0, :5] = -3.0                              # negative predictions clamp to 0 (nlspnsummary.py:178)
    d[1, :5] = 0.0
    p = tmp_path / "d.png"
    write_depth_png16(p, d)
    back = read_depth_png16(p)
    want = (np.clip(d, 0, None) * 256.0).astype(np.uint16).astype(np.float32) / 256.0   # truncation
    assert np.array_equal(back, want)
    assert float(np.abs(back - np.clip(d, 0, None)).max()) < 1.0 / 256.0
    from PIL import Image
    assert np.array(Image.open(p)).dtype == np.uint16
    with pytest.raises(ValueError):
        write_depth_png16(tmp_path / "x.png", np.full((2, 2), 300.0))
    with pytest.raises(ValueError):                # an 8-bit-looking file is rejected (kittidc.py:78-79)
        Image.fromarray(np.full((2, 2), 200, np.uint16)).save(tmp_path / "low.png")
        read_depth_png16(tmp_path / "low.png")


def test_geometry_dump(tmp_path):
    from nlspn_eccv20_b200.io import dump_geometry
    off, aff = np.zeros((2, 18, 3, 4), np.float32), np.ones((2, 9, 3, 4), np.float32)
    dump_geometry(tmp_path, off, aff, np.array([4.0], np.float32))
    assert np.load(tmp_path / "offset.npy").shape == (18, 3, 4)
    assert np.load(tmp_path / "aff.npy").shape == (9, 3, 4)
    assert float(np.load(tmp_path / "gamma.npy")[0]) == 4.0
    dump_geometry(tmp_path / "fixed", None, aff, np.array([1.0], np.float32))
    assert not (tmp_path / "fixed" / "offset.npy").exists()
