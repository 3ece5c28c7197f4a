"""CPU: the C-ABI library loads, exports every symbol include/nlspn_b200.h declares, and
rejects bad arguments with negative status codes BEFORE touching CUDA (no compute here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "nlspn_b200.h")).read()
    return sorted(set(re.findall(r"NLSPN_API[^;(]*?\b(nlspn_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from nlspn_eccv20_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ["nlspn_abi_version", "nlspn_last_error", "nlspn_prologue_fwd", "nlspn_propagate_fwd",
              "nlspn_backward", "nlspn_backward_workspace_bytes", "nlspn_dcn_forward",
              "nlspn_dcn_backward", "nlspn_dcn_forward_f64", "nlspn_dcn_backward_f64", "nlspn_debug_indices",
              "nlspn_device_info"]:
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    from nlspn_eccv20_b200 import _lib
    raw = ctypes.CDLL(_lib.lib_path())
    for s in declared_symbols():
        assert hasattr(raw, s), "libnlspn_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "python binding misses %s" % s
    assert lib.nlspn_abi_version() == 1


def test_library_has_sm100a_code_and_no_torch_dependency():
    import subprocess
    from nlspn_eccv20_b200 import _lib
    out = subprocess.run(["ldd", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        o = subprocess.run([cuobjdump, "-lelf", _lib.lib_path()], capture_output=True, text=True).stdout
        assert "sm_100a" in o


def test_validation_errors_are_negative_and_described(lib):
    P = ctypes.c_void_p
    one = P(16)  # never dereferenced: validation fails first
    # bad kernel size
    rc = lib.nlspn_prologue_fwd(one, None, one, None, one, 3, 0, 1, 4, 4, 4, one, one, None, one, None)
    assert rc == -3 and b"prop_kernel" in lib.nlspn_last_error()
    # missing required pointer
    rc = lib.nlspn_prologue_fwd(None, None, one, None, one, 3, 0, 1, 4, 4, 3, one, one, None, one, None)
    assert rc == -1
    # PRESERVE_INPUT without feat_fix
    rc = lib.nlspn_prologue_fwd(one, None, one, None, one, 3, 1, 1, 4, 4, 3, one, one, None, one, None)
    assert rc == -1 and b"feat_fix" in lib.nlspn_last_error()
    # bad affinity / bad shape
    assert lib.nlspn_prologue_fwd(one, None, one, None, one, 9, 0, 1, 4, 4, 3, one, one, None, one, None) == -6
    assert lib.nlspn_prologue_fwd(one, None, one, None, one, 3, 0, 0, 4, 4, 3, one, one, None, one, None) == -2
    # confidence needs >= 2 src planes
    assert lib.nlspn_propagate_fwd(one, one, one, None, 0, 1, 4, 4, 3, 5, one, 1, one, None) == -2
    # DCN outside the NLSPN domain: C=2, stride 2, wrong padding
    ints_ok = [3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64]
    def dcn(ints, C=1):
        return lib.nlspn_dcn_forward(one, one, one, one, one, *ints, 1, C, 4, 4, one, None)
    assert dcn(ints_ok, C=2) == -4
    assert dcn([3, 3, 2, 2, 1, 1, 1, 1, 1, 1, 64]) == -4
    assert dcn([3, 3, 1, 1, 0, 0, 1, 1, 1, 1, 64]) == -4
    assert dcn([3, 3, 1, 1, 1, 1, 1, 1, 2, 1, 64]) == -4
    assert lib.nlspn_backward_workspace_bytes(2, 8, 8, 3, 1) >= 4 * (3 + 9) * 2 * 64
    assert lib.nlspn_backward_workspace_bytes(2, 8, 8, 3, 40) >= 4 * 40 * 2 * 64
    assert lib.nlspn_backward_workspace_bytes(0, 8, 8, 3, 1) == 0
    # the flags-aware query: flags 0 = the plain query; DETERMINISTIC holds the exact CSR (8 B per corner of every tap)
    from nlspn_eccv20_b200 import _lib
    assert lib.nlspn_backward_workspace_bytes_ex(2, 8, 8, 3, 4, 0) == lib.nlspn_backward_workspace_bytes(2, 8, 8, 3, 4)
    assert lib.nlspn_backward_workspace_bytes_ex(2, 8, 8, 5, 4, _lib.FLAG_DETERMINISTIC) >= 8 * 4 * 24 * 2 * 64


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no file of the product package may reference it."""
    pkg = os.path.join(ROOT, "nlspn_eccv20_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from nlspn_eccv20_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="REQUIRED"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch
    from nlspn_eccv20_b200 import NLSPN
    m = NLSPN(prop_kernel=3, prop_time=2)
    x = torch.zeros(1, 1, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, torch.zeros(1, 24, 4, 4), x, x)
    assert list(m.state_dict().keys()) == ["aff_scale_const", "w", "b", "w_conf"]
    assert float(m.aff_scale_const) == 4.0 and m.aff_scale_const.requires_grad
    assert not m.w.requires_grad and tuple(m.w.shape) == (1, 1, 3, 3)


def test_only_test_infrastructure_touches_the_oracle():
    """tests/, __graft_entry__.py (build + smoke) and bench.py's CPU-baseline legs are the only places that
    may import or execute oracle/ (the product, tools/ and everything else must not)."""
    allowed = {os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")}
    bad = []
    for dp, dns, fs in os.walk(ROOT):
        dns[:] = [d for d in dns if d not in (".git", "tests", "oracle", "gpurun_out", "__pycache__", "baseline", ".pytest_cache")]
        for f in fs:
            p = os.path.join(dp, f)
            if f.endswith((".py", ".sh", ".cu", ".cuh", ".h")) and p not in allowed:
                txt = open(p, errors="ignore").read()
                if "import oracle" in txt or "from oracle" in txt or "liboracle" in txt or "oracle/_ref" in txt:
                    bad.append(os.path.relpath(p, ROOT))
    assert not bad, bad


def test_heads_entries_validate_before_touching_cuda(lib):
    """The head GEMM's C entries: sizes, the supported-domain query and the argument checks are
    host-only; the tcgen05 kernel itself is covered by tests/test_gpu_heads.py."""
    from nlspn_eccv20_b200 import _lib
    P = ctypes.c_void_p
    one = P(16)
    assert lib.nlspn_heads_packed_floats(4) == 0                       # prop_kernel must be 3, 5 or 7
    f3, f5, f7 = (lib.nlspn_heads_packed_floats(k) for k in (3, 5, 7))
    assert 0 < f3 < f5 and f7 > 0 and f3 % 4 == 0 and f5 % 4 == 0       # 16-byte granules
    assert lib.nlspn_heads_prologue_supported(1216, 3) == 1 and lib.nlspn_heads_prologue_supported(304, 5) == 1
    assert lib.nlspn_heads_prologue_supported(1218, 3) == 0            # W % 4 != 0: nine-tap form, no fused prologue
    assert lib.nlspn_heads_prologue_supported(1216, 7) == 0            # N = 448 accumulator columns do not fit one MMA
    with _lib.options(heads_rows=0):
        assert lib.nlspn_heads_prologue_supported(1216, 3) == 0
    # NULL output / NULL gamma
    rc = lib.nlspn_heads_prologue_fwd(one, one, one, one, one, one, None, None, 3, 0, 1, 8, 16, 3,
                                      one, one, None, one, one, None, one, None)
    assert rc == -1 and b"NULL" in lib.nlspn_last_error()
    # PRESERVE_INPUT without feat_fix
    rc = lib.nlspn_heads_prologue_fwd(one, one, one, one, one, one, None, one, 3, _lib.FLAG_PRESERVE_INPUT, 1, 8, 16, 3,
                                      one, one, None, one, one, None, one, None)
    assert rc == -1 and b"feat_fix" in lib.nlspn_last_error()
    # a flag the fused epilogue does not implement
    rc = lib.nlspn_heads_prologue_fwd(one, one, one, one, one, one, None, one, 3, _lib.FLAG_CONF_SAMPLED, 1, 8, 16, 3,
                                      one, one, None, one, one, None, one, None)
    assert rc == -4 and b"PRESERVE_INPUT and ALWAYS_CLIP" in lib.nlspn_last_error()
    # unknown affinity mode, bad kernel size
    rc = lib.nlspn_heads_prologue_fwd(one, one, one, one, one, one, None, one, 9, 0, 1, 8, 16, 3,
                                      one, one, None, one, one, None, one, None)
    assert rc == -6
    rc = lib.nlspn_heads_prologue_fwd(one, one, one, one, one, one, None, one, 3, 0, 1, 8, 16, 4,
                                      one, one, None, one, one, None, one, None)
    assert rc == -3
    # a width the fused form does not take: refused with a pointer to the two-call path
    rc = lib.nlspn_heads_prologue_fwd(one, one, one, one, one, one, None, one, 3, 0, 1, 8, 18, 3,
                                      one, one, None, one, one, None, one, None)
    assert rc == -2 and b"nlspn_heads_fwd + nlspn_prologue_fwd" in lib.nlspn_last_error()
    rc = lib.nlspn_heads_fwd(one, one, one, None, one, one, 1, 8, 16, 3, one, one, one, None)
    assert rc == -1


def test_heads_gradient_entries_validate_before_touching_cuda(lib):
    """The heads' backward entries (nlspn_heads_grad_prep / _wgrad / _dgrad_one / _dgrad_pack / _dgrad_wide): sizes, the
    supported-domain queries and the argument checks are host-only; the kernels are covered by tests/test_gpu_heads.py."""
    P = ctypes.c_void_p
    one = P(16)
    assert lib.nlspn_heads_wgrad_supported(1216, 3) == 1 and lib.nlspn_heads_wgrad_supported(304, 5) == 1
    assert lib.nlspn_heads_wgrad_supported(1218, 3) == 0 and lib.nlspn_heads_wgrad_supported(1216, 4) == 0
    assert lib.nlspn_heads_dgrad_supported(1216, 3) == 1 and lib.nlspn_heads_dgrad_supported(1216, 5) == 0
    assert lib.nlspn_heads_dgrad_packed_floats(3) == 9 * 4 * 128 * 8 and lib.nlspn_heads_dgrad_packed_floats(5) == 0
    # widths the TMA path does not take; bad prop_kernel; NULL and misaligned pointers
    assert lib.nlspn_heads_grad_prep(one, one, one, one, one, 1, 8, 18, 3, one, one, None) == -2
    assert b"W % 4" in lib.nlspn_last_error()
    assert lib.nlspn_heads_grad_prep(one, one, one, one, one, 1, 8, 16, 4, one, one, None) == -3
    assert lib.nlspn_heads_grad_prep(None, None, None, None, one, 1, 8, 16, 3, one, None, None) == -1
    assert lib.nlspn_heads_grad_prep(None, one, None, None, one, 1, 8, 16, 3, P(20), None, None) == -7
    assert lib.nlspn_heads_wgrad(one, one, one, None, one, 1, 8, 16, 3, one, None) == -1
    assert lib.nlspn_heads_wgrad(one, P(24), one, one, one, 1, 8, 16, 3, one, None) == -7
    assert lib.nlspn_heads_dgrad_one(None, one, one, 1, 8, 16, 3, one, one, None) == -1
    assert lib.nlspn_heads_dgrad_one(one, None, one, 1, 8, 16, 3, one, None, None) == -1          # d_id_fd1 wanted without w_id
    assert lib.nlspn_heads_dgrad_one(one, None, None, 1, 8, 16, 3, None, None, None) == 0         # nothing to do
    assert lib.nlspn_heads_dgrad_pack(one, one, one, 5, one, None) == -3
    assert lib.nlspn_heads_dgrad_pack(one, None, one, 3, one, None) == -1
    assert lib.nlspn_heads_dgrad_wide(one, one, 1, 8, 16, 5, one, one, None) == -3
    assert lib.nlspn_heads_dgrad_wide(one, one, 1, 8, 16, 3, None, None, None) == -1              # d_fe1 is required
    assert lib.nlspn_heads_dgrad_wide(one, one, 1, 8, 20, 3, None, P(40), None) == -7
