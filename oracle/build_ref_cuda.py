#!/usr/bin/env python
"""TEST INFRASTRUCTURE: builds the reference's own DCNv2 CUDA kernels for sm_100a.

    python oracle/build_ref_cuda.py        ->  oracle/_ref/DCN_ref.so   (git-ignored, travels via gpurun)

"Patched build" (SURVEY 8c): the reference translation unit
`/root/reference/src/model/deformconv/src/cuda/modulated_deform_conv_cuda.cu` (+ its
`modulated_deform_im2col_cuda.cuh`) does not compile against torch 2.11 as shipped, because
`AT_DISPATCH_FLOATING_TYPES(input.type(), ...)` (lines 93 and 224) no longer accepts a
`DeprecatedTypeProperties`.  This recipe compiles the sources *where they lie*: it writes a
two-line wrapper into a temporary directory that `#define`s nothing and `#include`s the reference
file after a `sed`-equivalent one-token substitution (`input.type()` -> `input.scalar_type()` at the
two dispatch sites, `x.type().is_cuda()` -> `x.is_cuda()` in the asserts) applied to a temporary
copy -- no reference source is written into this repository, only the compiled module.
The reference's own build system (setup.py: needs a visible GPU, builds DCNv1 and PS-RoI too) is
not run.  If /root/reference is absent (the GPU box) this script does nothing; the prebuilt
`oracle/_ref/DCN_ref.so` is used when present and the tests that need it skip otherwise.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/model/deformconv/src"
OUT = os.path.join(HERE, "_ref", "DCN_ref.so")


def build(force: bool = False) -> str | None:
    cu = os.path.join(REF_SRC, "cuda", "modulated_deform_conv_cuda.cu")
    cuh = os.path.join(REF_SRC, "cuda", "modulated_deform_im2col_cuda.cuh")
    if not (os.path.exists(cu) and os.path.exists(cuh)):
        return OUT if os.path.exists(OUT) else None
    binding = os.path.join(HERE, "ref_dcn_binding.cpp")
    newest = max(os.path.getmtime(p) for p in (cu, cuh, binding, __file__))
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    import torch
    from torch.utils.cpp_extension import include_paths, library_paths
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="dcn_ref_")
    try:
        os.makedirs(os.path.join(tmp, "cuda"))
        with open(cu) as f:
            text = f.read()
        text = text.replace("AT_DISPATCH_FLOATING_TYPES(input.type(),", "AT_DISPATCH_FLOATING_TYPES(input.scalar_type(),")
        text = re.sub(r"(\w+)\.type\(\)\.is_cuda\(\)", r"\1.is_cuda()", text)
        with open(os.path.join(tmp, "cuda", "modulated_deform_conv_cuda.cu"), "w") as f:
            f.write(text)
        # the .cuh is used unmodified, from where it lies
        os.symlink(cuh, os.path.join(tmp, "cuda", "modulated_deform_im2col_cuda.cuh"))
        inc = ["-I" + tmp, "-I" + sysconfig.get_paths()["include"]]
        for p in include_paths("cuda") if "device_type" in include_paths.__code__.co_varnames else include_paths(True):
            inc.append("-I" + p)
        libs = []
        for p in library_paths("cuda") if "device_type" in library_paths.__code__.co_varnames else library_paths(True):
            libs += ["-L" + p, "-Xlinker", "-rpath=" + p]
        abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
        cmd = ["nvcc", "-shared", "-O2", "-std=c++17", "-Xcompiler", "-fPIC",
               "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-DTORCH_EXTENSION_NAME=DCN_ref", "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi,
               "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
               "-D__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr", "-w",
               *inc, os.path.join(tmp, "cuda", "modulated_deform_conv_cuda.cu"), binding,
               *libs, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
               "-o", OUT]
        subprocess.run(cmd, check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "reference sources absent and no prebuilt oracle/_ref/DCN_ref.so")
