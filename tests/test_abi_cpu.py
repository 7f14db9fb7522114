"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, and the
ctypes struct mirrors have the sizes the C compiler computes."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "edgestyle_b200.h")


@pytest.fixture(scope="module")
def lib():
    from edgestyle_b200 import build, ext

    build.build()
    return ext.load()


def test_header_symbols_exported(lib):
    from edgestyle_b200 import ext

    src = open(HEADER).read()
    declared = set(re.findall(r"\b(es_[a-z0-9_]+)\s*\(", src))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in ext.EXPORTS, f"{name} has no ctypes signature in ext.py"
    assert lib.es_abi_version() == ext.ABI_VERSION == 2


def test_struct_sizes_match_c(tmp_path, lib):
    from edgestyle_b200 import ext
    import ctypes

    prog = tmp_path / "sz.c"
    prog.write_text(
        '#include <stdio.h>\n#include "edgestyle_b200.h"\n'
        'int main(){printf("%zu %zu %zu %zu\\n", sizeof(EsGemm), sizeof(EsAttention), sizeof(EsGroupNorm), sizeof(EsMergeBatch));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    want = [ctypes.sizeof(ext.EsGemm), ctypes.sizeof(ext.EsAttention), ctypes.sizeof(ext.EsGroupNorm),
            ctypes.sizeof(ext.EsMergeBatch)]
    assert [int(x) for x in out] == want


def test_bad_arguments_fail_loudly_without_gpu(lib):
    # argument validation happens before any CUDA call: a null descriptor is an error, not a crash
    assert lib.es_gemm(None, None) != 0
    assert b"null" in lib.es_last_error()


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from edgestyle_b200 import ext, ops

    a = torch.zeros(128, 64, dtype=torch.float16)
    with pytest.raises(ext.EdgeStyleNativeError):
        ops.gemm(a, a, 64, out=a.clone())
