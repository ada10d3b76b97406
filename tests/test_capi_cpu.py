"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/mmgl_b200.h declares, the ctypes signatures cover exactly that set, and the product path refuses to
run without CUDA (no CPU fallback, no route through oracle/)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mmgl_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmgl_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from mmgl_b200 import build
    return build.build()


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for must in ("mmgl_gemm_bf16", "mmgl_xattn_fwd", "mmgl_xattn_bwd", "mmgl_layernorm_fwd", "mmgl_layernorm_bwd",
                 "mmgl_bank_pack_fwd", "mmgl_bank_pack_bwd", "mmgl_gcn_concat_fwd", "mmgl_gcn_combine_bwd",
                 "mmgl_attn_fwd", "mmgl_attn_bwd", "mmgl_attn_bwd_workspace_bytes", "mmgl_rmsnorm_fwd", "mmgl_rmsnorm_bwd",
                 "mmgl_ce_fwd", "mmgl_ce_bwd", "mmgl_last_error_string", "mmgl_version", "mmgl_launch_count"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(built_lib)
    for name in _declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/mmgl_b200.h but not exported"
    assert handle.mmgl_version() == 1


def test_ctypes_signatures_match_header(built_lib):
    from mmgl_b200 import _capi
    assert sorted(_capi.SIGNATURES) == _declared_symbols()
    lib = _capi.lib()
    assert lib.mmgl_launch_count() >= 0
    assert isinstance(_capi.last_error(), str)


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors == sizeof of the C structs (compiled from the header with gcc)."""
    import subprocess
    import tempfile
    from mmgl_b200 import _capi
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "mmgl_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", '
                             'sizeof(mmgl_gemm_args), sizeof(mmgl_bank_args), sizeof(mmgl_bank_bwd_args), '
                             'sizeof(mmgl_attn_args));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_capi.GemmArgs), ctypes.sizeof(_capi.BankArgs), ctypes.sizeof(_capi.BankBwdArgs),
                     ctypes.sizeof(_capi.AttnArgs)]


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mmgl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle/"


def test_ops_fail_loudly_without_cuda():
    from mmgl_b200 import ops
    x = torch.zeros(2, 8, 64, dtype=torch.bfloat16)
    w = torch.zeros(64, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(x, w)
