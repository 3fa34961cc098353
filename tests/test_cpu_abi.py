"""CPU: the C-ABI library loads and exports every symbol include/ivf.h declares, the ctypes
signatures cover them, and the product path fails loudly without a GPU (no compute calls here)."""
import os
import re

import pytest
import torch

from common import REPO


def header_symbols():
    src = open(os.path.join(REPO, "include", "ivf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ivf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from interpreting_video_features_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libivf.so does not export %s" % s
    assert set(syms) == set(_lib.SIGNATURES), (set(syms) ^ set(_lib.SIGNATURES))


def test_version_and_tiling_helpers():
    from interpreting_video_features_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.ivf_version()
    # K stage selection: smallest padding, ties to the wider stage
    assert [lib.ivf_conv_bf16_kchunk(c) for c in (8, 16, 24, 32, 48, 64, 96, 112, 144, 160, 192, 832)] == \
        [16, 16, 32, 32, 64, 64, 32, 64, 32, 32, 64, 64]
    assert lib.ivf_conv_bf16_cin_pad(24) == 32 and lib.ivf_conv_bf16_cin_pad(96) == 96
    assert lib.ivf_conv_bf16_ntile(174) == 176 and lib.ivf_conv_bf16_ntile(384) == 192
    assert lib.ivf_conv_bf16_cout_pad(384) == 384 and lib.ivf_conv_bf16_cout_pad(288) == 288
    assert lib.ivf_conv_bf16_cout_pad(24) == 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from interpreting_video_features_b200 import _lib
    from interpreting_video_features_b200.pt import mask
    with pytest.raises(_lib.IvfError):
        _lib.handle()
    with pytest.raises(_lib.IvfError):
        mask.perturb_sequence(torch.zeros(1, 3, 4, 2, 2), torch.zeros(4))
    with pytest.raises(_lib.IvfError):
        mask.calc_tv_norm(torch.rand(8))
    # ivf_create itself reports the missing device instead of falling back
    import ctypes as C
    out = C.c_void_p()
    rc = _lib.load().ivf_create(0, C.byref(out))
    assert rc != 0 and b"no CPU fallback" in _lib.load().ivf_last_error()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(REPO, "interpreting_video_features_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(root, f)
