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
    # K stage: one 128-byte row (64 channels) unless the operand is narrower; tap pitch padded to 16
    assert [lib.ivf_conv_bf16_kchunk(c) for c in (8, 16, 24, 32, 48, 64, 96, 112, 144, 160, 192, 832)] == \
        [16, 16, 32, 32, 64, 64, 64, 64, 64, 64, 64, 64]
    assert lib.ivf_conv_bf16_cin_pad(24) == 32 and lib.ivf_conv_bf16_cin_pad(96) == 96
    assert lib.ivf_conv_bf16_cin_pad(8) == 16 and lib.ivf_conv_bf16_cin_pad(112) == 112
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


def test_slab_plan_diagnostic_needs_no_gpu():
    """ivf_conv_slab_plan: the wide stride-1 layers of C2 get the halo-slab kernel, the 14x14 stage and
    every 1x1x1 convolution the im2col kernel; the plan fits shared memory and TMEM."""
    import ctypes as C
    from interpreting_video_features_b200 import _lib
    lib = _lib.load()

    def plan(dhw, cin, cout, k, pf, flags=0):
        d = _lib.ConvDesc()
        d.n = 8
        d.id, d.ih, d.iw = dhw
        d.od, d.oh, d.ow = dhw
        d.cin, d.cout = cin, cout
        d.kd, d.kh, d.kw = k
        d.sd = d.sh = d.sw = 1
        d.pd, d.ph, d.pw = pf
        d.in_ld, d.out_ld = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
        d.dtype, d.flags = _lib.IVF_BF16, flags
        out = (C.c_int * 12)()
        return lib.ivf_conv_slab_plan(C.byref(d), 148, out), list(out)

    for args in [((8, 112, 112), 24, 64, (4, 4, 4), (1, 1, 1)), ((8, 112, 112), 64, 24, (4, 4, 4), (2, 2, 2), 8),
                 ((8, 56, 56), 64, 192, (3, 3, 3), (1, 1, 1)), ((8, 28, 28), 96, 128, (3, 3, 3), (1, 1, 1)),
                 ((8, 28, 28), 32, 16, (3, 3, 3), (1, 1, 1), 12)]:
        ok, (kch, bn, ntiles, mt, th, acc, a_st, b_st, tiles, smem, kwm, ncta) = plan(*args)
        assert ok == 1, args
        assert kch in (32, 64) and bn % 16 == 0 and bn * ntiles >= args[2]
        assert 1 <= mt <= 4 and acc in (1, 2) and a_st >= 2 and b_st >= 2 and tiles > 0
        assert kwm >= 1 and args[3][2] % kwm == 0 and kwm * bn <= 256
        assert acc * mt * ((kwm * bn + 31) // 32 * 32) <= 512          # TMEM columns
        assert smem <= 216 * 1024                                       # dynamic shared memory budget (incl. kw-merge exchange)
        assert th * (args[0][2] + args[3][2] - 1) <= mt * 128           # the tile's padded-width pixels fit
    assert plan((4, 14, 14), 96, 208, (3, 3, 3), (1, 1, 1))[0] == 1      # 14x14 stage
    assert plan((8, 56, 56), 64, 64, (1, 1, 1), (0, 0, 0))[0] == 0       # 1x1x1 -> im2col kernel (slab route opt-in)
    assert plan((2, 7, 7), 256, 832, (3, 3, 3), (1, 1, 1), 12)[0] == 1   # wide data gradient: 4+ N tiles
    assert plan((2, 4, 4), 64, 64, (3, 3, 3), (1, 1, 1))[0] == 0         # map narrower than 7 -> im2col kernel


def test_depth_stacking_is_a_soft_plan_request():
    """ivf_conv_desc.plan_ds = 2 (conv_slab.cu depth stacking): honoured where the layer can stack two output depths
    (even depth, front padding >= 1, N of whole 32-column TMEM blocks, 2 N <= 256), otherwise the layer silently
    runs unstacked - and the default is unstacked."""
    import ctypes as C
    from interpreting_video_features_b200 import _lib
    lib = _lib.load()

    def ds_of(dhw, cin, cout, k, pf, want):
        d = _lib.ConvDesc()
        d.n, (d.id, d.ih, d.iw), (d.od, d.oh, d.ow) = 8, dhw, dhw
        d.cin, d.cout, (d.kd, d.kh, d.kw), (d.pd, d.ph, d.pw) = cin, cout, k, pf
        d.sd = d.sh = d.sw = 1
        d.in_ld, d.out_ld, d.dtype, d.flags, d.plan_ds = max(cin, 32), cout, _lib.IVF_BF16, 3, want
        return lib.ivf_conv_slab_plan_ds(C.byref(d), 148)

    stem = ((8, 112, 112), 24, 64, (4, 4, 4), (1, 1, 1))
    assert ds_of(*stem, 2) == 2 and ds_of(*stem, 0) == 1 and ds_of(*stem, 1) == 1
    assert ds_of((8, 56, 56), 192, 64, (3, 3, 3), (1, 1, 1), 2) == 2          # Conv3d_2c data gradient
    assert ds_of((7, 56, 56), 192, 64, (3, 3, 3), (1, 1, 1), 2) == 1          # odd depth
    assert ds_of((8, 56, 56), 64, 192, (3, 3, 3), (1, 1, 1), 2) in (1, 2)     # N = 192: only as two N tiles of 96
    assert ds_of((8, 28, 28), 32, 16, (3, 3, 3), (1, 1, 1), 2) == 1           # bn = 16 is not a whole TMEM block
    assert ds_of((8, 30, 40), 32, 128, (1, 5, 5), (0, 2, 2), 2) == 1          # 2-D kernel: no depth taps to share


def test_shipped_tile_plans_are_accepted_by_the_kernel():
    """plans_sm100.json (measured on a B200, tools/write_plans.py): every shape key parses into the descriptor
    fields tune.py hashes, and every stored request is a plan ivf_conv_slab_plan still accepts for that shape -
    a table that drifted from the kernel's constraints would silently fall back to the cost model."""
    import ctypes as C
    import json
    import os
    from interpreting_video_features_b200 import _lib, tune
    path = os.path.join(os.path.dirname(tune.__file__), "plans_sm100.json")
    table = json.load(open(path))
    assert table["fields"] == list(tune._FIELDS)
    lib = _lib.load()
    n_req = 0
    for key, req in table["plans"].items():
        vals = [int(v) for v in key.split(",")]
        assert len(vals) == len(tune._FIELDS)
        d = _lib.ConvDesc()
        for f, v in zip(tune._FIELDS, vals):
            setattr(d, f, v)
        d.sd = d.sh = d.sw = 1
        assert tune.shape_key(d) == key
        if req is None:
            continue
        n_req += 1
        d.plan_kwm, d.plan_mt, d.plan_acc, d.plan_ncta, d.plan_ntiles = req[:5]
        d.plan_ds = req[5] if len(req) > 5 else 0
        out = (C.c_int * 12)()
        assert lib.ivf_conv_slab_plan(C.byref(d), 148, out) == 1, key
        got = (out[10], out[3], out[5], out[11], out[2])  # kwm, mt, acc, ncta, ntiles
        assert got == tuple(req[:5]), (key, req, got)
        if len(req) > 5:  # depth stacking (conv_slab.cu) as requested
            assert lib.ivf_conv_slab_plan_ds(C.byref(d), 148) == req[5], (key, req)
    assert n_req >= 1
