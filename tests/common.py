"""Shared helpers for the tests (state dicts, tolerances)."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLD = os.path.join(REPO, "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def i3d_state_dict(num_classes=174, kth=False, seed=0):
    """Seeded reference-keyed state dict (identical to the reference constructor's under the same
    seed: checked against /root/reference in tests/test_cpu_models.py)."""
    from interpreting_video_features_b200.pt.models import I3D_doubled, I3D_doubled_kth
    torch.manual_seed(seed)
    if kth:
        m = I3D_doubled_kth.Model(num_classes, last_stride=1, stride_mod_layers="", softMax=1, finalTimeLength=4)
    else:
        m = I3D_doubled.Model(num_classes, last_stride=1, stride_mod_layers="", softMax=1)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}, m.eval()


def rel_err(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def max_rel(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor + 1e-300)))


def assert_close_nan(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), what + ": NaN pattern differs"
    ok = ~np.isnan(a)
    np.testing.assert_allclose(a[ok], b[ok], rtol=rtol, atol=atol, err_msg=what)
