"""CPU: host-side logic of the product — 'same' padding, the weight packings the tcgen05 kernel
consumes (proved equivalent to the reference convolution with torch on CPU), clip sharding and the
world-size-2 gather over gloo."""
import os
import socket

import pytest
import torch
import torch.nn.functional as F

from interpreting_video_features_b200 import engine, ops, search


def test_same_pad_rule():
    # pt/models/I3D_doubled.py:77-101
    assert ops.same_pad(224, 7, 2) == (2, 3, 112)
    assert ops.same_pad(16, 7, 2) == (2, 3, 8)
    assert ops.same_pad(112, 3, 2) == (0, 1, 56)
    assert ops.same_pad(28, 3, 1) == (1, 1, 28)
    assert ops.same_pad(7, 2, 2) == (0, 1, 4)
    assert ops.same_pad(14, 2, 2) == (0, 0, 7)
    assert ops.same_pad(15, 3, 2) == (1, 1, 8)
    assert ops.same_pad(8, 1, 1) == (0, 0, 8)


def _s2d(x):
    """[n,c,t,h,w] -> [n, 8c, t/2, h/2, w/2], channel = ((a*2+b)*2+c')*C + ch (ivf.h IVF_PFMT_S2D_BF16)."""
    n, c, t, h, w = x.shape
    x = x.view(n, c, t // 2, 2, h // 2, 2, w // 2, 2)
    return x.permute(0, 3, 5, 7, 1, 2, 4, 6).reshape(n, 8 * c, t // 2, h // 2, w // 2)


def test_space_to_depth_stem_equals_strided_conv():
    torch.manual_seed(0)
    w = torch.randn(5, 3, 7, 7, 7)
    x = torch.randn(2, 3, 8, 12, 10)
    ref = F.conv3d(F.pad(x, (2, 3, 2, 3, 2, 3)), w, stride=2)
    w2 = engine.s2d_weight(w, 4)
    got = F.conv3d(F.pad(_s2d(x), (1, 2, 1, 2, 1, 2)), w2, stride=1)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


def test_flipped_conv_is_the_data_gradient():
    torch.manual_seed(1)
    w = torch.randn(6, 4, 3, 3, 3)
    x = torch.randn(1, 4, 5, 6, 7, requires_grad=True)
    y = F.conv3d(F.pad(x, (1, 1, 1, 1, 1, 1)), w)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    wd = w.flip(2, 3, 4).permute(1, 0, 2, 3, 4)  # what pack_dgrad lays out K-major
    got = F.conv3d(F.pad(gy, (1, 1, 1, 1, 1, 1)), wd)
    torch.testing.assert_close(got, gx, rtol=1e-4, atol=1e-4)


def test_space_to_depth_stem_data_gradient():
    torch.manual_seed(2)
    w = torch.randn(5, 3, 7, 7, 7)
    x = torch.randn(1, 3, 8, 12, 10, requires_grad=True)
    y = F.conv3d(F.pad(x, (2, 3, 2, 3, 2, 3)), w, stride=2)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    w2 = engine.s2d_weight(w, 4)
    wd = w2.flip(2, 3, 4).permute(1, 0, 2, 3, 4)
    got = F.conv3d(F.pad(gy, (2, 1, 2, 1, 2, 1)), wd)  # pad_front' = k-1-pf = 2, back 1
    torch.testing.assert_close(got, _s2d(gx.detach()), rtol=1e-4, atol=1e-4)


def test_shard_indices_cover_all_clips_once():
    for n in (0, 1, 7, 8, 1024):
        for w in (1, 2, 3, 8):
            got = sorted(i for r in range(w) for i in search.shard_indices(n, r, w))
            assert got == list(range(n))


def _gather_worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = search.shard_indices(n, rank, world)
    local = torch.tensor([[float(i), float(i) * 10] for i in idx]).reshape(len(idx), 2)
    full = search.gather_rows(local, idx, n, world)
    q.put((rank, full))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_gather_rows_world2_gloo(n):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    want = torch.tensor([[float(i), float(i) * 10] for i in range(n)])
    for _, full in outs:
        assert torch.equal(full, want)
