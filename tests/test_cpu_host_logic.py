"""CPU: host-side logic of the product — 'same' padding, the weight packings the tcgen05 kernel
consumes (proved equivalent to the reference convolution with torch on CPU), clip sharding and the
world-size-2 gather over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import packing_ref
from interpreting_video_features_b200 import ops, search


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
    w2 = packing_ref.s2d_weight(w, 4)
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
    w2 = packing_ref.s2d_weight(w, 4)
    wd = w2.flip(2, 3, 4).permute(1, 0, 2, 3, 4)
    got = F.conv3d(F.pad(gy, (2, 1, 2, 1, 2, 1)), wd)  # pad_front' = k-1-pf = 2, back 1
    torch.testing.assert_close(got, _s2d(gx.detach()), rtol=1e-4, atol=1e-4)


def test_pack_kernel_index_math_equals_torch_packers():
    """csrc/pack.cu's mapping (restated in numpy, packing_ref.emulate_pack_kernel) against the torch packers that
    the tests above prove equivalent to the reference convolutions: forward, data gradient, 3-D space-to-depth
    (the stem), 2-D space-to-depth with padded gate stacking (ConvLSTM), two data-gradient sources, fp32 layouts."""
    g = torch.Generator().manual_seed(3)
    w = torch.randn((5, 3, 3, 3, 3), generator=g)
    for dgrad, ref in ((0, packing_ref.pack_fwd), (1, packing_ref.pack_dgrad)):
        want = ref(w, "bf16", 16, 16).float().numpy()
        got = packing_ref.emulate_pack_kernel(w, dgrad, 0, (1, 1, 1), 0, 16, 16, 0, 0, np.zeros((16, 27, 16), np.float32))
        assert np.array_equal(torch.from_numpy(got).bfloat16().float().numpy(), want), dgrad
    # fp32 tap-major: forward [taps*ci, co], transposed-gather data gradient [taps*co, ci] (taps NOT flipped)
    got = packing_ref.emulate_pack_kernel(w, 0, 1, (1, 1, 1), 0, 5, 3, 0, 0, np.zeros((27, 3, 5), np.float32))
    assert np.array_equal(got.reshape(-1, 5), packing_ref.pack_fwd(w, "fp32").numpy())
    got = packing_ref.emulate_pack_kernel(w, 2, 1, (1, 1, 1), 0, 3, 5, 0, 0, np.zeros((27, 5, 3), np.float32))
    assert np.array_equal(got.reshape(-1, 3), packing_ref.pack_dgrad(w, "fp32").numpy())
    # the stem: 7x7x7 stride 2 -> 4x4x4 over 8*3 channels
    ws = torch.randn((4, 3, 7, 7, 7), generator=g)
    w2 = packing_ref.s2d_weight(ws, 4)
    for dgrad, ref, shp in ((0, packing_ref.pack_fwd, (16, 64, 32)), (1, packing_ref.pack_dgrad, (32, 64, 16))):
        want = ref(w2, "bf16", shp[0], shp[2]).float().numpy()
        got = packing_ref.emulate_pack_kernel(ws, dgrad, 0, (2, 2, 2), 0, shp[0], shp[2], 0, 0, np.zeros(shp, np.float32))
        assert np.array_equal(torch.from_numpy(got).bfloat16().float().numpy(), want), ("stem", dgrad)
    # ConvLSTM layer 1 x-convolution: four gates of hid 4 padded to he 8, input 4 -> 8 channels, 5x5 stride 2
    gates = [torch.randn((4, 4, 5, 5), generator=g) for _ in range(4)]
    wx = packing_ref.pad_gates(gates, 8, 8)
    w2 = packing_ref.s2d_weight_2d(wx[:, :, 0], 3).unsqueeze(2)  # [32, 32, 1, 3, 3]
    for dgrad, ref in ((0, packing_ref.pack_fwd), (1, packing_ref.pack_dgrad)):
        want = ref(w2, "bf16", 32, 32).float().numpy()
        got = np.zeros((32, 9, 32), np.float32)
        for gi, wg in enumerate(gates):
            off = (0, gi * 8) if dgrad else (gi * 8, 0)
            packing_ref.emulate_pack_kernel(wg.unsqueeze(2), dgrad, 0, (1, 2, 2), 8, 32, 32, off[0], off[1], got)
        assert np.array_equal(torch.from_numpy(got).bfloat16().float().numpy(), want), ("gates", dgrad)
    # two data-gradient sources (b0 | b1a,b2a): first K block padded to 64
    w0, w12 = torch.randn((24, 16, 1, 1, 1), generator=g), torch.randn((40, 16, 1, 1, 1), generator=g)
    want = packing_ref.pack_dgrad_two_sources(w0, w12, 16, 112).float().numpy()
    got = np.zeros((16, 1, 112), np.float32)
    packing_ref.emulate_pack_kernel(w0, 1, 0, (1, 1, 1), 0, 16, 112, 0, 0, got)
    packing_ref.emulate_pack_kernel(w12, 1, 0, (1, 1, 1), 0, 16, 112, 0, 64, got)
    assert np.array_equal(torch.from_numpy(got).bfloat16().float().numpy(), want)


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
    q.put((rank, full.numpy().copy()))  # by value: a shared-memory tensor's fd hand-over races with process exit
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
        assert torch.equal(torch.from_numpy(full), want)


def _assemble_worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, ncls, cam = 4, 3, [2, 1, 2]
    idx = search.shard_indices(n, rank, world)
    width = t + 2 + ncls + 4
    local = torch.stack([torch.arange(width, dtype=torch.float32) + 100.0 * i for i in idx]) if idx else \
        torch.zeros((0, width))
    stats = {}
    out = search.assemble_results(local, idx, n, world, t, ncls, cam, stats=stats)
    # by value (numpy), not as shared-memory tensors: the fd hand-over of a tensor races with this process's exit
    q.put((rank, {k: (v if k == "indices" else v.numpy().copy()) for k, v in out.items()}, stats))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 1])
def test_sharded_result_rows_world2_gloo(n):
    """The end of the sharded job (search.assemble_results): masks, scores, probabilities and low-resolution CAMs of
    two ranks gathered into clip order - ragged shards (5 clips over 2 ranks) and a rank with no clip at all."""
    import torch.multiprocessing as mp
    s_ = socket.socket()
    s_.bind(("127.0.0.1", 0))
    port = s_.getsockname()[1]
    s_.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_assemble_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    want = torch.stack([torch.arange(13, dtype=torch.float32) + 100.0 * i for i in range(n)])
    for _, out, stats in outs:
        out = {k: (v if k == "indices" else torch.from_numpy(v)) for k, v in out.items()}
        assert out["indices"] == list(range(n))
        assert torch.equal(out["time_mask"], want[:, :4]) and torch.equal(out["freeze_score"], want[:, 4])
        assert torch.equal(out["reverse_score"], want[:, 5]) and torch.equal(out["probs_orig"], want[:, 6:9])
        assert torch.equal(out["cam_lowres"], want[:, 9:].reshape(n, 2, 1, 2))
        assert stats["gathered_bytes"] == n * 13 * 4


def test_optimizer_chunk_table_covers_every_element_once():
    """ops.optim_table (the row table of ivf_optim_step_multi: one thread block per row): every element of every
    tensor in exactly one row of at most OPTIM_CHUNK elements, pointers advanced by whole elements, absent state
    buffers as NULL."""
    import torch
    from interpreting_video_features_b200 import ops
    sizes = [1, 4095, 4096, 4097, 3 * 4096 + 5]
    ps = [torch.zeros(n) for n in sizes]
    gs = [torch.zeros(n) for n in sizes]
    s1 = [torch.zeros(n) for n in sizes]
    table = ops.optim_table(ps, gs, s1, [None] * len(sizes), "cpu")
    assert table.dtype == torch.int64 and table.shape[1] == 5
    rows = table.tolist()
    assert len(rows) == sum((n + ops.OPTIM_CHUNK - 1) // ops.OPTIM_CHUNK for n in sizes)
    i = 0
    for p, g, a in zip(ps, gs, s1):
        covered = 0
        while covered < p.numel():
            pp, gp, ap, bp, n = rows[i]
            assert 0 < n <= ops.OPTIM_CHUNK and bp == 0
            assert pp == p.data_ptr() + 4 * covered and gp == g.data_ptr() + 4 * covered and ap == a.data_ptr() + 4 * covered
            covered += n
            i += 1
        assert covered == p.numel()
    assert i == len(rows)
