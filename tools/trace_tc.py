"""Phase timeline (globaltimer, ns) of CTA (0,0) of the im2col/1x1 tcgen05 kernel for a few 1x1x1 shapes.
GPU only:  IVF_TC_TRACE=1 python tools/trace_tc.py"""
import ctypes as C
import os
import sys

os.environ["IVF_TC_TRACE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interpreting_video_features_b200 import _lib, engine, ops  # noqa: E402
from interpreting_video_features_b200.ops import Act  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def read_trace():
    buf = (C.c_longlong * 64)()
    ops.check(lib.ivf_debug_read_scratch(_lib.handle(dev), buf, 64 * 8), "read")
    return list(buf)


def show(name, t):
    t0 = t[0]
    print("== %s: kernel(CTA0) %.2f us; setup done +%.2f us" % (name, (t[1] - t0) / 1e3, (t[2] - t0) / 1e3))
    for i in range(4):
        b = 8 + i * 8
        if t[b + 4] == 0:
            break
        print("  tile %d: producer +%.2f | mma has acc +%.2f | first operands +%.2f | mmas issued +%.2f | acc complete "
              "+%.2f | epilogue done +%.2f" % ((i,) + tuple((t[b + k] - t0) / 1e3 for k in range(6))))


def show_chunks(t):
    t0 = t[0]
    print("    tile 1 chunk 0 detail: before wait_read +%.3f | after +%.3f | after math/st.shared +%.3f | after fence "
          "+%.3f | after store issue +%.3f" % tuple((t[52 + k] - t0) / 1e3 for k in range(5)))
    for c in range(4):
        b = 40 + c * 3
        print("    tile 1 chunk %d: before tmem ld +%.3f | after ld +%.3f | after math+stores +%.3f" %
              ((c,) + tuple((t[b + k] - t0) / 1e3 for k in range(3))))


def run(name, n, dhw, cin, c0, c12, dgrad):
    g = torch.Generator().manual_seed(0)
    x = Act(torch.randn((n,) + dhw + (cin,), generator=g).to(dev).bfloat16(), n, *dhw, cin)
    if not dgrad:
        w = engine.pack_fwd(torch.randn((c0 + c12, cin, 1, 1, 1), generator=g).to(dev) * 0.05, "bf16")
        out = Act.empty(n, *dhw, c0 + 64, torch.bfloat16, dev)
        t12 = Act.empty(n, *dhw, c12, torch.bfloat16, dev)
        sc = torch.ones(c0 + c12, device=dev)
        f = lambda: ops.conv1x1_split(x, w, out.slice(0, c0), out2=t12, flags=ops.EP_RELU, scale=sc, shift=sc)
    else:
        w = engine.pack_dgrad_two_sources(torch.randn((c0, cin, 1, 1, 1), generator=g).to(dev) * 0.05,
                                          torch.randn((c12, cin, 1, 1, 1), generator=g).to(dev) * 0.05)
        dz = Act.empty(n, *dhw, c0 + 64, torch.bfloat16, dev, zero=True)
        g12 = Act.empty(n, *dhw, c12, torch.bfloat16, dev, zero=True)
        acc = Act.empty(n, *dhw, cin, torch.float32, dev, zero=True)
        gx = x.like()
        ms = torch.ones(cin, device=dev)
        f = lambda: ops.conv1x1_split(dz.slice(0, c0), w, gx, x2=g12, acc_in=acc, mask=x, mask_scale=ms)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    # flush L2, then one traced launch
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    junk.fill_(1)
    torch.cuda.synchronize()
    f()
    t = read_trace()
    show(name, t)
    show_chunks(t)


run("3c fwd b0|b1a|b2a (256 -> 128+160), 8x8x28x28", 8, (8, 28, 28), 256, 128, 160, False)
run("3c dgrad (128|160 -> 256)", 8, (8, 28, 28), 256, 128, 160, True)
run("2b-like fwd (64 -> 32+32), 8x8x56x56", 8, (8, 56, 56), 64, 32, 32, False)
run("4e fwd (528 -> 112+176), 8x4x14x14", 8, (4, 14, 14), 528, 112, 176, False)
run("4e dgrad (112|176 -> 528), 8x4x14x14", 8, (4, 14, 14), 528, 112, 176, True)
