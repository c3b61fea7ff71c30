"""Achieved HBM bandwidth of the streaming backward / reduction kernels at the S-rate shapes of the train step
(B=32, T=60300): python tools/bench_bwd_stream.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylish_tts_b200 import _lib as L  # noqa: E402
from stylish_tts_b200 import train_ops as T  # noqa: E402

d = torch.device("cuda:0")
B, S = 32, 60300


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, nbytes):
    print(f"{name:34s} {ms:8.4f} ms   {nbytes / ms / 1e6:8.1f} GB/s (algorithmic)")


for C in (32, 128):
    x = torch.randn(B, C, S, device=d)
    y = torch.randn(B, C, S, device=d)
    n = x.numel() * 4
    report(f"channel_sum C={C}", timeit(lambda: T.channel_sum(x)), n)
    report(f"row_moments C={C}", timeit(lambda: T.row_moments(x)), n)
    out = torch.empty(B * C, device=d)
    report(f"row_dot C={C}", timeit(lambda: L.call("sty_row_dot", x.data_ptr(), y.data_ptr(), out.data_ptr(), B * C, S,
                                                  L.stream_ptr())), 2 * n)
    sc, sh = torch.rand(B, C, device=d) + 0.5, torch.randn(B, C, device=d)
    al = torch.rand(C, device=d) + 0.5
    mean = torch.zeros(B, C, device=d)
    report(f"prologue_bwd_reduce snake C={C}",
           timeit(lambda: T.prologue_bwd(y, x, scale=sc, shift=sh, alpha=al, mask=None, act=L.ACT_SNAKE, center=mean,
                                         want_sums=True, sums_only=True)), 2 * n)
    c0, c1 = torch.randn(B, C, device=d), torch.randn(B, C, device=d)
    report(f"prologue_bwd_apply snake C={C}",
           timeit(lambda: T.prologue_bwd(y, x, scale=sc, shift=sh, alpha=al, mask=None, act=L.ACT_SNAKE,
                                         want_sums=False, c0=c0, c1=c1)), 3 * n)
    xg = x.clone().requires_grad_(True)
    gb = torch.randn(B, 2 * C, device=d, requires_grad=True)
    o = T.chan_ln(xg, gb=gb, eps=1e-6)
    report(f"chan_ln fwd C={C}", timeit(lambda: T.chan_ln(xg, gb=gb, eps=1e-6)), 2 * n)
    report(f"chan_ln bwd C={C}", timeit(lambda: torch.autograd.grad(o, (xg, gb), y, retain_graph=True)), 3 * n)
    w = torch.randn(C, 1, 7, device=d, requires_grad=True)
    bias = torch.randn(C, device=d, requires_grad=True)
    o2 = T.DwConvFn.apply(xg, w, bias, 7, 3)
    report(f"dwconv7 bwd C={C}", timeit(lambda: torch.autograd.grad(o2, (xg, w, bias), y, retain_graph=True)), 3 * n)
    del x, y, xg, o, o2
