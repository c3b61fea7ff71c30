"""Micro-benchmark of single C-ABI kernels at the BASELINE config-2 shapes
(B=16, S=60225 steps).  python tools/microbench.py [case ...] [--iters N]"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylish_tts_b200 import _lib as L  # noqa: E402
from stylish_tts_b200 import engine as E  # noqa: E402

B, S = 16, 60225
d = torch.device("cuda:0")


def conv_case(ci, co, k, dil=1, T=S, pro=None, out_act=0, ssq=False, res=False, shuffle=0, dwln=False):
    x = E.empty_bct(B, ci, T, d).normal_()  # pitch-padded rows, like the engine's S-rate activations (TMA path)
    w = E.ConvW(torch.randn(co, ci, k, device=d) / math.sqrt(ci * k), torch.randn(co, device=d))
    kw = {}
    if pro:
        kw["in_scale"] = torch.rand(B, ci, device=d) + 0.5
        if pro != "scale":
            kw["in_shift"] = torch.randn(B, ci, device=d)
        if pro == "snake":
            kw["in_act"] = L.ACT_SNAKE
            kw["in_alpha"] = torch.rand(ci, device=d) + 0.5
        elif pro == "leaky":
            kw["in_act"] = L.ACT_LEAKY02
    if out_act == L.ACT_SNAKE:
        kw["out_alpha"] = torch.rand(co, device=d) + 0.5
    if ssq:
        kw["out_sumsq"] = torch.zeros(B, co, device=d)
    if dwln:
        kw["dwln"] = (torch.randn(ci, 7, device=d) * 0.3, torch.randn(ci, device=d) * 0.1,
                      torch.randn(B, 2 * ci, device=d) * 0.3, 2 * ci, 1e-6)
    s = shuffle if shuffle > 1 else 1
    out = E.empty_bct(B, co // s, T * s, d)
    if res:
        kw["res"] = E.empty_bct(B, co // s, T * s, d).normal_()
    flops = 2.0 * B * T * ci * co * k
    byts = 4.0 * B * T * (ci + co + (co if res else 0))
    return (lambda: E.conv1d(x, w, dil=dil, out=out, out_act=out_act, shuffle=shuffle, **kw)), flops, byts


CASES = {
    "pw1": lambda: conv_case(32, 128, 1, out_act=L.ACT_SNAKE, ssq=True),
    "pw1dw": lambda: conv_case(32, 128, 1, out_act=L.ACT_SNAKE, ssq=True, dwln=True),
    "pw2": lambda: conv_case(128, 32, 1, pro="scale", res=True),
    "pw2_plain": lambda: conv_case(128, 32, 1),
    "k21": lambda: conv_case(32, 32, 21),
    "k21_96": lambda: conv_case(96, 32, 21),
    "k21_64": lambda: conv_case(32, 64, 21),
    "k11pro": lambda: conv_case(32, 32, 11, pro="snake", res=True),
    "k11d5pro": lambda: conv_case(32, 32, 11, dil=5, pro="snake"),
    "k11": lambda: conv_case(32, 32, 11),
    "ff2": lambda: conv_case(1024, 256, 1, T=803, pro="scale", res=True),
    "ff1": lambda: conv_case(256, 1024, 1, T=803, out_act=L.ACT_SNAKE, ssq=True),
    "ffn2": lambda: conv_case(512, 128, 3, T=258),
    "up2": lambda: conv_case(64, 160, 11, T=12045, shuffle=5),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=list(CASES))
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    for name in a.cases:
        fn, flops, byts = CASES[name]()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"{name:10s} {ms:8.4f} ms  {flops / ms / 1e9:7.2f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s (algorithmic)",
              flush=True)


if __name__ == "__main__":
    main()
