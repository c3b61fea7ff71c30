"""GPU-box micro-benchmark of the fused ConvNeXt block (both passes) at B=16, T=60225."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stylish_tts_b200 import _lib as L, engine as E

B, T, Cc, J = 16, 60225, 32, 128
d = torch.device("cuda:0")
g = torch.Generator(device=d).manual_seed(0)
x = E.empty_bct(B, Cc, T, d).normal_(generator=g)
y = E.empty_bct(B, Cc, T, d)
w1 = E.ConvW(torch.randn(J, Cc, 1, device=d) / math.sqrt(Cc), torch.randn(J, device=d) * 0.1)
w2 = E.ConvW(torch.randn(Cc, J, 1, device=d) / math.sqrt(J), torch.randn(Cc, device=d) * 0.1)
blk = dict(dw_w=torch.randn(Cc, 7, device=d) * 0.4, dw_b=torch.randn(Cc, device=d) * 0.1, norm="n", pw1=w1,
           snake=0.75 + 0.5 * torch.rand(J, device=d), grn_gamma=torch.randn(J, device=d) * 0.3, pw2=w2)
h = torch.randn(B, 2 * Cc, device=d) * 0.3
P = type("P", (), dict(fc_rows=2 * Cc, fc_off={"n": 0}))()
eng = E.SpeechEngine.__new__(E.SpeechEngine)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    eng.convnext(P, blk, x, h, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    eng.convnext(P, blk, x, h, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"fused convnext block: {ms:.4f} ms  ({3 * B * Cc * T * 4 / ms / 1e6:.0f} GB/s on 3u)")
