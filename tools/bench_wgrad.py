"""Time sty_conv1d_wgrad (tcgen05 vs fp32 FMA) at the training shapes (B=32, S-rate T=60225)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stylish_tts_b200 import train_ops as T, _lib as L

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

shapes = [(32, 32, 11, 1, 60225), (32, 32, 11, 5, 60225), (32, 32, 21, 1, 60225), (96, 32, 21, 1, 60225),
          (32, 128, 1, 1, 60225), (128, 32, 1, 1, 60225), (256, 1024, 1, 1, 803), (128, 512, 3, 1, 258)]
B = 32
for ci, co, k, dil, Tn in shapes:
    x = torch.randn(B, ci, Tn, device="cuda"); dy = torch.randn(B, co, Tn, device="cuda")
    tu = timeit(lambda: T.wgrad(x, dy, k, dil, umma=True))
    ts = float("nan")  # fp32-FMA column: pass --fma
    if "--fma" in sys.argv:
        ts = timeit(lambda: T.wgrad(x, dy, k, dil, umma=False), n=2)
    byts = B * Tn * (ci + co) * 4
    fl = 2.0 * B * Tn * ci * co * k
    print(f"wgrad ci={ci} co={co} k={k} d={dil} T={Tn}: tcgen05 {tu:.3f} ms ({byts/tu/1e6:.0f} GB/s, {fl/tu/1e9:.1f} TFLOP/s)   fp32-FMA {ts:.3f} ms")
    del x, dy
