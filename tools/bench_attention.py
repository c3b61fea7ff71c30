"""GPU box: conformer-shaped attention (B=16, 8 heads x 64, T=803) and denoiser-shaped (B=64, T=258): first vs second
generation kernel, CUDA-event times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stylish_tts_b200 import engine as E

d = torch.device("cuda:0")
for B, T in ((16, 803), (64, 258), (32, 804)):
    qkv = torch.randn(B, 3 * 512, T, device=d)
    for gen2 in (False, True):
        E.ATTENTION64 = gen2
        for _ in range(3):
            E.attention(qkv, 512, 512, 512, H=8, D=64, scale=0.125)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            E.attention(qkv, 512, 512, 512, H=8, D=64, scale=0.125)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 4.0 * B * 8 * T * T * 64
        print(f"B={B} T={T} gen{2 if gen2 else 1}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s logical ({3 * fl / ms / 1e9:.0f} bf16 MMA)")

# backward (B=32, T=804: the conformer attention of the training step)
from stylish_tts_b200 import train_ops as TO
for gen2 in (False, True):
    E.ATTENTION64 = gen2
    qkv = torch.randn(32, 3 * 512, 804, device=d, requires_grad=True)
    out = TO.AttentionFn.apply(qkv, 8, 64, None, None, 0.125)
    g = torch.randn_like(out)
    for _ in range(2):
        out.backward(g, retain_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out.backward(g, retain_graph=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"backward B=32 T=804 gen{2 if gen2 else 1}: {e0.elapsed_time(e1) / 5:.3f} ms")
