"""GPU box: per-kernel device time of the waveform discriminator (`disc`) forward + backward at training size."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stylish_tts_b200 import _lib as L, discriminator as D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
d = torch.device("cuda:0")
m = D.ContextFreeDiscriminator().to(d).train()
x = (0.1 * torch.randn(B, 241200, device=d)).requires_grad_(True)

def step():
    out = m(x)[0][0]
    out.square().mean().backward()
    m.zero_grad()
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"B={B}: disc fwd+bwd {e0.elapsed_time(e1):.1f} ms")
L.profile_log = []
step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for sig, a, b, info in L.profile_log:
    agg[sig][0] += a.elapsed_time(b); agg[sig][1] += 1
L.profile_log = None
print(f"sum of C-ABI kernel time {sum(v[0] for v in agg.values()):.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {v[0]:8.2f} ms  n={v[1]:3d}  {k}")
