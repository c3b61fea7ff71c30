"""Time the mel style encoder fwd+bwd at config-3 size (B=32, 80 x 804 mel) with a per-kernel table."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stylish_tts_b200 import synth, _lib as L
from stylish_tts_b200.style_encoder import MelStyleEncoder

torch.manual_seed(0)
m = synth.converge_spectral_(MelStyleEncoder(80, 64, 384, True)).cuda().train()
x = torch.randn(32, 1, 80, 804, device="cuda")

def step():
    for p in m.parameters(): p.grad = None
    m(x).square().sum().backward()

for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): step()
e1.record(); torch.cuda.synchronize()
print(f"style encoder fwd+bwd B=32: {e0.elapsed_time(e1)/3:.1f} ms")
L.profile_log = []
step(); torch.cuda.synchronize()
agg = {}
for sig, s0, s1, info in L.profile_log:
    n, tt = agg.get(sig, (0, 0.0)); agg[sig] = (n + 1, tt + s0.elapsed_time(s1))
L.profile_log = None
tot = sum(v[1] for v in agg.values())
print(f"sum of kernel time {tot:.1f} ms")
for sig, (n, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {100*tt/tot:6.2f}%  n={n:3d}  {tt:8.3f} ms  {sig}")
