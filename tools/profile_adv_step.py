"""GPU-box: per-kernel device time (CUDA events around every C-ABI call) of the discriminator part of a training step."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stylish_tts_b200 import _lib as L, discriminator as D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = torch.device("cuda:0")
mrd = [D.SpecDiscriminator().to(d) for _ in range(3)]
g = torch.Generator(device=d).manual_seed(0)
shapes = [(257, 1876), (513, 938), (1025, 469)]
tf = [torch.rand(B, 1, k, n, device=d, generator=g) for k, n in shapes]
pf = [(t * 0.9).requires_grad_(True) for t in tf]
import math, random
adv = D.AdversarialTerms(mrd0=mrd[0], mrd1=mrd[1], mrd2=mrd[2], device=d)

def step():
    # one evaluation of mrd0-2 on (target, prediction): generator term + backward to the spectrograms, then the
    # discriminator loss from the same scores and the backward of the stepped discriminator
    loss = adv(target_list=tf, pred_list=pf)
    loss.backward()
    adv.discriminator_backward(random.randrange(3), math.sqrt(B))
    for m in mrd:
        m.zero_grad()
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"B={B}: adversarial part (gen loss fwd+bwd, disc loss fwd+bwd) {e0.elapsed_time(e1):.1f} ms wall-on-device")
L.profile_log = []
step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for sig, a, b, info in L.profile_log:
    agg[sig][0] += a.elapsed_time(b); agg[sig][1] += 1
L.profile_log = None
tot = sum(v[0] for v in agg.values())
print(f"sum of C-ABI kernel time {tot:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOPN", "25"))]:
    print(f"  {v[0]:8.2f} ms  n={v[1]:3d}  {k}")
