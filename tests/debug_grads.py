import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_train_step import *
from collections import OrderedDict

sp, inp = case()
taps = {}
with torch.no_grad():
    sd32 = util.state_dict_of(sp)
    so.speech_predictor(sd32, inp["texts"], inp["text_lengths"], inp["alignment"], inp["pitch"], inp["energy"],
                        inp["voiced"], inp["style"], inp["denormal_pitch"], inp["draws"], taps=taps)
prior = (taps["har_spec"], taps["har_phase"])
a64, g64, d64, _ = oracle_grads(sp, inp, torch.float64, prior=prior)
a32, g32, d32, _ = oracle_grads(sp, inp, torch.float32, prior=prior)
dev = torch.device("cuda:0")
sp = sp.to(dev).train()
sp.regularisers = False
c = lambda t: t.to(dev)
style, pitch, energy = (c(inp[k]).clone().requires_grad_(True) for k in ("style", "pitch", "energy"))
out = sp(c(inp["texts"]), c(inp["text_lengths"]), c(inp["alignment"]), pitch, energy, c(inp["voiced"]), style,
         c(inp["denormal_pitch"]), prior=(c(prior[0]), c(prior[1])))
(out.audio * c(cotangent(out.audio.shape))).sum().backward()
print("audio gpu-o64", rel_l2(out.audio, a64), "o32-o64", rel_l2(a32, a64))
print("sat: frac |audio|>0.999:", float((a64.abs() > 0.999).double().mean()), "max", float(a64.abs().max()))
params = dict(sp.named_parameters())
groups = OrderedDict()
for n in g64:
    key = ".".join(n.split(".")[:3]) if n.startswith("generator.basegen") else ".".join(n.split(".")[:2])
    groups.setdefault(key, []).append(n)
for key, ns in groups.items():
    ref = torch.cat([g64[n].flatten() for n in ns])
    gpu = torch.cat([params[n].grad.flatten().double().cpu() for n in ns])
    o32 = torch.cat([g32[n].flatten().double() for n in ns])
    print(f"{key:55s} |g|={float(ref.norm()):.3e}  gpu-o64 {rel_l2(gpu, ref):.2e}  o32-o64 {rel_l2(o32, ref):.2e}")
for k, t in (("style", style), ("pitch", pitch), ("energy", energy)):
    print(k, "gpu-o64", rel_l2(t.grad, d64[k]), "o32-o64", rel_l2(d32[k], d64[k]))
