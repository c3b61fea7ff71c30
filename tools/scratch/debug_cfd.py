import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import disc_oracle as do
from stylish_tts_b200 import discriminator as D, engine as E
from tests import util
from tests.golden.make_disc_golden import state_dict_from_table
from tests.util import rel_l2
E.USE_UMMA = False
g = np.load(util.GOLDEN_DIR + "/discriminators.npz")
m = D.ContextFreeDiscriminator()
m.load_state_dict(state_dict_from_table(g["disc_names"], g["disc_shapes"]), strict=True)
training = len(sys.argv) > 1 and sys.argv[1] == "train"
gen = torch.Generator().manual_seed(3)
x = 0.2 * torch.randn(3, 1024 + 512 * 6, generator=gen)
names = {k for k, _ in m.named_parameters()}
sd64 = {k: (v.detach().double().requires_grad_(True) if k in names else v.detach().double().clone()
            if v.is_floating_point() else v.clone()) for k, v in m.state_dict().items()}
x64 = x.double().requires_grad_(True)
ref = do.context_free_discriminator(sd64, x64, bn_training=training)[0]
cot = torch.randn(ref.shape, generator=gen).double()
(ref * cot).sum().backward()
d = torch.device("cuda:0")
md = m.to(d); md.train(training)
xd = x.to(d).requires_grad_(True)
out = md(xd)[0][0]
print("fwd", rel_l2(out, ref))
(out * cot.float().to(d)).sum().backward()
print("d(x)", rel_l2(xd.grad, x64.grad))
e = (xd.grad.cpu().double() - x64.grad).abs()
print("   worst positions of d(x):", e.flatten().topk(8).indices.tolist(), "max", float(e.max()), "ref scale", float(x64.grad.abs().mean()))
pos = e[0]
print("   error by position mod 512 (mean):", [round(float(pos[i::512].mean()), 6) for i in (0, 1, 2, 3, 4, 5, 255, 256, 508, 509, 510, 511)])
for k, p in md.named_parameters():
    print(f"{rel_l2(p.grad, sd64[k].grad):10.3e}  {float(sd64[k].grad.norm()):10.3e}  {k}")
