"""Golden fixture of the UNMODIFIED reference MelStyleEncoder (build container only):

    python tests/golden/make_style_golden.py

Weights: the product shell's own seeded initialisation (regenerated identically anywhere) loaded into the
reference with strict=True.  Stored: eval-mode output; train-mode output (one spectral-norm power iteration),
per-parameter gradient norm + seeded probe dot for loss = <cotangent, style>, and the advanced u vectors."""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SHAPE = (2, 1, 80, 52)


def style_inputs(seed=31):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(SHAPE, generator=g), torch.randn(SHAPE[0], 64, generator=g)


def build(seed=17):
    from stylish_tts_b200.style_encoder import MelStyleEncoder
    from stylish_tts_b200 import synth
    torch.manual_seed(seed)
    return synth.converge_spectral_(MelStyleEncoder(80, 64, 384, True))


def build_pitch(seed=19):
    from stylish_tts_b200.style_encoder import PitchStyleEncoder
    from stylish_tts_b200 import synth
    torch.manual_seed(seed)
    return synth.converge_spectral_(PitchStyleEncoder(80, 64, 384, True, coarse_multiplier=1))


def pitch_inputs(seed=33):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(2, 80, 50, generator=g), 100 + 50 * torch.rand(2, 50, generator=g),
            torch.randn(2, 50, generator=g))


def probe(name, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(zlib.crc32(name.encode())))


def main():
    from oracle import ref_loader
    ref = ref_loader.build_model().speech_style_encoder
    mine = build()
    ref.load_state_dict(mine.state_dict(), strict=True)
    x, ct = style_inputs()
    ref.eval()
    with torch.no_grad():
        out_eval = ref(x)
    ref.train()
    out = ref(x)
    (out * ct).sum().backward()
    names, norms, dots = [], [], []
    for n, p in sorted(ref.named_parameters()):
        names.append(n), norms.append(float(p.grad.norm())), dots.append(float((p.grad * probe(n, p.shape)).sum()))
    sd = ref.state_dict()
    blob = dict(out_eval=out_eval.numpy(), out_train=out.detach().numpy(), names=np.array(names),
                norms=np.array(norms), dots=np.array(dots), u0=sd["shared.0.weight_u"].numpy(),
                u6=sd["shared.6.weight_u"].numpy(), v3=sd["shared.3.conv2.weight_v"].numpy())
    pe = ref_loader.build_model().pe_style_encoder
    minep = build_pitch()
    pe.load_state_dict(minep.state_dict(), strict=True)
    pe.eval()
    with torch.no_grad():
        blob["pe_out_eval"] = pe(*pitch_inputs()).numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "style_encoder.npz")
    np.savez_compressed(path, **blob)
    print(len(names), "parameters;", os.path.getsize(path) // 1024, "KiB; out", out_eval[0, :4])


if __name__ == "__main__":
    main()
