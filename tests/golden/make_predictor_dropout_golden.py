"""Golden outputs + gradients of the UNMODIFIED reference DurationPredictor / PitchEnergyPredictor in train()
mode (build container only):

    python tests/golden/make_predictor_dropout_golden.py

Every sampler the two modules call in train() — F.dropout, SDPA dropout_p, F.dropout1d, DropPath's
Tensor.bernoulli_ — is replaced for the run by the hash masks of the CUDA path
(oracle/dropout_oracle.patched_reference), which pins site placement, scaling and mask layout.
Storage as in make_predictor_grad_golden.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from make_predictor_grad_golden import build, cot, probe  # noqa: E402

SEED_DUR = 0x0F1E_2D3C_4B5A_6978
SEED_PE = 0xA5A5_0123_4567_89AB


def main():
    from oracle import dropout_oracle as do, ref_loader
    nets, inp, sty = build()
    ref = ref_loader.build_model()
    dpm, pem = ref.duration_predictor.train(), ref.pitch_energy_predictor.train()
    dpm.load_state_dict(nets.duration_predictor.state_dict(), strict=True)
    pem.load_state_dict(nets.pitch_energy_predictor.state_dict(), strict=True)
    blob = {}
    s1 = sty.clone().requires_grad_(True)
    with do.patched_reference(SEED_DUR, do.duration_predictor_sites()):
        out = dpm(inp["texts"], inp["text_lengths"], s1)
        (out * cot(out.shape, 41)).sum().backward()
    blob["dur_out"], blob["dur_dstyle"] = out.detach().numpy(), s1.grad.numpy()
    s2 = sty.clone().requires_grad_(True)
    with do.patched_reference(SEED_PE, do.pitch_energy_predictor_sites()):
        pitch, energy = pem(inp["texts"], inp["text_lengths"], inp["alignment"], s2)
        ((pitch * cot(pitch.shape, 42)).sum() + (energy * cot(energy.shape, 43)).sum()).backward()
    blob["pe_pitch"], blob["pe_energy"], blob["pe_dstyle"] = pitch.detach().numpy(), energy.detach().numpy(), s2.grad.numpy()
    for tag, m in (("dur", dpm), ("pe", pem)):
        names, norms, dots = [], [], []
        for n, p in sorted(m.named_parameters()):
            if p.grad is None:
                continue
            names.append(n), norms.append(float(p.grad.norm())), dots.append(float((p.grad * probe(n, p.shape)).sum()))
        blob[tag + "_names"], blob[tag + "_norms"], blob[tag + "_dots"] = np.array(names), np.array(norms), np.array(dots)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "predictor_grads_dropout.npz")
    np.savez_compressed(path, **blob)
    print(len(blob["dur_names"]), len(blob["pe_names"]), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
