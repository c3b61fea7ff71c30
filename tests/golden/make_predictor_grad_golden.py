"""Golden gradients of the UNMODIFIED reference DurationPredictor / PitchEnergyPredictor (build container only):

    python tests/golden/make_predictor_grad_golden.py

eval() mode (dropout / DropPath off — the deterministic setting of SURVEY §8d), loss = <cotangent, output>.
Stored per parameter: gradient L2 norm and its dot with a seeded probe; d(style) in full; the outputs."""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASE = dict(batch=2, tokens=20, iseed=8, dseed=6, pseed=7, sseed=9)


def probe(name, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(zlib.crc32(name.encode())))


def cot(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def build(conditioned=False):
    """conditioned: every utterance is padded (synth.speech_inputs all_padded) — the well-conditioned case on
    which gradient parity is asserted at kernel accuracy (``predictor_grads_wc.npz``)."""
    import stylish_tts_b200 as st
    from stylish_tts_b200 import synth
    nets = st.build_model(st.default_model_config())
    synth.randomize_(nets.duration_predictor, CASE["dseed"])
    synth.randomize_(nets.pitch_energy_predictor, CASE["pseed"])
    inp = synth.speech_inputs(CASE["batch"], CASE["tokens"], seed=CASE["iseed"], ragged=True,
                              all_padded=conditioned)
    sty = torch.randn(CASE["batch"], 64, generator=torch.Generator().manual_seed(CASE["sseed"]))
    return nets, inp, sty


def main():
    run(False, "predictor_grads.npz")
    run(True, "predictor_grads_wc.npz")


def run(conditioned, fname):
    from oracle import ref_loader
    nets, inp, sty = build(conditioned)
    ref = ref_loader.build_model()
    dpm, pem = ref.duration_predictor.eval(), ref.pitch_energy_predictor.eval()
    dpm.load_state_dict(nets.duration_predictor.state_dict(), strict=True)
    pem.load_state_dict(nets.pitch_energy_predictor.state_dict(), strict=True)
    blob = {}
    s1 = sty.clone().requires_grad_(True)
    out = dpm(inp["texts"], inp["text_lengths"], s1)
    (out * cot(out.shape, 41)).sum().backward()
    blob["dur_out"], blob["dur_dstyle"] = out.detach().numpy(), s1.grad.numpy()
    s2 = sty.clone().requires_grad_(True)
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
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), fname)
    np.savez_compressed(path, **blob)
    print(len(blob["dur_names"]), len(blob["pe_names"]), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
