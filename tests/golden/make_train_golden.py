"""Golden gradients of the UNMODIFIED reference SpeechPredictor (build container only):

    python tests/golden/make_train_golden.py

Two weight sets: the plain seeded random one (``train_grads.npz``; its phase head makes the gradient
ill-conditioned, see synth.condition_phase_head_) and the same with the conditioned phase head
(``train_grads_wc.npz``), on which gradient parity is asserted at kernel accuracy.
Setting: module.eval() with its BatchNorm1d switched to train() — i.e. batch statistics but the
stochastic regularisers (dropout, decoder box smoothing) off, the pinned configuration of
SURVEY.md §8(d) config 3 — loss = <cotangent, audio> with a seeded cotangent, harmonic-source draws
injected.  12.9 M gradients do not fit a fixture, so per parameter we store its L2 norm and its dot
product with a seeded probe (``probe(name, shape)``); the input gradients (style, pitch, energy)
and the audio are stored in full.
"""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASE = dict(batch=2, tokens=16, iseed=1, wseed=0, ragged=True, ct_seed=21)


def probe(name, shape):
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    return torch.randn(shape, generator=g)


def cotangent(shape, seed=CASE["ct_seed"]):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def main():
    run(False, "train_grads.npz")
    run(True, "train_grads_wc.npz")


def run(conditioned, fname):
    from oracle import ref_loader, ref_run
    import stylish_tts_b200 as st
    from stylish_tts_b200 import synth

    torch.set_num_threads(8)
    ref = ref_loader.build_model().speech_predictor.eval()
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.train()
    mine = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(mine, CASE["wseed"])
    if conditioned:
        synth.condition_phase_head_(mine)
    ref.load_state_dict(mine.state_dict(), strict=True)
    inp = synth.speech_inputs(CASE["batch"], CASE["tokens"], seed=CASE["iseed"], ragged=CASE["ragged"])
    style = inp["style"].clone().requires_grad_(True)
    pitch = inp["pitch"].clone().requires_grad_(True)
    energy = inp["energy"].clone().requires_grad_(True)
    with ref_run.injected_draws(inp["draws"]):
        out = ref(inp["texts"], inp["text_lengths"], inp["alignment"], pitch, energy, inp["voiced"], style,
                  inp["denormal_pitch"])
    audio = out.audio
    (audio * cotangent(audio.shape)).sum().backward()
    blob = dict(audio=audio.detach().numpy(), d_style=style.grad.numpy(), d_pitch=pitch.grad.numpy(),
                d_energy=energy.grad.numpy())
    names, norms, dots = [], [], []
    for name, p in sorted(ref.named_parameters()):
        if p.grad is None:
            continue
        names.append(name)
        norms.append(float(p.grad.norm()))
        dots.append(float((p.grad * probe(name, p.shape)).sum()))
    blob["names"] = np.array(names)
    blob["norms"] = np.array(norms, dtype=np.float64)
    blob["dots"] = np.array(dots, dtype=np.float64)
    bn = "generator.amp_conformer.layers.0.conv.net.4"
    sd = ref.state_dict()
    blob["bn_running_mean"] = sd[bn + ".running_mean"].numpy()
    blob["bn_running_var"] = sd[bn + ".running_var"].numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), fname)
    np.savez_compressed(path, **blob)
    print(len(names), "parameters with gradients;", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
