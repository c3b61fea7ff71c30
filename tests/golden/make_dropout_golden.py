"""Golden output + gradients of the UNMODIFIED reference SpeechPredictor in full train() mode
(build container only):

    python tests/golden/make_dropout_golden.py

train() = batch-statistics BatchNorm, every dropout site live, decoder box smoothing live.  The two samplers
the reference calls (F.dropout, F.scaled_dot_product_attention's dropout_p) are replaced for the run by the
hash masks of the CUDA path (oracle/dropout_oracle.patched_reference), and decoder.py's random.randint is
pinned to widths (7, 15) — so the fixture pins WHERE each mask is applied, its scaling and its memory layout.
Same storage scheme as make_train_golden.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from make_train_golden import CASE, cotangent, probe  # noqa: E402

SEED = 0x1234_5678_9ABC_DEF0
SMOOTHING = (7, 15)


def main():
    from oracle import dropout_oracle as do, ref_loader, ref_run
    import stylish_tts_b200 as st
    from stylish_tts_b200 import synth

    torch.set_num_threads(8)
    ref = ref_loader.build_model().speech_predictor.train()
    mine = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(mine, CASE["wseed"])
    synth.condition_phase_head_(mine)  # well-conditioned phase head: gradients comparable at kernel accuracy
    ref.load_state_dict(mine.state_dict(), strict=True)
    inp = synth.speech_inputs(CASE["batch"], CASE["tokens"], seed=CASE["iseed"], ragged=CASE["ragged"])
    style = inp["style"].clone().requires_grad_(True)
    pitch = inp["pitch"].clone().requires_grad_(True)
    energy = inp["energy"].clone().requires_grad_(True)
    sites = do.speech_predictor_sites()
    with ref_run.injected_draws(inp["draws"]), do.patched_reference(SEED, sites, SMOOTHING):
        out = ref(inp["texts"], inp["text_lengths"], inp["alignment"], pitch, energy, inp["voiced"], style,
                  inp["denormal_pitch"])
        audio = out.audio
        (audio * cotangent(audio.shape)).sum().backward()
    blob = dict(audio=audio.detach().numpy(), d_style=style.grad.numpy(), d_pitch=pitch.grad.numpy(),
                d_energy=energy.grad.numpy(), seed=np.array([SEED], dtype=np.uint64),
                smoothing=np.array(SMOOTHING))
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
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_grads_dropout.npz")
    np.savez_compressed(path, **blob)
    print(len(names), "parameters with gradients;", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
