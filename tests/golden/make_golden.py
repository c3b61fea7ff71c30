"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
Weights come from ``stylish_tts_b200.synth.randomize_`` (seeded; regenerated
identically on the GPU box), are loaded into the reference's own
``SpeechPredictor`` with ``strict=True`` and the reference forward is run on
seeded synthetic inputs with the harmonic-source random draws injected
(SURVEY.md F7).  Stored: the reference audio and tap tensors (large taps are
decimated in time by ``TAP_STRIDE`` to keep the fixtures small).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, ref_run  # noqa: E402
import stylish_tts_b200 as st  # noqa: E402
from stylish_tts_b200 import synth  # noqa: E402

TAP_STRIDE = 7
CASES = {
    # name: (batch, tokens, input seed, weight seed, ragged)
    "sp_b2_t16_ragged": (2, 16, 1, 0, True),
    "sp_b1_t12": (1, 12, 2, 3, False),
}


def decimate(t):
    return t[..., ::TAP_STRIDE].contiguous() if t.shape[-1] > 2000 else t


def main():
    torch.set_num_threads(8)
    mc = st.default_model_config()
    ref = ref_loader.build_model().speech_predictor.eval()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (b, t, iseed, wseed, ragged) in CASES.items():
        mine = st.build_model(mc).speech_predictor
        synth.randomize_(mine, wseed)
        ref.load_state_dict(mine.state_dict(), strict=True)
        inp = synth.speech_inputs(b, t, seed=iseed, ragged=ragged)
        taps = {}
        audio = ref_run.speech_predictor_forward(ref, inp, taps)
        blob = {"audio": audio.numpy()}
        for k, v in taps.items():
            blob["tap_" + k] = decimate(v).numpy()
        blob["meta"] = np.array([b, t, iseed, wseed, int(ragged), TAP_STRIDE])
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, {k: v.shape for k, v in blob.items()}, os.path.getsize(path) // 1024, "KiB")

    # duration / pitch-energy predictors (duration_predictor.py, pitch_energy_predictor.py)
    nets = ref_loader.build_model()
    dpm, pem = nets.duration_predictor.eval(), nets.pitch_energy_predictor.eval()
    mine = st.build_model(mc)
    synth.randomize_(mine.duration_predictor, 6)
    synth.randomize_(mine.pitch_energy_predictor, 7)
    dpm.load_state_dict(mine.duration_predictor.state_dict(), strict=True)
    pem.load_state_dict(mine.pitch_energy_predictor.state_dict(), strict=True)
    inp = synth.speech_inputs(2, 20, seed=8, ragged=True)
    sty = torch.randn(2, 64, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        dpred = dpm(inp["texts"], inp["text_lengths"], sty)
        pitch, energy = pem(inp["texts"], inp["text_lengths"], inp["alignment"], sty)
        from stylish_tts.train.utils import DurationProcessor as RefDP
        rdp = RefDP(16, 50)
        soft = rdp.prediction_to_duration(dpred, inp["text_lengths"])
        al = rdp(dpred, inp["text_lengths"])
    np.savez_compressed(os.path.join(out_dir, "predictors.npz"), dur_pred=dpred.numpy(),
                        pitch=pitch.numpy(), energy=energy.numpy(), soft_duration=soft.numpy(),
                        alignment=al.numpy(), style=sty.numpy())
    print("predictors", tuple(dpred.shape), tuple(pitch.shape), tuple(al.shape))

    # alignment golden (DurationProcessor.duration_to_alignment, utils.py:752-791)
    ref_loader.load()
    from stylish_tts.train.utils import DurationProcessor
    dp = DurationProcessor(16, 50)
    g = torch.Generator().manual_seed(5)
    dur = torch.randint(0, 12, (3, 20), generator=g).float()
    dur[1, 15:] = 0
    dur[2, 0] = 0.37  # fractional (soft) durations as produced by prediction_to_duration
    al = dp.duration_to_alignment(dur)
    np.savez_compressed(os.path.join(out_dir, "alignment.npz"), duration=dur.numpy(),
                        alignment=al.numpy())
    print("alignment", tuple(al.shape))


if __name__ == "__main__":
    main()
