"""Golden of the UNMODIFIED reference's composed acoustic training step (build container only):

    python tests/golden/make_acoustic_step_golden.py

Runs the reference's own ``AcousticStep(use_predicted_pe=False, predict_audio=True)``
(stage_type.py:61-180), ``step.mel_loss()``, ``step.multi_phase_loss()`` (stage_type.py:182-193) and
``LossLog.backwards_loss()`` (loss_log.py:82-94) on a fake ``train`` namespace built from the reference's own
objects (torchaudio MelSpectrogram ×2, MultiSpectrogram, MultiResolutionSTFTLoss, DurationProcessor,
NormalizationStats, the ``loss_weight`` block of config/config.yml), exactly as ``train_acoustic``
(stage_type.py:346-366) does minus the adversarial / SLM / no-op magphase terms, then ``.backward()``.

Pinned configuration: speech_predictor in eval() with its BatchNorm1d in train()
(batch statistics, stochastic regularisers off), speech_style_encoder in train() (spectral-norm power iteration
runs, as in a real step), harmonic-source draws injected, conditioned phase head (synth.condition_phase_head_).
Stored: every scalar the step logs, the backward scalar, energy / style / predicted audio, the gradient w.r.t.
the predicted audio, and for BOTH trained modules per-parameter gradient L2 norm + seeded probe dot.
"""
import logging
import os
import sys
import types
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASE = dict(batch=2, tokens=24, frames=76, wseed=0, sseed=5, iseed=3)


def probe(name, shape):
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    return torch.randn(shape, generator=g)


def make_batch(case=CASE, hop=300, harmonics=9):
    """The collated batch the reference's AcousticStep reads (audio_gt, text, text_length, pitch, alignment =
    integer durations (B,1,T)), seeded; plus the harmonic-source draws."""
    g = torch.Generator().manual_seed(case["iseed"])
    B, T, Fr = case["batch"], case["tokens"], case["frames"]
    text = torch.randint(1, 178, (B, T), generator=g)
    lengths = torch.full((B,), T, dtype=torch.long)
    if B > 1:
        lengths[1:] = torch.randint(T // 2, T, (B - 1,), generator=g)
    dur = torch.zeros((B, T), dtype=torch.long)
    for b in range(B):
        n = int(lengths[b])
        text[b, 0] = 0
        text[b, n - 1:] = 0
        d = torch.full((n,), 2, dtype=torch.long)
        total = Fr if b == 0 else int(Fr * n / T)
        extra = total - int(d.sum())
        idx = torch.randperm(n, generator=g)
        for j in range(extra):
            d[idx[j % n]] += 1
        dur[b, :n] = d
    L = Fr * hop
    audio = 0.1 * torch.randn(B, L, generator=g)
    pitch = 80.0 + 200.0 * torch.rand(B, Fr, generator=g)
    pitch = torch.nn.functional.avg_pool1d(pitch.unsqueeze(1), 5, 1, 2, count_include_pad=False).squeeze(1)
    for b in range(B):
        s = int(torch.randint(0, Fr - 8, (1,), generator=g))
        pitch[b, s:s + 8] = 0.0
    draws = {"rand_ini": torch.rand(B, harmonics, generator=g),
             "noise": torch.randn(B, L, harmonics, generator=g)}
    return dict(audio_gt=audio, text=text, text_length=lengths, pitch=pitch, alignment=dur.unsqueeze(1)), draws


def seeded_nets():
    """(speech_predictor, speech_style_encoder) of THIS repo with the fixture's seeded weights (CPU)."""
    import stylish_tts_b200 as st
    from stylish_tts_b200 import synth

    nets = st.build_model(st.default_model_config())
    sp, se = nets.speech_predictor, nets.speech_style_encoder
    synth.randomize_(sp, CASE["wseed"])
    synth.condition_phase_head_(sp)
    synth.randomize_(se, CASE["sseed"])
    synth.converge_spectral_(se)
    return sp, se


def main():
    from oracle import ref_loader, ref_run

    ref_loader.load()
    import torchaudio
    from stylish_tts.lib.config_loader import load_config_yaml
    import stylish_tts.train.train_context as tc
    from stylish_tts.train.stage_type import AcousticStep
    from stylish_tts.train.loss_log import build_loss_log
    from stylish_tts.train.losses import MultiResolutionSTFTLoss
    from stylish_tts.train.multi_spectrogram import MultiSpectrogram
    from stylish_tts.train.utils import DurationProcessor

    torch.set_num_threads(8)
    mc = ref_loader.model_config()
    cfg = load_config_yaml(os.path.join(ref_loader.REF_ROOT, "config", "config.yml"))
    ref = ref_loader.build_model()
    sp, se = seeded_nets()
    ref.speech_predictor.load_state_dict(sp.state_dict(), strict=True)
    ref.speech_style_encoder.load_state_dict(se.state_dict(), strict=True)
    ref.speech_predictor.eval()
    for m in ref.speech_predictor.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.train()
    ref.speech_style_encoder.train()

    train = types.SimpleNamespace(
        model=ref, model_config=mc, config=cfg, logger=logging.getLogger("golden"), writer=None,
        normalization=tc.NormalizationStats(),
        duration_processor=DurationProcessor(class_count=mc.duration_predictor.duration_classes,
                                             max_dur=mc.duration_predictor.max_duration),
        to_mel=torchaudio.transforms.MelSpectrogram(n_mels=mc.n_mels, n_fft=mc.n_fft, win_length=mc.win_length,
                                                    hop_length=mc.hop_length, sample_rate=mc.sample_rate),
        to_style_mel=torchaudio.transforms.MelSpectrogram(
            n_mels=mc.style_encoder.n_mels, n_fft=mc.style_encoder.n_fft, win_length=mc.style_encoder.win_length,
            hop_length=mc.style_encoder.hop_length, sample_rate=mc.sample_rate),
        multi_spectrogram=MultiSpectrogram(sample_rate=mc.sample_rate),
        stft_loss=MultiResolutionSTFTLoss(sample_rate=mc.sample_rate), generator_loss=None)

    raw, draws = make_batch()
    batch = ref_loader.Munch(**raw)
    log = build_loss_log(train)
    with ref_run.injected_draws(draws):
        step = AcousticStep(batch, train, log, use_predicted_pe=False, predict_audio=True)
    step.pred.audio.retain_grad()
    step.speech_style.retain_grad()
    step.mel_loss()
    step.multi_phase_loss()
    total = log.backwards_loss()
    total.backward()

    blob = dict(
        mel_loss=np.float64(log.metrics["mel"].item()), phase_loss=np.float64(log.metrics["multi_phase"].item()),
        backward_scalar=np.float64(total.item()), logged_total=np.float64(float(log.total())),
        weights=np.array([log.weight("mel"), log.weight("multi_phase")], dtype=np.float64),
        energy=step.energy.numpy(), mel_target=step.mel.numpy(), style=step.speech_style.detach().numpy(),
        audio=step.pred.audio.detach().numpy(), d_audio=step.pred.audio.grad.numpy(),
        d_style=step.speech_style.grad.numpy())
    for key, mod in (("sp", ref.speech_predictor), ("se", ref.speech_style_encoder)):
        names, norms, dots = [], [], []
        for name, p in sorted(mod.named_parameters()):
            if p.grad is None:
                continue
            names.append(name)
            norms.append(float(p.grad.norm()))
            dots.append(float((p.grad * probe(name, p.shape)).sum()))
        blob[key + "_names"] = np.array(names)
        blob[key + "_norms"] = np.array(norms, dtype=np.float64)
        blob[key + "_dots"] = np.array(dots, dtype=np.float64)
        print(key, len(names), "parameters with gradients")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "acoustic_step.npz")
    np.savez_compressed(path, **blob)
    print({k: float(blob[k]) for k in ("mel_loss", "phase_loss", "backward_scalar", "logged_total")},
          os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
