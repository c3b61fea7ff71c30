"""The drop-in claim of INTEGRATION.md, exercised in the build container (needs /root/reference, no GPU): the
UNMODIFIED reference's ``AcousticStep`` class (stage_type.py:61-180), its ``mel_loss`` / ``multi_phase_loss`` /
``generator_loss`` methods, ``LossLog.backwards_loss`` (loss_log.py:82-94) and ``.backward()`` are run against THIS
repo's modules and TrainContext objects — ``build_model`` modules for ``train.model``, ``spectral.MelSpectrogram`` /
``MultiSpectrogram`` / ``MultiResolutionSTFTLoss`` for ``train.to_mel`` … ``train.stft_loss``,
``discriminator.GeneratorLoss`` for ``train.generator_loss``, and the three functions INTEGRATION.md re-imports in
stage_type.py (``calculate_mel``, ``log_norm``, ``multi_phase_loss``).

Every C-ABI call is replaced by a recorder and the tensors stay on the CPU (``_lib.DRY_RUN``), so NO arithmetic is
performed or checked (tests/test_acoustic_step_golden.py does that on the GPU); what is checked is that the reference's
own step code drives our objects unchanged: call keywords, result types, shapes, and that its backward reaches every
trained parameter through our autograd functions."""
import collections
import logging
import os
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container)")


@pytest.fixture()
def dry_run(monkeypatch):
    from stylish_tts_b200 import _lib as L

    calls = []

    def recorder(name, *args):
        assert name in L._SIGNATURES, name
        assert len(args) == len(L._SIGNATURES[name]), (name, len(args), len(L._SIGNATURES[name]))
        calls.append(name)

    monkeypatch.setattr(L, "call", recorder)
    monkeypatch.setattr(L, "stream_ptr", lambda: 0)
    monkeypatch.setattr(L, "DRY_RUN", True)
    return calls


def _reference_side(monkeypatch):
    """the reference's stage module with the three imports INTEGRATION.md re-points, and a `train` namespace made of
    THIS repo's modules / TrainContext objects plus the reference's own helpers"""
    from oracle import ref_loader

    ref_loader.load()
    from stylish_tts.lib.config_loader import load_config_yaml
    import stylish_tts.train.train_context as tc       # (first: the reference's modules import each other in a cycle)
    import stylish_tts.train.stage_type as stage_type
    from stylish_tts.train.utils import DurationProcessor

    import stylish_tts_b200 as st
    from stylish_tts_b200 import discriminator as D
    from stylish_tts_b200 import spectral as b200
    from tests.golden.make_acoustic_step_golden import make_batch

    monkeypatch.setattr(stage_type, "calculate_mel", b200.calculate_mel)
    monkeypatch.setattr(stage_type, "log_norm", b200.log_norm)
    monkeypatch.setattr(stage_type, "multi_phase_loss", b200.multi_phase_loss)

    mc = ref_loader.model_config()                     # the reference's own pydantic ModelConfig
    cfg = load_config_yaml(os.path.join(ref_loader.REF_ROOT, "config", "config.yml"))
    nets = st.build_model(mc)
    for k in nets:
        nets[k].train()
    nets.speech_predictor.regularisers = False
    se_cfg = mc.style_encoder
    backward_calls = []
    train = types.SimpleNamespace(
        model=nets, model_config=mc, config=cfg, logger=logging.getLogger("dryrun"), writer=None,
        normalization=tc.NormalizationStats(),
        duration_processor=DurationProcessor(class_count=mc.duration_predictor.duration_classes,
                                             max_dur=mc.duration_predictor.max_duration),
        to_mel=b200.MelSpectrogram(n_mels=mc.n_mels, n_fft=mc.n_fft, win_length=mc.win_length,
                                   hop_length=mc.hop_length, sample_rate=mc.sample_rate),
        to_style_mel=b200.MelSpectrogram(n_mels=se_cfg.n_mels, n_fft=se_cfg.n_fft, win_length=se_cfg.win_length,
                                         hop_length=se_cfg.hop_length, sample_rate=mc.sample_rate),
        multi_spectrogram=b200.MultiSpectrogram(sample_rate=mc.sample_rate),
        stft_loss=b200.MultiResolutionSTFTLoss(sample_rate=mc.sample_rate),
        generator_loss=D.GeneratorLoss(mrd0=nets.mrd0, mrd1=nets.mrd1, mrd2=nets.mrd2, disc=nets.disc,
                                       pitch=nets.pitch_disc, duration=nets.dur_disc),
        # out of this engine's scope: the WavLM term is a stub that keeps the graph connected
        wavlm_loss=lambda target, pred: pred.float().mean() * 0.0,
        # what accelerate / the stage object do around a batch (train.py, stage.py:104-146)
        accelerator=types.SimpleNamespace(backward=lambda loss: (backward_calls.append(1), loss.backward())),
        stage=types.SimpleNamespace(optimizer=types.SimpleNamespace(zero_grad=lambda: None)))
    from stylish_tts.train.losses import DurationLoss

    classes = mc.duration_predictor.duration_classes
    train.duration_loss = DurationLoss(class_count=classes, weight=torch.ones(classes))
    raw, _ = make_batch()
    return stage_type, train, nets, mc, ref_loader.Munch(**raw), raw, backward_calls


def _trained(nets, keys):
    return [(f"{k}.{n}", p) for k in keys for n, p in nets[k].named_parameters()]


def test_reference_acoustic_step_drives_our_modules(dry_run, monkeypatch):
    stage_type, train, nets, mc, batch, raw, _ = _reference_side(monkeypatch)
    from stylish_tts.train.loss_log import build_loss_log

    sp, se = nets.speech_predictor, nets.speech_style_encoder
    log = build_loss_log(train)
    step = stage_type.AcousticStep(batch, train, log, use_predicted_pe=False, predict_audio=True)
    B, Fr = raw["pitch"].shape
    assert type(step.pred).__name__ == "DecoderPrediction" and step.pred.audio.shape == (B, 1, Fr * mc.hop_length)
    assert step.speech_style.shape == (B, mc.style_dim)
    assert step.mel.shape[:2] == (B, mc.n_mels) and step.energy.shape[0] == B
    assert len(step.pred_spec) == len(step.target_spec) == len(step.pred_phase) == len(step.pred_fft) == 3
    n_fwd = len(dry_run)
    step.mel_loss()
    step.multi_phase_loss()
    step.generator_loss(1)                              # stage_type.py:208-219: used=["mrd"], index=disc_index
    assert set(log.metrics) >= {"mel", "multi_phase", "generator"}
    total = log.backwards_loss()
    total.backward()
    cnt = collections.Counter(dry_run)
    # forward of the step went through our kernels' entry points ...
    for name in ("sty_spectrogram_fwd", "sty_mel_energy_fwd", "sty_conv1d_fwd", "sty_attention_lse_fwd",
                 "sty_source_fwd", "sty_istft_head_fwd", "sty_disc_first_fwd", "sty_disc_tail_fwd", "sty_tprls_fwd"):
        assert cnt[name] > 0, name
    # ... and the reference's backward reached every trained parameter through our autograd functions
    for name in ("sty_conv1d_wgrad", "sty_spectrogram_bwd", "sty_istft_head_bwd", "sty_disc_tail_bwd",
                 "sty_disc_first_dgrad", "sty_tprls_bwd"):
        assert cnt[name] > 0, name
    assert len(dry_run) > 2 * n_fwd > 0
    missing = [n for n, p in list(sp.named_parameters()) + list(se.named_parameters())
               if p.grad is None and "m_source.l_linear" not in n]
    assert not missing, missing[:5]
    for n, p in list(sp.named_parameters()) + list(se.named_parameters()):
        assert p.grad is None or p.grad.shape == p.shape, n
    # the generator step leaves the discriminators' parameters without gradients (constants of that step)
    assert all(p.grad is None for k in ("mrd0", "mrd1", "mrd2", "disc") for p in nets[k].parameters())


def test_reference_stage_functions_run_on_our_modules(dry_run, monkeypatch):
    """the reference's own per-batch stage functions, unmodified: ``train_acoustic`` (stage_type.py:346-373: mel,
    multi-phase, generator, SLM stub, backward) and ``train_textual`` (:416-447: predicted pitch / energy through
    ``pe_style_encoder`` + ``pitch_energy_predictor``, pitch discriminator term, pitch / energy losses)"""
    stage_type, train, nets, mc, batch, raw, backward_calls = _reference_side(monkeypatch)
    out = stage_type.train_acoustic(batch, nets, train, False, 2)
    assert len(backward_calls) == 1 and len(out) == 5
    log, target_spec, pred_spec, target_audio, pred_audio = out
    assert {"mel", "multi_phase", "generator", "slm"} <= set(log.metrics)
    assert len(target_spec) == len(pred_spec) == 3 and not pred_spec[0].requires_grad
    assert pred_audio[0].shape == raw["audio_gt"].shape
    missing = [n for n, p in _trained(nets, ("speech_predictor", "speech_style_encoder"))
               if p.grad is None and "m_source.l_linear" not in n]
    assert not missing, missing[:5]
    for k in nets:
        nets[k].zero_grad(set_to_none=True)
    n0 = len(dry_run)
    log, pitchcat, pred_pitchcat, _, _ = stage_type.train_textual(batch, nets, train, False, 0)
    assert len(backward_calls) == 2 and {"mel", "generator", "pitch", "energy"} <= set(log.metrics)
    B, Fr = raw["pitch"].shape
    assert pitchcat[0].shape == pred_pitchcat[0].shape == (B, 2, Fr)
    assert len(dry_run) > n0
    # the textual stage trains the pitch / energy predictor and its style encoder (stage_type.py:449-470)
    missing = [n for n, p in _trained(nets, ("pitch_energy_predictor", "pe_style_encoder")) if p.grad is None]
    assert not missing, missing[:5]


def test_reference_duration_stage_and_discriminator_half(dry_run, monkeypatch):
    """``train_duration`` (stage_type.py:494-553: duration style encoder + duration predictor, the reference's own
    DurationProcessor / DurationLoss, the duration discriminator term) and the discriminator half of
    ``Stage.train_batch`` (stage.py:125-146) with our ``DiscriminatorLoss`` called with the reference's keywords"""
    import math

    from stylish_tts_b200 import discriminator as D

    stage_type, train, nets, mc, batch, raw, backward_calls = _reference_side(monkeypatch)
    log, target_disc, pred_disc, _, _ = stage_type.train_duration(batch, nets, train, False, 0)
    assert len(backward_calls) == 1 and {"generator", "duration_ce", "duration"} <= set(log.metrics)
    assert target_disc[0].shape == pred_disc[0].shape == (raw["text"].shape[0], 1, raw["text"].shape[1])
    missing = [n for n, p in _trained(nets, ("duration_predictor", "duration_style_encoder")) if p.grad is None]
    assert not missing, missing[:5]
    assert all(p.grad is None for p in nets.dur_disc.parameters())
    # discriminator half of an acoustic batch, exactly as stage.py:125-146 calls it
    _, target_spec, pred_spec, target_audio, pred_audio = stage_type.train_acoustic(batch, nets, train, False, 1)
    for k in nets:
        nets[k].zero_grad(set_to_none=True)
    dl = D.DiscriminatorLoss(mrd0=nets.mrd0, mrd1=nets.mrd1, mrd2=nets.mrd2, disc=nets.disc, pitch=nets.pitch_disc,
                             duration=nets.dur_disc, device="cpu")
    d_loss = dl(target_list=target_spec, pred_list=pred_spec, target_audio=target_audio[0], pred_audio=pred_audio[0],
                used=["mrd0", "mrd1", "mrd2", "disc"], index=1)
    (d_loss * math.sqrt(raw["text"].shape[0])).backward()
    for key in ("mrd0", "mrd1", "mrd2", "disc"):
        missing = [n for n, p in nets[key].named_parameters() if p.grad is None]
        assert not missing, (key, missing[:5])
    assert all(p.grad is None for p in nets.speech_predictor.parameters())  # detached inputs
    assert isinstance(float(dl.get_disc_lr_multiplier("mrd1")), float)
