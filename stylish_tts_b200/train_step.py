"""The acoustic training step on the CUDA engine: the forward graph the reference builds in
``AcousticStep.__init__`` with ``use_predicted_pe=False, predict_audio=True`` (stage_type.py:61-180) plus its
``mel`` and ``multi_phase`` losses with ``LossLog.backwards_loss`` normalisation (stage_type.py:170-193,
loss_log.py:82-94) — SURVEY §8a row A3 / §8d config 3 — and, when a ``generator_loss`` is passed, the adversarial
term of ``step.generator_loss`` (stage_type.py:208-219); ``discriminator_step`` is the second half of
``Stage.train_batch`` (stage.py:125-146).  The SLM (WavLM) term is outside the hot path.

    fe = FrontEnd(model_config, mel_log_mean, mel_log_std)
    out = acoustic_step(batch, nets, fe)      # out.total is differentiable; out.pred.audio (B,1,L)
    out.total.backward(); optimizer.step()
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import _lib as L
from . import spectral
from .optim import acoustic_losses


class FrontEnd:
    """the TrainContext objects the step reads (train_context.py:155-178): to_mel, to_style_mel,
    multi_spectrogram, stft_loss and the mel normalisation constants"""

    def __init__(self, model_config, mel_log_mean=-4.0, mel_log_std=4.0):
        mc = model_config
        self.mean, self.std = float(mel_log_mean), float(mel_log_std)
        self.to_mel = spectral.MelSpectrogram(n_mels=mc.n_mels, n_fft=mc.n_fft, win_length=mc.win_length,
                                              hop_length=mc.hop_length, sample_rate=mc.sample_rate)
        se = mc.style_encoder
        self.to_style_mel = spectral.MelSpectrogram(n_mels=se.n_mels, n_fft=se.n_fft, win_length=se.win_length,
                                                    hop_length=se.hop_length, sample_rate=mc.sample_rate)
        self.multi_spectrogram = spectral.MultiSpectrogram(sample_rate=mc.sample_rate)
        self.stft_loss = spectral.MultiResolutionSTFTLoss(sample_rate=mc.sample_rate)


def alignment_from_durations(durations: torch.Tensor, frames: int = 0, multiplier: int = 1) -> torch.Tensor:
    """DurationProcessor.duration_to_alignment (utils.py:752-791) for the integer durations of a batch
    (stage_type.py:99-101): (B,T) -> soft alignment (B,T,F).  `frames` skips the device->host read of
    round(max sum) that the reference does with .item() (utils.py:759); `multiplier` =
    ModelConfig.coarse_multiplier scales durations and frame count (utils.py:759-761)."""
    dur = durations.to(torch.float32).contiguous()
    if multiplier != 1:
        dur = (dur * float(multiplier)).contiguous()
    B, Tn = dur.shape
    Fr = frames or int(dur.sum(dim=1).round().max().item())
    al = torch.empty((B, Tn, Fr), device=dur.device, dtype=torch.float32)
    L.call("sty_alignment_fwd", dur.data_ptr(), al.data_ptr(), B, Tn, Fr, L.stream_ptr())
    return al


def acoustic_step(batch, nets, fe: FrontEnd, *, w_mel=5.0, w_phase=8.0, source_draws=None, prior=None,
                  generator_loss=None, w_generator=1.0):
    """batch: audio_gt (B,L), text (B,T), text_length (B,), pitch (B,F), alignment (B,1,T) integer durations
    (the reference's collated batch, stage_type.py:61-106).  Returns total / mel / multi_phase losses and
    the prediction.  `prior` = (har_spec, har_phase) injects the harmonic prior (parity tests)."""
    audio_gt = batch.audio_gt
    cm = int(getattr(getattr(nets.speech_predictor, "model_config", None), "coarse_multiplier", 1) or 1)
    with torch.no_grad():
        mel, _ = spectral.calculate_mel(audio_gt, fe.to_mel, fe.mean, fe.std)
        style_mel, _ = spectral.calculate_mel(audio_gt, fe.to_style_mel, fe.mean, fe.std)
        energy = spectral.mel_energy(mel, fe.mean, fe.std)
        pitch = batch.pitch
        # alignment_fine of stage_type.py:111-115 (the speech predictor reads the fine one)
        alignment = alignment_from_durations(batch.alignment[:, 0, :], frames=pitch.shape[1], multiplier=cm)
        voiced = (pitch > 20).float()
    style = nets.speech_style_encoder(style_mel.unsqueeze(1))
    pred = nets.speech_predictor(batch.text, batch.text_length, alignment, pitch, energy, voiced, style, pitch,
                                 source_draws=source_draws, prior=prior)
    total, mel_loss, phase_loss, t_fft, p_fft = acoustic_losses(pred.audio.squeeze(1), audio_gt, fe.multi_spectrogram,
                                                                fe.stft_loss, w_mel=w_mel, w_phase=w_phase,
                                                                return_fft=True)
    gen_loss = None
    if generator_loss is not None:
        # step.generator_loss (stage_type.py:208-219): the adversarial term enters backwards_loss RAW (loss_log.py:85-86)
        gen_loss = generator_loss(target_list=t_fft, pred_list=p_fft, target_audio=audio_gt,
                                  pred_audio=pred.audio.squeeze(1))
        total = total + w_generator * gen_loss
    return SimpleNamespace(total=total, mel=mel_loss, multi_phase=phase_loss, generator=gen_loss, pred=pred,
                           style=style, energy=energy, mel_target=mel, target_fft=t_fft, pred_fft=p_fft)


def discriminator_step(out, batch, discriminator_loss, optimizers, *, disc_index: int, lr_source=None):
    """The second half of Stage.train_batch (stage.py:125-146): discriminator loss on the DETACHED spectrograms of
    the step just taken, backward of d_loss * sqrt(B), optimizer step of ``mrd{disc_index}`` (and ``disc`` when the
    loss carries a waveform discriminator).  `optimizers`: dict key -> FlatAdamW.  `lr_source`: the generator's
    FlatAdamW — when given, each stepped discriminator's learning rate is lr_gen x its gap-aware multiplier first
    (optimizers.py:54-65), all on the device."""
    import math

    scale = math.sqrt(batch.text.shape[0])
    if hasattr(discriminator_loss, "discriminator_backward"):
        # discriminator.AdversarialTerms: the scores of this batch are already on its tape (one evaluation serves
        # the generator term and this loss); only the stepped discriminator is back-propagated
        d_loss = discriminator_loss.discriminator_backward(disc_index, scale)
    else:
        d_loss = discriminator_loss(target_list=[t.detach() for t in out.target_fft],
                                    pred_list=[p.detach() for p in out.pred_fft], target_audio=batch.audio_gt,
                                    pred_audio=out.pred.audio.detach().squeeze(1))
        (d_loss * scale).backward()
    keys = [f"mrd{disc_index}"] + (["disc"] if "disc" in optimizers else [])
    for k in keys:
        if lr_source is not None:
            discriminator_loss.lr_control[k].apply(lr_source, optimizers[k])
        optimizers[k].step()
    for opt in optimizers.values():
        opt.zero_grad()
    return d_loss.detach()
