"""CPU oracle for the mel front-end and the multi-resolution STFT / phase losses.
TEST INFRASTRUCTURE ONLY (see oracle/speech_oracle.py for the import rules).

Plain PyTorch (CPU, fp32 or fp64) restatement of

* ``calculate_mel`` + torchaudio ``MelSpectrogram``      utils.py:825-834, train_context.py:155-169
* ``log_norm`` / ``raw_energy`` + log                    utils.py:73-85, stage_type.py:88-97
* ``MultiSpectrogram.calculate_single``                  multi_spectrogram.py:40-55
* ``MultiResolutionSTFTLoss``                            losses.py:27-38
* ``multi_phase_loss`` / ``differential_phase_loss``     losses.py:41-91
* ``LossLog.backwards_loss`` normalisation               loss_log.py:82-94

The mel filterbank arithmetic lives in a third-party dependency (torchaudio 2.x,
``torchaudio.functional.melscale_fbanks`` with mel_scale="htk", norm=None — the defaults both
reference call sites use); it is restated below from its published algorithm and pinned against
torchaudio itself through ``tests/golden/spectral.npz`` (made by tests/golden/make_spectral_golden.py
from the UNMODIFIED reference objects).  Everything is differentiable by autograd, which is what the
gradient-parity tests use.
"""
from __future__ import annotations

import math
from typing import List

import torch

RESOLUTIONS = ((512, 128, 512), (1024, 256, 1024), (2048, 512, 2048))  # multi_spectrogram.py:13-21


def hz_to_mel_htk(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_fbank(n_freqs: int, n_mels: int, sample_rate: int, f_min: float = 0.0, f_max=None,
              dtype=torch.float32) -> torch.Tensor:
    """(n_freqs, n_mels) triangular HTK filters, no area normalisation."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=dtype)
    m_pts = torch.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2, dtype=dtype)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    width = f_pts[1:] - f_pts[:-1]
    dist = f_pts[None, :] - freqs[:, None]
    falling = -dist[:, :-2] / width[:-1]
    rising = dist[:, 2:] / width[1:]
    return torch.clamp(torch.minimum(falling, rising), min=0.0)


def stft(audio, n_fft, hop, win):
    """torch.stft as the reference calls it (center=True, reflect, onesided, hann periodic)."""
    window = torch.hann_window(win, dtype=audio.dtype)
    return torch.stft(audio, n_fft=n_fft, hop_length=hop, win_length=win, window=window,
                      return_complex=True)


def mel_spectrogram(audio, *, n_fft, win, hop, n_mels, sample_rate):
    """torchaudio MelSpectrogram defaults: power 2, mel applied to the power spectrum."""
    p = stft(audio, n_fft, hop, win).abs() ** 2
    fb = mel_fbank(n_fft // 2 + 1, n_mels, sample_rate, dtype=audio.dtype)
    return torch.matmul(p.transpose(-1, -2), fb).transpose(-1, -2)


def calculate_mel(audio, *, n_fft, win, hop, n_mels, sample_rate, mean, std):
    """utils.py:825-834"""
    mel = mel_spectrogram(audio, n_fft=n_fft, win=win, hop=hop, n_mels=n_mels, sample_rate=sample_rate)
    mel = (torch.log(1e-5 + mel) - mean) / std
    return mel[:, :, : (mel.shape[-1] - mel.shape[-1] % 2)]


def log_energy(mel, mean, std):
    """stage_type.py:88-97: log(log_norm(mel.unsqueeze(1)).squeeze(1) + 1e-9) -> (B, F)"""
    x = torch.exp(mel.unsqueeze(1) * std + mean).norm(dim=2).squeeze(1)
    return torch.log(x + 1e-9)


def multi_spectrogram_single(audio, res, sample_rate, n_mels=128):
    """multi_spectrogram.py:40-55 -> (log1p mel-of-magnitude (B,1,M,N), masked phase (B,K,N), |X| (B,1,K,N))"""
    n_fft, hop, win = res
    X = stft(audio, n_fft, hop, win)
    fft_mag = X.abs()
    phase = (fft_mag > 1e-3).detach() * torch.angle(X)
    fb = mel_fbank(n_fft // 2 + 1, n_mels, sample_rate, dtype=audio.dtype)
    mag = torch.log1p(torch.matmul(fft_mag.transpose(-1, -2), fb).transpose(-1, -2))
    return mag.unsqueeze(1), phase, fft_mag.unsqueeze(1)


def spectral_convergence(target, pred):
    """losses.py:27-28"""
    return (target - pred).abs().sum() / (target.abs().sum() + 1e-6)


def stft_loss(target_list: List[torch.Tensor], pred_list: List[torch.Tensor]):
    """losses.py:30-38 (the value logged as "mel")"""
    loss = 0.0
    for t, p in zip(target_list, pred_list):
        loss = loss + spectral_convergence(t, p)
    return loss / len(target_list)


def anti_wrapping(d, w):
    """losses.py:41-43"""
    return (d - 2 * math.pi * torch.round(d / (2 * math.pi))).abs() * w


def differential_phase_loss(pred, target):
    """losses.py:46-84"""
    K = target.shape[1]
    base = math.exp(math.log(2.5) / (K // 2))
    w = torch.pow(base, torch.arange(K, dtype=pred.dtype))[None, :, None]
    loss = anti_wrapping(pred - target, w).mean()
    loss = loss + anti_wrapping(torch.diff(pred, dim=1) - torch.diff(target, dim=1), w[:, :-1]).mean()
    loss = loss + anti_wrapping(torch.diff(pred, dim=2) - torch.diff(target, dim=2), w).mean()
    return loss


def multi_phase_loss(pred_list, target_list):
    """losses.py:87-91"""
    loss = 0
    for p, t in zip(pred_list, target_list):
        loss = loss + differential_phase_loss(p, t)
    return loss / len(pred_list)


def backwards_total(losses: dict, weights: dict):
    """loss_log.py:82-94: sum_k w_k * loss_k / (loss_k.detach() + 1e-9) (none of the keys used here
    is in the un-normalised set {generator, align_loss})."""
    total = 0
    for k, v in losses.items():
        total = total + weights[k] * (v / (v.detach() + 1e-9))
    return total


def acoustic_spectral_losses(target_audio, pred_audio, sample_rate):
    """target/pred (B,L) -> dict(mel=, multi_phase=) exactly as AcousticStep wires them
    (stage_type.py:162-193)."""
    t_specs, p_specs, t_ph, p_ph = [], [], [], []
    for res in RESOLUTIONS:
        with torch.no_grad():
            tm, tp, _ = multi_spectrogram_single(target_audio, res, sample_rate)
        pm, pp, _ = multi_spectrogram_single(pred_audio, res, sample_rate)
        t_specs.append(tm), p_specs.append(pm), t_ph.append(tp), p_ph.append(pp)
    return dict(mel=stft_loss(t_specs, p_specs), multi_phase=multi_phase_loss(p_ph, t_ph))
