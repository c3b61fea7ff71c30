"""Mel front-end and multi-resolution STFT / phase losses on the sm_100a kernels
of ``csrc/spectral.cu`` (C ABI: ``sty_spectrogram_{fwd,bwd}``, ``sty_mel_energy_fwd``,
``sty_l1_sums_*``, ``sty_phase_loss_*``).

Drop-ins, same call signatures as the reference objects they replace:

=========================  ==========================================================
``MelSpectrogram``         torchaudio ``MelSpectrogram`` as built in train_context.py:155-169
``calculate_mel``          utils.py:825-834 (log-normalised mel, odd last frame dropped)
``log_norm``               utils.py:73-85 (+ the ``log(.+1e-9)`` of stage_type.py:97 via ``mel_energy``)
``MultiSpectrogram``       multi_spectrogram.py:25-81
``MultiResolutionSTFTLoss``  losses.py:17-38
``multi_phase_loss``       losses.py:41-91
=========================  ==========================================================

The predicted-audio branch is differentiable (``torch.autograd.Function`` whose backward is the
fused gradient kernel); the target branch never records a graph, like the reference's ``no_grad``.
There is no PyTorch/cuFFT fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib as L

_i32, _i64, _f32, _p = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class SpecArgs(C.Structure):
    """Mirror of ``sty_spectrogram_args``."""
    _fields_ = [
        ("audio", _p), ("audio_bs", _i64), ("window", _p), ("twiddle", _p),
        ("mag", _p), ("phase", _p), ("mel", _p),
        ("fb_start", _p), ("fb_len", _p), ("fb_off", _p), ("fb_w", _p),
        ("fbt_ptr", _p), ("fbt_mel", _p), ("fbt_w", _p),
        ("B", _i32), ("L", _i32), ("n_fft", _i32), ("hop", _i32), ("n_frames", _i32),
        ("n_mels", _i32), ("power", _i32), ("mel_mode", _i32),
        ("mel_eps", _f32), ("mel_mean", _f32), ("mel_std", _f32), ("phase_floor", _f32),
    ]


MEL_RAW, MEL_LOG1P, MEL_LOGNORM = 0, 1, 2


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """HTK mel triangles, norm=None: (n_freqs, n_mels) fp32.  Same arithmetic (fp32, same op order)
    as torchaudio.functional.melscale_fbanks, which both reference front-ends use with defaults
    (train_context.py:155-169, multi_spectrogram.py:31-37)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(zero, torch.min(down, up))


def centered_window(win_length: int, n_fft: int) -> torch.Tensor:
    """periodic hann(win_length) zero-padded to n_fft around the centre, as torch.stft does."""
    w = torch.hann_window(win_length, periodic=True, dtype=torch.float32)
    if win_length < n_fft:
        left = (n_fft - win_length) // 2
        w = torch.nn.functional.pad(w, (left, n_fft - win_length - left))
    return w


class SpectrogramPlan:
    """Device constants of one (n_fft, hop, window, mel filterbank) front-end."""

    def __init__(self, *, n_fft: int, hop: int, win_length: int, n_mels: int, sample_rate: int,
                 power: int, mel_mode: int = MEL_RAW, mel_eps: float = 1e-5, phase_floor: float = 1e-3,
                 f_min: float = 0.0, f_max: Optional[float] = None):
        self.n_fft, self.hop, self.win_length = n_fft, hop, win_length
        self.n_mels, self.sample_rate, self.power = n_mels, sample_rate, power
        self.mel_mode, self.mel_eps, self.phase_floor = mel_mode, mel_eps, phase_floor
        self.K = n_fft // 2 + 1
        self.window = centered_window(win_length, n_fft)
        q = torch.arange(n_fft // 2, dtype=torch.float64)
        ang = -2.0 * math.pi * q / n_fft
        self.twiddle = torch.stack([torch.cos(ang), torch.sin(ang)], 1).float().contiguous()
        fmax = float(sample_rate // 2) if f_max is None else f_max
        self.fb = melscale_fbanks(self.K, f_min, fmax, n_mels, sample_rate) if n_mels else None
        self._host = self._sparse(self.fb) if n_mels else None
        self._dev: Dict[str, dict] = {}

    @staticmethod
    def _sparse(fb: torch.Tensor):
        """triangles as (start, len, offset, weights) per filter + CSR by bin (for the backward)."""
        K, M = fb.shape
        start, length, off, w = [], [], [], []
        for m in range(M):
            nz = torch.nonzero(fb[:, m]).flatten()
            if nz.numel() == 0:
                start.append(0), length.append(0), off.append(len(w))
                continue
            s, e = int(nz[0]), int(nz[-1]) + 1
            start.append(s), length.append(e - s), off.append(len(w))
            w.extend(fb[s:e, m].tolist())
        ptr, mel, tw = [0], [], []
        for k in range(K):
            nz = torch.nonzero(fb[k]).flatten()
            mel.extend(int(v) for v in nz)
            tw.extend(fb[k, nz].tolist())
            ptr.append(len(mel))
        i32 = lambda v: torch.tensor(v if v else [0], dtype=torch.int32)
        f32 = lambda v: torch.tensor(v if v else [0.0], dtype=torch.float32)
        return dict(fb_start=i32(start), fb_len=i32(length), fb_off=i32(off), fb_w=f32(w),
                    fbt_ptr=i32(ptr), fbt_mel=i32(mel), fbt_w=f32(tw))

    def dev(self, device) -> dict:
        key = str(device)
        if key not in self._dev:
            d = dict(window=self.window.to(device), twiddle=self.twiddle.to(device))
            if self._host:
                d.update({k: v.to(device) for k, v in self._host.items()})
            self._dev[key] = d
        return self._dev[key]

    def n_frames(self, L_: int) -> int:
        return L_ // self.hop + 1

    def args(self, audio: torch.Tensor, n_frames: int, *, mean=0.0, std=1.0) -> SpecArgs:
        L.require_cuda(audio, "the audio of the spectral front-end")
        assert audio.dim() == 2 and audio.dtype == torch.float32 and audio.stride(1) == 1
        d = self.dev(audio.device)
        a = SpecArgs()
        a.audio, a.audio_bs = audio.data_ptr(), audio.stride(0)
        a.window, a.twiddle = d["window"].data_ptr(), d["twiddle"].data_ptr()
        if self.n_mels:
            for k in ("fb_start", "fb_len", "fb_off", "fb_w", "fbt_ptr", "fbt_mel", "fbt_w"):
                setattr(a, k, d[k].data_ptr())
        a.B, a.L, a.n_fft, a.hop, a.n_frames = audio.shape[0], audio.shape[1], self.n_fft, self.hop, n_frames
        a.n_mels, a.power, a.mel_mode = self.n_mels, self.power, self.mel_mode
        a.mel_eps, a.mel_mean, a.mel_std, a.phase_floor = self.mel_eps, mean, std, self.phase_floor
        return a

    def forward(self, audio, *, want_mag=False, want_phase=False, want_mel=True, n_frames=None,
                mean=0.0, std=1.0):
        nf = self.n_frames(audio.shape[1]) if n_frames is None else n_frames
        B = audio.shape[0]
        new = lambda c: torch.empty((B, c, nf), device=audio.device, dtype=torch.float32)
        mag = new(self.K) if want_mag else None
        phase = new(self.K) if want_phase else None
        mel = new(self.n_mels) if want_mel else None
        a = self.args(audio, nf, mean=mean, std=std)
        a.mag, a.phase, a.mel = L.ptr(mag), L.ptr(phase), L.ptr(mel)
        L.call("sty_spectrogram_fwd", C.byref(a), L.stream_ptr())
        return mag, phase, mel

    def backward(self, audio, n_frames, d_mag, d_phase, d_mel, *, mean=0.0, std=1.0):
        d_audio = torch.zeros_like(audio, memory_format=torch.contiguous_format)
        a = self.args(audio, n_frames, mean=mean, std=std)
        L.call("sty_spectrogram_bwd", C.byref(a), L.ptr(d_mel), L.ptr(d_phase), L.ptr(d_mag),
               d_audio.data_ptr(), d_audio.stride(0), L.stream_ptr())
        return d_audio


class _SpectrogramFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, audio, plan: SpectrogramPlan, want_mag: bool, want_phase: bool, want_mel: bool):
        audio = audio.contiguous()
        mag, phase, mel = plan.forward(audio, want_mag=want_mag, want_phase=want_phase, want_mel=want_mel)
        ctx.plan = plan
        ctx.n_frames = plan.n_frames(audio.shape[1])
        ctx.save_for_backward(audio)
        outs = tuple(t if t is not None else audio.new_empty(0) for t in (mag, phase, mel))
        ctx.mark_non_differentiable(*[o for o, w in zip(outs, (want_mag, want_phase, want_mel)) if not w])
        return outs

    @staticmethod
    def backward(ctx, d_mag, d_phase, d_mel):
        (audio,) = ctx.saved_tensors
        c = lambda g: None if g is None or g.numel() == 0 else g.contiguous()
        return ctx.plan.backward(audio, ctx.n_frames, c(d_mag), c(d_phase), c(d_mel)), None, None, None, None


def _as_2d(audio: torch.Tensor) -> Tuple[torch.Tensor, tuple]:
    lead = audio.shape[:-1]
    return audio.reshape(-1, audio.shape[-1]).to(torch.float32).contiguous(), tuple(lead)


class MelSpectrogram(nn.Module):
    """``to_mel`` / ``to_style_mel``: power mel spectrogram (..., time) -> (..., n_mels, frames) with
    torchaudio's defaults (hann, center/reflect, power 2, HTK, norm None, f_max sr/2)."""

    def __init__(self, *, n_mels, n_fft, win_length, hop_length, sample_rate):
        super().__init__()
        self.n_mels, self.n_fft, self.win_length = n_mels, n_fft, win_length
        self.hop_length, self.sample_rate = hop_length, sample_rate
        self.plan = SpectrogramPlan(n_fft=n_fft, hop=hop_length, win_length=win_length, n_mels=n_mels,
                                    sample_rate=sample_rate, power=2, mel_mode=MEL_RAW)
        self.plan_lognorm = SpectrogramPlan(n_fft=n_fft, hop=hop_length, win_length=win_length,
                                            n_mels=n_mels, sample_rate=sample_rate, power=2,
                                            mel_mode=MEL_LOGNORM, mel_eps=1e-5)

    def forward(self, audio):
        a2, lead = _as_2d(audio)
        if torch.is_grad_enabled() and audio.requires_grad:
            mel = _SpectrogramFn.apply(a2, self.plan, False, False, True)[2]
        else:
            mel = self.plan.forward(a2)[2]
        return mel.reshape(*lead, self.n_mels, mel.shape[-1])


@torch.no_grad()
def calculate_mel(audio, to_mel: MelSpectrogram, mean, std):
    """utils.py:825-834 in one kernel: STFT -> power -> mel -> (log(1e-5+.)-mean)/std, and the odd
    last frame is simply not computed."""
    a2, lead = _as_2d(audio)
    nf = to_mel.plan.n_frames(a2.shape[1])
    nf -= nf % 2
    mel = to_mel.plan_lognorm.forward(a2, n_frames=nf, mean=float(mean), std=float(std))[2]
    mel = mel.reshape(*lead, to_mel.n_mels, nf)
    length = torch.full([audio.shape[0]], nf, dtype=torch.long, device=audio.device)
    return mel, length


@torch.no_grad()
def mel_energy(mel, mean, std):
    """log(||exp(mel*std+mean)||_2 + 1e-9) over mel bins: (B, n_mels, F) -> (B, F)
    (log_norm + the log of stage_type.py:88-97)."""
    mel = mel.to(torch.float32).contiguous()
    L.require_cuda(mel, "the mel of mel_energy")
    B, M, Fr = mel.shape
    out = torch.empty((B, Fr), device=mel.device, dtype=torch.float32)
    L.call("sty_mel_energy_fwd", mel.data_ptr(), out.data_ptr(), B, M, Fr, float(mean), float(std),
           L.stream_ptr())
    return out


def log_norm(x, mean, std, dim=2):
    """utils.py:73-80: (B,1,n_mels,F) normalised log-mel -> L2 norm over mel bins (B,1,F)."""
    assert dim == 2 and x.dim() == 4 and x.shape[1] == 1
    return torch.exp(mel_energy(x[:, 0], mean, std)).sub_(1e-9).unsqueeze(1)


class Resolution:
    def __init__(self, *, fft, hop, window):
        self.fft, self.hop, self.window = fft, hop, window


# multi_spectrogram.py:13-21
resolutions = [Resolution(fft=512, hop=128, window=512), Resolution(fft=1024, hop=256, window=1024),
               Resolution(fft=2048, hop=512, window=2048)]
multi_spectrogram_count = len(resolutions)


class MultiSpectrogram(nn.Module):
    """multi_spectrogram.py:25-81: per resolution log1p(mel(|X|)) (B,1,128,N), masked phase (B,K,N) and
    |X| (B,1,K,N), for the target (no graph) and the prediction (differentiable)."""

    def __init__(self, resolutions=resolutions, *, sample_rate, n_mels=128):
        super().__init__()
        self.resolutions = list(resolutions)
        self.plans = [SpectrogramPlan(n_fft=r.fft, hop=r.hop, win_length=r.window, n_mels=n_mels,
                                      sample_rate=sample_rate, power=1, mel_mode=MEL_LOG1P,
                                      phase_floor=1e-3) for r in self.resolutions]

    def calculate_single(self, audio, index, item=None):
        a2, _ = _as_2d(audio)
        plan = self.plans[index]
        if torch.is_grad_enabled() and audio.requires_grad:
            fft_mag, phase, mag = _SpectrogramFn.apply(a2, plan, True, True, True)
        else:
            fft_mag, phase, mag = plan.forward(a2, want_mag=True, want_phase=True, want_mel=True)
        return mag.unsqueeze(1), phase, fft_mag.unsqueeze(1)

    def forward(self, *, target, pred):
        res = ([], [], [], [], [], [])
        for index in range(len(self.resolutions)):
            with torch.no_grad():
                t_mag, t_phase, t_fft = self.calculate_single(target, index)
            p_mag, p_phase, p_fft = self.calculate_single(pred, index)
            for lst, v in zip(res, (t_mag, p_mag, t_phase, p_phase, t_fft, p_fft)):
                lst.append(v)
        return res


class _L1RatioFn(torch.autograd.Function):
    """sum|t-p| / (sum|t| + 1e-6)   (losses.py:27-28)"""

    @staticmethod
    def forward(ctx, target, pred):
        target, pred = target.contiguous(), pred.contiguous()
        L.require_cuda(pred, "the STFT loss input")
        sums = torch.zeros(2, device=pred.device, dtype=torch.float32)
        L.call("sty_l1_sums_fwd", target.data_ptr(), pred.data_ptr(), pred.numel(), sums.data_ptr(),
               L.stream_ptr())
        ctx.save_for_backward(target, pred, sums)
        return sums[0] / (sums[1] + 1e-6)

    @staticmethod
    def backward(ctx, g):
        target, pred, sums = ctx.saved_tensors
        coef = (g / (sums[1] + 1e-6)).reshape(1).contiguous()
        d_pred = torch.empty_like(pred)
        L.call("sty_l1_sums_bwd", target.data_ptr(), pred.data_ptr(), pred.numel(), coef.data_ptr(),
               d_pred.data_ptr(), L.stream_ptr())
        return None, d_pred


class MultiResolutionSTFTLoss(nn.Module):
    """losses.py:17-38."""

    def __init__(self, *, sample_rate=None):
        super().__init__()

    def spectral_convergence_loss(self, target, pred):
        return _L1RatioFn.apply(target, pred)

    def forward(self, *, target_list, pred_list, log=None):
        loss = 0.0
        for target, pred in zip(target_list, pred_list):
            loss = loss + self.spectral_convergence_loss(target, pred)
        loss = loss / len(target_list)
        if log is not None:
            log.add_loss("mel", loss)
        return loss


class _PhaseLossFn(torch.autograd.Function):
    """differential_phase_loss (losses.py:46-84) on (B,K,N) phases."""

    @staticmethod
    def forward(ctx, pred, target):
        pred, target = pred.contiguous(), target.contiguous()
        L.require_cuda(pred, "the phase loss input")
        B, K, N = pred.shape
        sums = torch.zeros(3, device=pred.device, dtype=torch.float32)
        L.call("sty_phase_loss_fwd", pred.data_ptr(), target.data_ptr(), B, K, N, sums.data_ptr(),
               L.stream_ptr())
        ctx.save_for_backward(pred, target)
        # counts as Python floats: no host->device copy, so the call can be captured in a CUDA graph
        return sums[0] / float(B * K * N) + sums[1] / float(B * (K - 1) * N) + sums[2] / float(B * K * (N - 1))

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        B, K, N = pred.shape
        d_pred = torch.empty_like(pred)
        coef = g.reshape(1).to(torch.float32).contiguous()
        L.call("sty_phase_loss_bwd", pred.data_ptr(), target.data_ptr(), B, K, N, coef.data_ptr(),
               d_pred.data_ptr(), L.stream_ptr())
        return d_pred, None


def differential_phase_loss(pred, target, n_fft=None):
    return _PhaseLossFn.apply(pred, target)


def multi_phase_loss(pred_list: List[torch.Tensor], target_list: List[torch.Tensor], n_fft=None):
    loss = 0
    for pred, target in zip(pred_list, target_list):
        loss = loss + differential_phase_loss(pred, target, n_fft)
    return loss / len(pred_list)
