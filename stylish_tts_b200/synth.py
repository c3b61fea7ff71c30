"""Seeded synthetic weights and inputs (no datasets or checkpoints offline).

Used by bench.py, smoke() and the tests.  Shapes follow SURVEY.md §8(d):
tokens ``randint(1,178)`` with pad id 0 at both ends, 3 frames per token (+1
every 9th), pitch U(80,280) Hz with unvoiced runs, energy N(0,1), style N(0,1).
Everything is generated on the CPU with a ``torch.Generator`` so that the build
container (golden fixtures) and the GPU box regenerate identical tensors.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import nn


def randomize_(module: nn.Module, seed: int = 0) -> nn.Module:
    """Overwrite every parameter/buffer with seeded, non-degenerate values
    (zero-initialised reference parameters such as GRN gamma/beta or the prenet
    projection would otherwise hide bugs in those paths)."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    new: Dict[str, torch.Tensor] = {}

    def randn(shape, scale=1.0):
        return torch.randn(shape, generator=g) * scale

    for name in sorted(sd.keys()):
        t = sd[name]
        leaf = name.split(".")[-1]
        if not t.is_floating_point():
            new[name] = t.clone()
        elif ".stft." in name:
            new[name] = t.clone()  # fixed DFT bases
        elif leaf == "running_var":
            new[name] = 1.0 + 0.5 * torch.rand(t.shape, generator=g)
        elif leaf == "running_mean":
            new[name] = randn(t.shape, 0.1)
        elif leaf == "original0":
            continue  # weight_norm g: set after its v below
        elif leaf == "snake" or ".alpha" in name:
            new[name] = 0.75 + 0.5 * torch.rand(t.shape, generator=g)
        elif name.endswith("grn.gamma"):
            new[name] = randn(t.shape, 0.3)
        elif name.endswith("grn.beta"):
            new[name] = randn(t.shape, 0.1)
        elif leaf == "gamma" or (leaf == "weight" and t.dim() == 1):
            new[name] = 1.0 + randn(t.shape, 0.1)
        elif leaf in ("beta", "bias"):
            new[name] = randn(t.shape, 0.05)
        elif name.endswith("emb.weight"):
            new[name] = randn(t.shape, t.shape[1] ** -0.5)
        elif name.endswith(".fc.weight"):
            new[name] = randn(t.shape, 0.5 / math.sqrt(t.shape[1]))
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            new[name] = randn(t.shape, 1.0 / math.sqrt(fan_in))
        else:
            new[name] = randn(t.shape, 0.1)
    for name in sorted(sd.keys()):
        if name.endswith("original0"):
            v = new[name[:-1] + "1"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(sd[name].shape)
            new[name] = norm * (1.0 + 0.1 * torch.randn(sd[name].shape, generator=g))
    module.load_state_dict(new, strict=True)
    return module


@torch.no_grad()
def condition_phase_head_(speech_predictor: nn.Module, shift: float = 3.0) -> nn.Module:
    """Move the real part of the phase head away from zero (bias += shift).  phase = atan2(imag, real) has the
    derivative (-imag, real)/r^2; with purely random weights many bins have r ~ 0 and the gradient of everything
    upstream becomes ill-conditioned (fp32 vs fp64 of the SAME formula: 1.4e-2).  With the shift that difference
    is 3e-6, so gradient parity can be asserted at kernel accuracy.  Used by the gradient goldens / tests."""
    speech_predictor.generator.basegen.phase_output_real_conv.bias += shift
    return speech_predictor


@torch.no_grad()
def converge_spectral_(module: nn.Module, iters: int = 30) -> nn.Module:
    """Run the spectral-norm power iteration on every (weight_orig, weight_u, weight_v) triple so that a
    freshly initialised module is in the regime a trained checkpoint is in (sigma ~ largest singular value;
    a fresh module has random u, v and sigma ~ 0, i.e. weights blown up by 1e3+ per layer)."""
    sd = dict(module.named_parameters())
    bufs = dict(module.named_buffers())
    for name, W in sd.items():
        if not name.endswith(".weight_orig"):
            continue
        p = name[:-len(".weight_orig")]
        u, v = bufs[p + ".weight_u"], bufs[p + ".weight_v"]
        Wm = W.reshape(W.shape[0], -1)
        for _ in range(iters):
            v.copy_(torch.nn.functional.normalize(torch.mv(Wm.t(), u), dim=0, eps=1e-12))
            u.copy_(torch.nn.functional.normalize(torch.mv(Wm, v), dim=0, eps=1e-12))
    return module


def soft_alignment(duration: torch.Tensor) -> torch.Tensor:
    """Input generator: the soft (B,T,F) alignment matrix the reference's data
    path hands to speech_predictor (shape/meaning of utils.py:752-791)."""
    total = int(duration.sum(dim=1).round().max().item())
    upper = torch.cumsum(duration, dim=1)
    lower = upper - duration
    mid = ((lower + upper) / 2).unsqueeze(2)
    frames = torch.arange(total).view(1, 1, -1)
    x = frames - mid
    a = 1 - (x * 2 / (duration.unsqueeze(2) + 6)) ** 2
    keep = (frames > (lower - 3).unsqueeze(2)) * (frames < (upper + 3).unsqueeze(2))
    a = torch.clamp(a * keep, min=0.0)
    return torch.softmax(a, dim=1)


def speech_inputs(batch: int, tokens: int, *, seed: int = 1, ragged: bool = False, all_padded: bool = False,
                  frames_per_token: int = 3, hop: int = 300, n_symbols: int = 178,
                  style_dim: int = 64, harmonics: int = 9) -> Dict[str, torch.Tensor]:
    """Synthetic inputs of ``speech_predictor.forward`` for ``batch`` utterances of
    ``tokens`` symbols (incl. the two pads).  Returns CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    T = tokens
    texts = torch.randint(1, n_symbols, (batch, T), generator=g)
    if ragged:
        lengths = torch.randint(max(T // 2, 4), T + 1, (batch,), generator=g)
        # all_padded: no utterance fills the token axis.  (With a full-length row the style channels of
        # `prosody @ alignment` are EXACTLY constant over time and the towers' InstanceNorm turns their rounding
        # noise into O(1e-4) output noise — the reference's own fp32 is 8e-5 from fp64 there, 1e-5 otherwise.)
        lengths[0] = T - 3 if all_padded else T
    else:
        lengths = torch.full((batch,), T, dtype=torch.long)
    for b in range(batch):
        texts[b, 0] = 0
        texts[b, int(lengths[b]) - 1:] = 0
    dur = torch.full((batch, T), float(frames_per_token))
    dur[:, ::9] += 1.0
    tmask = torch.arange(T).unsqueeze(0) < lengths.unsqueeze(1)
    dur = dur * tmask
    alignment = soft_alignment(dur)
    F_ = alignment.shape[2]
    pitch = 80.0 + 200.0 * torch.rand(batch, F_, generator=g)
    # smooth a little so neighbouring frames are correlated, then carve unvoiced runs
    pitch = torch.nn.functional.avg_pool1d(pitch.unsqueeze(1), 5, 1, 2,
                                           count_include_pad=False).squeeze(1)
    run = max(F_ // 10, 1)
    for b in range(batch):
        for _ in range(2):
            s = int(torch.randint(0, max(F_ - run, 1), (1,), generator=g))
            pitch[b, s:s + run] = 0.0
    energy = torch.randn(batch, F_, generator=g)
    voiced = (pitch > 20).float()
    style = torch.randn(batch, style_dim, generator=g)
    L = F_ * hop
    draws = {
        "rand_ini": torch.rand(batch, harmonics, generator=g),
        "noise": torch.randn(batch, L, harmonics, generator=g),
    }
    return dict(texts=texts, text_lengths=lengths, alignment=alignment, pitch=pitch,
                energy=energy, voiced=voiced, style=style, denormal_pitch=pitch.clone(),
                draws=draws)
