"""CPU oracle for the discriminators and adversarial losses of the acoustic stage — SURVEY §8(f) rank 1, the row
the CUDA path widens into next.  TEST INFRASTRUCTURE ONLY (see speech_oracle.py for the rules).

Functional restatement against a state dict with the reference's key names; pinned by
tests/golden/make_disc_golden.py (the UNMODIFIED reference modules) / tests/test_disc_oracle.py.
No CUDA counterpart exists yet: this is step (a) "oracle and boundary" of that row.
Paths relative to /root/reference/src/stylish_tts/train/.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

from .speech_oracle import wn_weight

SD = Dict[str, torch.Tensor]
DISC_WEIGHT = 3  # losses.py:14


def _wn_conv2d(sd: SD, prefix: str, x, *, stride=(1, 1), padding=(0, 0)):
    return F.conv2d(x, wn_weight(sd, prefix), sd.get(prefix + ".bias"), stride=stride, padding=padding)


def spec_discriminator(sd: SD, y) -> List[torch.Tensor]:
    """models/discriminator.py:13-68 — five weight-normed Conv2d (3x9, stride (1,2) for the middle three; last 3x3)
    with LeakyReLU(0.1), each followed by its own 3x3 output conv; returns the five flattened score maps.
    y: (B, 1, bins, frames) magnitude spectrogram."""
    cfg = [((1, 1), (1, 4)), ((1, 2), (1, 4)), ((1, 2), (1, 4)), ((1, 2), (1, 4)), ((1, 1), (1, 1))]
    out = []
    for i, (stride, pad) in enumerate(cfg):
        y = F.leaky_relu(_wn_conv2d(sd, f"discriminators.{i}", y, stride=stride, padding=pad), 0.1)
        out.append(torch.flatten(_wn_conv2d(sd, f"out.{i}", y, padding=(1, 1)), 1, -1))
    return out


def pitch_discriminator(sd: SD, y, kernel: int) -> List[torch.Tensor]:
    """models/pitch_discriminator.py:7-68 — five weight-normed Conv1d (same padding) + LeakyReLU(0.1), each with its
    own output conv; `pitch_disc` (dim_in 2, kernel 21: pitch & energy curves) and `dur_disc` (dim_in 1, kernel 5:
    durations), models/models.py:81-82.  y: (B, dim_in, T) -> five (B, T) score maps"""
    out = []
    for i in range(5):
        y = F.leaky_relu(F.conv1d(y, wn_weight(sd, f"discriminators.{i}"), sd[f"discriminators.{i}.bias"],
                                  padding=kernel // 2), 0.1)
        out.append(torch.flatten(F.conv1d(y, wn_weight(sd, f"out.{i}"), sd[f"out.{i}.bias"], padding=kernel // 2),
                                 1, -1))
    return out


def _cf_block(sd: SD, prefix: str, x, *, kernel, stride=1, groups=1, bn_training=True):
    """ContextFreeBlock models/discriminator.py:94-116: Conv1d -> BatchNorm1d -> GELU(erf)"""
    x = F.conv1d(x, sd[prefix + ".net.0.weight"], sd.get(prefix + ".net.0.bias"), stride=stride,
                 padding=kernel // 2, groups=groups)
    bn = prefix + ".net.1"
    if bn_training:
        x = F.batch_norm(x, None, None, sd[bn + ".weight"], sd[bn + ".bias"], training=True, eps=1e-5)
    else:
        x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                         training=False, eps=1e-5)
    return F.gelu(x)


def context_free_discriminator(sd: SD, x, bn_training=True) -> List[torch.Tensor]:
    """models/discriminator.py:119-175 — waveform discriminator over 1024-sample windows (hop 512): four strided
    conv blocks, squeeze-excite style gate, a temporal (k7, k3; 8 groups) and a 'spectral' (k1, 8 groups) branch,
    fusion, 1x1 head.  x: (B, L) audio -> [(B, windows * features)]"""
    B = x.shape[0]
    x = x.unfold(1, 1024, 512)
    steps = x.shape[1]
    x = x.reshape(B * steps, 1, 1024)
    for i, (k, s) in enumerate(((11, 4), (11, 4), (7, 2), (5, 2))):
        x = _cf_block(sd, f"conv.{i}", x, kernel=k, stride=s, bn_training=bn_training)
    gate = torch.sigmoid(F.conv1d(x.mean(dim=2, keepdim=True), sd["attn.1.weight"], sd["attn.1.bias"]))
    x = x * gate
    t = _cf_block(sd, "temporal.0", x, kernel=7, groups=8, bn_training=bn_training)
    t = _cf_block(sd, "temporal.1", t, kernel=3, groups=8, bn_training=bn_training)
    s_ = _cf_block(sd, "spectral.0", x, kernel=1, groups=8, bn_training=bn_training)
    s_ = _cf_block(sd, "spectral.1", s_, kernel=1, groups=8, bn_training=bn_training)
    x = _cf_block(sd, "fusion", torch.cat([t, s_], dim=1), kernel=1, bn_training=bn_training)
    x = F.conv1d(x, sd["last.0.weight"], sd["last.0.bias"])
    x = F.conv1d(torch.relu(x), sd["last.2.weight"], sd["last.2.bias"])
    return [x.reshape(B, steps, -1).reshape(B, -1)]  # "(b t) c f -> b (t c f)"


# ----------------------------------------------------------------------------- losses (losses.py:233-373)
def discriminator_loss(real: List[torch.Tensor], gen: List[torch.Tensor]):
    """LSGAN discriminator terms, losses.py:251-262"""
    return sum(torch.mean((1 - dr) ** 2) + torch.mean(dg ** 2) for dr, dg in zip(real, gen))


def tprls_discriminator(real: List[torch.Tensor], gen: List[torch.Tensor], tau=0.04):
    """truncated pointwise relativistic least squares, discriminator side, losses.py:264-278"""
    loss = 0
    for dr, dg in zip(real, gen):
        m = torch.median(dr - dg)
        sel = (((dr - dg) - m) ** 2)[dr < dg + m]
        l_rel = torch.sum(sel) / (sel.numel() + 1e-9)
        loss = loss + tau - F.relu(tau - l_rel)
    return loss


def tprls_generator(real: List[torch.Tensor], gen: List[torch.Tensor], tau=0.04):
    """generator side, losses.py:356-363.  The reference zips (real, gen) into names (dg, dr), i.e. inside the loop
    'dr' is the GENERATED score and 'dg' the REAL one — restated literally; the mean over an empty selection is
    NaN there as well and relu(tau - NaN) propagates it (torch semantics kept)."""
    loss = 0
    for dg, dr in zip(real, gen):
        m = torch.median(dr - dg)
        l_rel = torch.mean((((dr - dg) - m) ** 2)[dr < dg + m])
        loss = loss + tau - F.relu(tau - l_rel)
    return loss


def generator_loss(gen: List[torch.Tensor]):
    """losses.py:339-343"""
    return sum(torch.mean((1 - dg) ** 2) for dg in gen)


def helper_discriminator(disc_fn, target, pred):
    """DiscriminatorLossHelper.forward losses.py:280-288 -> (loss, plain LSGAN part for the lr controller)"""
    real, gen = disc_fn(target), disc_fn(pred)
    d = discriminator_loss(real, gen)
    return d + tprls_discriminator(real, gen), d


def helper_generator(disc_fn, target, pred):
    """GeneratorLossHelper.forward losses.py:365-373; both discriminators return empty feature lists, so the
    feature-matching term (losses.py:345-354) is 0"""
    real, gen = disc_fn(target), disc_fn(pred)
    return generator_loss(gen) + tprls_generator(real, gen)


def acoustic_generator_loss(sds: Dict[str, SD], target_fft, pred_fft, target_audio, pred_audio):
    """GeneratorLoss.forward with used = the acoustic set (losses.py:316-327): mrd0..2 on the three
    magnitude spectrograms + 3 x the waveform discriminator"""
    loss = 0
    for i in range(3):
        loss = loss + helper_generator(lambda y, i=i: spec_discriminator(sds[f"mrd{i}"], y), target_fft[i], pred_fft[i])
    loss = loss + DISC_WEIGHT * helper_generator(lambda a: context_free_discriminator(sds["disc"], a), target_audio,
                                                 pred_audio)
    return loss


def acoustic_discriminator_loss(sds: Dict[str, SD], target_fft, pred_fft, target_audio, pred_audio):
    """DiscriminatorLoss.forward, acoustic set (losses.py:196-207)"""
    loss = 0
    for i in range(3):
        loss = loss + helper_discriminator(lambda y, i=i: spec_discriminator(sds[f"mrd{i}"], y), target_fft[i],
                                           pred_fft[i])[0]
    loss = loss + DISC_WEIGHT * helper_discriminator(lambda a: context_free_discriminator(sds["disc"], a),
                                                     target_audio, pred_audio)[0]
    return loss


def disc_lr_multiplier(last_loss: float, sub_count: int) -> float:
    """DiscriminatorLossHelper.get_disc_lr_multiplier losses.py:237-249"""
    ideal, f_max, h_min = 0.5 * sub_count, 4.0, 0.01
    x_max = x_min = 0.05 * sub_count
    x = abs(last_loss - ideal)
    if last_loss > ideal + x_max:
        return f_max
    if last_loss < ideal - x_min:
        return h_min
    if last_loss > ideal:
        return min(math.pow(f_max, x / x_max), f_max)
    return max(math.pow(h_min, x / x_min), h_min)
