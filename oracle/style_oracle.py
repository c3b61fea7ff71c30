"""CPU oracle of the mel style encoder.  TEST INFRASTRUCTURE ONLY (import rules: oracle/speech_oracle.py).

Functional PyTorch restatement of reference MelStyleEncoder (mel_style_encoder.py:66-152) against a state
dict with the reference's keys; spectral normalisation follows torch.nn.utils.spectral_norm (legacy hook,
n_power_iterations=1, eps=1e-12 — a third-party (PyTorch 2.x) algorithm restated here): in training mode u, v
are advanced in place in the state dict before sigma = u^T W v is formed, in eval mode they are used as stored.
Pinned by tests/golden/style_encoder.npz (made from the UNMODIFIED reference by
tests/golden/make_style_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def sn_weight(sd, prefix, training):
    W = sd[prefix + ".weight_orig"]
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    Wm = W.reshape(W.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(Wm.t(), u), dim=0, eps=1e-12))
            u.copy_(F.normalize(torch.mv(Wm, v), dim=0, eps=1e-12))
        u, v = u.clone(), v.clone()
    return W / torch.dot(u, torch.mv(Wm, v))


def res_block(sd, p, x, training):
    """ResBlk.forward mel_style_encoder.py:95-118 (normalize=False)"""
    half = (p + ".downsample_res.conv.weight_orig") in sd
    s = x
    if (p + ".conv1x1.weight_orig") in sd:
        s = F.conv2d(s, sn_weight(sd, p + ".conv1x1", training))
    if half:
        if s.shape[-1] % 2 != 0:
            s = torch.cat([s, s[..., -1].unsqueeze(-1)], dim=-1)
        s = F.avg_pool2d(s, 2)
    r = F.leaky_relu(x, 0.2)
    r = F.conv2d(r, sn_weight(sd, p + ".conv1", training), sd[p + ".conv1.bias"], padding=1)
    if half:
        r = F.conv2d(r, sn_weight(sd, p + ".downsample_res.conv", training), sd[p + ".downsample_res.conv.bias"],
                     stride=2, padding=1, groups=r.shape[1])
    r = F.leaky_relu(r, 0.2)
    r = F.conv2d(r, sn_weight(sd, p + ".conv2", training), sd[p + ".conv2.bias"], padding=1)
    return (s + r) / math.sqrt(2)


def mel_style_encoder(sd, x, training=False):
    """MelStyleEncoder.forward mel_style_encoder.py:146-152: (B,1,n_mels,F) -> (B,style_dim)"""
    h = F.conv2d(x, sn_weight(sd, "shared.0", training), sd["shared.0.bias"], padding=1)
    for i in range(1, 5):
        h = res_block(sd, f"shared.{i}", h, training)
    h = F.leaky_relu(h, 0.2)
    h = F.conv2d(h, sn_weight(sd, "shared.6", training), sd["shared.6.bias"])
    h = F.leaky_relu(h.mean(dim=(2, 3)), 0.2)
    return F.linear(h, sd["unshared.weight"], sd["unshared.bias"])


def pitch_style_encoder(sd, x, pitch, energy, training=False, coarse_multiplier=1):
    """PitchStyleEncoder.forward mel_style_encoder.py:188-206: pitch / energy are linearly resampled to
    frames // coarse_multiplier (the identity for 1) and stacked under the mel rows."""
    pitch, energy = pitch.unsqueeze(1), energy.unsqueeze(1)
    if coarse_multiplier != 1:
        pitch = F.interpolate(pitch, size=pitch.shape[2] // coarse_multiplier, mode="linear")
        energy = F.interpolate(energy, size=energy.shape[2] // coarse_multiplier, mode="linear")
    xc = torch.cat([x, pitch, energy], dim=1)
    w = torch._weight_norm(sd["preconv.parametrizations.weight.original1"],
                           sd["preconv.parametrizations.weight.original0"], 0)
    y = F.conv1d(xc, w, sd["preconv.bias"], padding=1)
    return mel_style_encoder(sd, y.unsqueeze(1), training)
