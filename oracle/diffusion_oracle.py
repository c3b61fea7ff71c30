"""CPU restatement of the diffusion style sampler named by BASELINE.json configs[3] / north_star.

TEST INFRASTRUCTURE ONLY — **restatement, parity unpinned**.  The reference repository contains NO implementation
of this component: its only trace is a legacy script that imports ``Modules.diffusion.sampler`` from an external,
unpinned StyleTTS 2 checkout (src/stylish_tts/tts/ttab/inference.py:58-77,131-138).  What
follows restates the published algorithm as SURVEY Appendix C summarises it:

* Karras sigma schedule (sigma_min 1e-4, sigma_max 3, rho 9 — the call site's constants, inference.py:131-138);
* ADPM2 sampler (two denoiser evaluations per step, ancestral noise);
* k-diffusion preconditioning with sigma_data = 0.2;
* a ``Transformer1d``-style denoiser: 256 style channels + 768 context features = 1024-wide tokens, 3 blocks of
  pre-LN multi-head attention (8 x 64) and a GELU feed-forward of multiplier 2, a learned-Fourier time embedding
  mapped by a 3-layer MLP and added to every token before each block, mean over tokens, 1x1 projection to 256.

Parameter names are this repository's own (``stylish_tts_b200.diffusion.StyleDenoiser`` owns the same keys).  The
only parity this oracle can anchor is self-consistency: the CUDA path against this file evaluated in fp64.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
SIGMA_DATA = 0.2
HEADS, HEAD_DIM = 8, 64


def karras_sigmas(num_steps: int, sigma_min=1e-4, sigma_max=3.0, rho=9.0, dtype=torch.float64) -> torch.Tensor:
    """sigma_i = (smax^(1/rho) + i/(N-1) (smin^(1/rho) - smax^(1/rho)))^rho, i = 0..N-1, then 0 appended"""
    i = torch.arange(num_steps, dtype=dtype)
    s = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([s, s.new_zeros(1)])


def time_embedding(sd: SD, c_noise: torch.Tensor) -> torch.Tensor:
    """learned Fourier features of c_noise -> 3-layer GELU MLP -> (B, 1024)"""
    f = c_noise[:, None] * sd["time.fourier"][None, :] * (2 * math.pi)
    e = torch.cat([c_noise[:, None], torch.sin(f), torch.cos(f)], dim=1)
    h = F.gelu(F.linear(e, sd["time.in.weight"], sd["time.in.bias"]))
    for i in range(2):
        h = F.gelu(F.linear(h, sd[f"time.mlp.{i}.weight"], sd[f"time.mlp.{i}.bias"]))
    return h


def block(sd: SD, p: str, h: torch.Tensor) -> torch.Tensor:
    """pre-LN MHSA (q: 1024 -> 512, kv: 1024 -> 1024, out: 512 -> 1024, scale 1/8) + residual; GELU FF x2 + residual"""
    B, T, C = h.shape
    n = F.layer_norm(h, (C,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    q = F.linear(n, sd[p + ".to_q.weight"])
    kv = F.linear(n, sd[p + ".to_kv.weight"])
    k, v = kv[..., :HEADS * HEAD_DIM], kv[..., HEADS * HEAD_DIM:]
    split = lambda t: t.reshape(B, T, HEADS, HEAD_DIM).transpose(1, 2)
    att = torch.softmax(split(q) @ split(k).transpose(-1, -2) * HEAD_DIM ** -0.5, dim=-1) @ split(v)
    att = att.transpose(1, 2).reshape(B, T, HEADS * HEAD_DIM)
    h = h + F.linear(att, sd[p + ".to_out.weight"], sd[p + ".to_out.bias"])
    ff = F.linear(F.gelu(F.linear(h, sd[p + ".ff.0.weight"], sd[p + ".ff.0.bias"])), sd[p + ".ff.1.weight"],
                  sd[p + ".ff.1.bias"])
    return h + ff


def network(sd: SD, x: torch.Tensor, c_noise: torch.Tensor, embedding: torch.Tensor, layers=3) -> torch.Tensor:
    """x (B,256) noisy style (already scaled by c_in), c_noise (B,), embedding (B,T,768) -> (B,256)"""
    B, T, _ = embedding.shape
    tok = torch.cat([x[:, None, :].expand(B, T, x.shape[1]), embedding], dim=2)
    m = time_embedding(sd, c_noise)
    for i in range(layers):
        tok = block(sd, f"blocks.{i}", tok + m[:, None, :])
    return F.linear(tok.mean(dim=1), sd["to_out.weight"], sd["to_out.bias"])


def denoise(sd: SD, x: torch.Tensor, sigma: torch.Tensor, embedding: torch.Tensor) -> torch.Tensor:
    """k-diffusion preconditioning: D(x, s) = c_skip x + c_out net(c_in x, ln(s)/4, embedding)"""
    s2, d2 = sigma * sigma, SIGMA_DATA * SIGMA_DATA
    c_skip = d2 / (s2 + d2)
    c_out = sigma * SIGMA_DATA / torch.sqrt(s2 + d2)
    c_in = 1.0 / torch.sqrt(s2 + d2)
    c_noise = (torch.log(sigma) * 0.25).expand(x.shape[0])
    return c_skip * x + c_out * network(sd, c_in * x, c_noise, embedding)


def adpm2_sample(sd: SD, noise: torch.Tensor, embedding: torch.Tensor, num_steps: int,
                 step_noise: List[torch.Tensor]) -> torch.Tensor:
    """ADPM2 (rho = 1): x0 = sigma_0 noise; per step two denoiser evaluations and one ancestral noise injection
    (``step_noise[i]``, supplied by the caller so that both arms draw the same numbers)."""
    sig = karras_sigmas(num_steps, dtype=noise.dtype)
    x = sig[0] * noise
    for i in range(num_steps - 1):
        s, s_next = sig[i], sig[i + 1]
        s_up = torch.sqrt(s_next ** 2 * (s ** 2 - s_next ** 2) / s ** 2)
        s_down = torch.sqrt(s_next ** 2 - s_up ** 2)
        s_mid = (s + s_down) / 2
        d = (x - denoise(sd, x, s, embedding)) / s
        x_mid = x + d * (s_mid - s)
        d_mid = (x_mid - denoise(sd, x_mid, s_mid, embedding)) / s_mid
        x = x + d_mid * (s_down - s) + step_noise[i] * s_up
    return x
