"""Diffusion style sampler (BASELINE.json configs[3], north_star "style-diffusion denoiser") on the sm_100a kernels.

**Restatement, parity unpinned**: the reference repository has no implementation of this component — only a legacy
script that imports it from an external StyleTTS 2 checkout (tts/ttab/inference.py:58-77,131-138).  The
network / sampler follow SURVEY Appendix C; ``oracle/diffusion_oracle.py`` is the fp64 restatement the CUDA path is
tested against (self-consistency is the only parity available).  Call contract kept from the call site:
``sampler(noise (B,256), embedding=(B,T,768), num_steps=N) -> style (B,256)``, Karras(1e-4, 3.0, rho 9), ADPM2.

Engine: tokens are rows ([B*T, 1024] fp32 master copy + bf16 hi|lo planes as GEMM operands).  Every dense layer is
``sty_gemm_split_fwd`` — TMA-fed tcgen05, bf16x3, bias / GELU / residual and the planes of its OUTPUT fused in the
epilogue; LayerNorm (+ time-embedding add) writes planes too; attention is the tcgen05 flash kernel reading
token-major q|k|v and writing planes.  Per denoiser evaluation: 12 GEMMs, 3 attentions, 3 LayerNorms, 2 tiny passes.
The (B,1024)-sized time MLP, the final (B,1024)->(B,256) projection and the sampler's (B,256) arithmetic are torch
ops on parameter-sized tensors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from .modules import Node

SIGMA_DATA = 0.2
HEADS, HEAD_DIM = 8, 64
ACT_NONE, ACT_GELU = 0, 5


def karras_sigmas(num_steps: int, sigma_min=1e-4, sigma_max=3.0, rho=9.0) -> List[float]:
    a, b = sigma_max ** (1 / rho), sigma_min ** (1 / rho)
    return [(a + i / (num_steps - 1) * (b - a)) ** rho for i in range(num_steps)] + [0.0]


def _planes(w: torch.Tensor) -> torch.Tensor:
    """fp32 (N,K) -> bf16 (2,N,K) hi | lo planes on the device"""
    w = w.detach().to(torch.float32).contiguous()
    out = torch.empty((2,) + tuple(w.shape), device=w.device, dtype=torch.bfloat16)
    L.call("sty_split_planes_fwd", w.data_ptr(), out.data_ptr(), w.numel(), L.stream_ptr())
    return out


class StyleDenoiser(nn.Module):
    """Transformer1d-style denoiser of SURVEY Appendix C: 256 style channels + 768 context features, 3 blocks."""

    def __init__(self, channels=256, context=768, layers=3, fourier=128):
        super().__init__()
        C = channels + context
        assert C == 1024, "the LayerNorm kernel is built for 1024-wide tokens"
        self.channels, self.context, self.layers, self.C = channels, context, layers, C
        lin = lambda i, o, bias=True: nn.Linear(i, o, bias=bias)
        self.time = Node(fourier=nn.Parameter(torch.randn(fourier)), mlp=[lin(C, C), lin(C, C)],
                         **{"in": lin(2 * fourier + 1, C)})
        self.blocks = nn.ModuleList([
            Node(norm=nn.LayerNorm(C), to_q=lin(C, HEADS * HEAD_DIM, False), to_kv=lin(C, 2 * HEADS * HEAD_DIM, False),
                 to_out=lin(HEADS * HEAD_DIM, C), ff=[lin(C, 2 * C), lin(2 * C, C)]) for _ in range(layers)])
        self.to_out = lin(C, channels)
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_key = None

    # ------------------------------------------------------------------ packed operand planes of the weights
    def packed(self):
        key = (L.param_epoch,) + tuple(p._version for p in self.parameters()) + (str(self.to_out.weight.device),)
        if self._packed is None or key != self._packed_key:
            P = {}
            for i, b in enumerate(self.blocks):
                P[f"{i}.qkv"] = _planes(torch.cat([b.to_q.weight, b.to_kv.weight], 0))
                P[f"{i}.out"] = _planes(b.to_out.weight)
                P[f"{i}.ff0"] = _planes(getattr(b.ff, "0").weight)
                P[f"{i}.ff1"] = _planes(getattr(b.ff, "1").weight)
            self._packed, self._packed_key = P, key
        return self._packed

    def time_embedding(self, c_noise):
        f = c_noise[:, None] * self.time.fourier[None, :] * (2 * math.pi)
        e = torch.cat([c_noise[:, None], torch.sin(f), torch.cos(f)], dim=1)
        h = F.gelu(getattr(self.time, "in")(e))
        for i in range(2):
            h = F.gelu(getattr(self.time.mlp, str(i))(h))
        return h.contiguous()

    @staticmethod
    def _gemm(a_planes, w_planes, bias, res, M, N, K, act=ACT_NONE, want_f32=True, want_planes=False):
        dev = a_planes.device
        out = torch.empty((M, N), device=dev, dtype=torch.float32) if want_f32 else None
        outp = torch.empty((2, M, N), device=dev, dtype=torch.bfloat16) if want_planes else None
        L.call("sty_gemm_split_fwd", a_planes.data_ptr(), w_planes.data_ptr(), L.ptr(bias), L.ptr(res), L.ptr(out),
               L.ptr(outp), M, N, K, act, L.stream_ptr())
        return out, outp

    @torch.no_grad()
    def forward(self, x, c_noise, embedding):
        """x (B,256) (already scaled by c_in), c_noise (B,), embedding (B,T,768) -> (B,256)"""
        if not x.is_cuda:
            raise RuntimeError("stylish_tts_b200: StyleDenoiser needs CUDA tensors (no CPU fallback)")
        B, T, Ce = embedding.shape
        C, HD = self.C, HEADS * HEAD_DIM
        M = B * T
        Mp = (M + 127) // 128 * 128
        dev = x.device
        P = self.packed()
        x = x.to(torch.float32).contiguous()
        embedding = embedding.to(torch.float32).contiguous()
        m = self.time_embedding(c_noise.to(torch.float32))
        tok = torch.empty((Mp, C), device=dev, dtype=torch.float32)
        L.call("sty_build_tokens_fwd", x.data_ptr(), embedding.data_ptr(), 1.0, tok.data_ptr(), B, T, self.channels,
               Ce, Mp, L.stream_ptr())
        for i, blk in enumerate(self.blocks):
            hm = torch.empty_like(tok)
            n_pl = torch.empty((2, Mp, C), device=dev, dtype=torch.bfloat16)
            L.call("sty_row_ln_split_fwd", tok.data_ptr(), m.data_ptr(), blk.norm.weight.data_ptr(),
                   blk.norm.bias.data_ptr(), 1e-5, hm.data_ptr(), n_pl.data_ptr(), Mp, M, T, C, L.stream_ptr())
            qkv, _ = self._gemm(n_pl, P[f"{i}.qkv"], None, None, Mp, 3 * HD, C)
            att_pl = torch.zeros((2, Mp, HD), device=dev, dtype=torch.bfloat16) if Mp != M else \
                torch.empty((2, Mp, HD), device=dev, dtype=torch.bfloat16)
            ws = torch.empty(int(L.load().sty_attention64_workspace_bytes(B, HEADS, T)), device=dev, dtype=torch.uint8)
            L.call("sty_attention64_tokens_fwd", qkv.data_ptr(), 3 * HD, att_pl.data_ptr(), Mp, B, HEADS, T,
                   HEAD_DIM ** -0.5, ws.data_ptr(), L.stream_ptr())
            h1, h1_pl = self._gemm(att_pl, P[f"{i}.out"], blk.to_out.bias, hm, Mp, C, HD, want_planes=True)
            ff0, ff1 = getattr(blk.ff, "0"), getattr(blk.ff, "1")
            _, f_pl = self._gemm(h1_pl, P[f"{i}.ff0"], ff0.bias, None, Mp, 2 * C, C, act=ACT_GELU, want_f32=False,
                                 want_planes=True)
            tok, _ = self._gemm(f_pl, P[f"{i}.ff1"], ff1.bias, h1, Mp, C, 2 * C)
        mean = torch.empty((B, C), device=dev, dtype=torch.float32)
        L.call("sty_token_mean_fwd", tok.data_ptr(), mean.data_ptr(), B, T, C, L.stream_ptr())
        return self.to_out(mean)

    def denoise(self, x, sigma: float, embedding):
        """k-diffusion preconditioning (sigma_data 0.2): D(x, s) = c_skip x + c_out net(c_in x, ln(s)/4, embedding)"""
        s2, d2 = sigma * sigma, SIGMA_DATA * SIGMA_DATA
        c_skip, c_out, c_in = d2 / (s2 + d2), sigma * SIGMA_DATA / math.sqrt(s2 + d2), 1.0 / math.sqrt(s2 + d2)
        c_noise = torch.full((x.shape[0],), math.log(sigma) * 0.25, device=x.device, dtype=torch.float32)
        return c_skip * x + c_out * self.forward(c_in * x, c_noise, embedding)


class DiffusionSampler(nn.Module):
    """ADPM2 sampler over a Karras schedule (the call site's constants, inference.py:131-138)."""

    def __init__(self, denoiser: StyleDenoiser, sigma_min=1e-4, sigma_max=3.0, rho=9.0):
        super().__init__()
        self.denoiser = denoiser
        self.schedule = (sigma_min, sigma_max, rho)

    @torch.no_grad()
    def forward(self, noise, *, embedding, num_steps: int, step_noise: Optional[List[torch.Tensor]] = None):
        sig = karras_sigmas(num_steps, *self.schedule)
        D = self.denoiser.denoise
        x = sig[0] * noise.to(torch.float32)
        for i in range(num_steps - 1):
            s, s_next = sig[i], sig[i + 1]
            s_up = math.sqrt(s_next ** 2 * (s ** 2 - s_next ** 2) / s ** 2)
            s_down = math.sqrt(s_next ** 2 - s_up ** 2)
            s_mid = (s + s_down) / 2
            d = (x - D(x, s, embedding)) / s
            x_mid = x + d * (s_mid - s)
            d_mid = (x_mid - D(x_mid, s_mid, embedding)) / s_mid
            eps = step_noise[i] if step_noise is not None else torch.randn_like(x)
            x = x + d_mid * (s_down - s) + eps * s_up
        return x
