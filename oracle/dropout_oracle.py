"""CPU restatement of the dropout masks of the CUDA path.  TEST INFRASTRUCTURE ONLY.

The product's training-mode masks are a stateless integer hash of (seed, site, element index)
(``stylish_tts_b200/csrc/common.cuh`` ``drop_keep``); this file restates that hash with numpy uint32
arithmetic (bit-exact), so the oracle — and, through ``patched_reference`` below, the UNMODIFIED reference
modules — can be run with exactly the masks the kernels draw.  That pins the placement, scaling and layout
of every dropout site (nn.Dropout, SDPA ``dropout_p``, Dropout1d, DropPath) against the reference; the
statistics of the hash itself (keep rate) are tested separately.
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch

_M = 0xFFFFFFFF


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x ^ (x >> np.uint32(16))
    x = x * np.uint32(0x7FEB352D)
    x = x ^ (x >> np.uint32(15))
    x = x * np.uint32(0x846CA68B)
    x = x ^ (x >> np.uint32(16))
    return x


def keep_mask(seed: int, site: int, n: int, p: float) -> np.ndarray:
    """bool[n]: element i is kept (common.cuh drop_keep)."""
    idx = np.arange(n, dtype=np.uint64)
    lo = (idx & np.uint64(_M)).astype(np.uint32)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    with np.errstate(over="ignore"):
        x = _mix32(lo ^ np.uint32(seed & _M))
        add = np.uint32((site * 0x9E3779B9 + (seed >> 32)) & _M)
        x = _mix32(x ^ (hi + add))
    thresh = int(round(float(np.float32(p)) * 2 ** 24))
    return (x >> np.uint32(8)) >= np.uint32(thresh)


def scale_mask(seed: int, site: int, shape, p: float, dtype=torch.float32) -> torch.Tensor:
    """keep / (1-p) as a tensor of ``shape`` (contiguous element order = the kernel's index)."""
    n = int(np.prod(shape))
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    m = keep_mask(seed, site, n, p).astype(np.float32) * inv
    return torch.from_numpy(m.reshape(tuple(shape))).to(dtype)


class Masks:
    """What the oracle functions call: ``masks(site, p, shape)`` -> keep/(1-p) tensor, or None when off."""

    def __init__(self, seed: int, dtype=torch.float32):
        self.seed, self.dtype = int(seed), dtype
        self.used = []

    def __call__(self, site, p, shape):
        self.used.append(site)
        return scale_mask(self.seed, site, shape, p, self.dtype)


def speech_predictor_sites(n_layers=8):
    """the dropout sites of SpeechPredictor.forward in the order the reference executes them, with the layout
    of the tensor they act on: 'bct' channel-major, 'bnc' token-major (mask transposed), 'attn' (B,H,T,T)"""
    seq = [(1 + i, "bct") for i in range(3)]
    for i in range(n_layers):
        s0 = 16 + 4 * i
        seq += [(s0, "attn"), (s0 + 1, "bct"), (s0 + 2, "bct"), (s0 + 3, "bct")]
    # the generator's conformer has none: its blocks are built with dropout 0.0 (conformer.py:284-296)
    return seq


def duration_predictor_sites(n_layer=3):
    """DurationPredictor.forward in train(): text encoder, cross-attention probabilities, then per ConvNeXt
    block its DropPath ('path', drawn with Tensor.bernoulli_) and the Dropout1d after it ('chan')"""
    seq = speech_predictor_sites() + [(80, "attn")]
    for i in range(n_layer):
        seq += [(84 + 2 * i, "path"), (85 + 2 * i, "chan")]
    return seq


def pitch_energy_predictor_sites():
    """PitchEnergyPredictor.forward in train(): text encoder, 3 prosody-encoder layers, 2 towers x 4 blocks x 2"""
    seq = speech_predictor_sites()
    for i in range(3):
        s0 = 80 + 4 * i
        seq += [(s0, "attn"), (s0 + 1, "bct"), (s0 + 2, "bct"), (s0 + 3, "bct")]
    seq += [(96 + j, "bct") for j in range(16)]
    return seq


@contextlib.contextmanager
def patched_reference(seed: int, sites, smoothing=(0, 0)):
    """Run UNMODIFIED reference modules in train() mode with the hash masks: replaces ``torch.nn.functional
    .dropout`` and ``F.scaled_dot_product_attention`` (the two samplers the reference calls) for the duration,
    consuming ``sites`` in order; ``random.randint`` is pinned so that decoder.py:54-57 picks ``smoothing``
    = (F0 width, N width), and ``Tensor.to('cuda')`` of its box filter is mapped to the CPU."""
    import random
    import torch.nn.functional as F

    it = iter(sites)
    real_dropout, real_sdpa, real_randint, real_to = F.dropout, F.scaled_dot_product_attention, random.randint, \
        torch.Tensor.to
    real_dropout1d, real_bernoulli = F.dropout1d, torch.Tensor.bernoulli_

    def dropout1d(x, p=0.5, training=True, inplace=False):  # nn.Dropout1d: one draw per (b, c)
        if not training or p == 0.0:
            return x
        site, layout = next(it)
        assert layout == "chan", (site, layout)
        return x * scale_mask(seed, site, tuple(x.shape[:2]), p, x.dtype).unsqueeze(-1)

    def bernoulli_(self, keep_prob=0.5, *, generator=None):  # DropPath's per-sample draw (conv_next.py:138-143)
        site, layout = next(it)
        assert layout == "path", (site, layout)
        m = keep_mask(seed, site, self.numel(), 1.0 - keep_prob).astype(np.float32)
        return self.copy_(torch.from_numpy(m).reshape(self.shape).to(self.dtype))

    def dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        site, layout = next(it)
        assert layout in ("bct", "bnc"), (site, layout)
        if layout == "bnc":
            B, N, Cc = x.shape
            m = scale_mask(seed, site, (B, Cc, N), p, x.dtype).transpose(1, 2)
        else:
            m = scale_mask(seed, site, tuple(x.shape), p, x.dtype)
        return x * m

    def sdpa(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kw):
        sc = q.shape[-1] ** -0.5 if scale is None else scale
        s = torch.matmul(q, k.transpose(-1, -2)) * sc
        if attn_mask is not None:
            s = s + attn_mask
        pr = torch.softmax(s, dim=-1)
        if dropout_p > 0.0:
            site, layout = next(it)
            assert layout == "attn"
            pr = pr * scale_mask(seed, site, tuple(pr.shape), dropout_p, pr.dtype)
        return torch.matmul(pr, v)

    draws = iter([[0, 7, 15].index(smoothing[0]), [0, 7, 15, 31].index(smoothing[1])])

    def to(self, *a, **kw):
        if a and a[0] == "cuda":
            return self
        return real_to(self, *a, **kw)

    F.dropout, F.scaled_dot_product_attention, F.dropout1d = dropout, sdpa, dropout1d
    random.randint = lambda lo, hi: next(draws)
    torch.Tensor.to = to
    torch.Tensor.bernoulli_ = bernoulli_
    try:
        yield
        left = list(it)
        assert not left, f"reference did not reach dropout sites {left}"
    finally:
        F.dropout, F.scaled_dot_product_attention, random.randint = real_dropout, real_sdpa, real_randint
        F.dropout1d, torch.Tensor.bernoulli_ = real_dropout1d, real_bernoulli
        torch.Tensor.to = real_to
