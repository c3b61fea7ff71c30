"""Spectrogram discriminators and the adversarial losses of the acoustic stage on the sm_100a kernels
(SURVEY 8f rank 1).

``SpecDiscriminator`` is the drop-in for the reference's ``mrd0 / mrd1 / mrd2`` (train/models/discriminator.py:13-69,
built at models.py:44-46): same constructor, same state-dict keys (``discriminators.{i}`` / ``out.{i}``,
``parametrizations.weight.original0/1`` + ``bias``), ``forward(y (B,1,bins,frames)) -> (five flattened score maps, [])``.
Images are row-channel (B, bins+2, C, frames) like the style encoder's, so each 3xK Conv2d is the stride-1
Conv1d kernel over three stacked rows (tcgen05 path for the 32 -> 32 layers); the stride-(1,2) layers run on the
space-to-depth rearrangement of their input (2C channels, 5 taps: no wasted MACs), which the LeakyReLU(0.1)
kernel produces in the same pass (``sty_leaky_s2d``).

``GeneratorLoss`` / ``DiscriminatorLoss`` mirror train/losses.py:166-373 for the spectrogram discriminators:
LSGAN terms + TPRLS (median by radix select on the device, masked sums; no ``.item()``, no boolean-mask gather),
the moving average that drives the discriminator learning rate stays in device memory (``optim.DiscriminatorLR``).
``ContextFreeDiscriminator`` is the drop-in for the waveform discriminator ``disc`` (discriminator.py:119-175).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn
from torch.autograd import Function
from torch.nn.utils.parametrizations import weight_norm

from . import _lib as L
from .optim import DiscriminatorLR
from .style_encoder import RowConvFn

DISC_WEIGHT = 3.0  # losses.py:14
TAU = 0.04         # losses.py:270,359


class LeakyS2dFn(Function):
    """(N,C,W) -> (N, C*s, ceil(W/s)): LeakyReLU(slope) + space-to-depth by s along W (s = 1: activation only)"""

    @staticmethod
    def forward(ctx, x, s, slope):
        x = x.contiguous()
        N, Cc, W = x.shape
        y = torch.empty((N, Cc * s, (W + s - 1) // s), device=x.device, dtype=torch.float32)
        L.call("sty_leaky_s2d_fwd", x.data_ptr(), y.data_ptr(), N, Cc, W, s, slope, L.stream_ptr())
        ctx.save_for_backward(x)
        ctx.s, ctx.slope = s, slope
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        N, Cc, W = x.shape
        dx = torch.empty_like(x)
        L.call("sty_leaky_s2d_bwd", dy.contiguous().data_ptr(), x.data_ptr(), dx.data_ptr(), N, Cc, W, ctx.s, ctx.slope,
               L.stream_ptr())
        return dx, None, None


def leaky_image(img, s=1, slope=0.1):
    """row-channel image (B,Hp,C,W) -> (B,Hp,C*s,ceil(W/s))"""
    B, Hp, Cc, W = img.shape
    y = LeakyS2dFn.apply(img.reshape(B * Hp, Cc, W), s, slope)
    return y.view(B, Hp, Cc * s, y.shape[2])


def _skip(mode, what):
    return mode is not None and not getattr(mode, what)


class FirstConvFn(Function):
    """Conv2d(1 -> 32, 3x9, pad (1,4)) of the raw spectrogram: y (B,bins,W) -> row-channel h (B,bins+2,32,W).
    864 FMAs per pixel against a 128-byte output write: fp32 FMA kernels (csrc/disc_ops.cu), output-bound."""

    @staticmethod
    def forward(ctx, y, w4, bias, mode):
        y = y.contiguous()
        B, bins, W = y.shape
        w = w4.detach().contiguous()
        h = torch.empty((B, bins + 2, 32, W), device=y.device, dtype=torch.float32)
        L.call("sty_disc_first_fwd", y.data_ptr(), w.data_ptr(), bias.detach().contiguous().data_ptr(), h.data_ptr(),
               B, bins, W, L.stream_ptr())
        ctx.save_for_backward(y, w)
        ctx.mode = mode
        return h

    @staticmethod
    def backward(ctx, dh):
        y, w = ctx.saved_tensors
        B, bins, W = y.shape
        dh = dh.contiguous()
        dy = dw = db = None
        if ctx.needs_input_grad[0] and not _skip(ctx.mode, "first_input"):
            dy = torch.empty_like(y)
            L.call("sty_disc_first_dgrad", dh.data_ptr(), w.data_ptr(), dy.data_ptr(), B, bins, W, L.stream_ptr())
        if ctx.needs_input_grad[1] and not _skip(ctx.mode, "weights"):
            dw = torch.empty((32, 1, 3, 9), device=y.device, dtype=torch.float32)
            db = torch.empty((32,), device=y.device, dtype=torch.float32)
            L.call("sty_disc_first_wgrad", y.data_ptr(), dh.data_ptr(), dw.data_ptr(), db.data_ptr(), B, bins, W,
                   L.stream_ptr())
        return dy, dw, db, None


class TailFn(Function):
    """What hangs off a pre-activation h (B,Hp,32,W) besides the next 32 -> 32 conv, in one pass:
    a = LeakyReLU_0.1(h) (discriminator.py:59), score = Conv2d(32 -> 1, 3x3)(a) (:60-61), and the next layer's input
    (kind 1: a itself; kind 2: its space-to-depth along W for a stride-(1,2) layer; kind 0: none).  The backward is one
    pass too: dh = leaky'(h) * (score-conv data gradient + d(next))."""

    @staticmethod
    def forward(ctx, h, ws4, bs, kind, mode):
        h = h.contiguous()
        B, Hp, Cc, W = h.shape
        assert Cc == 32
        w = ws4.detach().contiguous()
        score = torch.empty((B, Hp - 2, W), device=h.device, dtype=torch.float32)
        nxt = None
        if kind == 1:
            nxt = torch.empty_like(h)
        elif kind == 2:
            nxt = torch.empty((B, Hp, 64, (W + 1) // 2), device=h.device, dtype=torch.float32)
        L.call("sty_disc_tail_fwd", h.data_ptr(), w.data_ptr(), bs.detach().contiguous().data_ptr(), score.data_ptr(),
               L.ptr(nxt), kind, B, Hp, W, L.stream_ptr())
        ctx.save_for_backward(h, w)
        ctx.kind, ctx.mode = kind, mode
        if nxt is None:
            nxt = score.new_zeros(())  # placeholder output
            ctx.mark_non_differentiable(nxt)
        return score, nxt

    @staticmethod
    def backward(ctx, dscore, dnext):
        h, w = ctx.saved_tensors
        B, Hp, _, W = h.shape
        kind = ctx.kind
        dscore = None if dscore is None else dscore.contiguous()
        if kind != 0 and dnext is None:
            dnext = torch.zeros((B, Hp, 32, W) if kind == 1 else (B, Hp, 64, (W + 1) // 2), device=h.device)
        dh = dw = db = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty_like(h)
            L.call("sty_disc_tail_bwd", h.data_ptr(), w.data_ptr(), L.ptr(dscore),
                   L.ptr(dnext.contiguous() if kind != 0 else None), dh.data_ptr(), kind, B, Hp, W, L.stream_ptr())
        if ctx.needs_input_grad[1] and not _skip(ctx.mode, "weights"):
            dw = torch.zeros((1, 32, 3, 3), device=h.device, dtype=torch.float32)
            db = torch.zeros((1,), device=h.device, dtype=torch.float32)
            if dscore is not None:
                L.call("sty_disc_score_wgrad", h.data_ptr(), dscore.data_ptr(), dw.data_ptr(), db.data_ptr(), B, Hp, W,
                       L.stream_ptr())
        return dh, dw, db, None, None


def stride2_weight(w4: torch.Tensor) -> torch.Tensor:
    """(Co,Ci,3,9) stride-(1,2) pad-(1,4) kernel -> (Co,2Ci,3,5) stride-1 pad-(1,2) kernel on the space-to-depth
    input: y[w'] = sum_k w[k] x[2w'+k-4] = sum_{j,p} w[2j+p] xs[p][w'+j-2]; tap k = 9 does not exist (zero).
    Plain tensor ops on the weight, so autograd carries the gradient back to the 3x9 parameter."""
    Co, Ci, R, K = w4.shape
    assert K == 9
    wp = torch.nn.functional.pad(w4, (0, 1))                      # k = 0..9
    return wp.reshape(Co, Ci, R, 5, 2).permute(0, 1, 4, 2, 3).reshape(Co, Ci * 2, R, 5)


class BackwardMode:
    """Which gradients a backward pass through a discriminator is run for.  The reference evaluates every
    discriminator twice per batch on the same inputs with the same weights — once inside ``GeneratorLoss`` (weight
    gradients computed and thrown away by ``zero_grad``, stage.py:127) and once inside ``DiscriminatorLoss``
    (detached inputs).  ``AdversarialTerms`` keeps ONE forward and walks its tape twice; this switch tells the conv
    backward which half of its work the current walk needs."""

    def __init__(self):
        self.weights = True       # weight / bias gradients
        self.first_input = True   # gradient w.r.t. the spectrogram itself (first layer's data gradient)

    def set(self, *, weights=True, first_input=True):
        mode = self

        class _Ctx:
            def __enter__(self):
                self.saved = (mode.weights, mode.first_input)
                mode.weights, mode.first_input = weights, first_input

            def __exit__(self, *exc):
                mode.weights, mode.first_input = self.saved

        return _Ctx()


class SpecDiscriminator(nn.Module):
    """Drop-in for reference SpecDiscriminator (discriminator.py:13-69)."""

    def __init__(self):
        super().__init__()
        self.backward_mode = BackwardMode()
        c2 = lambda ci, co, k, stride, pad: weight_norm(nn.Conv2d(ci, co, kernel_size=k, stride=stride, padding=pad))
        self.discriminators = nn.ModuleList([
            c2(1, 32, (3, 9), 1, (1, 4)), c2(32, 32, (3, 9), (1, 2), (1, 4)), c2(32, 32, (3, 9), (1, 2), (1, 4)),
            c2(32, 32, (3, 9), (1, 2), (1, 4)), c2(32, 32, (3, 3), 1, (1, 1))])
        self.out = nn.ModuleList([c2(32, 1, 3, 1, 1) for _ in range(5)])
        self._masks = {}

    def _row_mask(self, B, Hp, W, device):
        key = (B, Hp, W, str(device))
        if key not in self._masks:
            n = torch.arange(B * Hp - 2, device=device)
            self._masks[key] = ((n % Hp) < (Hp - 2)).float()[:, None].expand(-1, W).contiguous()
        return self._masks[key]

    @staticmethod
    def _w(conv):
        p = conv.parametrizations.weight
        return torch._weight_norm(p.original1, p.original0, 0)

    fused = True  # False: the first version (generic conv kernels for every layer); kept for A/B tests

    def forward(self, y):
        L.require_cuda(y, "the input of SpecDiscriminator")
        if not self.fused:
            return self._forward_generic(y)
        B, one, K, N = y.shape
        assert one == 1
        mode = self.backward_mode
        conv = lambda t, w4, b: RowConvFn.apply(
            t, w4, b, None, dict(row_mask=self._row_mask(t.shape[0], t.shape[1], t.shape[3], t.device), mode=mode,
                                 fold_free=True, wide=True))
        result: List[torch.Tensor] = []
        d0 = self.discriminators[0]
        h = FirstConvFn.apply(y[:, 0].to(torch.float32), self._w(d0), d0.bias, mode)
        for i in range(5):
            kind = 2 if i < 3 else (1 if i == 3 else 0)   # layers 1-3 are stride (1,2): space-to-depth input
            score, nxt = TailFn.apply(h, self._w(self.out[i]), self.out[i].bias, kind, mode)
            result.append(score.reshape(B, -1))               # torch.flatten(out, 1, -1), discriminator.py:61
            if i == 4:
                break
            d = self.discriminators[i + 1]
            h = conv(nxt, stride2_weight(self._w(d)) if i < 3 else self._w(d), d.bias)
        return result, []

    def _forward_generic(self, y):
        B, one, K, N = y.shape
        assert one == 1
        Hp = K + 2
        img = torch.zeros((B, Hp, 1, N), device=y.device, dtype=torch.float32)
        img[:, 1:K + 1, 0, :] = y[:, 0].to(torch.float32)
        conv = lambda t, w4, b, first=False: RowConvFn.apply(
            t, w4, b, None, dict(row_mask=self._row_mask(t.shape[0], t.shape[1], t.shape[3], t.device),
                                 mode=self.backward_mode, first=first))
        result: List[torch.Tensor] = []
        h = conv(img, self._w(self.discriminators[0]), self.discriminators[0].bias, True)
        for i in range(5):
            a = leaky_image(h)                                   # LeakyReLU(0.1), discriminator.py:59
            # score conv 32 -> 1: zero-padded to 16 output channels so that forward, data gradient and weight
            # gradient all take the tensor-core kernels; channel 0 is the score map
            wo = torch.nn.functional.pad(self._w(self.out[i]), (0, 0, 0, 0, 0, 0, 0, 15))
            bo = torch.nn.functional.pad(self.out[i].bias, (0, 15))
            o = conv(a, wo, bo)                                  # (B,Hp,16,W)
            result.append(o[:, 1:K + 1, 0, :].reshape(B, -1))    # torch.flatten(out, 1, -1)
            if i == 4:
                break
            d = self.discriminators[i + 1]
            if i < 3:   # stride (1,2): space-to-depth of the SAME activation, 5-tap stride-1 conv
                h = conv(leaky_image(h, s=2), stride2_weight(self._w(d)), d.bias)
            else:
                h = conv(a, self._w(d), d.bias)
        return result, []


class PitchDiscriminator(nn.Module):
    """Drop-in for reference PitchDiscriminator (pitch_discriminator.py:6-68; ``pitch_disc`` = (2, 64, k21),
    ``dur_disc`` = (1, 64, k5), models.py:78-83): five weight-normed 'same' Conv1d + LeakyReLU(0.1), each with its
    own k-tap score conv; forward(y (B,dim_in,T)) -> (five (B,T) score maps, [])."""

    def __init__(self, *, dim_in, dim_hidden, kernel):
        super().__init__()
        pad = kernel // 2
        c1 = lambda ci, co: weight_norm(nn.Conv1d(ci, co, kernel_size=kernel, padding=pad))
        self.discriminators = nn.ModuleList([c1(dim_in, dim_hidden)] + [c1(dim_hidden, dim_hidden) for _ in range(4)])
        self.out = nn.ModuleList([c1(dim_hidden, 1) for _ in range(5)])

    def forward(self, y):
        from . import train_ops as T

        L.require_cuda(y, "the input of PitchDiscriminator")
        w = SpecDiscriminator._w
        result = []
        h = y.to(torch.float32).contiguous()
        for i, d in enumerate(self.discriminators):
            h = LeakyS2dFn.apply(T.conv(h, w(d), d.bias), 1, 0.1)
            result.append(T.conv(h, w(self.out[i]), self.out[i].bias).flatten(1))
        return result, []


# ------------------------------------------------------------------------------ waveform discriminator (`disc`)
class SegmentScaleFn(Function):
    """y[b, c, w*P + j] = x[b, c, w*P + j] * g[b, c, w]  (the gate of ContextFreeDiscriminator, discriminator.py:167-168,
    on windows laid end to end with pitch P)"""

    @staticmethod
    def forward(ctx, x, g, P):
        x, g = x.contiguous(), g.contiguous()
        y = torch.empty_like(x)
        L.call("sty_row_scale_fwd", x.data_ptr(), g.data_ptr(), y.data_ptr(), g.numel(), P, 1.0, L.stream_ptr())
        ctx.save_for_backward(x, g)
        ctx.P = P
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dg = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            L.call("sty_row_scale_fwd", dy.data_ptr(), g.data_ptr(), dx.data_ptr(), g.numel(), ctx.P, 1.0, L.stream_ptr())
        if ctx.needs_input_grad[1]:
            dg = torch.empty_like(g)
            L.call("sty_row_dot", dy.data_ptr(), x.data_ptr(), dg.data_ptr(), g.numel(), ctx.P, L.stream_ptr())
        return dx, dg, None


class SegmentMeanFn(Function):
    """(B, C, W*P) with Tw data positions per window (zeros in the gaps) -> per-window mean (B, C, W)
    (AdaptiveAvgPool1d(1), discriminator.py:139)"""

    @staticmethod
    def forward(ctx, x, P, Tw):
        x = x.contiguous()
        B, Cc, T_ = x.shape
        y = torch.empty((B, Cc, T_ // P), device=x.device, dtype=torch.float32)
        L.call("sty_segment_sum_fwd", x.data_ptr(), y.data_ptr(), y.numel(), P, 1.0 / Tw, L.stream_ptr())
        ctx.meta = (x.shape, P, Tw)
        return y

    @staticmethod
    def backward(ctx, g):
        shape, P, Tw = ctx.meta
        g = g.contiguous()
        dx = torch.empty(shape, device=g.device, dtype=torch.float32)  # gap positions: ignored by the masked producer
        L.call("sty_row_scale_fwd", None, g.data_ptr(), dx.data_ptr(), g.numel(), P, 1.0 / Tw, L.stream_ptr())
        return dx, None, None


def strided_weight(w: torch.Tensor, s: int) -> torch.Tensor:
    """(Co, C, K) stride-s pad-K//2 Conv1d kernel -> (Co, C*s, K') stride-1 'same' kernel on the space-to-depth input
    xs[c*s + p][u] = x[c][s*u + p]:  y[u] = sum_k w[k] x[s*u + k - pad] with k - pad = s*j + p, |j| <= J = ceil(pad/s).
    Plain tensor ops, so autograd carries the gradient back to the (Co, C, K) parameter."""
    if s == 1:
        return w
    Co, Cc, K = w.shape
    pad = K // 2
    J = -(-pad // s)
    Kp = 2 * J + 1
    left = s * J - pad
    wp = torch.nn.functional.pad(w, (left, s * Kp - K - left))
    return wp.reshape(Co, Cc, Kp, s).permute(0, 1, 3, 2).reshape(Co, Cc * s, Kp)


def grouped_as_dense(w: torch.Tensor, groups: int) -> torch.Tensor:
    """(Co, Ci/g, K) grouped kernel -> block-diagonal (Co, Ci, K): the grouped layers are 1-8 % of the discriminator's
    MACs, so they run as dense tensor-core convs on a weight with zero off-diagonal blocks"""
    if groups == 1:
        return w
    Co, Cg, K = w.shape
    eye = torch.eye(groups, device=w.device, dtype=w.dtype).view(groups, 1, groups, 1, 1)
    return (w.view(groups, Co // groups, 1, Cg, K) * eye).reshape(Co, groups * Cg, K)


class _CFBlock(nn.Module):
    """parameter holder with the reference's names (ContextFreeBlock, discriminator.py:94-116: net.0 conv, net.1 BN)"""

    def __init__(self, ci, co, *, kernel, stride=1, groups=1, bias=False):
        super().__init__()
        self.net = nn.Sequential(nn.Conv1d(ci, co, kernel, stride, kernel // 2, groups=groups, bias=bias),
                                 nn.BatchNorm1d(co), nn.GELU())
        self.kernel, self.stride, self.groups = kernel, stride, groups


class ContextFreeDiscriminator(nn.Module):
    """Drop-in for the reference's waveform discriminator `disc` (discriminator.py:119-175): 1024-sample windows (hop
    512) -> four strided Conv1d + BatchNorm + GELU blocks -> squeeze-excite gate -> temporal (k7, k3; 8 groups) and
    'spectral' (k1; 8 groups) branches -> fusion -> 1x1 head; forward(x (B, L)) -> ([(B, windows*16)], []).

    Every block is `conv -> BN -> GELU`; BN + GELU run as the PROLOGUE of the consuming convolution (batch statistics
    from `sty_row_moments`, their backward terms from `sty_prologue_bwd_*`), strided layers run as stride-1
    tensor-core convs on the space-to-depth input with phase-pooled BatchNorm statistics, grouped layers as
    block-diagonal dense convs.  The torch children only own the parameters / buffers (same state-dict keys).

    Layout: the reference folds the windows into the batch axis — (B*470, C, 16) at the last level, millions of
    16-element rows.  Here the windows of an utterance lie END TO END on the time axis with a zero gap after each one
    (pitch 1280 / 320 / 80 / 40 / 20 for 1024 / 256 / 64 / 32 / 16 data positions: the pitch divides by every stride, so
    space-to-depth keeps windows aligned, and the gap is wider than every kernel's reach), i.e. (B, C, 470*pitch): the
    long-row regime all conv / norm kernels are built for.  Producers keep the gaps zero (`out_mask`), consumers mask
    their prologue (`in_mask`), BatchNorm counts only the data positions (`bn_frac` = 0.8)."""

    PITCH = (1280, 320, 80, 40, 20)
    DATA = (1024, 256, 64, 32, 16)

    def __init__(self):
        super().__init__()
        d = 64
        self.conv = nn.ModuleList([_CFBlock(1, d, kernel=11, stride=4), _CFBlock(d, 2 * d, kernel=11, stride=4),
                                   _CFBlock(2 * d, 4 * d, kernel=7, stride=2), _CFBlock(4 * d, 4 * d, kernel=5, stride=2)])
        self.attn = nn.Sequential(nn.AdaptiveAvgPool1d(1), nn.Conv1d(4 * d, 4 * d, 1), nn.Sigmoid())
        self.temporal = nn.Sequential(_CFBlock(4 * d, 4 * d, kernel=7, groups=8, bias=True),
                                      _CFBlock(4 * d, 4 * d, kernel=3, groups=8, bias=True))
        self.spectral = nn.Sequential(_CFBlock(4 * d, 12 * d, kernel=1, groups=8, bias=True),
                                      _CFBlock(12 * d, 4 * d, kernel=1, groups=8, bias=True))
        self.fusion = _CFBlock(8 * d, 4 * d, kernel=1, bias=True)
        self.last = nn.Sequential(nn.Conv1d(4 * d, 8 * d, 1, 1), nn.ReLU(), nn.Conv1d(8 * d, 1, 1))
        self._masks = {}
        self.backward_mode = BackwardMode()

    def _mask(self, B, windows, level, device):
        key = (B, windows, level, str(device))
        if key not in self._masks:
            pos = torch.arange(windows * self.PITCH[level], device=device) % self.PITCH[level]
            self._masks[key] = (pos < self.DATA[level]).float().unsqueeze(0).expand(B, -1).contiguous()
        return self._masks[key]

    def _bn(self, *blocks, group=1):
        """prologue arguments: BatchNorm + GELU of `blocks` (side by side on the channel axis) ahead of the next conv"""
        from ._lib import ACT_GELU

        bns = [b.net[1] for b in blocks]
        if self.training:
            with torch.no_grad():
                for bn in bns:
                    bn.num_batches_tracked += 1
        cat = lambda ts: ts[0] if len(ts) == 1 else torch.cat(ts)
        return dict(bn_w=cat([bn.weight for bn in bns]), bn_b=cat([bn.bias for bn in bns]), norm="batch",
                    in_act=ACT_GELU, bn_group=group, bn_eval=not self.training, eps=bns[0].eps,
                    bn_frac=self.DATA[0] / self.PITCH[0],
                    bn_buffers=[(bn.running_mean, bn.running_var) for bn in bns])

    def forward(self, x):
        from . import train_ops as T
        from ._lib import ACT_RELU

        L.require_cuda(x, "the input of ContextFreeDiscriminator")
        B = x.shape[0]
        conv = lambda *a, **kw: T.conv(*a, mode=self.backward_mode, **kw)
        win = x.to(torch.float32).unfold(1, 1024, 512)                       # (B, W, 1024), discriminator.py:160
        W = win.shape[1]
        h = torch.nn.functional.pad(win, (0, self.PITCH[0] - self.DATA[0])).reshape(B, 1, W * self.PITCH[0])
        prev = None
        for lvl, blk in enumerate(self.conv, start=1):                        # conv -> BN -> GELU, strides 4 4 2 2
            s, m = blk.stride, self._mask(B, W, lvl, x.device)
            xs = LeakyS2dFn.apply(h, s, 1.0)                                   # space-to-depth only (slope 1)
            w = strided_weight(blk.net[0].weight, s)
            if prev is None:
                h = conv(xs, w, None, out_mask=m, first=True)
            else:
                h = conv(xs, w, None, in_mask=m, in_mask_post=True, out_mask=m, **self._bn(prev, group=s))
            prev = blk
        m = self._mask(B, W, 4, x.device)
        P, Tw = self.PITCH[4], self.DATA[4]
        mk = dict(in_mask=m, in_mask_post=True, out_mask=m)
        # x3 = GELU(BN(h)) is needed as a tensor (pooled, gated, read by two branches): identity 1x1 conv carries it
        eye = torch.eye(h.shape[1], device=h.device, dtype=torch.float32).unsqueeze(-1)
        x3 = conv(h, eye, None, **mk, **self._bn(prev))
        gate = torch.sigmoid(conv(SegmentMeanFn.apply(x3, P, Tw), self.attn[1].weight, self.attn[1].bias))
        xg = SegmentScaleFn.apply(x3, gate, P)                                 # discriminator.py:167-168
        branches = []
        for seq in (self.temporal, self.spectral):
            b0, b1 = seq[0], seq[1]
            y0 = conv(xg, grouped_as_dense(b0.net[0].weight, b0.groups), b0.net[0].bias, out_mask=m)
            branches.append(conv(y0, grouped_as_dense(b1.net[0].weight, b1.groups), b1.net[0].bias, **mk,
                                   **self._bn(b0)))
        f = conv(torch.cat(branches, 1), self.fusion.net[0].weight, self.fusion.net[0].bias, **mk,
                   **self._bn(self.temporal[1], self.spectral[1]))
        l0 = conv(f, self.last[0].weight, self.last[0].bias, **mk, **self._bn(self.fusion))
        out = conv(l0, self.last[2].weight, self.last[2].bias, in_act=ACT_RELU)     # (B, 1, W*P)
        return [out.view(B, W, P)[:, :, :Tw].reshape(B, -1)], []               # "(b t) c f -> b (t c f)"


# ---------------------------------------------------------------------------------------------- losses
class _SqMeanFn(Function):
    """mean((c - x)^2) with the reduction on the device kernel (losses.py:257-259,341)"""

    @staticmethod
    def forward(ctx, x, c):
        x = x.contiguous()
        out = torch.zeros((), device=x.device, dtype=torch.float32)
        L.call("sty_sqdiff_sum_fwd", x.data_ptr(), x.numel(), float(c), out.data_ptr(), L.stream_ptr())
        ctx.save_for_backward(x)
        ctx.c = c
        return out / x.numel()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return (x - ctx.c) * (g * (2.0 / x.numel())), None


class _TprlsFn(Function):
    """tau - relu(tau - l_rel), l_rel = sum_{a < b + m} ((a-b) - m)^2 / (count + eps), m = median(a - b)
    (losses.py:264-278 with eps = 1e-9; :356-363 with the plain mean, eps = 0)."""

    @staticmethod
    def forward(ctx, a, b, eps):
        a, b = a.contiguous(), b.contiguous()
        n = a.numel()
        ws = torch.empty(int(L.load().sty_tprls_workspace_bytes()) // 4 + 4, device=a.device, dtype=torch.int32)
        sums = torch.empty(3, device=a.device, dtype=torch.float32)
        med = torch.empty(1, device=a.device, dtype=torch.float32)
        L.call("sty_tprls_fwd", a.data_ptr(), b.data_ptr(), n, ws.data_ptr(), sums.data_ptr(), med.data_ptr(),
               L.stream_ptr())
        l_rel = sums[0] / (sums[1] + eps)
        ctx.save_for_backward(a, b, sums, med, l_rel)
        ctx.eps = eps
        return TAU - torch.relu(TAU - l_rel)

    @staticmethod
    def backward(ctx, g):
        a, b, sums, med, l_rel = ctx.saved_tensors
        gate = (l_rel < TAU).to(torch.float32) * g   # d/dl [tau - relu(tau - l)]
        denom = sums[1] + ctx.eps
        coef = torch.stack([gate * 2.0 / denom, gate * (-2.0) * sums[2] / denom]).contiguous()
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        da = torch.empty_like(a) if need_a else None
        db = torch.empty_like(b) if need_b else None
        flag = torch.empty(1, device=a.device, dtype=torch.int32)
        L.call("sty_tprls_bwd", a.data_ptr(), b.data_ptr(), a.numel(), med.data_ptr(), coef.data_ptr(), L.ptr(da),
               L.ptr(db), flag.data_ptr(), L.stream_ptr())
        return da, db, None


def lsgan_discriminator(real: List[torch.Tensor], gen: List[torch.Tensor]):
    """losses.py:251-262"""
    return sum(_SqMeanFn.apply(dr, 1.0) + _SqMeanFn.apply(dg, 0.0) for dr, dg in zip(real, gen))


def lsgan_generator(gen: List[torch.Tensor]):
    """losses.py:339-343"""
    return sum(_SqMeanFn.apply(dg, 1.0) for dg in gen)


def tprls_discriminator(real, gen):
    """losses.py:264-278"""
    return sum(_TprlsFn.apply(dr, dg, 1e-9) for dr, dg in zip(real, gen))


def tprls_generator(real, gen):
    """losses.py:356-363 — the reference zips (real, gen) into the names (dg, dr): inside its loop 'dr' is the
    GENERATED score; restated literally (median of gen - real, mask gen < real + m, plain mean)"""
    return sum(_TprlsFn.apply(dg_gen, dr_real, 0.0) for dr_real, dg_gen in zip(real, gen))


class GeneratorLoss(nn.Module):
    """GeneratorLoss of train/losses.py:291-337, same constructor and call keywords: for the acoustic set
    (``used`` without a pitch / duration discriminator) the sum over mrd0..2 of [LSGAN + TPRLS on the generated / real
    scores] + disc_weight x the same on the waveform discriminator; ``used=["pitch_disc"]`` / ``["dur_disc"]`` route
    to the curve discriminators (losses.py:316-321).  Gradients flow into `pred_list` / `pred_audio` only; the
    discriminators' parameters are treated as constants (the reference computes their gradients here and discards
    them with zero_grad, stage.py:127)."""

    def __init__(self, *, mrd0, mrd1, mrd2, disc: Optional[nn.Module] = None, pitch: Optional[nn.Module] = None,
                 duration: Optional[nn.Module] = None):
        super().__init__()
        self.mrd = nn.ModuleList([mrd0, mrd1, mrd2])
        self.disc, self.pitch, self.duration = disc, pitch, duration

    @staticmethod
    def _one(model, target, pred):
        with torch.no_grad():
            real, _ = model(target)
        gen, _ = model(pred)
        return lsgan_generator(gen) + tprls_generator(real, gen)

    def forward(self, *, target_list, pred_list, target_audio=None, pred_audio=None, used=None, index=0):
        curve = _curve_model(self, used)
        models = [curve] if curve is not None else list(self.mrd) + ([self.disc] if self.disc is not None else [])
        req = [p.requires_grad for m in models for p in m.parameters()]
        for m in models:
            m.requires_grad_(False)
        try:
            if curve is not None:
                loss = self._one(curve, target_list[0], pred_list[0])
            else:
                loss = sum(self._one(m, t, p) for m, t, p in zip(self.mrd, target_list, pred_list))
                if self.disc is not None:
                    loss = loss + DISC_WEIGHT * self._one(self.disc, target_audio, pred_audio)
        finally:
            it = iter(req)
            for m in models:
                for p in m.parameters():
                    p.requires_grad_(next(it))
        return loss


def _curve_model(owner, used):
    """losses.py:196-200,316-321: `used` containing pitch_disc / dur_disc selects that discriminator alone"""
    if used is not None and "pitch_disc" in used:
        if owner.pitch is None:
            raise ValueError("stylish_tts_b200: this loss was built without a pitch discriminator")
        return owner.pitch
    if used is not None and "dur_disc" in used:
        if owner.duration is None:
            raise ValueError("stylish_tts_b200: this loss was built without a duration discriminator")
        return owner.duration
    return None


class DiscriminatorLoss(nn.Module):
    """DiscriminatorLoss of train/losses.py:166-230, same constructor and call keywords, on DETACHED inputs, plus
    the moving average of the plain LSGAN part per discriminator that drives its learning rate (losses.py:280-288;
    device-resident, ``get_disc_lr_multiplier`` / ``state_dict`` as in the reference)."""

    def __init__(self, *, mrd0, mrd1, mrd2, disc: Optional[nn.Module] = None, pitch: Optional[nn.Module] = None,
                 duration: Optional[nn.Module] = None, device="cuda"):
        super().__init__()
        self.mrd = nn.ModuleList([mrd0, mrd1, mrd2])
        self.disc, self.pitch, self.duration = disc, pitch, duration
        self.lr_control = {f"mrd{i}": DiscriminatorLR(5, device) for i in range(3)}
        self.lr_control["disc"] = DiscriminatorLR(1, device)
        self.lr_control["pitch_disc"] = DiscriminatorLR(5, device)
        self.lr_control["dur_disc"] = DiscriminatorLR(5, device)

    def get_disc_lr_multiplier(self, key):
        return self.lr_control[key].multiplier()

    def state_dict(self, *args, **kwargs):  # losses.py:209-214
        state = {}
        for key, c in self.lr_control.items():
            state[f"discriminators.{key}.last_loss"] = float(c.last_loss)
            state[f"discriminators.{key}.weight"] = 1
        return state

    def load_state_dict(self, state_dict, strict=True):  # losses.py:216-220
        for key, c in self.lr_control.items():
            k = f"discriminators.{key}.last_loss"
            if k in state_dict:
                c.last_loss.fill_(float(state_dict[k]))
        return state_dict

    def _one(self, key, model, target, pred):
        real, _ = model(target.detach())
        gen, _ = model(pred.detach())
        d = lsgan_discriminator(real, gen)
        self.lr_control[key].update(d)
        return d + tprls_discriminator(real, gen)

    def forward(self, *, target_list, pred_list, target_audio=None, pred_audio=None, used=None, index=0):
        curve = _curve_model(self, used)
        if curve is not None:
            return self._one("pitch_disc" if curve is self.pitch else "dur_disc", curve, target_list[0], pred_list[0])
        loss = sum(self._one(f"mrd{i}", m, t, p) for i, (m, t, p) in enumerate(zip(self.mrd, target_list, pred_list)))
        if self.disc is not None:
            loss = loss + DISC_WEIGHT * self._one("disc", self.disc, target_audio, pred_audio)
        return loss


class AdversarialTerms(nn.Module):
    """Both adversarial halves of an acoustic batch from ONE evaluation of the discriminators.

    Stage.train_batch (stage.py:104-146) runs ``GeneratorLoss`` (inside the generator's backward) and then
    ``DiscriminatorLoss`` on the detached spectrograms; between the two only the generator's parameters change, so
    both see the same discriminator outputs for the same (target, prediction) pair.  Here the discriminators run
    once per batch with their tape kept:

    * ``generator_loss(...)`` (same keywords as ``GeneratorLoss.forward``) returns the generator term; its backward
      walks the tape for the gradient w.r.t. the predicted spectrograms only (no weight gradients — the reference
      computes and discards them);
    * ``discriminator_backward(index, scale)`` evaluates ``DiscriminatorLoss`` from the stored scores (value of all
      three + moving averages, losses.py:196-207,280-288) and back-propagates ``scale x`` the term of the ONE
      discriminator that is stepped (``mrd{index}``, stage.py:141-143; the others' gradients are zeroed unused in
      the reference), without the data gradient of the first layer.

    5.3 discriminator-forward equivalents per batch instead of 9 + 2 discarded weight-gradient passes; the numbers
    are identical to the two-evaluation classes above (tests/test_gpu_discriminators.py)."""

    def __init__(self, *, mrd0, mrd1, mrd2, disc: Optional[nn.Module] = None, device="cuda"):
        super().__init__()
        self.mrd = nn.ModuleList([mrd0, mrd1, mrd2])
        self.disc = disc
        self.lr_control = {f"mrd{i}": DiscriminatorLR(5, device) for i in range(3)}
        self.lr_control["disc"] = DiscriminatorLR(1, device)
        self._state = None

    def _models(self):
        return list(self.mrd) + ([self.disc] if self.disc is not None else [])

    def _modes(self, **kw):
        import contextlib
        stack = contextlib.ExitStack()
        for m in self._models():
            mode = getattr(m, "backward_mode", None)
            if mode is not None:
                stack.enter_context(mode.set(**kw))
        return stack

    class _GeneratorTermFn(Function):
        @staticmethod
        def forward(ctx, owner, targets, *preds):
            models = owner._models()
            with torch.enable_grad():
                leafs = [p.detach().requires_grad_(True) for p in preds]
                real, gen, total = [], [], 0.0
                for k, (m, t, p) in enumerate(zip(models, targets, leafs)):
                    r, _ = m(t.detach())
                    g, _ = m(p)
                    term = lsgan_generator(g) + tprls_generator([x.detach() for x in r], g)
                    total = total + (term if k < 3 else DISC_WEIGHT * term)
                    real.append(r)
                    gen.append(g)
            owner._state = SimpleState(leafs=leafs, real=real, gen=gen, gen_loss=total)
            ctx.owner = owner
            return total.detach().clone()

        @staticmethod
        def backward(ctx, g):
            st = ctx.owner._state
            if st is None:
                raise RuntimeError("AdversarialTerms: generator backward after discriminator_backward released the tape")
            with ctx.owner._modes(weights=False):
                grads = torch.autograd.grad(st.gen_loss, st.leafs, retain_graph=True)
            return (None, None) + tuple(x * g for x in grads)

    def generator_loss(self, *, target_list, pred_list, target_audio=None, pred_audio=None, used=None, index=0):
        """same keywords as the reference's GeneratorLoss.forward (`used` / `index` name the acoustic set there)"""
        if used is not None and ("pitch_disc" in used or "dur_disc" in used):
            raise ValueError("stylish_tts_b200: AdversarialTerms carries the acoustic set (mrd0-2, disc) only")
        targets, preds = list(target_list), list(pred_list)
        if self.disc is not None:
            targets.append(target_audio)
            preds.append(pred_audio)
        return self._GeneratorTermFn.apply(self, targets, *preds)

    forward = generator_loss

    def release(self) -> None:
        """drop the kept evaluation (activations of mrd0-2 / disc on target and prediction) when no discriminator
        half follows the generator term, e.g. in validation"""
        self._state = None

    def discriminator_backward(self, index: int, scale: float = 1.0) -> torch.Tensor:
        st = self._state
        if st is None:
            raise RuntimeError("AdversarialTerms: discriminator_backward needs the generator_loss of the same batch")
        keys = [f"mrd{i}" for i in range(3)] + (["disc"] if self.disc is not None else [])
        models = self._models()
        stepped = {index} | ({3} if self.disc is not None else set())
        total, back = 0.0, 0.0
        for k, key in enumerate(keys):
            with torch.set_grad_enabled(k in stepped):
                d = lsgan_discriminator(st.real[k], st.gen[k])
                self.lr_control[key].update(d)
                term = d + tprls_discriminator(st.real[k], st.gen[k])
                term = term if k < 3 else DISC_WEIGHT * term
            total = total + term.detach()
            if k in stepped:
                back = back + term
        params = [p for k in sorted(stepped) for p in models[k].parameters()]
        with self._modes(first_input=False):
            torch.autograd.backward(back * scale, inputs=params)
        self._state = None
        return total


class SimpleState:
    def __init__(self, **kw):
        self.__dict__.update(kw)
