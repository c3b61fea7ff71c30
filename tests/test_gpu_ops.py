"""Per-kernel parity: every C-ABI entry point against the matching piece of the CPU
oracle (oracle/speech_oracle.py) or the plain PyTorch fp32 op it restates, on seeded
inputs.  Tolerance: fp32 GPU vs fp32 CPU, rel-L2 <= 2e-5 unless stated (the north-star
end-to-end tolerance is 1e-3)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import speech_oracle as so
from stylish_tts_b200 import _lib as L
from stylish_tts_b200 import engine as E
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-5


def dev():
    return torch.device("cuda:0")


def g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("ci,co,k,dil,T", [
    (32, 32, 21, 1, 1000), (96, 32, 21, 1, 517), (32, 32, 11, 3, 700), (32, 32, 11, 5, 300),
    (32, 128, 1, 1, 999), (128, 32, 1, 1, 260), (128, 512, 3, 1, 258), (512, 128, 3, 1, 40),
    (131, 128, 3, 1, 75), (256, 384, 11, 1, 50), (1, 1, 3, 1, 33), (128, 1, 1, 1, 64),
    (20, 24, 7, 2, 130),
])
def test_conv1d_plain(ci, co, k, dil, T):
    gen = g(ci * 1000 + co + k)
    B = 3
    x = torch.randn(B, ci, T, generator=gen)
    w = torch.randn(co, ci, k, generator=gen) / math.sqrt(ci * k)
    b = torch.randn(co, generator=gen)
    ref = F.conv1d(x, w, b, padding=(k - 1) * dil // 2, dilation=dil)
    out = E.conv1d(x.to(dev()), E.ConvW(w.to(dev()), b.to(dev())), dil=dil)
    assert rel_l2(out, ref) < TOL


def test_conv1d_prologue_epilogue():
    """AdaIN affine + snake prologue, mask, residual, scale, strided in/out views."""
    gen = g(7)
    B, ci, co, T, k = 2, 32, 32, 333, 11
    big = torch.randn(B, 96, T, generator=gen)
    x = big[:, 32:64]
    w = torch.randn(co, ci, k, generator=gen) / math.sqrt(ci * k)
    bias = torch.randn(co, generator=gen)
    sc = torch.randn(B, ci, generator=gen)
    sh = torch.randn(B, ci, generator=gen)
    al = 0.5 + torch.rand(ci, generator=gen)
    res = torch.randn(B, co, T, generator=gen)
    mask = (torch.rand(B, T, generator=gen) > 0.3).float()
    xin = sc[:, :, None] * x + sh[:, :, None]
    xin = so.snake(xin, al.view(1, -1, 1))
    ref = F.conv1d(xin, w, bias, padding=5)
    ref = F.leaky_relu(ref, 0.2) * mask[:, None] * 0.7 + 0.3 * res
    d = dev()
    bigd = big.to(d)
    outbuf = torch.zeros(B, 64, T, device=d)
    E.conv1d(bigd[:, 32:64], E.ConvW(w.to(d), bias.to(d)), in_scale=sc.to(d), in_shift=sh.to(d),
             in_alpha=al.to(d), in_act=L.ACT_SNAKE, out_act=L.ACT_LEAKY02, out_mask=mask.to(d),
             res=res.to(d), out_scale=0.7, res_scale=0.3, out=outbuf[:, 16:48])
    assert rel_l2(outbuf[:, 16:48], ref) < TOL
    assert float(outbuf[:, :16].abs().max()) == 0.0 and float(outbuf[:, 48:].abs().max()) == 0.0


def test_conv1d_in_mask_relu_leaky():
    gen = g(8)
    B, ci, co, T, k = 2, 128, 128, 77, 5
    x = torch.randn(B, ci, T, generator=gen)
    w = torch.randn(co, ci, k, generator=gen) / math.sqrt(ci * k)
    bias = torch.randn(co, generator=gen)
    mask = (torch.arange(T)[None] < torch.tensor([[T], [T // 2]])).float()
    ref = torch.relu(F.conv1d(x * mask[:, None], w, bias, padding=2))
    d = dev()
    out = E.conv1d(x.to(d), E.ConvW(w.to(d), bias.to(d)), in_mask=mask.to(d), out_act=L.ACT_RELU)
    assert rel_l2(out, ref) < TOL
    sc = torch.randn(B, ci, generator=gen)
    sh = torch.randn(B, ci, generator=gen)
    ref2 = F.conv1d(F.leaky_relu(sc[:, :, None] * x + sh[:, :, None], 0.2), w, bias, padding=2)
    out2 = E.conv1d(x.to(d), E.ConvW(w.to(d), bias.to(d)), in_scale=sc.to(d), in_shift=sh.to(d),
                    in_act=L.ACT_LEAKY02)
    assert rel_l2(out2, ref2) < TOL


@pytest.mark.parametrize("s,ci,co,T", [(3, 256, 384, 50), (5, 64, 160, 203)])
def test_conv1d_pixel_shuffle(s, ci, co, T):
    gen = g(s)
    B = 2
    x = torch.randn(B, ci, T, generator=gen)
    w = torch.randn(co, ci, 11, generator=gen) / math.sqrt(ci * 11)
    b = torch.randn(co, generator=gen)
    ref = so.pixel_shuffle_1d(F.conv1d(x, w, b, padding=5), s)
    d = dev()
    out = E.conv1d(x.to(d), E.ConvW(w.to(d), b.to(d)), shuffle=s)
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < TOL


def test_conv1d_snake_epilogue_sumsq_and_grn():
    """pwconv1 + Snake + GRN statistics + GRN-scaled pwconv2 (conv_next.py:85-90)."""
    gen = g(9)
    B, Cc, T = 2, 32, 1234
    inter = 4 * Cc
    y = torch.randn(B, Cc, T, generator=gen)
    w1 = torch.randn(inter, Cc, generator=gen) / math.sqrt(Cc)
    b1 = torch.randn(inter, generator=gen) * 0.1
    al = 0.75 + 0.5 * torch.rand(inter, generator=gen)
    gam = torch.randn(inter, generator=gen) * 0.3
    bet = torch.randn(inter, generator=gen) * 0.1
    w2 = torch.randn(Cc, inter, generator=gen) / math.sqrt(inter)
    b2 = torch.randn(Cc, generator=gen) * 0.1
    res = torch.randn(B, Cc, T, generator=gen)
    h = so.snake(F.linear(y.transpose(1, 2), w1, b1), al.view(1, 1, -1))
    hg = so.grn(h, gam.view(1, 1, -1), bet.view(1, 1, -1))
    ref = res + F.linear(hg, w2, b2).transpose(1, 2)
    d = dev()
    sumsq = torch.zeros(B, inter, device=d)
    hb = E.conv1d(y.to(d), E.ConvW(w1.unsqueeze(-1).to(d), b1.to(d)), out_act=L.ACT_SNAKE,
                  out_alpha=al.to(d), out_sumsq=sumsq)
    assert rel_l2(hb, h.transpose(1, 2)) < TOL
    assert rel_l2(sumsq, (h ** 2).sum(1)) < TOL
    gs = torch.empty_like(sumsq)
    gamd = gam.to(d)
    L.call("sty_grn_scale_fwd", sumsq.data_ptr(), gamd.data_ptr(), gs.data_ptr(), B, inter,
           L.stream_ptr())
    b2f = b2 + w2 @ bet
    xr = res.to(d)
    E.conv1d(hb, E.ConvW(w2.unsqueeze(-1).to(d), b2f.to(d)), in_scale=gs, res=xr, out=xr)
    assert rel_l2(xr, ref) < TOL


@pytest.mark.parametrize("Cc,T", [(32, 1000), (64, 300), (128, 258), (256, 75)])
def test_chan_layernorm(Cc, T):
    gen = g(Cc)
    B = 3
    x = torch.randn(B, Cc, T, generator=gen) * 2 + 0.5
    r = torch.randn(B, Cc, T, generator=gen)
    gm = 1 + 0.1 * torch.randn(Cc, generator=gen)
    bt = 0.1 * torch.randn(Cc, generator=gen)
    mask = (torch.rand(B, T, generator=gen) > 0.2).float()
    d = dev()
    ref = torch.relu(so.channel_layernorm(x + r, gm, bt, 1e-4)) * mask[:, None]
    out = E.chan_layernorm(x.to(d), gm.to(d), bt.to(d), eps=1e-4, res=r.to(d), mask=mask.to(d),
                           act=L.ACT_RELU)
    assert rel_l2(out, ref) < TOL
    # adaptive (1+gamma) per batch, eps 1e-6, in place
    gb = torch.randn(B, 2 * Cc + 5, generator=gen) * 0.3
    ref2 = (1 + gb[:, None, :Cc]) * F.layer_norm(x.transpose(1, 2), (Cc,), eps=1e-6) + gb[:, None, Cc:2 * Cc]
    gbd = gb.to(d)
    xd = x.to(d)
    E.chan_layernorm(xd, gbd, gbd[:, Cc:], eps=1e-6, g_bs=gb.shape[1], plus_one=True, out=xd)
    assert rel_l2(xd, ref2.transpose(1, 2)) < TOL


@pytest.mark.parametrize("Cc,T", [(32, 6000), (131, 75), (195, 803)])
def test_instnorm_affine(Cc, T):
    gen = g(Cc + T)
    B = 2
    x = torch.randn(B, Cc, T, generator=gen) * 3
    x[:, 0] += 150.0  # F0-in-Hz like channel: mean^2 >> var
    gb = torch.randn(B, 2 * Cc + 3, generator=gen) * 0.3
    d = dev()
    sc, sh = E.instnorm_affine(x.to(d), gb.to(d), gb.shape[1])
    ref = (1 + gb[:, :Cc, None]) * F.instance_norm(x, eps=1e-5) + gb[:, Cc:2 * Cc, None]
    out = sc.cpu()[:, :, None] * x + sh.cpu()[:, :, None]
    assert rel_l2(out, ref) < 5e-5


@pytest.mark.parametrize("Cc,T", [(32, 3000), (64, 500), (128, 240), (256, 80)])
def test_dwconv_ln(Cc, T):
    gen = g(Cc * 3)
    B = 2
    big = torch.randn(B, Cc + 8, T, generator=gen)
    w = torch.randn(Cc, 1, 7, generator=gen) * 0.4
    b = torch.randn(Cc, generator=gen) * 0.1
    gb = torch.randn(B, 2 * Cc, generator=gen) * 0.3
    x = big[:, :Cc]
    dwc = F.conv1d(x, w, b, padding=3, groups=Cc).transpose(1, 2)
    ref = ((1 + gb[:, None, :Cc]) * F.layer_norm(dwc, (Cc,), eps=1e-6) + gb[:, None, Cc:]).transpose(1, 2)
    d = dev()
    bigd = big.to(d)
    y = torch.empty(B, Cc, T, device=d)
    wd, bd, gbd = w.reshape(Cc, 7).contiguous().to(d), b.to(d), gb.to(d)  # keep alive
    L.call("sty_dwconv_ln_fwd", bigd.data_ptr(), bigd.stride(0), wd.data_ptr(), bd.data_ptr(),
           gbd.data_ptr(), 2 * Cc, y.data_ptr(), y.stride(0), B, Cc, T, 1e-6, L.stream_ptr())
    assert rel_l2(y, ref) < TOL


def test_dwconv1d_bn_swish():
    gen = g(31)
    B, Cc, T, K = 2, 512, 90, 31
    x = torch.randn(B, Cc, T, generator=gen)
    w = torch.randn(Cc, 1, K, generator=gen) / math.sqrt(K)
    b = torch.randn(Cc, generator=gen) * 0.1
    ps = 1 + 0.2 * torch.randn(Cc, generator=gen)
    pt = 0.1 * torch.randn(Cc, generator=gen)
    c = F.conv1d(F.pad(x, (15, 15)), w, b, groups=Cc) * ps[None, :, None] + pt[None, :, None]
    ref = c * torch.sigmoid(c)
    d = dev()
    y = torch.empty(B, Cc, T, device=d)
    E.dwconv1d(x.to(d), w.reshape(Cc, K).contiguous().to(d), b.to(d), K=K, pad_left=15, out=y,
               post_scale=ps.to(d), post_shift=pt.to(d), act=L.ACT_SWISH)
    assert rel_l2(y, ref) < TOL


@pytest.mark.parametrize("T,lens", [(258, [258, 131, 200]), (40, [40, 7, 1])])
def test_attention_text(T, lens):
    """RoPE(8 of 16 dims) + -1e4 mask + softmax, 8 heads x 16 (text_encoder.py:233-272)."""
    gen = g(T)
    B, H, D = len(lens), 8, 16
    qkv = torch.randn(B, 3 * H * D, T, generator=gen)
    lengths = torch.tensor(lens)
    q, k, v = qkv[:, :128], qkv[:, 128:256], qkv[:, 256:]
    qh, kh, vh = (so.heads_split(t.contiguous(), H) for t in (q, k, v))
    qh, kh = so.rope(qh, 8), so.rope(kh, 8)
    mask = so.sequence_mask(lengths, T).float()
    am = mask[:, None, :, None] * mask[:, None, None, :]
    scores = qh @ kh.transpose(2, 3) / 4.0 + (1 - am) * -1e4
    ref = (torch.softmax(scores, -1) @ vh).transpose(2, 3).reshape(B, H * D, T)
    d = dev()
    c = torch.empty(T, 4, device=d)
    s = torch.empty(T, 4, device=d)
    L.call("sty_rope_table", c.data_ptr(), s.data_ptr(), T, 8, 10000.0, L.stream_ptr())
    out = E.attention(qkv.to(d), 128, 128, 128, H=H, D=D, lengths=lengths.to(d), rope=(c, s, 8),
                      scale=0.25)
    # valid query rows must match tightly; padded rows are garbage-in/garbage-out in the
    # reference too but still deterministic, so compare everything
    assert rel_l2(out, ref) < 5e-5


@pytest.mark.parametrize("second_gen", [True, False])
@pytest.mark.parametrize("T", [203, 803, 64, 129, 258, 1030])
def test_attention_conformer(T, second_gen, monkeypatch):
    """8 heads x 64, no mask: the tcgen05 flash kernels (bf16x3) against fp64 softmax attention.
    second_gen: pre-split operands + cp.async.bulk tiles (csrc/attention64.cu); else the first kernel."""
    monkeypatch.setattr(E, "ATTENTION64", second_gen)
    gen = g(64)
    B, H, D = 2, 8, 64
    qkv = torch.randn(B, 3 * H * D, T, generator=gen) * 1.5
    q, k, v = (so.heads_split(t.contiguous().double(), H) for t in (qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:]))
    ref = (torch.softmax(q @ k.transpose(2, 3) * D ** -0.5, -1) @ v).transpose(2, 3).reshape(B, H * D, T)
    out = E.attention(qkv.to(dev()), 512, 512, 512, H=H, D=D, scale=D ** -0.5)
    assert rel_l2(out, ref) < 3e-5, rel_l2(out, ref)


def test_bmm_glu_embed_mask_linear():
    gen = g(5)
    d = dev()
    A = torch.randn(3, 128, 258, generator=gen)
    Bm = torch.rand(3, 258, 803, generator=gen)
    out = torch.zeros(3, 131, 803, device=d)
    Ad, Bd = A.to(d), Bm.to(d)
    L.call("sty_bmm_fwd", Ad.data_ptr(), Ad.stride(0), Bd.data_ptr(), Bd.stride(0), out.data_ptr(),
           out.stride(0), 3, 128, 803, 258, L.stream_ptr())
    assert rel_l2(out[:, :128], A @ Bm) < TOL
    assert float(out[:, 128:].abs().max()) == 0.0

    x = torch.randn(2, 1024, 77, generator=gen)
    y = torch.empty(2, 512, 77, device=d)
    xd = x.to(d)
    L.call("sty_glu_fwd", xd.data_ptr(), y.data_ptr(), 2, 512, 77, L.stream_ptr())
    assert rel_l2(y, x[:, :512] * torch.sigmoid(x[:, 512:])) < TOL

    tok = torch.randint(0, 178, (3, 50), generator=gen)
    lens = torch.tensor([50, 20, 1])
    emb = torch.randn(178, 128, generator=gen)
    o = torch.empty(3, 128, 50, device=d)
    td, ld, ed = tok.to(d), lens.to(d), emb.to(d)
    L.call("sty_embed_fwd", td.data_ptr(), ld.data_ptr(), ed.data_ptr(), o.data_ptr(), 3, 50, 128,
           178, math.sqrt(128.0), L.stream_ptr())
    m = so.sequence_mask(lens, 50).float()
    ref = (F.embedding(tok, emb) * math.sqrt(128.0)).transpose(1, 2) * m[:, None]
    assert torch.equal(o.cpu(), ref)  # gather * scale * {0,1}: bit-exact
    mo = torch.empty(3, 50, device=d)
    L.call("sty_sequence_mask_fwd", ld.data_ptr(), mo.data_ptr(), 3, 50, L.stream_ptr())
    assert torch.equal(mo.cpu(), m)

    s = torch.randn(4, 64, generator=gen)
    W = torch.randn(300, 64, generator=gen)
    bb = torch.randn(300, generator=gen)
    ho = torch.empty(4, 300, device=d)
    sd_, Wd, bd = s.to(d), W.to(d), bb.to(d)
    L.call("sty_linear_rows_fwd", sd_.data_ptr(), Wd.data_ptr(), bd.data_ptr(), ho.data_ptr(), 4, 64,
           300, L.stream_ptr())
    assert rel_l2(ho, F.linear(s, W, bb)) < TOL


@pytest.mark.parametrize("ci,co,k,dil,T", [
    (32, 32, 21, 1, 1000), (96, 32, 21, 1, 700), (32, 64, 21, 1, 600), (32, 32, 11, 3, 777),
    (32, 32, 11, 5, 640), (32, 128, 1, 1, 1111), (128, 32, 1, 1, 513), (256, 1024, 1, 1, 803),
    (256, 1536, 1, 1, 515), (256, 384, 11, 1, 803),
    (1024, 256, 1, 1, 803), (128, 256, 21, 1, 803), (64, 160, 11, 1, 600), (48, 16, 3, 1, 515),
    (384, 1152, 3, 1, 101), (1920, 384, 5, 1, 101), (128, 128, 3, 1, 64),  # short rows (style encoder, T < 128)
])
def test_conv1d_tensor_core_bf16x3(ci, co, k, dil, T):
    """tcgen05 path (bf16 hi/lo split, 3 MMAs, fp32 accumulate): ~16-bit operands."""
    gen = g(ci * 7 + co + k)
    B = 2
    x = torch.randn(B, ci, T, generator=gen)
    w = torch.randn(co, ci, k, generator=gen) / math.sqrt(ci * k)
    b = torch.randn(co, generator=gen)
    ref = F.conv1d(x.double(), w.double(), b.double(), padding=(k - 1) * dil // 2, dilation=dil)
    cw = E.ConvW(w.to(dev()), b.to(dev()))
    assert cw.split is not None
    before = L.launches
    out = E.conv1d(x.to(dev()), cw, dil=dil)
    assert rel_l2(out, ref) < 3e-5, rel_l2(out, ref)


def test_conv1d_tensor_core_fused_ops():
    """prologue (affine+snake), epilogue (snake, sumsq, residual, shuffle) on the tcgen05 path."""
    gen = g(77)
    B, ci, co, T, k = 2, 32, 32, 900, 11
    x = torch.randn(B, ci, T, generator=gen)
    w = torch.randn(co, ci, k, generator=gen) / math.sqrt(ci * k)
    bias = torch.randn(co, generator=gen)
    sc, sh = torch.randn(B, ci, generator=gen), torch.randn(B, ci, generator=gen)
    al = 0.5 + torch.rand(ci, generator=gen)
    res = torch.randn(B, co, T, generator=gen)
    xin = so.snake(sc[:, :, None] * x + sh[:, :, None], al.view(1, -1, 1))
    ref = F.conv1d(xin.double(), w.double(), bias.double(), padding=5).float() + res
    d = dev()
    cw = E.ConvW(w.to(d), bias.to(d))
    r = res.to(d)
    E.conv1d(x.to(d), cw, in_scale=sc.to(d), in_shift=sh.to(d), in_alpha=al.to(d),
             in_act=L.ACT_SNAKE, res=r, out=r)
    assert rel_l2(r, ref) < 3e-5
    # snake epilogue + sum of squares (pwconv1 of the ConvNeXt block)
    w1 = torch.randn(128, 32, 1, generator=gen) / math.sqrt(32)
    b1 = torch.randn(128, generator=gen) * 0.1
    a1 = 0.75 + 0.5 * torch.rand(128, generator=gen)
    h = so.snake(F.conv1d(x.double(), w1.double(), b1.double()), a1.view(1, -1, 1).double())
    ssq = torch.zeros(B, 128, device=d)
    hb = E.conv1d(x.to(d), E.ConvW(w1.to(d), b1.to(d)), out_act=L.ACT_SNAKE, out_alpha=a1.to(d),
                  out_sumsq=ssq)
    assert rel_l2(hb, h) < 3e-5
    assert rel_l2(ssq, (h ** 2).sum(2)) < 3e-5
    # pixel shuffle store
    w2 = torch.randn(160, 32, 11, generator=gen) / math.sqrt(32 * 11)
    b2 = torch.randn(160, generator=gen)
    ref2 = so.pixel_shuffle_1d(F.conv1d(x.double(), w2.double(), b2.double(), padding=5), 5)
    out2 = E.conv1d(x.to(d), E.ConvW(w2.to(d), b2.to(d)), shuffle=5)
    assert rel_l2(out2, ref2) < 3e-5


@pytest.mark.parametrize("Cc,T", [(32, 1000), (64, 700)])
def test_convnext_front_fused_into_pointwise_conv(Cc, T):
    """depthwise k7 + LayerNorm + adaptive affine computed by the tcgen05 conv's producer warps,
    then pwconv1 + Snake + GRN sums (conv_next.py:82-86), on a strided input view."""
    gen = g(Cc + 5)
    B, inter = 2, 4 * Cc
    big = torch.randn(B, Cc + 16, T, generator=gen)
    x = big[:, :Cc]
    dw_w = torch.randn(Cc, 1, 7, generator=gen) * 0.4
    dw_b = torch.randn(Cc, generator=gen) * 0.1
    gb = torch.randn(B, 2 * Cc + 4, generator=gen) * 0.3
    w1 = torch.randn(inter, Cc, generator=gen) / math.sqrt(Cc)
    b1 = torch.randn(inter, generator=gen) * 0.1
    al = 0.75 + 0.5 * torch.rand(inter, generator=gen)
    d = F.conv1d(x.double(), dw_w.double(), dw_b.double(), padding=3, groups=Cc).transpose(1, 2)
    y = (1 + gb[:, None, :Cc].double()) * F.layer_norm(d, (Cc,), eps=1e-6) + gb[:, None, Cc:2 * Cc].double()
    h = so.snake(F.linear(y, w1.double(), b1.double()), al.view(1, 1, -1).double()).transpose(1, 2)
    dv = dev()
    bigd, gbd = big.to(dv), gb.to(dv)
    ssq = torch.zeros(B, inter, device=dv)
    cw = E.ConvW(w1.unsqueeze(-1).to(dv), b1.to(dv))
    dww, dwb, ald = dw_w.reshape(Cc, 7).contiguous().to(dv), dw_b.to(dv), al.to(dv)
    hb = E.conv1d(bigd[:, :Cc], cw, out_act=L.ACT_SNAKE, out_alpha=ald, out_sumsq=ssq,
                  dwln=(dww, dwb, gbd, gb.shape[1], 1e-6))
    assert rel_l2(hb, h) < 3e-5
    assert rel_l2(ssq, (h ** 2).sum(2)) < 3e-5


@pytest.mark.parametrize("T", [900, 1504])
def test_conv1d_tensor_core_gelu_prologue_and_post_mask(T):
    """generic-activation prologue (IN_MODE 5: affine + GELU) on the tcgen05 path, with the "negative mask value =
    zero AFTER the prologue" convention of end-to-end window layouts (waveform discriminator) and an output mask;
    T = 1504 has 16-byte aligned rows (TMA producers), 900 takes the load path as well (T < 512 rule aside)"""
    gen = g(78)
    B, ci, co, k = 2, 64, 48, 5
    x = torch.randn(B, ci, T, generator=gen)
    w = torch.randn(co, ci, k, generator=gen) / math.sqrt(ci * k)
    bias = torch.randn(co, generator=gen)
    sc, sh = torch.randn(B, ci, generator=gen), torch.randn(B, ci, generator=gen)
    keep = ((torch.arange(T) % 20) < 16).float().unsqueeze(0).expand(B, -1).contiguous()
    xin = F.gelu(sc[:, :, None] * x.double() + sh[:, :, None]) * keep[:, None, :]
    ref = F.conv1d(xin, w.double(), bias.double(), padding=k // 2) * keep[:, None, :]
    d = dev()
    calls = []
    orig = L.call
    L.call = lambda name, *a: (calls.append((name, a)), orig(name, *a))[1]
    try:
        out = E.conv1d(x.to(d), E.ConvW(w.to(d), bias.to(d)), in_scale=sc.to(d), in_shift=sh.to(d),
                       in_act=L.ACT_GELU, in_mask=(2 * keep - 1).to(d), out_mask=keep.to(d))
    finally:
        L.call = orig
    assert calls[0][1][0]._obj.w_split  # tensor-core path requested (and taken: same result on the FMA path below)
    assert rel_l2(out, ref) < 3e-5, rel_l2(out, ref)
    out_fma = E.conv1d(x.to(d), E.ConvW(w.to(d), bias.to(d)), in_scale=sc.to(d), in_shift=sh.to(d),
                       in_act=L.ACT_GELU, in_mask=(2 * keep - 1).to(d), out_mask=keep.to(d), umma=False)
    assert rel_l2(out_fma, ref) < 2e-6, rel_l2(out_fma, ref)


@pytest.mark.parametrize("umma", [True, False])
def test_conv1d_epilogue_moments_feed_adain(umma):
    """out_sum / out_sumsq accumulated by the producing conv + sty_moments_affine_fwd == the two-pass
    InstanceNorm statistics of sty_instnorm_affine_fwd (ada_norm.py:129-140)"""
    gen = g(77)
    B, ci, co, k, T = 3, 32, 32, 11, 1500
    x = torch.randn(B, ci, T, generator=gen)
    w = torch.randn(co, ci, k, generator=gen) / math.sqrt(ci * k)
    b = torch.randn(co, generator=gen) * 0.5
    res = torch.randn(B, co, T, generator=gen)
    gb = torch.randn(B, 2 * co, generator=gen) * 0.3
    cw = E.ConvW(w.to(dev()), b.to(dev()))
    mom = torch.zeros(2, B, co, device=dev())
    y = E.conv1d(x.to(dev()), cw, res=res.to(dev()), out_sum=mom[0], out_sumsq=mom[1], umma=umma)
    ref = F.conv1d(x.double(), w.double(), b.double(), padding=5) + res.double()
    assert rel_l2(y, ref) < 3e-5
    assert rel_l2(mom[0], ref.sum(-1)) < 1e-4 and rel_l2(mom[1], (ref * ref).sum(-1)) < 1e-4
    sc, sh = E.moments_affine(mom, gb.to(dev()), 2 * co, T)
    sc2, sh2 = E.instnorm_affine(y, gb.to(dev()), 2 * co)
    assert rel_l2(sc, sc2) < 2e-5 and rel_l2(sh, sh2) < 5e-5


@pytest.mark.parametrize("T,B", [(1000, 2), (2301, 3), (60225, 2)])
def test_convnext_block_fused_no_intermediate(T, B):
    """The whole GeneratorConvNeXtBlock (conv_next.py:80-93, GRN :7-18) as ONE call that never stores the
    4C-wide intermediate (csrc/convnext_fused.cu: TMA-fed, two passes) against the fp64 oracle block; odd
    lengths (pitch-padded rows, partial last tile), several tiles per CTA at T = 60 225, and the output written
    into a channel slice of a wider buffer like the engine does."""
    gen = g(T + B)
    Cc, inter, sdim = 32, 128, 64
    p = "blk"
    sd = {
        p + ".dwconv.weight": torch.randn(Cc, 1, 7, generator=gen) * 0.4,
        p + ".dwconv.bias": torch.randn(Cc, generator=gen) * 0.1,
        p + ".norm.fc.weight": torch.randn(2 * Cc, sdim, generator=gen) * 0.05,
        p + ".norm.fc.bias": torch.randn(2 * Cc, generator=gen) * 0.1,
        p + ".pwconv1.weight": torch.randn(inter, Cc, generator=gen) / math.sqrt(Cc),
        p + ".pwconv1.bias": torch.randn(inter, generator=gen) * 0.1,
        p + ".snake": 0.75 + 0.5 * torch.rand(1, 1, inter, generator=gen),
        p + ".grn.gamma": torch.randn(1, 1, inter, generator=gen) * 0.3,
        p + ".grn.beta": torch.randn(1, 1, inter, generator=gen) * 0.1,
        p + ".pwconv2.weight": torch.randn(Cc, inter, generator=gen) / math.sqrt(inter),
        p + ".pwconv2.bias": torch.randn(Cc, generator=gen) * 0.1,
    }
    x = torch.randn(B, Cc, T, generator=gen)
    style = torch.randn(B, sdim, generator=gen)
    ref = so.convnext_block({k: v.double() for k, v in sd.items()}, p, x.double(), style.double())
    dv = dev()
    w2 = sd[p + ".pwconv2.weight"]
    blk = dict(dw_w=sd[p + ".dwconv.weight"].reshape(Cc, 7).contiguous().to(dv), dw_b=sd[p + ".dwconv.bias"].to(dv),
               norm="n", pw1=E.ConvW(sd[p + ".pwconv1.weight"].unsqueeze(-1).to(dv), sd[p + ".pwconv1.bias"].to(dv)),
               snake=sd[p + ".snake"].reshape(-1).contiguous().to(dv),
               grn_gamma=sd[p + ".grn.gamma"].reshape(-1).contiguous().to(dv),
               pw2=E.ConvW(w2.unsqueeze(-1).to(dv),
                           (sd[p + ".pwconv2.bias"] + w2 @ sd[p + ".grn.beta"].reshape(-1)).to(dv)))
    hfc = (style @ sd[p + ".norm.fc.weight"].t() + sd[p + ".norm.fc.bias"]).to(dv).contiguous()  # gamma | beta rows
    P = type("P", (), dict(fc_rows=2 * Cc, fc_off={"n": 0}))()
    eng = E.SpeechEngine.__new__(E.SpeechEngine)
    xd = E.empty_bct(B, Cc, T, dv).copy_(x.to(dv))
    wide = E.empty_bct(B, 3 * Cc, T, dv).fill_(7.0)
    calls = []
    orig = L.call
    L.call = lambda name, *a: (calls.append(name), orig(name, *a))[1]
    try:
        y = eng.convnext(P, blk, xd, hfc, out=wide[:, Cc:2 * Cc])
    finally:
        L.call = orig
    assert calls == ["sty_convnext_fused_fwd"], calls
    assert y.data_ptr() == wide[:, Cc:2 * Cc].data_ptr()
    assert rel_l2(y, ref) < 5e-5, rel_l2(y, ref)
    assert float((wide[:, :Cc] - 7.0).abs().max()) == 0.0 and float((wide[:, 2 * Cc:] - 7.0).abs().max()) == 0.0
    assert rel_l2(xd, x) == 0.0  # the input is not modified
    # and the two-kernel path (in place) agrees
    y2 = eng.convnext(P, blk, x.to(dv).contiguous().clone(), hfc) if T % 4 else None
    if y2 is not None:
        assert rel_l2(y2, ref) < 5e-5
