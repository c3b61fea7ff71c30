"""Backward kernels, one primitive at a time: each ``train_ops`` Function (forward AND backward on
the CUDA kernels, through the C ABI) against plain PyTorch fp64 autograd of the op it restates, on
seeded inputs.  Tolerance on every gradient: rel-L2 <= 2e-4 (tensor-core bf16x3 path included)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import speech_oracle as so
from stylish_tts_b200 import _lib as L
from stylish_tts_b200 import train_ops as T
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-4


def g(seed):
    return torch.Generator().manual_seed(seed)


def rn(gen, *shape, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale)


def leaf(t):
    return t.double().requires_grad_(True), t.cuda().requires_grad_(True)


def check(named_pairs, out_ref, out_gpu, ct):
    """backprop the cotangent ct through both graphs and compare every leaf's gradient"""
    assert rel_l2(out_gpu, out_ref) < TOL, ("forward", rel_l2(out_gpu, out_ref))
    (out_ref * ct.double()).sum().backward()
    (out_gpu * ct.cuda()).sum().backward()
    for name, (r, c) in named_pairs.items():
        assert c.grad is not None, name
        e = rel_l2(c.grad, r.grad)
        assert e < TOL, (name, e)


def act_ref(a, act, alpha=None):
    if act == L.ACT_SNAKE:
        return so.snake(a, alpha.view(1, -1, 1))
    if act == L.ACT_LEAKY02:
        return F.leaky_relu(a, 0.2)
    if act == L.ACT_RELU:
        return torch.relu(a)
    if act == L.ACT_SWISH:
        return a * torch.sigmoid(a)
    return a


@pytest.mark.parametrize("ci,co,k,dil,T_,umma", [
    (32, 48, 5, 1, 300, True), (32, 32, 21, 1, 700, True), (128, 64, 3, 1, 258, True),
    (131, 128, 3, 1, 75, False), (20, 24, 7, 2, 130, True), (256, 1024, 1, 1, 203, True),
])
def test_conv_plain_masks_residual(ci, co, k, dil, T_, umma):
    gen = g(1)
    B = 2
    x, w, b, r = rn(gen, B, ci, T_), rn(gen, co, ci, k, scale=0.1), rn(gen, co), rn(gen, B, co, T_)
    lens = torch.tensor([T_, T_ - 37])
    mask = so.sequence_mask(lens, T_).float()
    (xr, xc), (wr, wc), (br, bc), (rr, rc) = leaf(x), leaf(w), leaf(b), leaf(r)
    m64 = mask.double()[:, None]
    ref = 0.7 * F.conv1d(xr * m64, wr, br, padding=(k - 1) * dil // 2, dilation=dil) * m64 + 0.5 * rr
    out = T.conv(xc, wc, bc, res=rc, dil=dil, in_mask=mask.cuda(), out_mask=mask.cuda(), out_scale=0.7,
                 res_scale=0.5, umma=umma)
    check(dict(x=(xr, xc), w=(wr, wc), b=(br, bc), res=(rr, rc)), ref, out, rn(gen, B, co, T_))


@pytest.mark.parametrize("ci,co,k,dil,T_,act,umma", [
    (32, 32, 11, 3, 900, L.ACT_SNAKE, True), (32, 32, 11, 1, 640, L.ACT_SNAKE, True),
    (131, 128, 3, 1, 90, L.ACT_LEAKY02, False), (128, 128, 3, 1, 200, L.ACT_LEAKY02, False),
])
def test_conv_adain_prologue(ci, co, k, dil, T_, act, umma):
    """AdaIN -> activation -> conv (ada_norm.py:109-120,176-192), incl. an F0-like channel in Hz"""
    gen = g(2)
    B = 2
    x = rn(gen, B, ci, T_)
    x[:, -1] = 200.0 + 50.0 * x[:, -1]
    w, b, gb = rn(gen, co, ci, k, scale=0.1), rn(gen, co), rn(gen, B, 2 * ci, scale=0.3)
    al = 1.0 + 0.2 * rn(gen, ci)
    (xr, xc), (wr, wc), (br, bc), (gr, gc), (ar, ac) = leaf(x), leaf(w), leaf(b), leaf(gb), leaf(al)
    n = F.instance_norm(xr, eps=1e-5)
    a = (1 + gr[:, :ci, None]) * n + gr[:, ci:, None]
    ref = F.conv1d(act_ref(a, act, ar), wr, br, padding=(k - 1) * dil // 2, dilation=dil)
    out = T.conv(xc, wc, bc, gb=gc, alpha=ac if act == L.ACT_SNAKE else None, dil=dil, in_act=act,
                 norm="instance", eps=1e-5, umma=umma)
    pairs = dict(x=(xr, xc), w=(wr, wc), b=(br, bc), gb=(gr, gc))
    if act == L.ACT_SNAKE:
        pairs["alpha"] = (ar, ac)
    check(pairs, ref, out, rn(gen, B, co, T_))


def test_conv_batchnorm_swish_prologue():
    """training-mode BatchNorm1d -> Swish -> pointwise conv (conformer.py:183-186)"""
    gen = g(3)
    B, ci, co, T_ = 3, 64, 32, 150
    x, w, b = rn(gen, B, ci, T_) + 0.5, rn(gen, co, ci, 1, scale=0.1), rn(gen, co)
    bw, bb = 1.0 + 0.1 * rn(gen, ci), 0.1 * rn(gen, ci)
    (xr, xc), (wr, wc), (br, bc), (wr2, wc2), (br2, bc2) = leaf(x), leaf(w), leaf(b), leaf(bw), leaf(bb)
    rm, rv = torch.zeros(ci).double(), torch.ones(ci).double()
    a = F.batch_norm(xr, rm, rv, wr2, br2, training=True, momentum=0.1, eps=1e-5)
    ref = F.conv1d(a * torch.sigmoid(a), wr, br)
    bufs = (torch.zeros(ci).cuda(), torch.ones(ci).cuda())
    out = T.conv(xc, wc, bc, bn_w=wc2, bn_b=bc2, in_act=L.ACT_SWISH, norm="batch", eps=1e-5, bn_buffers=bufs)
    check(dict(x=(xr, xc), w=(wr, wc), b=(br, bc), bn_w=(wr2, wc2), bn_b=(br2, bc2)), ref, out,
          rn(gen, B, co, T_))
    assert rel_l2(bufs[0], rm) < 1e-5 and rel_l2(bufs[1], rv) < 1e-5


@pytest.mark.parametrize("act", [L.ACT_SWISH, L.ACT_RELU])
def test_conv_activation_prologue_no_norm(act):
    gen = g(4)
    B, ci, co, T_ = 2, 64, 32, 130
    x, w, b = rn(gen, B, ci, T_), rn(gen, co, ci, 3, scale=0.1), rn(gen, co)
    mask = so.sequence_mask(torch.tensor([T_, 77]), T_).float()
    (xr, xc), (wr, wc), (br, bc) = leaf(x), leaf(w), leaf(b)
    ref = F.conv1d(act_ref(xr * mask.double()[:, None], act), wr, br, padding=1)
    out = T.conv(xc, wc, bc, in_act=act, in_mask=mask.cuda())
    check(dict(x=(xr, xc), w=(wr, wc), b=(br, bc)), ref, out, rn(gen, B, co, T_))


def test_conv_pixel_shuffle():
    gen = g(5)
    B, ci, co, s, T_ = 2, 64, 160, 5, 140
    x, w, b = rn(gen, B, ci, T_), rn(gen, co, ci, 11, scale=0.1), rn(gen, co)
    (xr, xc), (wr, wc), (br, bc) = leaf(x), leaf(w), leaf(b)
    ref = so.pixel_shuffle_1d(F.conv1d(xr, wr, br, padding=5), s)
    out = T.conv(xc, wc, bc, shuffle=s)
    check(dict(x=(xr, xc), w=(wr, wc), b=(br, bc)), ref, out, rn(gen, B, co // s, T_ * s))


@pytest.mark.parametrize("Cc,T_", [(32, 700), (64, 300), (256, 130)])
def test_convnext_tail(Cc, T_):
    """pwconv1 -> Snake -> GRN -> pwconv2 -> + residual (conv_next.py:85-93)"""
    gen = g(6)
    B, J = 2, 4 * Cc
    y, xres = rn(gen, B, Cc, T_), rn(gen, B, Cc, T_)
    w1, b1 = rn(gen, J, Cc, scale=Cc ** -0.5), 0.1 * rn(gen, J)
    w2, b2 = rn(gen, Cc, J, scale=J ** -0.5), 0.1 * rn(gen, Cc)
    al, gam, bet = 1.0 + 0.2 * rn(gen, J), 0.3 * rn(gen, J), 0.1 * rn(gen, J)
    L_ = [leaf(t) for t in (y, xres, w1, b1, w2, b2, al, gam, bet)]
    (yr, yc), (xr, xc), (w1r, w1c), (b1r, b1c), (w2r, w2c), (b2r, b2c), (ar, ac), (gr, gc), (ber, bec) = L_
    h = so.snake(F.linear(yr.transpose(1, 2), w1r, b1r), ar.view(1, 1, -1))
    hg = so.grn(h, gr.view(1, 1, -1), ber.view(1, 1, -1))
    ref = F.linear(hg, w2r, b2r).transpose(1, 2) + xr
    b2f = b2c + w2c @ bec  # GRN beta folded into the bias (plain autograd on (C,)-sized tensors)
    out = T.ConvNeXtTailFn.apply(yc, xc, w1c, b1c, ac, gc, w2c, b2f, True)
    names = "y xres w1 b1 w2 b2 alpha gamma beta".split()
    check(dict(zip(names, L_)), ref, out, rn(gen, B, Cc, T_))


@pytest.mark.parametrize("Cc,T_,adaptive", [(128, 258, False), (32, 1500, False), (256, 130, True), (32, 900, True)])
def test_chan_layernorm(Cc, T_, adaptive):
    gen = g(7)
    B = 2
    x, r = rn(gen, B, Cc, T_), rn(gen, B, Cc, T_)
    mask = so.sequence_mask(torch.tensor([T_, T_ - 30]), T_).float()
    (xr, xc), (rr, rc) = leaf(x), leaf(r)
    if adaptive:
        gb = rn(gen, B, 2 * Cc + 10, scale=0.3)  # rows of a wider style-FC output
        gr, gc = leaf(gb)
        n = F.layer_norm((xr + rr).transpose(1, 2), (Cc,), eps=1e-5).transpose(1, 2)
        ref = (1 + gr[:, :Cc, None]) * n + gr[:, Cc:2 * Cc, None]
        out = T.chan_ln(xc, res=rc, gb=gc[:, :2 * Cc], eps=1e-5)
        pairs = dict(x=(xr, xc), res=(rr, rc), gb=(gr, gc))
    else:
        gm, bt = 1 + 0.1 * rn(gen, Cc), 0.1 * rn(gen, Cc)
        (gr, gc), (br, bc) = leaf(gm), leaf(bt)
        ref = torch.relu(so.channel_layernorm(xr + rr, gr, br, 1e-4)) * mask.double()[:, None]
        out = T.chan_ln(xc, res=rc, gamma=gc, beta=bc, eps=1e-4, mask=mask.cuda(), act=L.ACT_RELU)
        pairs = dict(x=(xr, xc), res=(rr, rc), gamma=(gr, gc), beta=(br, bc))
    check(pairs, ref, out, rn(gen, B, Cc, T_))


@pytest.mark.parametrize("Cc,K,T_", [(32, 7, 800), (512, 31, 203), (1, 3, 100)])
def test_dwconv(Cc, K, T_):
    gen = g(8)
    B = 2
    x, w, b = rn(gen, B, Cc, T_), rn(gen, Cc, 1, K, scale=0.3), rn(gen, Cc)
    (xr, xc), (wr, wc), (br, bc) = leaf(x), leaf(w), leaf(b)
    ref = F.conv1d(xr, wr, br, padding=K // 2, groups=Cc)
    out = T.DwConvFn.apply(xc, wc, bc, K, K // 2)
    check(dict(x=(xr, xc), w=(wr, wc), b=(br, bc)), ref, out, rn(gen, B, Cc, T_))


@pytest.mark.parametrize("H,D,T_,lens,use_rope", [(8, 16, 258, [258, 200], True), (8, 16, 40, [40, 17], True),
                                                  (8, 64, 203, None, False), (8, 64, 804, None, False),
                                                  (4, 64, 64, None, False), (2, 64, 385, None, False)])
def test_attention(H, D, T_, lens, use_rope):
    """forward + backward against torch SDPA autograd in fp64; the 64-wide unmasked cases take the tcgen05 forward
    and the tcgen05 backward (dQ kernel + dK/dV kernel, csrc/attention64.cu)"""
    gen = g(9)
    B, n = 2, H * D
    qkv = rn(gen, B, 3 * n, T_)
    qr, qc = leaf(qkv)
    q, k, v = (so.heads_split(t.contiguous(), H) for t in (qr[:, :n], qr[:, n:2 * n], qr[:, 2 * n:]))
    am, lengths, rope = None, None, None
    if use_rope:
        q, k = so.rope(q, 8), so.rope(k, 8)
        c = torch.empty(T_, 4, device="cuda")
        s = torch.empty(T_, 4, device="cuda")
        L.call("sty_rope_table", c.data_ptr(), s.data_ptr(), T_, 8, 10000.0, L.stream_ptr())
        rope = (c, s, 8)
    if lens is not None:
        lengths = torch.tensor(lens)
        m = so.sequence_mask(lengths, T_).double()
        am = ((1 - m[:, None, :, None] * m[:, None, None, :]) * -1e4)
        lengths = lengths.cuda()
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=am, scale=1.0 / math.sqrt(D))
    ref = o.permute(0, 1, 3, 2).reshape(B, n, T_)
    out = T.AttentionFn.apply(qc, H, D, lengths, rope, 1.0 / math.sqrt(D))
    check(dict(qkv=(qr, qc)), ref, out, rn(gen, B, n, T_))


def test_glu_embed_bmm_linear_rows():
    gen = g(10)
    B, Cc, T_, Fr = 2, 64, 50, 120
    x = rn(gen, B, 2 * Cc, T_)
    xr, xc = leaf(x)
    check(dict(x=(xr, xc)), F.glu(xr, dim=1), T.GluFn.apply(xc), rn(gen, B, Cc, T_))

    emb = rn(gen, 30, Cc)
    tok = torch.randint(0, 30, (B, T_), generator=gen)
    lens = torch.tensor([T_, 31])
    er, ec = leaf(emb)
    m = so.sequence_mask(lens, T_).double()[:, None]
    ref = (er[tok] * math.sqrt(Cc)).transpose(1, 2) * m
    out = T.EmbedFn.apply(ec, tok.cuda(), lens.cuda(), math.sqrt(Cc))
    check(dict(emb=(er, ec)), ref, out, rn(gen, B, Cc, T_))

    mu, al = rn(gen, B, Cc, T_), torch.softmax(rn(gen, B, T_, Fr), 1)
    mr, mc = leaf(mu)
    check(dict(mu=(mr, mc)), mr @ al.double(), T.BmmAlignFn.apply(mc, al.cuda()), rn(gen, B, Cc, Fr))

    s, W, b = rn(gen, B, 64), rn(gen, 300, 64, scale=0.1), rn(gen, 300)
    (sr, sc), (Wr, Wc), (br, bc) = leaf(s), leaf(W), leaf(b)
    check(dict(s=(sr, sc), W=(Wr, Wc), b=(br, bc)), F.linear(sr, Wr, br), T.LinearRowsFn.apply(sc, Wc, bc),
          rn(gen, B, 300))


def test_istft_head():
    """exp / atan2 head + literal conv-iSTFT + tanh (generator.py:782-799,896)"""
    from stylish_tts_b200.modules import stft_buffers

    gen = g(11)
    B, Hs, S = 2, 32, 333
    bufs = stft_buffers(64, 64)
    sd = {"s." + k: v.double() for k, v in bufs.state_dict().items()}
    la, ri = rn(gen, B, Hs, S, scale=0.5), rn(gen, B, 2 * Hs, S)
    (lr, lc), (rr, rc) = leaf(la), leaf(ri)
    phase = torch.atan2(rr[:, Hs:], rr[:, :Hs])
    logamp = F.pad(lr, (0, 1), mode="replicate")
    phase = F.pad(phase, (0, 1), mode="replicate")
    spec_full = torch.zeros(B, 33, S + 1, dtype=torch.float64)
    ph_full = torch.zeros(B, 33, S + 1, dtype=torch.float64)
    spec_full = torch.cat([torch.exp(logamp), spec_full[:, Hs:]], 1)
    ph_full = torch.cat([phase, ph_full[:, Hs:]], 1)
    ref = torch.tanh(so.stft_inverse(sd, "s", spec_full, torch.cos(ph_full), torch.sin(ph_full)))
    b_re = bufs.weight_backward_real.reshape(-1, 64).contiguous().cuda()
    b_im = bufs.weight_backward_imag.reshape(-1, 64).contiguous().cuda()
    out = T.IstftHeadFn.apply(lc, rc, b_re, b_im, 4)
    check(dict(logamp=(lr, lc), ri=(rr, rc)), ref, out, rn(gen, B, 1, 4 * S))


@pytest.mark.parametrize("ci,co,k,dil,T_", [
    (32, 32, 21, 1, 1000), (32, 32, 11, 3, 777), (32, 32, 11, 5, 640), (96, 32, 21, 1, 700), (32, 64, 21, 1, 600),
    (32, 128, 1, 1, 1111), (128, 32, 1, 1, 513), (256, 1024, 1, 1, 203), (1024, 256, 1, 1, 203),
    (128, 512, 3, 1, 258), (512, 128, 3, 1, 258), (128, 128, 5, 1, 258), (64, 160, 11, 1, 600),
    (128, 256, 21, 1, 300), (48, 24, 3, 1, 515), (256, 384, 11, 1, 100), (32, 32, 7, 2, 90),
])
@pytest.mark.parametrize("pro", [False, True])
def test_wgrad_tensor_core_vs_fp64(ci, co, k, dil, T_, pro):
    """sty_conv1d_wgrad on the tcgen05 path (taps folded into the M side) against fp64 autograd, and against
    the fp32 FMA kernel where that one is built"""
    gen = g(12)
    B = 3
    x, dy = rn(gen, B, ci, T_), rn(gen, B, co, T_)
    lens = torch.tensor([T_, T_ - 21, T_ // 2])
    mask = so.sequence_mask(lens, T_).float()
    kw = {}
    xin = x.double()
    gy = dy.double()
    if pro:
        sc, sh, al = 1 + 0.3 * rn(gen, B, ci), 0.2 * rn(gen, B, ci), 1 + 0.2 * rn(gen, ci)
        xin = so.snake(sc.double()[:, :, None] * (xin * mask.double()[:, None]) + sh.double()[:, :, None],
                       al.double().view(1, -1, 1))
        gy = gy * mask.double()[:, None] * 0.7
        kw = dict(in_scale=sc.cuda(), in_shift=sh.cuda(), in_alpha=al.cuda(), in_act=L.ACT_SNAKE,
                  in_mask=mask.cuda(), out_mask=mask.cuda(), out_scale=0.7)
    w = torch.zeros(co, ci, k, dtype=torch.float64, requires_grad=True)
    (F.conv1d(xin, w, padding=(k - 1) * dil // 2, dilation=dil) * gy).sum().backward()
    dw_tc = T.wgrad(x.cuda(), dy.cuda(), k, dil, umma=True, **kw)
    assert rel_l2(dw_tc, w.grad) < 5e-5, rel_l2(dw_tc, w.grad)
    if k in (1, 3, 5, 7, 11, 21):
        dw_simt = T.wgrad(x.cuda(), dy.cuda(), k, dil, umma=False, **kw)
        assert rel_l2(dw_simt, w.grad) < 2e-5
