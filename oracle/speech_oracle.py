"""CPU oracle for the speech_predictor forward path.  TEST INFRASTRUCTURE ONLY.

A functional PyTorch (CPU, fp32 or fp64) restatement of the reference's
algorithm, written against a *state dict* with the reference's key names so no
reference module is needed at run time (``/root/reference`` does not exist on
the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` legs may import this; the product package
``stylish_tts_b200`` never does.

Pinned against the unmodified reference: ``tests/golden/make_golden.py`` runs
the real reference modules (imported from /root/reference in the build
container) on seeded inputs and commits their outputs under ``tests/golden``;
``tests/test_oracle_golden.py`` checks every function below against them (and
against the live reference when it is mounted).

Random draws of the harmonic source (reference generator.py:345,440,509) are
INPUTS here (``draws``), see SURVEY.md F7.

Each function cites the reference file:line it restates (paths relative to
/root/reference/src/stylish_tts/train/).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# Training-mode dropout: ``MASKS`` is None (eval / regularisers off) or a callable (site, p, shape) -> keep/(1-p)
# tensor — oracle/dropout_oracle.Masks, the numpy restatement of the kernels' mask hash.  Site ids are the ones
# stylish_tts_b200/train_engine.py uses.
MASKS = None


def _drop(x, site, p, layout="bct"):
    """dropout site on x (B,C,T): 'bct' nn.Dropout (one mask value per element), 'chan' nn.Dropout1d (per (b,c)),
    'path' DropPath (per sample); 'bnc': nn.Dropout on a token-major (B,N,C) tensor, mask laid out (B,C,N)"""
    if MASKS is None or p <= 0.0:
        return x
    if layout == "bnc":
        B, N, Cc = x.shape
        return x * MASKS(site, p, (B, Cc, N)).transpose(1, 2)
    if layout == "chan":
        return x * MASKS(site, p, tuple(x.shape[:2])).unsqueeze(-1)
    if layout == "path":
        return x * MASKS(site, p, (x.shape[0],)).reshape(-1, *([1] * (x.dim() - 1)))
    return x * MASKS(site, p, tuple(x.shape))


def _p(sd: SD, name: str) -> torch.Tensor:
    return sd[name]


def wn_weight(sd: SD, prefix: str) -> torch.Tensor:
    """weight_norm parametrization: w = g * v / ||v|| over all dims but 0
    (torch.nn.utils.parametrizations.weight_norm, used at decoder.py:36-50,
    ada_norm.py:16-83,160-173)."""
    k0 = prefix + ".parametrizations.weight.original0"
    if k0 in sd:
        g = sd[k0]
        v = sd[prefix + ".parametrizations.weight.original1"]
        return torch._weight_norm(v, g, 0)
    return sd[prefix + ".weight"]


def conv1d(sd: SD, prefix: str, x, *, padding=0, dilation=1, groups=1):
    w = wn_weight(sd, prefix)
    b = sd.get(prefix + ".bias")
    return F.conv1d(x, w, b, padding=padding, dilation=dilation, groups=groups)


def linear(sd: SD, prefix: str, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def sequence_mask(length, max_length):
    """utils.py:54-58"""
    x = torch.arange(max_length, dtype=length.dtype)
    return x.unsqueeze(0) < length.unsqueeze(1)


# ---------------------------------------------------------------------------
# text encoder
# ---------------------------------------------------------------------------
def channel_layernorm(x, gamma, beta, eps=1e-4):
    """models/text_encoder.py:24-33 — statistics over dim 1, eps 1e-4."""
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    x = (x - mean) * torch.rsqrt(var + eps)
    return x * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def rope(x, d_rot: int, base: float = 10000.0):
    """models/text_encoder.py:89-168 on (B,H,T,d): rotate-half on the first
    d_rot features, theta_i = base^(-2i/d_rot)."""
    T = x.shape[2]
    theta = 1.0 / (base ** (torch.arange(0, d_rot, 2, dtype=torch.float32) / d_rot))
    idx = torch.arange(T, dtype=torch.float32)
    ang = torch.einsum("n,d->nd", idx, theta)
    ang = torch.cat([ang, ang], dim=1).to(x.dtype)  # (T, d_rot)
    cos, sin = ang.cos()[None, None], ang.sin()[None, None]
    xr, xp = x[..., :d_rot], x[..., d_rot:]
    h = d_rot // 2
    neg_half = torch.cat([-xr[..., h:], xr[..., :h]], dim=-1)
    xr = xr * cos + neg_half * sin
    return torch.cat([xr, xp], dim=-1)


def heads_split(x, n_heads):
    """models/text_encoder.py:224-231: (B, H*d, T) -> (B, H, T, d)."""
    B, C, T = x.shape
    return x.view(B, n_heads, C // n_heads, T).permute(0, 1, 3, 2)


def mha(sd: SD, prefix: str, x, c, n_heads: int, attn_mask=None, drop=None):
    """models/text_encoder.py:214-297.  attn_mask is the 0/1 keep mask (B,1,T,T); masked entries get the
    additive value -1e4.  ``drop`` = (site, p): SDPA's dropout on the probabilities (:270-275)."""
    q = conv1d(sd, prefix + ".conv_q", x)
    k = conv1d(sd, prefix + ".conv_k", c)
    v = conv1d(sd, prefix + ".conv_v", c)
    B, C, T = q.shape
    d = C // n_heads
    d_rot = int(d * 0.5)
    q, k, v = heads_split(q, n_heads), heads_split(k, n_heads), heads_split(v, n_heads)
    q, k = rope(q, d_rot), rope(k, d_rot)
    scores = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(d)
    if attn_mask is not None:
        add = torch.zeros_like(attn_mask, dtype=q.dtype)
        add.masked_fill_(~attn_mask.to(torch.bool), -1e4)
        scores = scores + add
    p = torch.softmax(scores, dim=-1)
    if drop is not None:
        p = _drop(p, drop[0], drop[1])
    o = torch.matmul(p, v)  # (B,H,T,d)
    o = o.transpose(2, 3).contiguous().view(B, C, T)
    return conv1d(sd, prefix + ".conv_o", o)


def text_encoder(sd: SD, prefix: str, tokens, lengths, *, n_heads=8, n_layers=8,
                 kernel_size=3, taps=None, p_dropout=0.2):
    """models/text_encoder.py:434-463 (+ prenet :79-86, Encoder :378-394,
    FFN :325-330).  Returns (mu, x, mask)."""
    emb = sd[prefix + ".emb.weight"]
    C = emb.shape[1]
    x = F.embedding(tokens, emb) * math.sqrt(C)
    x = x.transpose(1, -1)
    mask = sequence_mask(lengths, x.size(2)).unsqueeze(1).to(x.dtype)
    # prenet
    x_org = x
    for i in range(3):
        x = conv1d(sd, f"{prefix}.prenet.conv_layers.{i}", x * mask, padding=2)
        x = channel_layernorm(x, sd[f"{prefix}.prenet.norm_layers.{i}.gamma"],
                              sd[f"{prefix}.prenet.norm_layers.{i}.beta"])
        x = _drop(torch.relu(x), 1 + i, 0.5)  # ConvReluNorm.relu_drop, p_dropout=0.5 (:63,418)
    x = x_org + conv1d(sd, prefix + ".prenet.proj", x)
    x = x * mask
    if taps is not None:
        taps["prenet"] = x
    # encoder
    attn_mask = mask.unsqueeze(2) * mask.unsqueeze(-1)
    e = prefix + ".encoder"
    pad = kernel_size // 2
    for i in range(n_layers):
        x = x * mask
        s0 = 16 + 4 * i
        y = mha(sd, f"{e}.attn_layers.{i}", x, x, n_heads, attn_mask, drop=(s0, p_dropout))
        y = _drop(y, s0 + 1, p_dropout)
        x = channel_layernorm(x + y, sd[f"{e}.norm_layers_1.{i}.gamma"],
                              sd[f"{e}.norm_layers_1.{i}.beta"])
        y = conv1d(sd, f"{e}.ffn_layers.{i}.conv_1", x * mask, padding=pad)
        y = _drop(torch.relu(y), s0 + 2, p_dropout)
        y = conv1d(sd, f"{e}.ffn_layers.{i}.conv_2", y * mask, padding=pad)
        y = _drop(y * mask, s0 + 3, p_dropout)
        x = channel_layernorm(x + y, sd[f"{e}.norm_layers_2.{i}.gamma"],
                              sd[f"{e}.norm_layers_2.{i}.beta"])
        if taps is not None and i == 0:
            taps["enc_layer0"] = x
    x = x * mask
    mu = conv1d(sd, prefix + ".proj_m", x) * mask
    return mu, x, mask


# ---------------------------------------------------------------------------
# style-adaptive norms and blocks
# ---------------------------------------------------------------------------
def adain(sd: SD, prefix: str, x, s, eps=1e-5):
    """models/ada_norm.py:129-140 — InstanceNorm1d(affine=False) over time
    (biased variance, eps 1e-5) then (1+gamma)*x+beta."""
    h = linear(sd, prefix + ".fc", s).unsqueeze(-1)
    gamma, beta = torch.chunk(h, 2, dim=1)
    return (1 + gamma) * F.instance_norm(x, eps=eps) + beta


def adaln(sd: SD, prefix: str, x, s, eps):
    """models/ada_norm.py:203-211 — x is (B,T,C)."""
    h = linear(sd, prefix + ".fc", s).unsqueeze(1)  # (B,1,2C)
    gamma, beta = torch.chunk(h, 2, dim=-1)
    x = F.layer_norm(x, (x.shape[-1],), eps=eps)
    return (1 + gamma) * x + beta


def snake(x, alpha):
    """x + (1/a) sin^2(a x)  (ada_norm.py:114,117; conv_next.py:77-78)"""
    return x + (1 / alpha) * (torch.sin(alpha * x) ** 2)


def decoder_block(sd: SD, prefix: str, x, s, drop=(0, 0.0)):
    """models/ada_norm.py:143-192; ``drop`` = (site, p) of the two nn.Dropout before conv1 / conv2 (:183,186)."""
    h = adain(sd, prefix + ".norm1", x, s)
    h = F.leaky_relu(h, 0.2)
    h = conv1d(sd, prefix + ".conv1", _drop(h, drop[0], drop[1]), padding=1)
    h = adain(sd, prefix + ".norm2", h, s)
    h = F.leaky_relu(h, 0.2)
    h = conv1d(sd, prefix + ".conv2", _drop(h, drop[0] + 1, drop[1]), padding=1)
    sc = x
    if (prefix + ".conv1x1.parametrizations.weight.original0") in sd or \
            (prefix + ".conv1x1.weight") in sd:
        sc = conv1d(sd, prefix + ".conv1x1", x)
    return (h + sc) / math.sqrt(2)


def box_smooth(x, width):
    """decoder.py:58-75: zero-padded moving average of odd ``width`` over (B,F)"""
    if not width:
        return x
    w = torch.ones(1, 1, width, dtype=x.dtype)
    return F.conv1d(x.unsqueeze(1), w, padding=width // 2).squeeze(1) / width


def decoder(sd: SD, prefix: str, asr, f0_curve, n, s, voiced, taps=None, smoothing=(0, 0)):
    """models/decoder.py:52-90; ``smoothing`` = the (F0, N) box widths the train-only branch :53-75 drew."""
    f0_curve, n = box_smooth(f0_curve, smoothing[0]), box_smooth(n, smoothing[1])
    f0 = conv1d(sd, prefix + ".F0_conv", f0_curve.unsqueeze(1), padding=1)
    nn_ = conv1d(sd, prefix + ".N_conv", n.unsqueeze(1), padding=1)
    vo = conv1d(sd, prefix + ".voiced_conv", voiced.unsqueeze(1), padding=1)
    x = torch.cat([asr, f0, nn_, vo], dim=1)
    x = decoder_block(sd, prefix + ".encode", x, s)
    if taps is not None:
        taps["dec_encode"] = x
    asr_res = conv1d(sd, prefix + ".asr_res.0", asr)
    for i in range(4):
        x = torch.cat([x, asr_res, f0, nn_, vo], dim=1)
        x = decoder_block(sd, f"{prefix}.decode.{i}", x, s)
    return x, f0_curve


def generator_block(sd: SD, prefix: str, x, s):
    """models/ada_norm.py:109-120 — AdaptiveGeneratorBlock, k11, dil (1,3,5)."""
    for i, d in enumerate((1, 3, 5)):
        a1 = sd[f"{prefix}.alpha1.{i}"]
        a2 = sd[f"{prefix}.alpha2.{i}"]
        xt = adain(sd, f"{prefix}.adain1.{i}", x, s)
        xt = snake(xt, a1)
        xt = conv1d(sd, f"{prefix}.convs1.{i}", xt, padding=5 * d, dilation=d)
        xt = adain(sd, f"{prefix}.adain2.{i}", xt, s)
        xt = snake(xt, a2)
        xt = conv1d(sd, f"{prefix}.convs2.{i}", xt, padding=5)
        x = xt + x
    return x


def grn(x, gamma, beta):
    """models/conv_next.py:15-18 on (B,T,C): L2 over time, mean over C."""
    gx = torch.norm(x, p=2, dim=1, keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    return gamma * (x * nx) + beta + x


def convnext_block(sd: SD, prefix: str, x, s, taps=None):
    """models/conv_next.py:80-93 — GeneratorConvNeXtBlock on (B,C,T)."""
    C = x.shape[1]
    r = x
    x = conv1d(sd, prefix + ".dwconv", x, padding=3, groups=C)
    x = x.transpose(1, 2)
    x = adaln(sd, prefix + ".norm", x, s, 1e-6)
    x = linear(sd, prefix + ".pwconv1", x)
    x = snake(x, sd[prefix + ".snake"])
    if taps is not None:
        taps["cnx_h"] = x
    x = grn(x, sd[prefix + ".grn.gamma"], sd[prefix + ".grn.beta"])
    x = linear(sd, prefix + ".pwconv2", x)
    return r + x.transpose(1, 2)


# ---------------------------------------------------------------------------
# conformer (models/conformer.py)
# ---------------------------------------------------------------------------
def _swish(x):
    return x * torch.sigmoid(x)


def conformer_ff(sd: SD, prefix: str, x, s):
    """Scale(0.5, PreNorm(FeedForward)) conformer.py:59-95.  The block's nn.Dropout modules all have p = 0.0:
    Conformer.__init__ does not forward the rates generator.py:824-826 asks for (conformer.py:284-296)."""
    h = adaln(sd, prefix + ".fn.norm", x, s, 1e-5)
    h = linear(sd, prefix + ".fn.fn.net.0", h)
    h = _swish(h)
    h = linear(sd, prefix + ".fn.fn.net.3", h)
    return 0.5 * h


def conformer_attn(sd: SD, prefix: str, x, s, heads=8):
    """PreNorm(Attention) conformer.py:99-143: 8 heads x 64, scale 1/8, no mask."""
    h = adaln(sd, prefix + ".norm", x, s, 1e-5)
    q = F.linear(h, sd[prefix + ".fn.to_q.weight"])
    kv = F.linear(h, sd[prefix + ".fn.to_kv.weight"])
    k, v = kv.chunk(2, dim=-1)
    B, N, I = q.shape
    d = I // heads

    def sp(t):
        return t.view(B, N, heads, d).permute(0, 2, 1, 3)

    q, k, v = sp(q), sp(k), sp(v)
    p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    o = torch.matmul(p, v).permute(0, 2, 1, 3).reshape(B, N, I)
    return linear(sd, prefix + ".fn.to_out", o)


def conformer_conv(sd: SD, prefix: str, x, s, bn_training=False):
    """ConformerConvModule conformer.py:160-193 (eval BatchNorm: running stats; ``bn_training``:
    batch statistics, as nn.BatchNorm1d does in train mode)."""
    h = adaln(sd, prefix + ".norm", x, s, 1e-5).transpose(1, 2)  # (B,C,N)
    h = conv1d(sd, prefix + ".net.1", h)
    a, g = h.chunk(2, dim=1)
    h = a * torch.sigmoid(g)
    C = h.shape[1]
    h = F.pad(h, (15, 15))
    h = conv1d(sd, prefix + ".net.3.conv", h, groups=C)
    bn = prefix + ".net.4"
    if bn_training:
        h = F.batch_norm(h, None, None, sd[bn + ".weight"], sd[bn + ".bias"], training=True, eps=1e-5)
    else:
        h = F.batch_norm(h, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"],
                         sd[bn + ".bias"], training=False, eps=1e-5)
    h = _swish(h)
    h = conv1d(sd, prefix + ".net.6", h)
    return h.transpose(1, 2)


def conformer_block(sd: SD, prefix: str, x, s, bn_training=False):
    """conformer.py:242-250 — note attention reads the block INPUT x."""
    x_ff1 = conformer_ff(sd, prefix + ".ff1", x, s) + x
    x = conformer_attn(sd, prefix + ".attn", x, s)
    x = x + x_ff1
    x = conformer_conv(sd, prefix + ".conv", x, s, bn_training) + x
    x = conformer_ff(sd, prefix + ".ff2", x, s) + x
    return adaln(sd, prefix + ".post_norm", x, s, 1e-5)


# ---------------------------------------------------------------------------
# harmonic source prior + conv-STFT (models/generator.py:295-510,711-729; stft.py)
# ---------------------------------------------------------------------------
def source_prior(sd: SD, prefix: str, pitch, voiced, draws, *, hop=300, sr=24000,
                 harmonics=9, sine_amp=0.1, noise_std=0.003, vthresh=10.0):
    """generator.py:719-723 + SourceModuleHnNSF.forward :496-510 +
    SineGen.forward/_f02sine :336-447.  draws = {"rand_ini": (B,9) uniform,
    "noise": (B,L,9) normal}.  Returns the merged excitation (B,L)."""
    frames = pitch.shape[1]
    f0 = F.interpolate((pitch * voiced)[:, None], scale_factor=float(hop), mode="linear")
    f0 = f0.transpose(1, 2)  # (B,L,1)
    mult = torch.arange(1, harmonics + 1, dtype=f0.dtype).view(1, 1, -1)
    fn = f0 * mult
    rad = (fn / sr) % 1
    rand_ini = draws["rand_ini"].clone().to(f0.dtype)
    rand_ini[:, 0] = 0
    rad[:, 0, :] = rad[:, 0, :] + rand_ini
    rad = F.interpolate(rad.transpose(1, 2), size=frames, mode="linear").transpose(1, 2)
    phase = torch.cumsum(rad, dim=1) * 2 * torch.pi
    phase = F.interpolate(phase.transpose(1, 2) * hop, scale_factor=float(hop),
                          mode="linear").transpose(1, 2)
    sines = torch.sin(phase) * sine_amp
    uv = (f0 > vthresh).to(f0.dtype)
    noise_amp = uv * noise_std + (1 - uv) * sine_amp / 3
    noise = noise_amp * draws["noise"].to(f0.dtype)
    sines = sines * uv + noise
    merged = torch.tanh(linear(sd, prefix + ".m_source.l_linear", sines))
    return merged.squeeze(2)


def stft_transform(sd: SD, prefix: str, wave, hop=4):
    """models/stft.py:98-136 (center, replicate pad n_fft/2)."""
    wr = sd[prefix + ".weight_forward_real"]
    wi = sd[prefix + ".weight_forward_imag"]
    n_fft = wr.shape[-1]
    x = F.pad(wave, (n_fft // 2, n_fft // 2), mode="replicate").unsqueeze(1)
    re = F.conv1d(x, wr, stride=hop)
    im = F.conv1d(x, wi, stride=hop)
    mag = torch.sqrt(re ** 2 + im ** 2 + 1e-14)
    return mag, re / mag, im / mag


def stft_inverse(sd: SD, prefix: str, mag, x, y, hop=4):
    """models/stft.py:138-187 — literal: no window-envelope normalisation."""
    wr = sd[prefix + ".weight_backward_real"]
    wi = sd[prefix + ".weight_backward_imag"]
    n_fft = wr.shape[-1]
    rr = F.conv_transpose1d(mag * x, wr, stride=hop)
    ir = F.conv_transpose1d(mag * y, wi, stride=hop)
    w = rr - ir
    return w[..., n_fft // 2: -(n_fft // 2)]


def harmonic_prior(sd: SD, prefix: str, pitch, voiced, draws, hidden=32):
    """generator.py:711-729 -> (har_spec, har_phase), each (B,hidden,S)."""
    prior = source_prior(sd, prefix, pitch, voiced, draws)
    mag, hx, hy = stft_transform(sd, prefix + ".stft", prior)
    har_spec = mag[:, :hidden, :-1]
    har_phase = torch.atan2(hy, hx)[:, :hidden, :-1]
    return har_spec, har_phase, prior


# ---------------------------------------------------------------------------
# Generator / MultiGenerator / SpeechPredictor
# ---------------------------------------------------------------------------
def layer_norm_ct(x, w, b, eps=1e-6):
    """nn.LayerNorm over channels applied to (B,C,T) via transposes
    (generator.py:756-758,770-772,776-778)."""
    return F.layer_norm(x.transpose(1, 2), (x.shape[1],), w, b, eps).transpose(1, 2)


def pixel_shuffle_1d(x, s):
    """einops 'b (c s) t -> b c (t s)' (generator.py:746)."""
    B, CS, T = x.shape
    return x.view(B, CS // s, s, T).permute(0, 1, 3, 2).reshape(B, CS // s, T * s)


def basegen(sd: SD, prefix: str, mel, style, har_spec, har_phase, *, amp_layers=5,
            conv_layers=8, rates=(3, 5, 5), hidden=32, taps=None):
    """Generator.forward generator.py:731-799 given the harmonic prior."""
    g = prefix
    logamp_prior = conv1d(sd, g + ".amp_prior_conv", har_spec, padding=10)
    logamp_prior = generator_block(sd, g + ".amp_prior_block", logamp_prior, style)
    phase_prior = conv1d(sd, g + ".phase_prior_conv", har_phase, padding=10)
    phase_prior = generator_block(sd, g + ".phase_prior_block", phase_prior, style)
    if taps is not None:
        taps["logamp_prior"] = logamp_prior
        taps["phase_prior"] = phase_prior
    for i in range(amp_layers):
        mel = convnext_block(sd, f"{g}.amp_convnext.{i}", mel, style)
    if taps is not None:
        taps["amp_convnext"] = mel
    for i, s in enumerate(rates):
        mel = conv1d(sd, f"{g}.upconvs.{i}", mel, padding=5)
        mel = pixel_shuffle_1d(mel, s)
        mel = convnext_block(sd, f"{g}.upblocks.{i}", mel, style)
    if taps is not None:
        taps["upsampled"] = mel
    logamp = layer_norm_ct(mel, sd[g + ".amp_final_layer_norm.weight"],
                           sd[g + ".amp_final_layer_norm.bias"])
    logamp = conv1d(sd, g + ".amp_output_conv", logamp, padding=10)
    phase_in = torch.cat([mel, logamp_prior, phase_prior], dim=1)
    phase = conv1d(sd, g + ".phase_input_conv", phase_in, padding=10)
    phase = layer_norm_ct(phase, sd[g + ".phase_norm.weight"], sd[g + ".phase_norm.bias"])
    for i in range(conv_layers):
        phase = convnext_block(sd, f"{g}.phase_convnext.{i}", phase, style)
    phase = layer_norm_ct(phase, sd[g + ".phase_final_layer_norm.weight"],
                          sd[g + ".phase_final_layer_norm.bias"])
    real = conv1d(sd, g + ".phase_output_real_conv", phase, padding=10)
    imag = conv1d(sd, g + ".phase_output_imag_conv", phase, padding=10)
    if taps is not None:
        taps["logamp"] = logamp
        taps["real"] = real
        taps["imag"] = imag
    phase = torch.atan2(imag, real)
    logamp = F.pad(logamp, (0, 1), mode="replicate")
    phase = F.pad(phase, (0, 1), mode="replicate")
    spec = torch.exp(logamp)
    B, _, N = spec.shape
    bins = sd[g + ".stft.weight_backward_real"].shape[0]
    spec_full = torch.zeros(B, bins, N, dtype=spec.dtype)
    spec_full[:, :hidden] = spec
    phase_full = torch.zeros(B, bins, N, dtype=spec.dtype)
    phase_full[:, :hidden] = phase
    return stft_inverse(sd, g + ".stft", spec_full, torch.cos(phase_full),
                        torch.sin(phase_full))


def multi_generator(sd: SD, prefix: str, mel, style, pitch, voiced, draws=None, *,
                    prior=None, taps=None, bn_training=False):
    """MultiGenerator.forward generator.py:884-901.  Either ``draws`` (RNG of
    the source) or an injected ``prior=(har_spec, har_phase)`` must be given."""
    x = conv1d(sd, prefix + ".amp_input_conv", mel, padding=10)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd[prefix + ".amp_norm.weight"],
                     sd[prefix + ".amp_norm.bias"], 1e-6)
    if taps is not None:
        taps["amp_norm"] = x.transpose(1, 2)
    x = conformer_block(sd, prefix + ".amp_conformer.layers.0", x, style, bn_training)
    x = x.transpose(1, 2)
    if taps is not None:
        taps["conformer"] = x
    if prior is None:
        with torch.no_grad():  # generator.py:711 — the whole prior branch is built without a graph
            har_spec, har_phase, wave = harmonic_prior(sd, prefix + ".basegen", pitch, voiced, draws)
        if taps is not None:
            taps["prior_wave"] = wave
    else:
        har_spec, har_phase = prior
    if taps is not None:
        taps["har_spec"] = har_spec
        taps["har_phase"] = har_phase
    audio = basegen(sd, prefix + ".basegen", x, style, har_spec, har_phase, taps=taps)
    return torch.tanh(audio)


def speech_predictor(sd: SD, texts, text_lengths, alignment, pitch, energy, voiced, style,
                     denormal_pitch, draws=None, *, prior=None,
                     taps: Optional[dict] = None, bn_training=False, smoothing=(0, 0)):
    """SpeechPredictor.forward speech_predictor.py:47-73 -> audio (B,1,L).  Training-mode regularisers:
    set ``MASKS`` (dropout) and ``smoothing`` (decoder.py:53-75)."""
    mu, _, _ = text_encoder(sd, "text_encoder", texts, text_lengths, taps=taps)
    if taps is not None:
        taps["text_encoding"] = mu
    asr = mu @ alignment
    mel, _ = decoder(sd, "decoder", asr, pitch, energy, style, voiced, taps=taps, smoothing=smoothing)
    if taps is not None:
        taps["decoder"] = mel
    return multi_generator(sd, "generator", mel, style, denormal_pitch, voiced, draws,
                           prior=prior, taps=taps, bn_training=bn_training)


# ---------------------------------------------------------------------------
# alignment (utils.py:752-791)
# ---------------------------------------------------------------------------
def duration_to_alignment(duration, multiplier=1):
    """DurationProcessor.duration_to_alignment utils.py:752-791."""
    total = int(duration.sum(dim=1).round().max().long().item()) * multiplier
    duration = duration * multiplier
    upper = torch.cumsum(duration, dim=1)
    lower = upper - duration
    mean = ((lower + upper) / 2).unsqueeze(2)
    seq = torch.arange(round(total)).view(1, 1, -1)
    x = seq - mean
    a = 1 - (x * 2 / (duration.unsqueeze(2) + 6)) ** 2
    lower = lower - 3
    upper = upper + 3
    m = (seq > lower.unsqueeze(2)) * (seq < upper.unsqueeze(2))
    a = torch.clamp(a * m, min=0.0)
    return torch.softmax(a, dim=1)


def to_dtype(sd: SD, dtype) -> SD:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# ---------------------------------------------------------------------------
# duration / pitch-energy predictors and the inference graph (E7-E9, A1)
# ---------------------------------------------------------------------------
def adaptive_convnext_block(sd: SD, prefix: str, x, s, site=0, p=0.0):
    """models/conv_next.py:125-141 — AdaptiveConvNeXtBlock, GELU(erf); DropPath(p) on the branch (:130,138-153)."""
    C = x.shape[1]
    r = x
    x = conv1d(sd, prefix + ".dwconv", x, padding=3, groups=C)
    x = x.transpose(1, 2)
    x = adaln(sd, prefix + ".norm", x, s, 1e-6)
    x = linear(sd, prefix + ".pwconv1", x)
    x = F.gelu(x)
    x = grn(x, sd[prefix + ".grn.gamma"], sd[prefix + ".grn.beta"])
    x = linear(sd, prefix + ".pwconv2", x)
    return r + _drop(x.transpose(1, 2), site, p, "path")


def mha_qc(sd: SD, prefix: str, x, c, n_heads, attn_mask):
    """MultiHeadAttention.forward with distinct query / context inputs (text_encoder.py:214-222)."""
    return mha(sd, prefix, x, c, n_heads, attn_mask)


def duration_predictor(sd: SD, texts, text_lengths, style, *, n_layer=3, taps=None, last_dropout=0.5):
    """models/duration_predictor.py:58-87 -> (B,T,classes) monotone logits.  With ``MASKS`` set: train() mode
    (cross-attention dropout 0.5 :40, DropPath 0.5 :25, Dropout1d(last_dropout) :30,79)."""
    enc, _, _ = text_encoder(sd, "text_encoder", texts, text_lengths)
    enc = enc.transpose(1, 2)  # the reference's "b t c -> b c t" rearrange of a (B,C,T) tensor
    mask = sequence_mask(text_lengths, enc.size(1)).unsqueeze(1).to(enc.dtype)
    q = adaln(sd, "query_norm", enc, style, 1e-5).transpose(1, 2)
    k = adaln(sd, "key_norm", enc, style, 1e-5).transpose(1, 2)
    am = mask.unsqueeze(2) * mask.unsqueeze(-1)
    att = mha(sd, "cross_attention", q, k, 8, am, drop=(80, 0.5))
    C = att.shape[1]
    att = conv1d(sd, "cross_post.0", att, padding=2, groups=C)
    att = F.silu(att)
    att = conv1d(sd, "cross_post.2", att)
    pros = (att + enc.transpose(1, 2)) / math.sqrt(2.0)
    if taps is not None:
        taps["dur_cross"] = pros
    for i in range(n_layer):
        pros = adaptive_convnext_block(sd, f"conv_next.{i}", pros, style, 84 + 2 * i, 0.5)
        pros = pros * mask
        pros = _drop(pros, 85 + 2 * i, last_dropout, "chan")
    pros = pros.transpose(1, 2)
    d = linear(sd, "duration_proj.linear_layer", pros)
    d = torch.cat([d[:, :, :1], torch.abs(d)[:, :, 1:]], dim=2)
    d = -torch.abs(torch.cumsum(d, dim=2))
    return d * mask.transpose(1, 2)


def prosody_encoder(sd: SD, prefix: str, x, style, lengths, *, n_layers=3, n_heads=2, p=0.2):
    """models/prosody_encoder.py:63-81 -> (B,T,d_model+style); dropout p (pitch_energy_predictor.py:27)."""
    mask = sequence_mask(lengths, x.size(2)).unsqueeze(1).to(x.dtype)
    am = mask.unsqueeze(2) * mask.unsqueeze(-1)
    st = style.unsqueeze(2).expand(x.shape[0], -1, x.shape[2])
    x = torch.cat([x, st], dim=1)
    for i in range(n_layers):
        x = x * mask
        s0 = 80 + 4 * i
        y = mha(sd, f"{prefix}.attn_layers.{i}", x, x, n_heads, am, drop=(s0, p))
        y = _drop(y, s0 + 1, p)
        x = adaln(sd, f"{prefix}.norm_layers_1.{i}", (x + y).transpose(1, 2), style, 1e-5).transpose(1, 2)
        y = conv1d(sd, f"{prefix}.ffn_layers.{i}.conv_1", x * mask)
        y = _drop(torch.relu(y), s0 + 2, p)
        y = _drop(conv1d(sd, f"{prefix}.ffn_layers.{i}.conv_2", y * mask) * mask, s0 + 3, p)
        x = adaln(sd, f"{prefix}.norm_layers_2.{i}", (x + y).transpose(1, 2), style, 1e-5).transpose(1, 2)
        x = conv1d(sd, f"{prefix}.proj_layers.{i}", x)
        x = torch.cat([x, st], dim=1)
    x = x * mask
    return x.transpose(-1, -2)


def pitch_energy_predictor(sd: SD, texts, text_lengths, alignment, style, taps=None, dropout=0.2):
    """models/pitch_energy_predictor.py:62-82 -> (pitch (B,F), energy (B,F)); ``dropout`` = the towers'
    AdaptiveDecoderBlock dropout_p (:22), live when ``MASKS`` is set."""
    enc, _, _ = text_encoder(sd, "text_encoder", texts, text_lengths)
    pros = prosody_encoder(sd, "prosody_encoder", enc, style, text_lengths)
    if taps is not None:
        taps["prosody"] = pros
    x = pros.transpose(1, 2) @ alignment
    f0 = x
    for i in range(4):
        f0 = decoder_block(sd, f"F0.{i}", f0, style, (96 + 2 * i, dropout))
    f0 = conv1d(sd, "F0_proj", f0)
    n = x
    for i in range(4):
        n = decoder_block(sd, f"N.{i}", n, style, (104 + 2 * i, dropout))
    n = conv1d(sd, "N_proj", n)
    return f0.squeeze(1), n.squeeze(1)


CLASS_TO_DUR = [1, 2, 3, 4, 5, 6, 7, 9, 12, 15, 18, 22, 27, 32, 38, 46]


def prediction_to_duration(pred, text_lengths):
    """DurationProcessor.prediction_to_duration / class_to_dur_soft utils.py:726-750."""
    table = torch.tensor(CLASS_TO_DUR, dtype=pred.dtype)
    p = torch.softmax(pred, dim=-1)
    soft = (p * table).sum(dim=-1) / (p.sum(dim=-1) + 1e-9)
    return soft * sequence_mask(text_lengths, pred.shape[1])


def synthesize(sds, texts, text_lengths, speech_style, pe_style, duration_style, draws_fn, taps=None):
    """ExportModel.forward export_model.py:40-63, batched.  `sds` = dict of the three state dicts;
    `draws_fn(frames)` supplies the source noise once the (data-dependent) length is known."""
    dur_pred = duration_predictor(sds["duration_predictor"], texts, text_lengths, duration_style)
    dur = prediction_to_duration(dur_pred, text_lengths)
    alignment = duration_to_alignment(dur)
    pitch, energy = pitch_energy_predictor(sds["pitch_energy_predictor"], texts, text_lengths,
                                           alignment, pe_style)
    voiced = (pitch > 20).to(pitch.dtype)
    audio = speech_predictor(sds["speech_predictor"], texts, text_lengths, alignment, pitch, energy,
                             voiced, speech_style, pitch, draws_fn(alignment.shape[2]), taps=taps)
    return audio, dict(dur_pred=dur_pred, duration=dur, alignment=alignment, pitch=pitch, energy=energy)
