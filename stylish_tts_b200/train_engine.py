"""Differentiable forward of ``speech_predictor`` for training (configs 3 and 5).

Same stage order as ``engine.SpeechEngine.forward`` (reference speech_predictor.py:47-73 ->
text_encoder.py:434-463 -> decoder.py:77-90 -> generator.py:884-901, 710-799) but built from the
autograd primitives of ``train_ops`` (forward AND backward on our CUDA kernels) applied to the
module's LIVE parameters, so ``loss.backward()`` fills ``.grad`` of every ``nn.Parameter`` and of the
``pitch`` / ``energy`` / ``style`` inputs (stage_type.py:415-448 trains the PE predictor through them).

Training-mode semantics: BatchNorm1d of the conformer conv module uses batch statistics and updates
its running buffers (conformer.py:183); the harmonic prior is computed without a graph exactly as the
reference does under ``torch.no_grad()`` (generator.py:711-729).  The stochastic regularisers are active
when the module is in ``train()`` mode and ``module.regularisers`` is on (the default): every nn.Dropout /
SDPA dropout_p / Dropout1d / DropPath site of the reference is a ``train_ops.dropout`` / attention-dropout
site whose mask is a stateless hash of (device seed, site, element) — see ``DropoutRng`` — and the decoder's
random box smoothing (decoder.py:53-75) draws its widths from Python's ``random`` like the reference.
``begin_step()`` advances the seed and the widths once per forward (``auto_step``), or is called by
``runtime.GraphedAcousticStep`` before each replay.
"""
from __future__ import annotations

import math
import random
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _lib as L
from ._lib import ACT_LEAKY02, ACT_RELU, ACT_SNAKE, ACT_SWISH
from . import train_ops as T

INV_SQRT2 = 1.0 / math.sqrt(2.0)


class TrainGraph:
    """One differentiable forward over the live parameters of a ``SpeechPredictor`` shell."""

    def __init__(self, module, engine=None):
        self.m = module
        self.engine = engine  # the inference engine (speech predictor: harmonic prior kernels)
        self.rebind()
        self.mc = module.model_config
        names = sorted(k[:-len(".fc.weight")] for k in self.P if k.endswith(".fc.weight"))
        self.fc_names = names
        self.fc_off, off = {}, 0
        for n in names:
            self.fc_off[n] = off
            off += self.P[n + ".fc.weight"].shape[0]
        self.fc_rows = off
        self._rope: Dict[tuple, tuple] = {}
        self.rng: Optional[T.DropoutRng] = None  # created on first stochastic forward
        self.auto_step = True    # forward() calls begin_step() itself (False: a graph runner does)
        self.active = None       # the DropoutRng while a stochastic forward is being built, else None
        self.smooth_w = None     # (2, 1, 31) box filters of the decoder's F0 / N smoothing (decoder.py:53-75)
        self.smooth = (0, 0)

    def rebind(self):
        """(re)read the module's parameters and buffers: nn.Module.to() / .cuda() / .double() REBIND buffer tensors
        (and may rebind parameters), so the shells call this from ``_apply``."""
        self.P: Dict[str, torch.Tensor] = dict(self.m.named_parameters())
        self.Bf: Dict[str, torch.Tensor] = dict(self.m.named_buffers())
        self._rope = {}
        self.smooth_w = None

    # ---------------------------------------------------------------- stochastic regularisers
    def stochastic(self) -> bool:
        return bool(self.m.training and getattr(self.m, "regularisers", True))

    def begin_step(self, device=None):
        """next step's randomness: new mask seed; new F0 / N smoothing widths from ``random`` in the reference's
        order (decoder.py:54-57).  Host -> device writes happen here, never inside forward()."""
        device = device or next(iter(self.P.values())).device
        if self.rng is None:
            self.rng = T.DropoutRng(getattr(self.m, "regulariser_seed", 0), device)
        elif self.auto_step:
            self.rng = self.rng.fork()  # eager: a seed cell (and smoothing filter) per forward, see fork()
        else:
            self.rng.advance()
        if any(k.startswith("decoder.F0_conv") for k in self.P):
            f0 = [0, 7, 15][random.randint(0, 2)]
            nn_ = [0, 7, 15, 31][random.randint(0, 3)]
            self.smooth = (f0, nn_)
            rows = []
            for wd in (f0, nn_):
                wd = wd or 1
                rows.append([1.0 / wd if abs(i - 15) <= wd // 2 else 0.0 for i in range(31)])
            host = torch.tensor(rows, dtype=torch.float32).reshape(2, 1, 31)
            if self.smooth_w is None or self.auto_step:
                self.smooth_w = host.to(device)
            else:
                self.smooth_w.copy_(host)

    def start_forward(self, device):
        if self.stochastic():
            if self.auto_step or self.rng is None:
                self.begin_step(device)
            self.active = self.rng
        else:
            self.active = None

    def drop(self, x, site, p, **kw):
        return T.dropout(x, self.active, site, p, **kw)

    def rope(self, Tn: int, d_head: int, device):
        key = (Tn, d_head, str(device))
        if key not in self._rope:
            d_rot = int(d_head * 0.5)
            c = torch.empty((Tn, d_rot // 2), device=device, dtype=torch.float32)
            s = torch.empty_like(c)
            L.call("sty_rope_table", c.data_ptr(), s.data_ptr(), Tn, d_rot, 10000.0, L.stream_ptr())
            self._rope[key] = (c, s, d_rot)
        return self._rope[key]

    def prior_pack(self):
        """the constants ``SpeechEngine.harmonic_prior`` reads (no gradient flows through the prior)"""
        P, Bf, bg = self.P, self.Bf, "generator.basegen"
        return SimpleNamespace(
            mc=self.mc, hidden_s=self.mc.n_fft // 2 // 8,
            src_w=P[bg + ".m_source.l_linear.weight"].detach().reshape(-1).contiguous(),
            src_b=P[bg + ".m_source.l_linear.bias"].detach().contiguous(),
            stft_f_re=Bf[bg + ".stft.weight_forward_real"].reshape(-1, 64).contiguous(),
            stft_f_im=Bf[bg + ".stft.weight_forward_imag"].reshape(-1, 64).contiguous())

    # ---------------------------------------------------------------- parameters
    def w(self, prefix):
        k0 = prefix + ".parametrizations.weight.original0"
        if k0 in self.P:  # weight_norm: g * v/||v||, differentiable w.r.t. both
            return torch._weight_norm(self.P[prefix + ".parametrizations.weight.original1"], self.P[k0], 0)
        return self.P[prefix + ".weight"]

    def b(self, prefix):
        return self.P.get(prefix + ".bias")

    def lin(self, prefix):
        return self.P[prefix + ".weight"].unsqueeze(-1), self.P.get(prefix + ".bias")

    def gb(self, h, name, C):
        o = self.fc_off[name]
        return h[:, o:o + 2 * C]

    # ---------------------------------------------------------------- stages
    def text_encoder(self, texts, lengths):
        P, t = self.P, "text_encoder"
        te = self.mc.text_encoder
        Cc, H = te.hidden_dim, te.heads
        B, Tn = texts.shape
        dev = texts.device
        x0 = T.EmbedFn.apply(P[t + ".emb.weight"], texts, lengths, math.sqrt(Cc))
        mask = torch.empty((B, Tn), device=dev, dtype=torch.float32)
        L.call("sty_sequence_mask_fwd", lengths.data_ptr(), mask.data_ptr(), B, Tn, L.stream_ptr())
        h = x0
        pd = float(te.dropout)  # Encoder / MHA / FFN p_dropout (text_encoder.py:427); the prenet's is 0.5 (:418)
        for i in range(3):
            y = T.conv(h, self.w(f"{t}.prenet.conv_layers.{i}"), self.b(f"{t}.prenet.conv_layers.{i}"),
                       in_mask=mask)
            h = T.chan_ln(y, gamma=P[f"{t}.prenet.norm_layers.{i}.gamma"],
                          beta=P[f"{t}.prenet.norm_layers.{i}.beta"], eps=1e-4, act=ACT_RELU)
            h = self.drop(h, 1 + i, 0.5)
        x = T.conv(h, self.w(t + ".prenet.proj"), self.b(t + ".prenet.proj"), out_mask=mask, res=x0)
        D = Cc // H
        rope = self.rope(Tn, D, dev)
        e = t + ".encoder"
        for i in range(te.layers):
            a = f"{e}.attn_layers.{i}"
            wqkv = torch.cat([P[a + ".conv_q.weight"], P[a + ".conv_k.weight"], P[a + ".conv_v.weight"]], 0)
            bqkv = torch.cat([P[a + ".conv_q.bias"], P[a + ".conv_k.bias"], P[a + ".conv_v.bias"]], 0)
            qkv = T.conv(x, wqkv, bqkv, in_mask=mask)
            s0 = 16 + 4 * i
            att = T.AttentionFn.apply(qkv, H, D, lengths, rope, 1.0 / math.sqrt(D),
                                      (self.active, s0, pd) if self.active is not None else None)
            y = self.drop(T.conv(att, self.w(a + ".conv_o"), self.b(a + ".conv_o")), s0 + 1, pd)
            x1 = T.chan_ln(y, res=x, gamma=P[f"{e}.norm_layers_1.{i}.gamma"],
                           beta=P[f"{e}.norm_layers_1.{i}.beta"], eps=1e-4)
            f = f"{e}.ffn_layers.{i}"
            hh = T.conv(x1, self.w(f + ".conv_1"), self.b(f + ".conv_1"), in_mask=mask)
            hh = self.drop(hh, s0 + 2, pd)  # relu -> drop == drop -> relu (the keep factor is >= 0)
            y2 = T.conv(hh, self.w(f + ".conv_2"), self.b(f + ".conv_2"), in_act=ACT_RELU, in_mask=mask,
                        out_mask=mask)
            y2 = self.drop(y2, s0 + 3, pd)
            x = T.chan_ln(y2, res=x1, gamma=P[f"{e}.norm_layers_2.{i}.gamma"],
                          beta=P[f"{e}.norm_layers_2.{i}.beta"], eps=1e-4, mask=mask)
        return T.conv(x, self.w(t + ".proj_m"), self.b(t + ".proj_m"), out_mask=mask), mask

    def style_fc(self, style):
        fc_w = torch.cat([self.P[n + ".fc.weight"] for n in self.fc_names], 0)
        fc_b = torch.cat([self.P[n + ".fc.bias"] for n in self.fc_names], 0)
        return T.LinearRowsFn.apply(style, fc_w, fc_b)

    def decoder_block(self, p, x, h, drop=None):
        """AdaptiveDecoderBlock (ada_norm.py:143-192); fp32 FMA convs (F0 channel in Hz).
        drop = (site, p): the two nn.Dropout in front of conv1 / conv2 (:183,186) — AdaIN + LeakyReLU + mask are
        then materialised by one kernel instead of living in the conv prologue."""
        ci = x.shape[1]
        w1 = self.w(p + ".conv1")
        co = w1.shape[0]
        has_sc = (p + ".conv1x1.parametrizations.weight.original0") in self.P
        short = T.conv(x, self.w(p + ".conv1x1"), None, umma=False) if has_sc else x
        if drop is not None and self.active is not None and drop[1] > 0.0:
            a1 = T.AdaInDropFn.apply(x, self.gb(h, p + ".norm1", ci), 1e-5, ACT_LEAKY02, self.active, drop[0], drop[1])
            t1 = T.conv(a1, w1, self.b(p + ".conv1"), umma=False)
            a2 = T.AdaInDropFn.apply(t1, self.gb(h, p + ".norm2", co), 1e-5, ACT_LEAKY02, self.active, drop[0] + 1,
                                     drop[1])
            return T.conv(a2, self.w(p + ".conv2"), self.b(p + ".conv2"), res=short, out_scale=INV_SQRT2,
                          res_scale=INV_SQRT2, umma=False)
        t1 = T.conv(x, w1, self.b(p + ".conv1"), gb=self.gb(h, p + ".norm1", ci), in_act=ACT_LEAKY02,
                    norm="instance", eps=1e-5, umma=False)
        return T.conv(t1, self.w(p + ".conv2"), self.b(p + ".conv2"), gb=self.gb(h, p + ".norm2", co),
                      in_act=ACT_LEAKY02, norm="instance", eps=1e-5, res=short, out_scale=INV_SQRT2,
                      res_scale=INV_SQRT2, umma=False)

    def decoder(self, mu, alignment, pitch, energy, voiced, h):
        d = "decoder"
        B, Fr = pitch.shape
        asr = T.BmmAlignFn.apply(mu, alignment)
        if self.active is not None:
            # train-only random box smoothing of F0 and N (decoder.py:53-75): always the same 31-tap kernel, the
            # drawn width lives in the device-side filter (width 0 = identity), so a captured graph follows it
            pitch = T.DwConvFn.apply(pitch.reshape(B, 1, Fr), self.smooth_w[0:1], None, 31, 15).reshape(B, Fr)
            energy = T.DwConvFn.apply(energy.reshape(B, 1, Fr), self.smooth_w[1:2], None, 31, 15).reshape(B, Fr)
        side = []
        for src, n in ((pitch, "F0_conv"), (energy, "N_conv"), (voiced, "voiced_conv")):
            side.append(T.DwConvFn.apply(src.reshape(B, 1, Fr), self.w(f"{d}.{n}"), self.b(f"{d}.{n}"), 3, 1))
        asr_res = T.conv(asr, self.w(d + ".asr_res.0"), self.b(d + ".asr_res.0"), umma=False)
        x = self.decoder_block(d + ".encode", torch.cat([asr] + side, 1), h)
        for i in range(4):
            x = self.decoder_block(f"{d}.decode.{i}", torch.cat([x, asr_res] + side, 1), h)
        return x

    def conformer(self, x, h):
        P, Bf = self.P, self.Bf
        c = "generator.amp_conformer.layers.0"
        Cc = x.shape[1]

        # No dropout here even in train(): generator.py:824-826 asks for 0.2, but Conformer.__init__ does not
        # forward attn/ff/conv_dropout to its ConformerBlocks (conformer.py:284-296), so every nn.Dropout of the
        # block has p = 0.0 in the reference (pinned by tests/golden/train_grads_dropout.npz).

        def ff(p, xin):
            n = T.chan_ln(xin, gb=self.gb(h, p + ".fn.norm", Cc), eps=1e-5)
            u = T.conv(n, *self.lin(p + ".fn.fn.net.0"))
            w3, b3 = self.lin(p + ".fn.fn.net.3")
            return T.conv(u, w3, b3, in_act=ACT_SWISH, res=xin, out_scale=0.5)

        x_ff1 = ff(c + ".ff1", x)
        n = T.chan_ln(x, gb=self.gb(h, c + ".attn.norm", Cc), eps=1e-5)
        wqkv = torch.cat([P[c + ".attn.fn.to_q.weight"], P[c + ".attn.fn.to_kv.weight"]], 0).unsqueeze(-1)
        qkv = T.conv(n, wqkv, None)
        att = T.AttentionFn.apply(qkv, 8, 64, None, None, 64 ** -0.5)
        x2 = T.conv(att, *self.lin(c + ".attn.fn.to_out"), res=x_ff1)
        n = T.chan_ln(x2, gb=self.gb(h, c + ".conv.norm", Cc), eps=1e-5)
        g = T.conv(n, P[c + ".conv.net.1.weight"], P[c + ".conv.net.1.bias"])
        gl = T.GluFn.apply(g)
        dw = T.DwConvFn.apply(gl, P[c + ".conv.net.3.conv.weight"], P[c + ".conv.net.3.conv.bias"], 31, 15)
        bn = c + ".conv.net.4"
        x3 = T.conv(dw, P[c + ".conv.net.6.weight"], P[c + ".conv.net.6.bias"], bn_w=P[bn + ".weight"],
                    bn_b=P[bn + ".bias"], in_act=ACT_SWISH, norm="batch", eps=1e-5,
                    bn_buffers=(Bf[bn + ".running_mean"], Bf[bn + ".running_var"]), res=x2,
                    bn_eval=not self.m.training)
        if self.m.training and bn + ".num_batches_tracked" in Bf:
            Bf[bn + ".num_batches_tracked"].add_(1)
        x4 = ff(c + ".ff2", x3)
        return T.chan_ln(x4, gb=self.gb(h, c + ".post_norm", Cc), eps=1e-5)

    def convnext(self, p, x, h):
        """GeneratorConvNeXtBlock (conv_next.py:80-93)"""
        P = self.P
        Cc = x.shape[1]
        d = T.DwConvFn.apply(x, P[p + ".dwconv.weight"], P[p + ".dwconv.bias"], 7, 3)
        y = T.chan_ln(d, gb=self.gb(h, p + ".norm", Cc), eps=1e-6)
        w2 = P[p + ".pwconv2.weight"]
        b2f = P[p + ".pwconv2.bias"] + w2 @ P[p + ".grn.beta"].reshape(-1)  # GRN beta folded into the bias
        return T.ConvNeXtTailFn.apply(y, x, P[p + ".pwconv1.weight"], P[p + ".pwconv1.bias"],
                                      P[p + ".snake"].reshape(-1), P[p + ".grn.gamma"].reshape(-1), w2, b2f, True)

    def gen_block(self, p, x, h):
        """AdaptiveGeneratorBlock (ada_norm.py:109-120)"""
        P = self.P
        Cc = x.shape[1]
        for i, dl in enumerate((1, 3, 5)):
            xt = T.conv(x, self.w(f"{p}.convs1.{i}"), self.b(f"{p}.convs1.{i}"),
                        gb=self.gb(h, f"{p}.adain1.{i}", Cc), alpha=P[f"{p}.alpha1.{i}"].reshape(-1), dil=dl,
                        in_act=ACT_SNAKE, norm="instance", eps=1e-5)
            x = T.conv(xt, self.w(f"{p}.convs2.{i}"), self.b(f"{p}.convs2.{i}"),
                       gb=self.gb(h, f"{p}.adain2.{i}", Cc), alpha=P[f"{p}.alpha2.{i}"].reshape(-1),
                       in_act=ACT_SNAKE, norm="instance", eps=1e-5, res=x)
        return x

    def generator(self, mel, h, pitch, voiced, noise, prior):
        P, Bf, g = self.P, self.Bf, "generator"
        mc = self.mc
        x = T.conv(mel, self.w(g + ".amp_input_conv"), self.b(g + ".amp_input_conv"))
        x = T.chan_ln(x, gamma=P[g + ".amp_norm.weight"], beta=P[g + ".amp_norm.bias"], eps=1e-6)
        x = self.conformer(x, h)
        if prior is None:
            with torch.no_grad():
                prior = self.engine.harmonic_prior(self.prior_pack(), pitch.detach(), voiced.detach(), noise)
        har_spec, har_phase = prior
        bg = g + ".basegen"
        lp = T.conv(har_spec, self.w(bg + ".amp_prior_conv"), self.b(bg + ".amp_prior_conv"))
        lp = self.gen_block(bg + ".amp_prior_block", lp, h)
        pp = T.conv(har_phase, self.w(bg + ".phase_prior_conv"), self.b(bg + ".phase_prior_conv"))
        pp = self.gen_block(bg + ".phase_prior_block", pp, h)
        for i in range(mc.generator.conv_layers - 3):
            x = self.convnext(f"{bg}.amp_convnext.{i}", x, h)
        for i, r in enumerate((3, 5, 5)):
            x = T.conv(x, self.w(f"{bg}.upconvs.{i}"), self.b(f"{bg}.upconvs.{i}"), shuffle=r)
            x = self.convnext(f"{bg}.upblocks.{i}", x, h)
        la = T.chan_ln(x, gamma=P[bg + ".amp_final_layer_norm.weight"], beta=P[bg + ".amp_final_layer_norm.bias"],
                       eps=1e-6)
        logamp = T.conv(la, self.w(bg + ".amp_output_conv"), self.b(bg + ".amp_output_conv"))
        ph = T.conv(torch.cat([x, lp, pp], 1), self.w(bg + ".phase_input_conv"), self.b(bg + ".phase_input_conv"))
        ph = T.chan_ln(ph, gamma=P[bg + ".phase_norm.weight"], beta=P[bg + ".phase_norm.bias"], eps=1e-6)
        for i in range(mc.generator.conv_layers):
            ph = self.convnext(f"{bg}.phase_convnext.{i}", ph, h)
        ph = T.chan_ln(ph, gamma=P[bg + ".phase_final_layer_norm.weight"],
                       beta=P[bg + ".phase_final_layer_norm.bias"], eps=1e-6)
        wri = torch.cat([P[bg + ".phase_output_real_conv.weight"], P[bg + ".phase_output_imag_conv.weight"]], 0)
        bri = torch.cat([P[bg + ".phase_output_real_conv.bias"], P[bg + ".phase_output_imag_conv.bias"]], 0)
        ri = T.conv(ph, wri, bri)
        b_re = Bf[bg + ".stft.weight_backward_real"].reshape(-1, 64)
        b_im = Bf[bg + ".stft.weight_backward_imag"].reshape(-1, 64)
        return T.IstftHeadFn.apply(logamp, ri, b_re.contiguous(), b_im.contiguous(), mc.hop_length // 75)

    # ---------------------------------------------------------------- forward
    def forward(self, texts, text_lengths, alignment, pitch, energy, voiced, style, denormal_pitch, *,
                source_draws=None, prior=None):
        dev = texts.device
        L.require_cuda(dev, "inputs")
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
        texts = texts.to(torch.int64).contiguous()
        lengths = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
        alignment, pitch, energy = f32(alignment), f32(pitch), f32(energy)
        voiced, style, denormal_pitch = f32(voiced), f32(style), f32(denormal_pitch)
        self.start_forward(dev)
        h = self.style_fc(style)
        mu, _ = self.text_encoder(texts, lengths)
        mel = self.decoder(mu, alignment, pitch, energy, voiced, h)
        noise = None if source_draws is None else f32(source_draws["noise"])
        return self.generator(mel, h, denormal_pitch, voiced, noise, prior)


def _prep(dev, texts, text_lengths, *floats):
    L.require_cuda(dev, "inputs")
    f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
    return (texts.to(torch.int64).contiguous(), text_lengths.to(device=dev, dtype=torch.int64).contiguous(),
            *[f32(t) for t in floats])


class DurationTrainGraph(TrainGraph):
    """Differentiable DurationPredictor.forward (duration_predictor.py:58-87).  train() mode: text-encoder
    dropout, SDPA dropout 0.5 of the cross attention (:40), DropPath 0.5 of every ConvNeXt block
    (conv_next.py:130) and Dropout1d(last_dropout) after it (:79) — sites 80, 84+2i, 85+2i."""

    def forward(self, texts, text_lengths, style):
        P = self.P
        texts, lengths, style = _prep(texts.device, texts, text_lengths, style)
        self.start_forward(texts.device)
        on = self.active is not None
        dcfg = self.mc.duration_predictor
        h = self.style_fc(style)
        enc, mask = self.text_encoder(texts, lengths)
        B, Cc, Tn = enc.shape
        qn = T.chan_ln(enc, gb=self.gb(h, "query_norm", Cc), eps=1e-5)
        kn = T.chan_ln(enc, gb=self.gb(h, "key_norm", Cc), eps=1e-5)
        a = "cross_attention"
        q = T.conv(qn, self.w(a + ".conv_q"), self.b(a + ".conv_q"))
        wkv = torch.cat([P[a + ".conv_k.weight"], P[a + ".conv_v.weight"]], 0)
        bkv = torch.cat([P[a + ".conv_k.bias"], P[a + ".conv_v.bias"]], 0)
        kv = T.conv(kn, wkv, bkv)
        H = 8
        D = Cc // H
        att = T.AttentionFn.apply(torch.cat([q, kv], 1), H, D, lengths, self.rope(Tn, D, enc.device),
                                  1.0 / math.sqrt(D), (self.active, 80, 0.5) if on else None)
        att = T.conv(att, self.w(a + ".conv_o"), self.b(a + ".conv_o"))
        dw = T.DwConvFn.apply(att, self.w("cross_post.0"), self.b("cross_post.0"), 5, 2)
        pros = T.conv(dw, self.w("cross_post.2"), self.b("cross_post.2"), in_act=ACT_SWISH, res=enc,
                      out_scale=INV_SQRT2, res_scale=INV_SQRT2)
        m3 = mask.unsqueeze(1)
        for i in range(self.mc.duration_predictor.n_layer):  # AdaptiveConvNeXtBlock, then * mask
            p = f"conv_next.{i}"
            d = T.DwConvFn.apply(pros, P[p + ".dwconv.weight"], P[p + ".dwconv.bias"], 7, 3)
            y = T.chan_ln(d, gb=self.gb(h, p + ".norm", Cc), eps=1e-6)
            w2 = P[p + ".pwconv2.weight"]
            b2f = P[p + ".pwconv2.bias"] + w2 @ P[p + ".grn.beta"].reshape(-1)
            if on:  # (residual + DropPath(branch)) * mask, then Dropout1d over (b, c)
                br = T.ConvNeXtTailFn.apply(y, torch.zeros_like(pros), P[p + ".pwconv1.weight"],
                                            P[p + ".pwconv1.bias"], None, P[p + ".grn.gamma"].reshape(-1), w2, b2f,
                                            True, L.ACT_GELU, mask)
                pros = self.drop(br, 84 + 2 * i, 0.5, group=Cc * Tn, res=pros * m3)
                pros = self.drop(pros, 85 + 2 * i, float(dcfg.last_dropout), group=Tn)
                continue
            pros = T.ConvNeXtTailFn.apply(y, pros * m3, P[p + ".pwconv1.weight"], P[p + ".pwconv1.bias"], None,
                                          P[p + ".grn.gamma"].reshape(-1), w2, b2f, True, L.ACT_GELU, mask)
        logits = T.conv(pros, P["duration_proj.linear_layer.weight"].unsqueeze(-1),
                        P["duration_proj.linear_layer.bias"])  # (B,NC,T)
        # monotone head on the (B,NC,T) class scores (16 x T values per utterance): duration_predictor.py:82-86
        d = torch.cat([logits[:, :1], logits[:, 1:].abs()], 1)
        d = -torch.cumsum(d, 1).abs()
        return d.transpose(1, 2) * mask.unsqueeze(2)


class PitchEnergyTrainGraph(TrainGraph):
    """Differentiable PitchEnergyPredictor.forward (pitch_energy_predictor.py:62-82, prosody_encoder.py:63-81)."""

    def forward(self, texts, text_lengths, alignment, style):
        P = self.P
        texts, lengths, alignment, style = _prep(texts.device, texts, text_lengths, alignment, style)
        self.start_forward(texts.device)
        on = self.active is not None
        pp = 0.2  # ProsodyEncoder(dropout=0.2) pitch_energy_predictor.py:27; sites 80+4i .. 83+4i
        pb = float(self.mc.pitch_energy_predictor.dropout)  # AdaptiveDecoderBlock dropout_p (:22,33-56); sites 96..
        h = self.style_fc(style)
        enc, mask = self.text_encoder(texts, lengths)
        B, dm, Tn = enc.shape
        sdim = style.shape[1]
        Ch = dm + sdim
        st = style.unsqueeze(2).expand(B, sdim, Tn)
        m3 = mask.unsqueeze(1)
        H = 2
        D = Ch // H
        rope = self.rope(Tn, D, enc.device)
        pe = "prosody_encoder"
        x = torch.cat([enc, st], 1)
        for i in range(3):
            a = f"{pe}.attn_layers.{i}"
            wqkv = torch.cat([P[a + ".conv_q.weight"], P[a + ".conv_k.weight"], P[a + ".conv_v.weight"]], 0)
            bqkv = torch.cat([P[a + ".conv_q.bias"], P[a + ".conv_k.bias"], P[a + ".conv_v.bias"]], 0)
            qkv = T.conv(x, wqkv, bqkv, in_mask=mask)
            s0 = 80 + 4 * i
            att = T.AttentionGenericFn.apply(qkv[:, :Ch], qkv[:, Ch:2 * Ch], qkv[:, 2 * Ch:], H, D, lengths, rope,
                                             1.0 / math.sqrt(D), (self.active, s0, pp) if on else None)
            y = self.drop(T.conv(att, self.w(a + ".conv_o"), self.b(a + ".conv_o")), s0 + 1, pp)
            x1 = T.chan_ln(y, res=x * m3, gb=self.gb(h, f"{pe}.norm_layers_1.{i}", Ch), eps=1e-5)
            f = f"{pe}.ffn_layers.{i}"
            hh = T.conv(x1, self.w(f + ".conv_1"), self.b(f + ".conv_1"), in_mask=mask)
            hh = self.drop(hh, s0 + 2, pp)  # commutes with the ReLU of conv_2's prologue
            y2 = T.conv(hh, self.w(f + ".conv_2"), self.b(f + ".conv_2"), in_act=ACT_RELU, in_mask=mask, out_mask=mask)
            y2 = self.drop(y2, s0 + 3, pp)
            x2 = T.chan_ln(y2, res=x1, gb=self.gb(h, f"{pe}.norm_layers_2.{i}", Ch), eps=1e-5)
            xp = T.conv(x2, self.w(f"{pe}.proj_layers.{i}"), self.b(f"{pe}.proj_layers.{i}"))
            x = torch.cat([xp, st], 1)
        pros = x * m3
        xa = T.BmmAlignFn.apply(pros, alignment)
        outs = []
        for ti, (tower, proj) in enumerate((("F0", "F0_proj"), ("N", "N_proj"))):
            z = xa
            for i in range(4):
                z = self.decoder_block(f"{tower}.{i}", z, h, drop=(96 + 8 * ti + 2 * i, pb))
            outs.append(T.conv(z, self.w(proj), self.b(proj), umma=False).squeeze(1))
        return outs[0], outs[1]
