"""Forward engine of ``speech_predictor``: packs the shell's parameters into
kernel-friendly device buffers and drives the sm_100a kernels through the
C ABI.  PyTorch is used for device memory and the stream only.

Stage order follows the reference forward graph
(speech_predictor.py:47-73 -> text_encoder.py:434-463 -> decoder.py:77-90 ->
generator.py:884-901 -> generator.py:710-799), see DESIGN.md for the kernel
each stage maps to and where normalisations are fused.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional

import torch

from . import _lib as L
from ._lib import ACT_LEAKY02, ACT_NONE, ACT_RELU, ACT_SNAKE, ACT_SWISH, ConvArgs

INV_SQRT2 = 1.0 / math.sqrt(2.0)
# tensor-core (tcgen05, bf16x3) path for eligible convs; STYLISH_B200_UMMA=0 forces fp32 FMA
USE_UMMA = os.environ.get("STYLISH_B200_UMMA", "1") != "0"
UMMA_MIN_T = 64
WIDE_MIN_ELEMS = 65536  # conv1d(wide=True): rows x steps from which short rows still go to the tensor cores
# InstanceNorm statistics of the S-rate AdaINs accumulated in the producing conv's epilogue (one pass less)
FUSE_STATS = os.environ.get("STYLISH_B200_FUSE_STATS", "1") != "0"
# output-rate ConvNeXt blocks as one fused two-pass kernel that never stores the 4C-wide intermediate
FUSE_CONVNEXT = os.environ.get("STYLISH_B200_FUSE_CONVNEXT", "1") != "0"


def empty_bct(B: int, C: int, T: int, device) -> torch.Tensor:
    """(B,C,T) fp32 activation whose rows start on 16-byte boundaries (row pitch = T rounded up to 4 floats):
    what the TMA loads of the tensor-core convs need (`cp.async.bulk.tensor`: base and strides multiples of
    16 bytes).  T = 75 * frames is odd for an odd frame count (60 225 at the BASELINE shapes)."""
    Tp = (T + 3) & ~3
    return torch.empty((B, C, Tp), device=device, dtype=torch.float32)[:, :, :T]


def _rows_aligned(t: torch.Tensor) -> bool:
    """TMA-compatible (B,C,T) view: 16-byte aligned base, strides multiples of 4 floats, row pitch >= T rounded up to 4"""
    return (t.stride(2) == 1 and t.data_ptr() % 16 == 0 and t.stride(0) % 4 == 0 and t.stride(1) % 4 == 0
            and t.stride(1) >= ((t.shape[2] + 3) & ~3))


def split_bf16(w_oik: torch.Tensor) -> torch.Tensor:
    """(CO,CI,K) fp32 -> bf16 (hi, lo) pair in the UMMA K-major layout [K][2][CI/8][CO][8]."""
    co, ci, k = w_oik.shape
    hi = w_oik.to(torch.bfloat16)
    lo = (w_oik - hi.float()).to(torch.bfloat16)

    def lay(t):
        return t.permute(2, 1, 0).reshape(k, ci // 8, 8, co).permute(0, 1, 3, 2)

    return torch.stack([lay(hi), lay(lo)], dim=1).contiguous()


# --------------------------------------------------------------------------
# thin op wrappers (tensor -> raw pointers)
# --------------------------------------------------------------------------
class ConvW:
    """A conv/linear weight pre-packed as (CI, K, CO) plus its bias."""

    __slots__ = ("w", "bias", "CI", "K", "CO", "split")

    def __init__(self, w_oik: torch.Tensor, bias: Optional[torch.Tensor]):
        co, ci, k = w_oik.shape
        self.w = w_oik.permute(1, 2, 0).contiguous()
        self.bias = None if bias is None else bias.contiguous()
        self.CI, self.K, self.CO = ci, k, co
        ok = ci % 16 == 0 and co % 16 == 0 and co >= 16
        self.split = split_bf16(w_oik) if ok else None


def conv1d(x, cw: ConvW, *, dil=1, out=None, res=None, in_scale=None, in_shift=None,
           in_alpha=None, in_act=ACT_NONE, in_mask=None, out_mask=None, out_act=ACT_NONE,
           out_alpha=None, out_sumsq=None, out_sum=None, shuffle=0, out_scale=1.0, res_scale=1.0, umma=True,
           dwln=None, wide=False):
    B, CI, T = x.shape
    assert CI == cw.CI, (CI, cw.CI)
    x_bs, x_cs = L._bct(x, "x")
    s = shuffle if shuffle > 1 else 1
    if out is None:  # same row pitch policy as the input: pitch-padded in, pitch-padded out
        alloc = empty_bct if x_cs != T else (lambda *a: torch.empty(a[:3], device=a[3], dtype=torch.float32))
        out = alloc(B, cw.CO // s, T * s, x.device)
    assert out.shape == (B, cw.CO // s, T * s), (out.shape, (B, cw.CO // s, T * s))
    y_bs, y_cs = L._bct(out, "out")
    a = ConvArgs()
    a.x, a.x_bs, a.x_cs = x.data_ptr(), x_bs, x_cs
    a.w, a.w_bs = cw.w.data_ptr(), 0
    a.bias = L.ptr(cw.bias)
    a.y, a.y_bs, a.y_cs = out.data_ptr(), y_bs, y_cs
    if res is not None:
        assert res.shape == out.shape
        r_bs, r_cs = L._bct(res, "res")
        a.res, a.r_bs, a.r_cs = res.data_ptr(), r_bs, r_cs
    a.in_scale, a.in_shift, a.in_alpha = L.ptr(in_scale), L.ptr(in_shift), L.ptr(in_alpha)
    a.in_mask, a.out_mask, a.out_alpha = L.ptr(in_mask), L.ptr(out_mask), L.ptr(out_alpha)
    a.out_sumsq = L.ptr(out_sumsq)
    a.out_sum = L.ptr(out_sum)
    a.B, a.CI, a.CO, a.T, a.K, a.dil = B, CI, cw.CO, T, cw.K, dil
    a.pad = (cw.K - 1) * dil // 2
    a.in_act, a.out_act, a.shuffle = in_act, out_act, shuffle
    a.out_scale, a.res_scale = out_scale, res_scale
    # `wide`: thousands of short rows (row-stacked images) — the tensor-core kernel wins below UMMA_MIN_T too
    if umma and USE_UMMA and cw.split is not None and (T >= UMMA_MIN_T or (wide and T >= 16 and B * T >= WIDE_MIN_ELEMS)):
        a.w_split = cw.split.data_ptr()
    if dwln is not None:  # fused ConvNeXt front (dw_w, dw_b, gamma|beta rows, row stride, eps)
        assert a.w_split, "fused ConvNeXt front needs the tensor-core path"
        a.dw_w, a.dw_b, a.dw_gb = dwln[0].data_ptr(), dwln[1].data_ptr(), dwln[2].data_ptr()
        a.dw_gb_bs, a.dw_eps = dwln[3], dwln[4]
    L.call("sty_conv1d_fwd", C.byref(a), L.stream_ptr())
    return out


def chan_layernorm(x, gamma, beta, *, eps, res=None, g_bs=0, plus_one=False, out=None, mask=None,
                   act=ACT_NONE):
    B, Cc, T = x.shape
    x_bs, x_cs = L._bct(x, "x")
    if res is not None:
        assert res.stride() == x.stride()
    if out is None:
        out = empty_bct(B, Cc, T, x.device) if x_cs != T else torch.empty((B, Cc, T), device=x.device,
                                                                          dtype=torch.float32)
    y_bs, y_cs = L._bct(out, "out")
    if x_cs == T and y_cs == T:
        L.call("sty_chan_layernorm_fwd", x.data_ptr(), L.ptr(res), x_bs, gamma.data_ptr(),
               beta.data_ptr(), g_bs, int(plus_one), out.data_ptr(), y_bs, L.ptr(mask), B, Cc, T,
               eps, act, L.stream_ptr())
    else:  # pitch-padded rows
        L.call("sty_chan_layernorm_pitched_fwd", x.data_ptr(), L.ptr(res), x_bs, x_cs, gamma.data_ptr(),
               beta.data_ptr(), g_bs, int(plus_one), out.data_ptr(), y_bs, y_cs, L.ptr(mask), B, Cc, T,
               eps, act, L.stream_ptr())
    return out


def instnorm_affine(x, gb, gb_bs, eps=1e-5):
    B, Cc, T = x.shape
    x_bs, x_cs = L._bct(x, "x")
    scale = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    shift = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    L.call("sty_instnorm_affine_fwd", x.data_ptr(), x_bs, x_cs, gb.data_ptr(), gb_bs,
           scale.data_ptr(), shift.data_ptr(), B, Cc, T, eps, L.stream_ptr())
    return scale, shift


def moments_affine(mom, gb, gb_bs, T, eps=1e-5):
    """AdaIN affine from the (2, B, C) moments (sum, sum of squares) a producer conv accumulated."""
    _, B, Cc = mom.shape
    scale = torch.empty((B, Cc), device=mom.device, dtype=torch.float32)
    shift = torch.empty((B, Cc), device=mom.device, dtype=torch.float32)
    L.call("sty_moments_affine_fwd", mom[0].data_ptr(), mom[1].data_ptr(), gb.data_ptr(), gb_bs,
           scale.data_ptr(), shift.data_ptr(), B, Cc, T, eps, L.stream_ptr())
    return scale, shift


def dwconv1d(x, w, bias, *, K, pad_left, out, post_scale=None, post_shift=None, act=ACT_NONE):
    B, Cc, T = x.shape
    x_bs, x_cs = L._bct(x, "x")
    y_bs, y_cs = L._bct(out, "out")
    L.call("sty_dwconv1d_fwd", x.data_ptr(), x_bs, x_cs, w.data_ptr(), L.ptr(bias),
           L.ptr(post_scale), L.ptr(post_shift), out.data_ptr(), y_bs, y_cs, B, Cc, T, K, pad_left,
           act, L.stream_ptr())
    return out


# 64-wide unmasked attention through the pre-split / bulk-copy kernel (csrc/attention64.cu); 0 = first-generation kernel
ATTENTION64 = os.environ.get("STYLISH_B200_ATTENTION64", "1") != "0"
# harmonic-prior branch of the generator on a side stream, concurrent with the text encoder / decoder / conformer
OVERLAP_PRIOR = os.environ.get("STYLISH_B200_OVERLAP_PRIOR", "1") != "0"


def attention64(q_ptr, k_ptr, v_ptr, qkv_bs, out, B, H, T, scale, lse=None):
    """q, k, v: device pointers of (B, H*64, T) fp32 tensors with batch stride qkv_bs -> out (B, H*64, T)"""
    ws = torch.empty(int(L.load().sty_attention64_workspace_bytes(B, H, T)), device=out.device, dtype=torch.uint8)
    L.call("sty_attention64_fwd", q_ptr, k_ptr, v_ptr, qkv_bs, out.data_ptr(), out.stride(0), B, H, T, scale,
           L.ptr(lse), ws.data_ptr(), L.stream_ptr())
    return out


def attention(qkv, n_q, n_k, n_v, *, H, D, lengths=None, rope=None, scale):
    """qkv: (B, n_q+n_k+n_v, T) fused projection output."""
    B, _, T = qkv.shape
    bs, cs = L._bct(qkv, "qkv")
    assert cs == T and n_q == n_k == n_v == H * D
    out = torch.empty((B, H * D, T), device=qkv.device, dtype=torch.float32)
    q = qkv.data_ptr()
    k = q + 4 * n_q * T
    v = k + 4 * n_k * T
    if D == 64 and rope is None and lengths is None and T >= 64 and ATTENTION64:
        attention64(q, k, v, bs, out, B, H, T, scale)
        return out
    rc, rs, d_rot = (None, None, 0) if rope is None else (rope[0].data_ptr(), rope[1].data_ptr(),
                                                          rope[2])
    L.call("sty_attention_fwd", q, k, v, bs, out.data_ptr(), out.stride(0), L.ptr(lengths), rc, rs,
           d_rot, B, H, D, T, scale, L.stream_ptr())
    return out


class PackBase:
    """Device-resident, kernel-layout copy of a module's parameters: helpers, the packed
    style FCs and the text encoder every hot-path module owns (`text_encoder.*`)."""

    def __init__(self, module, device):
        self.sd = sd = {k: v.detach().to(device=device, dtype=torch.float32) if v.is_floating_point()
                        else v.detach().to(device) for k, v in module.state_dict().items()}
        self.device = device
        mc = module.model_config
        self.mc = mc
        te = mc.text_encoder
        self.n_layers, self.n_heads, self.hidden = te.layers, te.heads, te.hidden_dim
        cv = self.cv

        # ---- all style FCs packed row-wise into one matrix --------------------
        fc_names = sorted(k[:-len(".fc.weight")] for k in sd if k.endswith(".fc.weight"))
        self.fc_off: Dict[str, int] = {}
        rows, biases, off = [], [], 0
        for n in fc_names:
            w = sd[n + ".fc.weight"]
            self.fc_off[n] = off
            rows.append(w)
            biases.append(sd[n + ".fc.bias"])
            off += w.shape[0]
        self.fc_w = torch.cat(rows, 0).contiguous()
        self.fc_b = torch.cat(biases, 0).contiguous()
        self.fc_rows = off

        # ---- text encoder -------------------------------------------------------
        t = "text_encoder"
        self.emb = sd[t + ".emb.weight"].contiguous()
        self.prenet = [(cv(f"{t}.prenet.conv_layers.{i}"),
                        sd[f"{t}.prenet.norm_layers.{i}.gamma"].contiguous(),
                        sd[f"{t}.prenet.norm_layers.{i}.beta"].contiguous()) for i in range(3)]
        self.prenet_proj = cv(t + ".prenet.proj")
        self.enc = []
        e = t + ".encoder"
        for i in range(self.n_layers):
            a = f"{e}.attn_layers.{i}"
            self.enc.append(dict(
                qkv=self.qkv(a), o=cv(a + ".conv_o"),
                n1=(sd[f"{e}.norm_layers_1.{i}.gamma"].contiguous(),
                    sd[f"{e}.norm_layers_1.{i}.beta"].contiguous()),
                f1=cv(f"{e}.ffn_layers.{i}.conv_1"), f2=cv(f"{e}.ffn_layers.{i}.conv_2"),
                n2=(sd[f"{e}.norm_layers_2.{i}.gamma"].contiguous(),
                    sd[f"{e}.norm_layers_2.{i}.beta"].contiguous())))
        self.proj_m = cv(t + ".proj_m")
        self.rope_cache: Dict[tuple, tuple] = {}

    # weight helpers ----------------------------------------------------------
    def wn(self, prefix):
        sd = self.sd
        k0 = prefix + ".parametrizations.weight.original0"
        if k0 in sd:
            return torch._weight_norm(sd[prefix + ".parametrizations.weight.original1"], sd[k0], 0)
        return sd[prefix + ".weight"]

    def cv(self, prefix):
        return ConvW(self.wn(prefix), self.sd.get(prefix + ".bias"))

    def lin(self, prefix, bias=True):
        return ConvW(self.sd[prefix + ".weight"].unsqueeze(-1),
                     self.sd.get(prefix + ".bias") if bias else None)

    def qkv(self, a):
        """fused q|k|v projection of a reference MultiHeadAttention (text_encoder.py:195-203)"""
        sd = self.sd
        w = torch.cat([sd[a + ".conv_q.weight"], sd[a + ".conv_k.weight"], sd[a + ".conv_v.weight"]], 0)
        b = torch.cat([sd[a + ".conv_q.bias"], sd[a + ".conv_k.bias"], sd[a + ".conv_v.bias"]], 0)
        return ConvW(w, b)

    def dblock(self, p):
        """AdaptiveDecoderBlock (ada_norm.py:143-192)"""
        has_sc = (p + ".conv1x1.parametrizations.weight.original0") in self.sd
        return dict(c1=self.cv(p + ".conv1"), c2=self.cv(p + ".conv2"),
                    sc=self.cv(p + ".conv1x1") if has_sc else None, n1=p + ".norm1", n2=p + ".norm2")

    def rope(self, T: int, d_head: int = 0):
        d = d_head or (self.hidden // self.n_heads)
        key = (T, d)
        if key not in self.rope_cache:
            d_rot = int(d * 0.5)
            c = torch.empty((T, d_rot // 2), device=self.device, dtype=torch.float32)
            s = torch.empty_like(c)
            L.call("sty_rope_table", c.data_ptr(), s.data_ptr(), T, d_rot, 10000.0, L.stream_ptr())
            self.rope_cache[key] = (c, s, d_rot)
        return self.rope_cache[key]


class Packed(PackBase):
    """speech_predictor: text encoder + decoder + generator."""

    def __init__(self, module, device):
        super().__init__(module, device)
        sd, mc = self.sd, self.mc
        wn, cv, lin, dblock = self.wn, self.cv, self.lin, self.dblock

        # ---- decoder --------------------------------------------------------------
        d = "decoder"

        self.dec_encode = dblock(d + ".encode")
        self.dec_decode = [dblock(f"{d}.decode.{i}") for i in range(4)]
        self.dec_side = [(wn(f"{d}.{n}").reshape(1, 3).contiguous(), sd[f"{d}.{n}.bias"].contiguous())
                         for n in ("F0_conv", "N_conv", "voiced_conv")]
        self.asr_res = cv(d + ".asr_res.0")

        # ---- generator front + conformer ---------------------------------------------
        g = "generator"
        self.amp_input = cv(g + ".amp_input_conv")
        self.amp_norm = (sd[g + ".amp_norm.weight"].contiguous(), sd[g + ".amp_norm.bias"].contiguous())
        c = g + ".amp_conformer.layers.0"

        def ff(p):
            return dict(norm=p + ".fn.norm", w0=lin(p + ".fn.fn.net.0"), w3=lin(p + ".fn.fn.net.3"))

        bn = c + ".conv.net.4"
        bn_scale = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + 1e-5)
        bn_shift = sd[bn + ".bias"] - sd[bn + ".running_mean"] * bn_scale
        wqkv = torch.cat([sd[c + ".attn.fn.to_q.weight"], sd[c + ".attn.fn.to_kv.weight"]], 0)
        self.conf = dict(
            ff1=ff(c + ".ff1"), ff2=ff(c + ".ff2"),
            attn_norm=c + ".attn.norm", qkv=ConvW(wqkv.unsqueeze(-1), None),
            out=lin(c + ".attn.fn.to_out"),
            conv_norm=c + ".conv.norm", pw1=cv(c + ".conv.net.1"),
            dw_w=sd[c + ".conv.net.3.conv.weight"].reshape(-1, 31).contiguous(),
            dw_b=sd[c + ".conv.net.3.conv.bias"].contiguous(),
            bn_scale=bn_scale.contiguous(), bn_shift=bn_shift.contiguous(),
            pw2=cv(c + ".conv.net.6"), post_norm=c + ".post_norm")

        # ---- basegen ----------------------------------------------------------------------
        bg = g + ".basegen"

        def cnx(p):
            w2 = sd[p + ".pwconv2.weight"]
            b2 = sd[p + ".pwconv2.bias"] + w2 @ sd[p + ".grn.beta"].reshape(-1)
            return dict(dw_w=sd[p + ".dwconv.weight"].reshape(-1, 7).contiguous(),
                        dw_b=sd[p + ".dwconv.bias"].contiguous(), norm=p + ".norm",
                        pw1=lin(p + ".pwconv1"), snake=sd[p + ".snake"].reshape(-1).contiguous(),
                        grn_gamma=sd[p + ".grn.gamma"].reshape(-1).contiguous(),
                        pw2=ConvW(w2.unsqueeze(-1), b2))

        def gblock(p):
            return [dict(c1=cv(f"{p}.convs1.{i}"), c2=cv(f"{p}.convs2.{i}"),
                         n1=f"{p}.adain1.{i}", n2=f"{p}.adain2.{i}",
                         a1=sd[f"{p}.alpha1.{i}"].reshape(-1).contiguous(),
                         a2=sd[f"{p}.alpha2.{i}"].reshape(-1).contiguous(), dil=dl)
                    for i, dl in enumerate((1, 3, 5))]

        self.amp_convnext = [cnx(f"{bg}.amp_convnext.{i}")
                             for i in range(mc.generator.conv_layers - 3)]
        self.upconvs = [cv(f"{bg}.upconvs.{i}") for i in range(3)]
        self.upblocks = [cnx(f"{bg}.upblocks.{i}") for i in range(3)]
        self.rates = (3, 5, 5)
        self.src_w = sd[bg + ".m_source.l_linear.weight"].reshape(-1).contiguous()
        self.src_b = sd[bg + ".m_source.l_linear.bias"].contiguous()
        self.amp_prior_conv = cv(bg + ".amp_prior_conv")
        self.phase_prior_conv = cv(bg + ".phase_prior_conv")
        self.amp_prior_block = gblock(bg + ".amp_prior_block")
        self.phase_prior_block = gblock(bg + ".phase_prior_block")
        self.phase_input = cv(bg + ".phase_input_conv")
        self.amp_output = cv(bg + ".amp_output_conv")
        wri = torch.cat([sd[bg + ".phase_output_real_conv.weight"],
                         sd[bg + ".phase_output_imag_conv.weight"]], 0)
        bri = torch.cat([sd[bg + ".phase_output_real_conv.bias"],
                         sd[bg + ".phase_output_imag_conv.bias"]], 0)
        self.phase_out_ri = ConvW(wri, bri)
        self.phase_norm = (sd[bg + ".phase_norm.weight"].contiguous(),
                           sd[bg + ".phase_norm.bias"].contiguous())
        self.phase_convnext = [cnx(f"{bg}.phase_convnext.{i}")
                               for i in range(mc.generator.conv_layers)]
        self.amp_final_ln = (sd[bg + ".amp_final_layer_norm.weight"].contiguous(),
                             sd[bg + ".amp_final_layer_norm.bias"].contiguous())
        self.phase_final_ln = (sd[bg + ".phase_final_layer_norm.weight"].contiguous(),
                               sd[bg + ".phase_final_layer_norm.bias"].contiguous())
        self.stft_f_re = sd[bg + ".stft.weight_forward_real"].reshape(-1, 64).contiguous()
        self.stft_f_im = sd[bg + ".stft.weight_forward_imag"].reshape(-1, 64).contiguous()
        self.stft_b_re = sd[bg + ".stft.weight_backward_real"].reshape(-1, 64).contiguous()
        self.stft_b_im = sd[bg + ".stft.weight_backward_imag"].reshape(-1, 64).contiguous()
        self.hidden_s = mc.n_fft // 2 // 8  # 32 channels at the STFT-frame rate


class SpeechEngine:
    def __init__(self, module):
        L.load()  # fail loudly if the CUDA library is missing
        self.module = module
        self._packed: Optional[Packed] = None
        self._packed_key = None
        self._side_streams = {}

    # parameters are repacked when any of them changed (optimizer step, load_state_dict)
    def packed(self, device) -> Packed:
        key = (str(device), L.param_epoch) + tuple(p._version for p in self.module.parameters()) + \
            tuple(b._version for b in self.module.buffers())
        if self._packed is None or key != self._packed_key:
            self._packed = Packed(self.module, device)
            self._packed_key = key
        return self._packed

    # ------------------------------------------------------------------ stages
    @staticmethod
    def _gb(P: Packed, h, name):
        """pointer view of the (gamma|beta) rows of style FC `name` inside h (B, J)."""
        return h[:, P.fc_off[name]:]

    def text_encoder(self, P, texts, lengths, taps=None, mu_out=None):
        B, T = texts.shape
        Cc = P.hidden
        dev = texts.device
        x0 = torch.empty((B, Cc, T), device=dev, dtype=torch.float32)
        L.call("sty_embed_fwd", texts.data_ptr(), lengths.data_ptr(), P.emb.data_ptr(),
               x0.data_ptr(), B, T, Cc, P.emb.shape[0], math.sqrt(Cc), L.stream_ptr())
        mask = torch.empty((B, T), device=dev, dtype=torch.float32)
        L.call("sty_sequence_mask_fwd", lengths.data_ptr(), mask.data_ptr(), B, T, L.stream_ptr())
        h = x0
        for cw, gm, bt in P.prenet:
            y = conv1d(h, cw, in_mask=mask)
            h = chan_layernorm(y, gm, bt, eps=1e-4, act=ACT_RELU)
        x = conv1d(h, P.prenet_proj, out_mask=mask, res=x0)  # x0 is already masked
        if taps is not None:
            taps["prenet"] = x.clone()
        rope = P.rope(T)
        H = P.n_heads
        D = Cc // H
        for i, ly in enumerate(P.enc):
            qkv = conv1d(x, ly["qkv"], in_mask=mask)
            att = attention(qkv, Cc, Cc, Cc, H=H, D=D, lengths=lengths, rope=rope,
                            scale=1.0 / math.sqrt(D))
            y = conv1d(att, ly["o"])
            x1 = chan_layernorm(y, *ly["n1"], eps=1e-4, res=x)
            hh = conv1d(x1, ly["f1"], in_mask=mask, out_act=ACT_RELU)
            y2 = conv1d(hh, ly["f2"], in_mask=mask, out_mask=mask)
            x = chan_layernorm(y2, *ly["n2"], eps=1e-4, res=x1, mask=mask)
        mu = conv1d(x, P.proj_m, out_mask=mask, out=mu_out)
        return mu, x, mask

    def _decoder_block(self, P, blk, x, h, out):
        J = P.fc_rows
        sc1, sh1 = instnorm_affine(x, self._gb(P, h, blk["n1"]), J)
        # fp32 FMA path: the F0 side channel is in Hz (|x| ~ 1e2 next to O(1) features), so the
        # bf16x3 operand split would cost ~1e-4 relative here (measured); the decoder is ~2 % of FLOPs
        t1 = conv1d(x, blk["c1"], in_scale=sc1, in_shift=sh1, in_act=ACT_LEAKY02, umma=False)
        sc2, sh2 = instnorm_affine(t1, self._gb(P, h, blk["n2"]), J)
        short = conv1d(x, blk["sc"], umma=False) if blk["sc"] is not None else x
        return conv1d(t1, blk["c2"], in_scale=sc2, in_shift=sh2, in_act=ACT_LEAKY02, res=short,
                      out_scale=INV_SQRT2, res_scale=INV_SQRT2, out=out, umma=False)

    def decoder(self, P: Packed, mu, alignment, pitch, energy, voiced, h, taps=None):
        B, Cm, T = mu.shape
        Fr = alignment.shape[2]
        dev = mu.device
        res_dim = P.asr_res.CO
        cat0 = torch.empty((B, Cm + 3, Fr), device=dev, dtype=torch.float32)
        cat_a = torch.empty((B, Cm + res_dim + 3, Fr), device=dev, dtype=torch.float32)
        cat_b = torch.empty_like(cat_a)
        # asr = text_encoding @ alignment, written straight into the concat buffer
        L.call("sty_bmm_fwd", mu.data_ptr(), mu.stride(0), alignment.data_ptr(),
               alignment.stride(0), cat0.data_ptr(), cat0.stride(0), B, Cm, Fr, T, L.stream_ptr())
        asr = cat0[:, :Cm]
        sides = (pitch, energy, voiced)
        for j, (src, (w, b)) in enumerate(zip(sides, P.dec_side)):
            s3 = src.reshape(B, 1, Fr)
            dwconv1d(s3, w, b, K=3, pad_left=1, out=cat0[:, Cm + j:Cm + j + 1])
            for cat in (cat_a, cat_b):
                c0 = Cm + res_dim + j
                dwconv1d(s3, w, b, K=3, pad_left=1, out=cat[:, c0:c0 + 1])
        for cat in (cat_a, cat_b):
            conv1d(asr, P.asr_res, out=cat[:, Cm:Cm + res_dim], umma=False)
        x = self._decoder_block(P, P.dec_encode, cat0, h, cat_a[:, :Cm])
        if taps is not None:
            taps["dec_encode"] = x.clone()
        cur, nxt = cat_a, cat_b
        for i, blk in enumerate(P.dec_decode):
            last = i == len(P.dec_decode) - 1
            out = torch.empty((B, Cm, Fr), device=dev, dtype=torch.float32) if last else nxt[:, :Cm]
            x = self._decoder_block(P, blk, cur, h, out)
            cur, nxt = nxt, cur
        return x

    def _ada_ln(self, P, x, h, name, eps):
        gb = self._gb(P, h, name)
        Cc = x.shape[1]
        return chan_layernorm(x, gb, gb[:, Cc:], eps=eps, g_bs=P.fc_rows, plus_one=True)

    def conformer(self, P: Packed, x, h):
        cf = P.conf
        B, Cc, T = x.shape

        def ff(blk, xin):
            n = self._ada_ln(P, xin, h, blk["norm"], 1e-5)
            u = conv1d(n, blk["w0"], out_act=ACT_SWISH)
            return conv1d(u, blk["w3"], res=xin, out_scale=0.5, res_scale=1.0)

        x_ff1 = ff(cf["ff1"], x)
        n = self._ada_ln(P, x, h, cf["attn_norm"], 1e-5)
        qkv = conv1d(n, cf["qkv"])
        att = attention(qkv, 512, 512, 512, H=8, D=64, scale=64 ** -0.5)
        x2 = conv1d(att, cf["out"], res=x_ff1)
        n = self._ada_ln(P, x2, h, cf["conv_norm"], 1e-5)
        g = conv1d(n, cf["pw1"])
        half = g.shape[1] // 2
        gl = torch.empty((B, half, T), device=x.device, dtype=torch.float32)
        L.call("sty_glu_fwd", g.data_ptr(), gl.data_ptr(), B, half, T, L.stream_ptr())
        d = torch.empty_like(gl)
        dwconv1d(gl, cf["dw_w"], cf["dw_b"], K=31, pad_left=15, out=d, post_scale=cf["bn_scale"],
                 post_shift=cf["bn_shift"], act=ACT_SWISH)
        x3 = conv1d(d, cf["pw2"], res=x2)
        x4 = ff(cf["ff2"], x3)
        return self._ada_ln(P, x4, h, cf["post_norm"], 1e-5)

    def convnext(self, P: Packed, blk, x, h, out=None):
        """GeneratorConvNeXtBlock (conv_next.py:80-93) -> (B,C,T).  At the output rate (C = 32, pitch-padded rows)
        the whole block is ONE C-ABI call that never stores the 4C-wide intermediate (csrc/convnext_fused.cu) and
        writes a new tensor (`out`, or a fresh pitch-padded one); elsewhere the block runs in place on x."""
        B, Cc, T = x.shape
        J = P.fc_rows
        gb = self._gb(P, h, blk["norm"])
        inter = blk["pw1"].CO
        if (FUSE_CONVNEXT and USE_UMMA and Cc == 32 and inter == 128 and T >= 512 and blk["pw1"].split is not None
                and blk["pw2"].split is not None and _rows_aligned(x)):
            y = out if out is not None else empty_bct(B, Cc, T, x.device)
            assert _rows_aligned(y) and y.data_ptr() != x.data_ptr()
            ws = torch.empty((2, B, inter), device=x.device, dtype=torch.float32)
            L.call("sty_convnext_fused_fwd", x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), y.stride(0),
                   y.stride(1), blk["dw_w"].data_ptr(), blk["dw_b"].data_ptr(), gb.data_ptr(), J, 1e-6,
                   blk["pw1"].split.data_ptr(), blk["pw1"].bias.data_ptr(), blk["snake"].data_ptr(),
                   blk["grn_gamma"].data_ptr(), blk["pw2"].split.data_ptr(), blk["pw2"].bias.data_ptr(),
                   ws[0].data_ptr(), ws[1].data_ptr(), B, Cc, inter, T, L.stream_ptr())
            return y
        if out is not None:  # unfused path works in place: start from a copy in the caller's buffer
            out.copy_(x)
            x = out
        sumsq = torch.zeros((B, inter), device=x.device, dtype=torch.float32)
        fused_front = (USE_UMMA and blk["pw1"].split is not None and Cc <= 64 and Cc % 16 == 0
                       and T >= UMMA_MIN_T and x.stride(2) == 1)
        if fused_front:  # depthwise k7 + LN + AdaLN computed by the pointwise conv's producer warps
            hb = conv1d(x, blk["pw1"], out_act=ACT_SNAKE, out_alpha=blk["snake"], out_sumsq=sumsq,
                        dwln=(blk["dw_w"], blk["dw_b"], gb, J, 1e-6))
        else:
            y = torch.empty((B, Cc, T), device=x.device, dtype=torch.float32)
            L.call("sty_dwconv_ln_fwd", x.data_ptr(), x.stride(0), blk["dw_w"].data_ptr(),
                   blk["dw_b"].data_ptr(), gb.data_ptr(), J, y.data_ptr(), y.stride(0), B, Cc, T, 1e-6,
                   L.stream_ptr())
            hb = conv1d(y, blk["pw1"], out_act=ACT_SNAKE, out_alpha=blk["snake"], out_sumsq=sumsq)
        gs = torch.empty_like(sumsq)
        L.call("sty_grn_scale_fwd", sumsq.data_ptr(), blk["grn_gamma"].data_ptr(), gs.data_ptr(), B,
               inter, L.stream_ptr())
        conv1d(hb, blk["pw2"], in_scale=gs, res=x, out=x)
        return x

    def gen_block(self, P: Packed, blocks, x, h, mom=None):
        """AdaptiveGeneratorBlock, in place on x (ada_norm.py:109-120).  The InstanceNorm statistics of
        every AdaIN are accumulated (sum, sum of squares per (b,c)) by the conv that produces its input
        (`mom` = those of x, from the caller's conv) instead of a separate pass over the tensor."""
        J = P.fc_rows
        B, Cc, T = x.shape
        fused = mom is not None and FUSE_STATS
        for blk in blocks:
            if fused:
                sc1, sh1 = moments_affine(mom, self._gb(P, h, blk["n1"]), J, T)
                mom_t = torch.zeros((2, B, Cc), device=x.device, dtype=torch.float32)
                mom = torch.zeros((2, B, Cc), device=x.device, dtype=torch.float32)
            else:
                sc1, sh1 = instnorm_affine(x, self._gb(P, h, blk["n1"]), J)
                mom_t = None
            xt = conv1d(x, blk["c1"], dil=blk["dil"], in_scale=sc1, in_shift=sh1, in_act=ACT_SNAKE,
                        in_alpha=blk["a1"], out_sum=None if mom_t is None else mom_t[0],
                        out_sumsq=None if mom_t is None else mom_t[1])
            if fused:
                sc2, sh2 = moments_affine(mom_t, self._gb(P, h, blk["n2"]), J, T)
            else:
                sc2, sh2 = instnorm_affine(xt, self._gb(P, h, blk["n2"]), J)
            conv1d(xt, blk["c2"], in_scale=sc2, in_shift=sh2, in_act=ACT_SNAKE, in_alpha=blk["a2"],
                   res=x, out=x, out_sum=mom[0] if fused else None, out_sumsq=mom[1] if fused else None)
        return x

    def harmonic_prior(self, P: Packed, pitch, voiced, noise, taps=None):
        B, Fr = pitch.shape
        mc = P.mc
        hop = mc.hop_length
        Lw = Fr * hop
        H = P.src_w.numel()
        dev = pitch.device
        if noise is None:
            # the reference draws these with torch.randn (generator.py:440)
            noise = torch.randn((B, Lw, H), device=dev, dtype=torch.float32)
        assert noise.shape == (B, Lw, H) and noise.is_contiguous()
        work = torch.empty((B, H, Fr), device=dev, dtype=torch.float64)
        wave = torch.empty((B, Lw), device=dev, dtype=torch.float32)
        L.call("sty_source_fwd", pitch.data_ptr(), voiced.data_ptr(), noise.data_ptr(),
               P.src_w.data_ptr(), P.src_b.data_ptr(), work.data_ptr(), wave.data_ptr(), B, Fr, hop,
               H, float(mc.sample_rate), 0.1, 0.003, 10.0, L.stream_ptr())
        hop_s = hop // 75
        S = Lw // hop_s
        spec = empty_bct(B, P.hidden_s, S, dev)
        phase = empty_bct(B, P.hidden_s, S, dev)
        L.call("sty_stft_pitched_fwd", wave.data_ptr(), P.stft_f_re.data_ptr(), P.stft_f_im.data_ptr(),
               spec.data_ptr(), phase.data_ptr(), spec.stride(0), spec.stride(1), B, Lw, 64, hop_s, P.hidden_s,
               L.stream_ptr())
        if taps is not None:
            taps["prior_wave"] = wave
        return spec, phase

    def prior_branch(self, P: Packed, pin, h, pitch, voiced, noise, prior=None, taps=None):
        """harmonic source -> STFT -> the two prior convs + AdaptiveGeneratorBlocks, written into channels [Hs, 3 Hs)
        of the phase-head input `pin`.  Depends on pitch / voiced / style only (generator.py:719-760), so forward()
        runs it on a side stream next to the text encoder / decoder / conformer."""
        dev = pin.device
        B, Hs = pin.shape[0], P.hidden_s
        if prior is None:
            har_spec, har_phase = self.harmonic_prior(P, pitch, voiced, noise, taps)
        else:  # injected (parity tests): same pitch-padded layout as the computed prior
            har_spec, har_phase = (empty_bct(*t.shape, dev).copy_(t) for t in prior)
        if taps is not None:
            taps["har_spec"], taps["har_phase"] = har_spec, har_phase
        mom = torch.zeros((2, 2, B, Hs), device=dev, dtype=torch.float32)
        lp = conv1d(har_spec, P.amp_prior_conv, out=pin[:, Hs:2 * Hs], out_sum=mom[0, 0], out_sumsq=mom[0, 1])
        self.gen_block(P, P.amp_prior_block, lp, h, mom[0])
        pp = conv1d(har_phase, P.phase_prior_conv, out=pin[:, 2 * Hs:], out_sum=mom[1, 0], out_sumsq=mom[1, 1])
        self.gen_block(P, P.phase_prior_block, pp, h, mom[1])
        if taps is not None:
            taps["logamp_prior"], taps["phase_prior"] = lp.clone(), pp.clone()

    def generator(self, P: Packed, mel, h, pitch, voiced, noise, prior=None, taps=None, pin=None, join=None):
        """pin / join: the phase-head input whose prior channels forward() is already filling on a side stream, and
        the callable that makes the current stream wait for it"""
        B, _, Fr = mel.shape
        dev = mel.device
        x = conv1d(mel, P.amp_input)
        x = chan_layernorm(x, *P.amp_norm, eps=1e-6)
        if taps is not None:
            taps["amp_norm"] = x.clone()
        x = self.conformer(P, x, h)
        if taps is not None:
            taps["conformer"] = x.clone()
        Hs = P.hidden_s
        S = Fr * P.mc.hop_length // (P.mc.hop_length // 75)
        if pin is None:
            # phase-head input: [upsampled mel | logamp prior | phase prior] in one buffer
            pin = empty_bct(B, 3 * Hs, S, dev)
            self.prior_branch(P, pin, h, pitch, voiced, noise, prior, taps)
        for blk in P.amp_convnext:
            x = self.convnext(P, blk, x, h)
        if taps is not None:
            taps["amp_convnext"] = x.clone()
        for i, (cw, blk, r) in enumerate(zip(P.upconvs, P.upblocks, P.rates)):
            last = i == len(P.rates) - 1
            if last:  # output rate: pitch-padded rows, the block writes straight into the phase-head input
                x = conv1d(x, cw, shuffle=r, out=empty_bct(B, cw.CO // r, x.shape[2] * r, dev))
                x = self.convnext(P, blk, x, h, out=pin[:, :Hs])
            else:
                x = conv1d(x, cw, shuffle=r)
                x = self.convnext(P, blk, x, h)
        if taps is not None:
            taps["upsampled"] = x.clone()
        la = chan_layernorm(x, *P.amp_final_ln, eps=1e-6)
        logamp = conv1d(la, P.amp_output)
        if join is not None:
            join()  # the prior channels of `pin` are complete
        ph = conv1d(pin, P.phase_input)
        ph = chan_layernorm(ph, *P.phase_norm, eps=1e-6, out=ph)
        for blk in P.phase_convnext:
            ph = self.convnext(P, blk, ph, h)
        ph = chan_layernorm(ph, *P.phase_final_ln, eps=1e-6, out=ph)
        ri = conv1d(ph, P.phase_out_ri)  # (B, 2*Hs, S): real | imag
        if taps is not None:
            taps["logamp"], taps["real"], taps["imag"] = logamp, ri[:, :Hs], ri[:, Hs:]
        hop_s = P.mc.hop_length // 75
        audio = torch.empty((B, 1, S * hop_s), device=dev, dtype=torch.float32)
        L.call("sty_istft_head_pitched_fwd", logamp.data_ptr(), logamp.stride(0), logamp.stride(1), ri.data_ptr(),
               ri[:, Hs:].data_ptr(), ri.stride(0), ri.stride(1), P.stft_b_re.data_ptr(),
               P.stft_b_im.data_ptr(), audio.data_ptr(), B, S, Hs, 64, hop_s, L.stream_ptr())
        return audio

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, texts, text_lengths, alignment, pitch, energy, voiced, style,
                denormal_pitch, *, source_draws=None, prior=None, taps=None):
        dev = texts.device
        L.require_cuda(dev, "inputs")
        P = self.packed(dev)
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
        texts = texts.to(torch.int64).contiguous()
        text_lengths = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
        alignment, pitch, energy = f32(alignment), f32(pitch), f32(energy)
        voiced, style, denormal_pitch = f32(voiced), f32(style), f32(denormal_pitch)
        B = texts.shape[0]
        h = torch.empty((B, P.fc_rows), device=dev, dtype=torch.float32)
        L.call("sty_linear_rows_fwd", style.data_ptr(), P.fc_w.data_ptr(), P.fc_b.data_ptr(),
               h.data_ptr(), B, style.shape[1], P.fc_rows, L.stream_ptr())
        noise = None
        if source_draws is not None:
            noise = f32(source_draws["noise"])
        pin = join = None
        if OVERLAP_PRIOR and taps is None and prior is None:
            # the harmonic-prior branch (source, STFT, 2 x (k21 conv + AdaptiveGeneratorBlock): S-rate kernels that fill
            # the GPU) has no dependence on the text: it runs on a side stream while the text encoder, decoder and
            # conformer (frame-rate kernels with small grids) leave most SMs idle.  Captured graphs keep the fork / join.
            main = torch.cuda.current_stream()
            side = self._side_streams.get(dev)
            if side is None:
                side = self._side_streams[dev] = torch.cuda.Stream(device=dev)
            pin = empty_bct(B, 3 * P.hidden_s, pitch.shape[1] * P.mc.hop_length // (P.mc.hop_length // 75), dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.prior_branch(P, pin, h, denormal_pitch, voiced, noise)
            join = lambda: main.wait_stream(side)
        mu, _, _ = self.text_encoder(P, texts, text_lengths, taps)
        if taps is not None:
            taps["text_encoding"] = mu
        mel = self.decoder(P, mu, alignment, pitch, energy, voiced, h, taps)
        if taps is not None:
            taps["decoder"] = mel
        return self.generator(P, mel, h, denormal_pitch, voiced, noise, prior=prior, taps=taps, pin=pin, join=join)


# ==========================================================================
# duration predictor, pitch/energy predictor, alignment, text -> wav
# ==========================================================================
def attention_generic(q, k, v, *, H, D, lengths=None, rope=None, scale):
    """q: (B,H*D,T) view; k, v: views sharing one batch stride (e.g. halves of a fused k|v buffer)."""
    B, _, T = q.shape
    assert q.stride(1) == T and k.stride(1) == T and v.stride(1) == T and k.stride(0) == v.stride(0)
    out = torch.empty((B, H * D, T), device=q.device, dtype=torch.float32)
    rc, rs, d_rot = (None, None, 0) if rope is None else (rope[0].data_ptr(), rope[1].data_ptr(),
                                                          rope[2])
    L.call("sty_attention_generic_fwd", q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(),
           k.stride(0), out.data_ptr(), out.stride(0), L.ptr(lengths), rc, rs, d_rot, B, H, D, T,
           scale, L.stream_ptr())
    return out


class _EngineBase:
    pack_cls = None

    def __init__(self, module):
        L.load()
        self.module = module
        self._packed = None
        self._packed_key = None

    def packed(self, device):
        key = (str(device), L.param_epoch) + tuple(p._version for p in self.module.parameters()) + \
            tuple(b._version for b in self.module.buffers())
        if self._packed is None or key != self._packed_key:
            self._packed = self.pack_cls(self.module, device)
            self._packed_key = key
        return self._packed

    @staticmethod
    def style_fc(P, style):
        B = style.shape[0]
        h = torch.empty((B, P.fc_rows), device=style.device, dtype=torch.float32)
        L.call("sty_linear_rows_fwd", style.data_ptr(), P.fc_w.data_ptr(), P.fc_b.data_ptr(),
               h.data_ptr(), B, style.shape[1], P.fc_rows, L.stream_ptr())
        return h

    # shared stage implementations (same kernels as the speech predictor)
    text_encoder = SpeechEngine.text_encoder
    _gb = staticmethod(SpeechEngine._gb)
    _ada_ln = SpeechEngine._ada_ln
    _decoder_block = SpeechEngine._decoder_block


class PackedDuration(PackBase):
    def __init__(self, module, device):
        super().__init__(module, device)
        sd, cv = self.sd, self.cv
        self.q = cv("cross_attention.conv_q")
        wkv = torch.cat([sd["cross_attention.conv_k.weight"], sd["cross_attention.conv_v.weight"]], 0)
        bkv = torch.cat([sd["cross_attention.conv_k.bias"], sd["cross_attention.conv_v.bias"]], 0)
        self.kv = ConvW(wkv, bkv)
        self.o = cv("cross_attention.conv_o")
        self.post_dw_w = self.wn("cross_post.0").reshape(-1, 5).contiguous()
        self.post_dw_b = sd["cross_post.0.bias"].contiguous()
        self.post_pw = cv("cross_post.2")
        self.blocks = []
        for i in range(self.mc.duration_predictor.n_layer):
            p = f"conv_next.{i}"
            w2 = sd[p + ".pwconv2.weight"]
            b2 = sd[p + ".pwconv2.bias"] + w2 @ sd[p + ".grn.beta"].reshape(-1)
            self.blocks.append(dict(
                dw_w=sd[p + ".dwconv.weight"].reshape(-1, 7).contiguous(),
                dw_b=sd[p + ".dwconv.bias"].contiguous(), norm=p + ".norm", pw1=self.lin(p + ".pwconv1"),
                grn_gamma=sd[p + ".grn.gamma"].reshape(-1).contiguous(),
                pw2=ConvW(w2.unsqueeze(-1), b2)))
        self.proj = self.lin("duration_proj.linear_layer")


class DurationEngine(_EngineBase):
    """DurationPredictor.forward (duration_predictor.py:58-87) -> (B,T,classes)."""
    pack_cls = PackedDuration

    @torch.no_grad()
    def forward(self, texts, text_lengths, style, taps=None):
        dev = texts.device
        L.require_cuda(dev, "inputs")
        P = self.packed(dev)
        texts = texts.to(torch.int64).contiguous()
        lengths = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
        style = style.to(device=dev, dtype=torch.float32).contiguous()
        B, T = texts.shape
        h = self.style_fc(P, style)
        enc, _, mask = self.text_encoder(P, texts, lengths)  # (B,C,T)
        Cc = enc.shape[1]
        qn = self._ada_ln(P, enc, h, "query_norm", 1e-5)
        kn = self._ada_ln(P, enc, h, "key_norm", 1e-5)
        qkv = torch.empty((B, 3 * Cc, T), device=dev, dtype=torch.float32)
        conv1d(qn, P.q, out=qkv[:, :Cc])
        conv1d(kn, P.kv, out=qkv[:, Cc:])
        H = 8
        D = Cc // H
        att = attention(qkv, Cc, Cc, Cc, H=H, D=D, lengths=lengths, rope=P.rope(T, D),
                        scale=1.0 / math.sqrt(D))
        att = conv1d(att, P.o)
        dw = torch.empty_like(att)
        dwconv1d(att, P.post_dw_w, P.post_dw_b, K=5, pad_left=2, out=dw, act=ACT_SWISH)
        pros = conv1d(dw, P.post_pw, res=enc, out_scale=INV_SQRT2, res_scale=INV_SQRT2)
        if taps is not None:
            taps["dur_cross"] = pros.clone()
        J = P.fc_rows
        for blk in P.blocks:  # AdaptiveConvNeXtBlock (GELU) then * mask  (conv_next.py:125-141)
            gb = self._gb(P, h, blk["norm"])
            y = torch.empty_like(pros)
            L.call("sty_dwconv_ln_fwd", pros.data_ptr(), pros.stride(0), blk["dw_w"].data_ptr(),
                   blk["dw_b"].data_ptr(), gb.data_ptr(), J, y.data_ptr(), y.stride(0), B, Cc, T, 1e-6,
                   L.stream_ptr())
            inter = blk["pw1"].CO
            sumsq = torch.zeros((B, inter), device=dev, dtype=torch.float32)
            hb = conv1d(y, blk["pw1"], out_act=L.ACT_GELU, out_sumsq=sumsq)
            gs = torch.empty_like(sumsq)
            L.call("sty_grn_scale_fwd", sumsq.data_ptr(), blk["grn_gamma"].data_ptr(), gs.data_ptr(), B,
                   inter, L.stream_ptr())
            # (residual + block) * mask: the residual is already masked from the second block on,
            # for the first one the mask is applied to the sum by masking both terms
            if blk is P.blocks[0]:
                resm = torch.empty_like(pros)
                L.call("sty_scale_mask_fwd", pros.data_ptr(), mask.data_ptr(), resm.data_ptr(), B, Cc, T,
                       1.0, L.stream_ptr())
                pros = resm
            conv1d(hb, blk["pw2"], in_scale=gs, out_mask=mask, res=pros, out=pros)
        logits = conv1d(pros, P.proj)  # (B,NC,T)
        NC = logits.shape[1]
        out = torch.empty((B, T, NC), device=dev, dtype=torch.float32)
        L.call("sty_duration_head_fwd", logits.data_ptr(), lengths.data_ptr(), out.data_ptr(), B, NC, T,
               L.stream_ptr())
        return out


class PackedPE(PackBase):
    def __init__(self, module, device):
        super().__init__(module, device)
        sd, cv = self.sd, self.cv
        pe = "prosody_encoder"
        self.layers = []
        for i in range(3):
            self.layers.append(dict(
                qkv=self.qkv(f"{pe}.attn_layers.{i}"), o=cv(f"{pe}.attn_layers.{i}.conv_o"),
                n1=f"{pe}.norm_layers_1.{i}", f1=cv(f"{pe}.ffn_layers.{i}.conv_1"),
                f2=cv(f"{pe}.ffn_layers.{i}.conv_2"), n2=f"{pe}.norm_layers_2.{i}",
                proj=cv(f"{pe}.proj_layers.{i}")))
        self.f0 = [self.dblock(f"F0.{i}") for i in range(4)]
        self.n = [self.dblock(f"N.{i}") for i in range(4)]
        self.f0_proj = cv("F0_proj")
        self.n_proj = cv("N_proj")


class PitchEnergyEngine(_EngineBase):
    """PitchEnergyPredictor.forward (pitch_energy_predictor.py:62-82) -> pitch (B,F), energy (B,F)."""
    pack_cls = PackedPE

    @torch.no_grad()
    def forward(self, texts, text_lengths, alignment, style, taps=None):
        dev = texts.device
        L.require_cuda(dev, "inputs")
        P = self.packed(dev)
        texts = texts.to(torch.int64).contiguous()
        lengths = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
        style = style.to(device=dev, dtype=torch.float32).contiguous()
        alignment = alignment.to(device=dev, dtype=torch.float32).contiguous()
        B, T = texts.shape
        Fr = alignment.shape[2]
        h = self.style_fc(P, style)
        dm = P.proj_m.CO
        sdim = style.shape[1]
        Ch = dm + sdim
        # x = cat[text encoding, style]: the style rows are constant over time; the encoder's last
        # projection writes straight into the first dm channels
        x = torch.empty((B, Ch, T), device=dev, dtype=torch.float32)
        x2 = torch.empty_like(x)
        for buf in (x, x2):
            L.call("sty_broadcast_rows_fwd", style.data_ptr(), buf[:, dm:].data_ptr(), buf.stride(0), B,
                   sdim, T, L.stream_ptr())
        _, _, mask = self.text_encoder(P, texts, lengths, mu_out=x[:, :dm])
        H = 2
        D = Ch // H
        rope = P.rope(T, D)
        cur, nxt = x, x2
        for ly in P.layers:
            qkv = conv1d(cur, ly["qkv"], in_mask=mask)
            att = attention_generic(qkv[:, :Ch], qkv[:, Ch:2 * Ch], qkv[:, 2 * Ch:], H=H, D=D,
                                    lengths=lengths, rope=rope, scale=1.0 / math.sqrt(D))
            y = conv1d(att, ly["o"], in_mask=None, res=None)
            # x*mask + y, AdaLN
            xm = torch.empty_like(cur)
            L.call("sty_scale_mask_fwd", cur.data_ptr(), mask.data_ptr(), xm.data_ptr(), B, Ch, T, 1.0,
                   L.stream_ptr())
            gb = self._gb(P, h, ly["n1"])
            x1 = chan_layernorm(y, gb, gb[:, Ch:], eps=1e-5, res=xm, g_bs=P.fc_rows, plus_one=True)
            hh = conv1d(x1, ly["f1"], in_mask=mask, out_act=ACT_RELU)
            y2 = conv1d(hh, ly["f2"], in_mask=mask, out_mask=mask)
            gb = self._gb(P, h, ly["n2"])
            x2n = chan_layernorm(y2, gb, gb[:, Ch:], eps=1e-5, res=x1, g_bs=P.fc_rows, plus_one=True)
            conv1d(x2n, ly["proj"], out=nxt[:, :dm])
            cur, nxt = nxt, cur
        pros = torch.empty_like(cur)  # final x * mask, (B, Ch, T)
        L.call("sty_scale_mask_fwd", cur.data_ptr(), mask.data_ptr(), pros.data_ptr(), B, Ch, T, 1.0,
               L.stream_ptr())
        if taps is not None:
            taps["prosody"] = pros.transpose(1, 2).clone()
        xa = torch.empty((B, Ch, Fr), device=dev, dtype=torch.float32)
        L.call("sty_bmm_fwd", pros.data_ptr(), pros.stride(0), alignment.data_ptr(), alignment.stride(0),
               xa.data_ptr(), xa.stride(0), B, Ch, Fr, T, L.stream_ptr())
        outs = []
        for tower, proj in ((P.f0, P.f0_proj), (P.n, P.n_proj)):
            z = xa
            for blk in tower:
                out = torch.empty((B, blk["c2"].CO, Fr), device=dev, dtype=torch.float32)
                z = self._decoder_block(P, blk, z, h, out)
            outs.append(conv1d(z, proj).reshape(B, Fr))
        return outs[0], outs[1]


CLASS_TO_DUR = (1, 2, 3, 4, 5, 6, 7, 9, 12, 15, 18, 22, 27, 32, 38, 46)


def soft_durations(pred, text_lengths):
    """DurationProcessor.prediction_to_duration (utils.py:745-750): class scores (B,T,NC) -> durations (B,T)"""
    dev = pred.device
    B, T, NC = pred.shape
    pred = pred.contiguous()
    lengths = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
    table = torch.tensor(CLASS_TO_DUR[:NC], device=dev, dtype=torch.float32)
    dur = torch.empty((B, T), device=dev, dtype=torch.float32)
    total = torch.zeros((1,), device=dev, dtype=torch.int32)
    L.call("sty_soft_duration_fwd", pred.data_ptr(), lengths.data_ptr(), table.data_ptr(), dur.data_ptr(),
           total.data_ptr(), B, T, NC, L.stream_ptr())
    return dur


def duration_to_alignment(pred, text_lengths, multiplier=1):
    """DurationProcessor.forward (utils.py:804-807): class scores (B,T,NC) -> (alignment (B,T,F),
    durations (B,T)).  The frame count is data dependent: one device->host read, like the
    reference's `.item()` (utils.py:759).  ``multiplier`` (ModelConfig.coarse_multiplier) scales the frame
    count and the durations before the alignment is laid out (utils.py:759-761)."""
    dev = pred.device
    B, T, NC = pred.shape
    pred = pred.contiguous()
    lengths = text_lengths.to(device=dev, dtype=torch.int64).contiguous()
    table = torch.tensor(CLASS_TO_DUR[:NC], device=dev, dtype=torch.float32)
    dur = torch.empty((B, T), device=dev, dtype=torch.float32)
    total = torch.zeros((1,), device=dev, dtype=torch.int32)
    L.call("sty_soft_duration_fwd", pred.data_ptr(), lengths.data_ptr(), table.data_ptr(), dur.data_ptr(),
           total.data_ptr(), B, T, NC, L.stream_ptr())
    Fr = int(total.item()) * int(multiplier)
    if Fr <= 0:
        raise RuntimeError("stylish_tts_b200: predicted durations sum to zero frames")
    al = torch.empty((B, T, Fr), device=dev, dtype=torch.float32)
    dur_m = dur if multiplier == 1 else (dur * float(multiplier)).contiguous()
    L.call("sty_alignment_fwd", dur_m.data_ptr(), al.data_ptr(), B, T, Fr, L.stream_ptr())
    return al, dur
