"""Host-side mirror of the reference's model-build API (hot path only).

``build_model(model_config)`` returns a mapping whose hot-path entries are
``nn.Module`` shells with the SAME parameter / buffer names and shapes as the
reference modules (reference: src/stylish_tts/train/models/models.py:29-85,
state-dict layout SURVEY.md Appendix D), so reference checkpoints load with
``load_state_dict(strict=True)``.  The shells own parameters only; their
``forward`` hands raw device pointers to the sm_100a kernels in
``csrc/`` through the C-ABI (``include/stylish_b200.h``).  There is no
PyTorch/CPU fallback: calling ``forward`` without the CUDA library raises.

The module tree is described declaratively (``Node`` = named container) rather
than class-per-layer: the shells carry no arithmetic of their own.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import nn
from torch.nn.utils.parametrizations import weight_norm

STYLE_DIM_DEFAULT = 64


class Node(nn.Module):
    """Named container: children may be modules, Parameters or buffers."""

    def __init__(self, **children):
        super().__init__()
        for name, child in children.items():
            self.put(name, child)

    def put(self, name, child):
        if isinstance(child, nn.Parameter):
            self.register_parameter(name, child)
        elif isinstance(child, torch.Tensor):
            self.register_buffer(name, child)
        elif isinstance(child, (list, tuple)):
            self.add_module(name, seq(child))
        else:
            self.add_module(name, child)
        return self


def seq(items) -> Node:
    """Index-named children ("0", "1", ...), like ModuleList/Sequential keys."""
    n = Node()
    for i, it in enumerate(items):
        if it is not None:
            n.put(str(i), it)
    return n


def ones(*shape):
    return nn.Parameter(torch.ones(*shape))


def zeros(*shape):
    return nn.Parameter(torch.zeros(*shape))


def conv(ci, co, k, *, bias=True, wn=False, groups=1):
    c = nn.Conv1d(ci, co, k, groups=groups, bias=bias)
    return weight_norm(c) if wn else c


def chan_norm(c):
    # reference text_encoder.LayerNorm: parameters gamma/beta (text_encoder.py:15-22)
    return Node(gamma=ones(c), beta=zeros(c))


def ada_fc(style_dim, c):
    # AdaptiveInstance / AdaptiveLayerNorm: fc = Linear(style, 2*C) (ada_norm.py:129-133,195-201)
    return Node(fc=nn.Linear(style_dim, 2 * c))


# --------------------------------------------------------------------------
# text encoder (reference text_encoder.py:397-463)
# --------------------------------------------------------------------------
def text_encoder_tree(inter_dim, cfg) -> Node:
    c, f, k, L = cfg.hidden_dim, cfg.filter_channels, cfg.kernel_size, cfg.layers
    emb = nn.Embedding(cfg.tokens, c)
    nn.init.normal_(emb.weight, 0.0, c ** -0.5)
    proj = conv(c, c, 1)
    nn.init.zeros_(proj.weight)
    nn.init.zeros_(proj.bias)
    prenet = Node(
        conv_layers=[conv(c, c, 5) for _ in range(3)],
        norm_layers=[chan_norm(c) for _ in range(3)],
        proj=proj,
    )

    def mha():
        m = Node(conv_q=conv(c, c, 1), conv_k=conv(c, c, 1), conv_v=conv(c, c, 1),
                 conv_o=conv(c, c, 1))
        for n in ("conv_q", "conv_k", "conv_v"):
            nn.init.xavier_uniform_(getattr(m, n).weight)
        return m

    encoder = Node(
        attn_layers=[mha() for _ in range(L)],
        norm_layers_1=[chan_norm(c) for _ in range(L)],
        ffn_layers=[Node(conv_1=conv(c, f, k), conv_2=conv(f, c, k)) for _ in range(L)],
        norm_layers_2=[chan_norm(c) for _ in range(L)],
    )
    return Node(emb=emb, prenet=prenet, encoder=encoder, proj_m=conv(c, inter_dim, 1))


# --------------------------------------------------------------------------
# decoder (reference decoder.py:7-90, ada_norm.py:143-192)
# --------------------------------------------------------------------------
def decoder_block_tree(dim_in, dim_out, style_dim) -> Node:
    n = Node(
        conv1=conv(dim_in, dim_out, 3, wn=True),
        conv2=conv(dim_out, dim_out, 3, wn=True),
        norm1=ada_fc(style_dim, dim_in),
        norm2=ada_fc(style_dim, dim_out),
    )
    if dim_in != dim_out:
        n.put("conv1x1", conv(dim_in, dim_out, 1, bias=False, wn=True))
    return n


def decoder_tree(dim_in, style_dim, hidden, residual) -> Node:
    return Node(
        encode=decoder_block_tree(dim_in + 3, hidden, style_dim),
        decode=[decoder_block_tree(hidden + 3 + residual, hidden, style_dim) for _ in range(4)],
        F0_conv=conv(1, 1, 3, wn=True),
        N_conv=conv(1, 1, 3, wn=True),
        voiced_conv=conv(1, 1, 3, wn=True),
        asr_res=[conv(dim_in, residual, 1, wn=True)],
    )


# --------------------------------------------------------------------------
# vocoder (reference generator.py:513-901, conv_next.py:57-93, conformer.py)
# --------------------------------------------------------------------------
def convnext_tree(dim, style_dim) -> Node:
    inter = 4 * dim
    return Node(
        snake=ones(1, 1, inter),
        dwconv=conv(dim, dim, 7, groups=dim),
        norm=ada_fc(style_dim, dim),
        pwconv1=nn.Linear(dim, inter),
        grn=Node(gamma=zeros(1, 1, inter), beta=zeros(1, 1, inter)),
        pwconv2=nn.Linear(inter, dim),
    )


def gen_block_tree(ch, style_dim, k=11) -> Node:
    return Node(
        convs1=[conv(ch, ch, k, wn=True) for _ in range(3)],
        convs2=[conv(ch, ch, k, wn=True) for _ in range(3)],
        adain1=[ada_fc(style_dim, ch) for _ in range(3)],
        adain2=[ada_fc(style_dim, ch) for _ in range(3)],
        alpha1=[ones(1, ch, 1) for _ in range(3)],
        alpha2=[ones(1, ch, 1) for _ in range(3)],
    )


def conformer_block_tree(dim, style_dim, heads=8, dim_head=64, ff_mult=4, expansion=2,
                         kernel=31) -> Node:
    inner = heads * dim_head
    cin = dim * expansion

    def ff():
        net = Node()
        net.put("0", nn.Linear(dim, dim * ff_mult))
        net.put("3", nn.Linear(dim * ff_mult, dim))
        # Scale(PreNorm(FeedForward)) -> keys  ffN.fn.fn.net.K / ffN.fn.norm.fc
        return Node(fn=Node(fn=Node(net=net), norm=ada_fc(style_dim, dim)))

    attn = Node(
        fn=Node(to_q=nn.Linear(dim, inner, bias=False),
                to_kv=nn.Linear(dim, inner * 2, bias=False),
                to_out=nn.Linear(inner, dim)),
        norm=ada_fc(style_dim, dim),
    )
    net = Node()
    net.put("1", conv(dim, cin * 2, 1))
    net.put("3", Node(conv=conv(cin, cin, kernel, groups=cin)))
    net.put("4", nn.BatchNorm1d(cin))
    net.put("6", conv(cin, dim, 1))
    cmod = Node(norm=ada_fc(style_dim, dim), net=net)
    return Node(ff1=ff(), attn=attn, conv=cmod, ff2=ff(), post_norm=ada_fc(style_dim, dim))


def stft_buffers(n_fft, win_length):
    """Windowed DFT bases of the conv-STFT (reference stft.py:39-96): periodic
    hann, forward cos/-sin, backward cos/sin scaled by 1/n_fft (no OLA
    normalisation, no doubling of interior bins)."""
    bins = n_fft // 2 + 1
    window = torch.hann_window(win_length, periodic=True, dtype=torch.float32)
    if win_length < n_fft:
        window = torch.nn.functional.pad(window, (0, n_fft - win_length))
    elif win_length > n_fft:
        window = window[:n_fft]
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(bins, dtype=torch.float64)
    ang = 2.0 * math.pi * torch.outer(k, n) / n_fft
    w64 = window.double()
    f_re = (torch.cos(ang) * w64).float().unsqueeze(1)
    f_im = (-torch.sin(ang) * w64).float().unsqueeze(1)
    b_re = (torch.cos(ang) * (w64 / n_fft)).float().unsqueeze(1)
    b_im = (torch.sin(ang) * (w64 / n_fft)).float().unsqueeze(1)
    return Node(window=window, weight_forward_real=f_re, weight_forward_imag=f_im,
                weight_backward_real=b_re, weight_backward_imag=b_im)


def basegen_tree(*, style_dim, n_fft, win_length, hop_length, scale, scalehop, hidden_dim,
                 input_dim, io_k, conv_layers, upsample_rates) -> Node:
    amp_layers = conv_layers - len(upsample_rates)
    upconvs, upblocks = [], []
    after = input_dim
    for s in upsample_rates:
        before, after = after, after // 2
        upconvs.append(conv(before, after * s, 11))
        upblocks.append(convnext_tree(after, style_dim))
    h = hidden_dim
    g = Node(
        amp_convnext=[convnext_tree(input_dim, style_dim) for _ in range(amp_layers)],
        upconvs=upconvs,
        upblocks=upblocks,
        m_source=Node(l_linear=nn.Linear(9, 1)),
        amp_prior_conv=conv(h, h, io_k),
        phase_prior_conv=conv(h, h, io_k),
        amp_prior_block=gen_block_tree(h, style_dim),
        phase_prior_block=gen_block_tree(h, style_dim),
        phase_input_conv=conv(3 * h, h, io_k),
        amp_output_conv=conv(h, h, io_k),
        phase_output_real_conv=conv(h, h, io_k),
        phase_output_imag_conv=conv(h, h, io_k),
        phase_norm=nn.LayerNorm(h, eps=1e-6),
        phase_convnext=[convnext_tree(h, style_dim) for _ in range(conv_layers)],
        amp_final_layer_norm=nn.LayerNorm(h, eps=1e-6),
        phase_final_layer_norm=nn.LayerNorm(h, eps=1e-6),
    )
    # reference Generator._init_weights (generator.py:705-708): every Conv1d inside
    for m in g.modules():
        if isinstance(m, nn.Conv1d):
            with torch.no_grad():
                w = torch.empty_like(m.weight)
                nn.init.trunc_normal_(w, std=0.02)
                if torch.nn.utils.parametrize.is_parametrized(m, "weight"):
                    m.weight = w  # routed through weight_norm's right_inverse
                else:
                    m.weight.copy_(w)
                if m.bias is not None:
                    m.bias.zero_()
    g.put("stft", stft_buffers(n_fft // scale, win_length // scale))
    return g


def generator_tree(*, style_dim, n_fft, win_length, hop_length, cfg) -> Node:
    hidden = n_fft // 2
    return Node(
        amp_input_conv=conv(cfg.input_dim, hidden, cfg.io_conv_kernel_size),
        amp_norm=nn.LayerNorm(hidden, eps=1e-6),
        amp_conformer=Node(layers=[conformer_block_tree(hidden, style_dim)
                                   for _ in range(cfg.conformer_layers)]),
        basegen=basegen_tree(
            style_dim=style_dim, n_fft=n_fft, win_length=win_length, hop_length=hop_length,
            scale=8, scalehop=75, hidden_dim=n_fft // 2 // 8, input_dim=hidden,
            io_k=cfg.io_conv_kernel_size, conv_layers=cfg.conv_layers,
            upsample_rates=[3, 5, 5]),
    )


class DecoderPrediction:
    """Same result object as the reference (utils.py:643-653)."""

    def __init__(self, *, audio, magnitude=None, phase=None):
        self.audio = audio
        self.magnitude = magnitude
        self.phase = phase


class SpeechPredictor(nn.Module):
    """Drop-in for reference SpeechPredictor (speech_predictor.py:11-73).

    forward(texts, text_lengths, alignment, pitch, energy, voiced, style,
            denormal_pitch) -> DecoderPrediction(audio (B,1,L))

    ``train()`` mode runs the reference's stochastic regularisers (dropout sites, decoder box smoothing);
    set ``regularisers = False`` for a deterministic training forward, ``regulariser_seed`` seeds the masks.
    """
    regularisers = True
    regulariser_seed = 0

    def __init__(self, model_config):
        super().__init__()
        mc = model_config
        self.model_config = mc
        self.text_encoder = text_encoder_tree(mc.inter_dim, mc.text_encoder)
        self.decoder = decoder_tree(mc.inter_dim, mc.style_dim, mc.decoder.hidden_dim,
                                    mc.decoder.residual_dim)
        self.generator = generator_tree(style_dim=mc.style_dim, n_fft=mc.n_fft,
                                        win_length=mc.win_length, hop_length=mc.hop_length,
                                        cfg=mc.generator)
        self._engine = None
        self._train_graph = None

    def engine(self):
        from .engine import SpeechEngine

        if self._engine is None:
            self._engine = SpeechEngine(self)
        return self._engine

    def train_graph(self):
        from .train_engine import TrainGraph

        if self._train_graph is None:
            self._train_graph = TrainGraph(self, self.engine())
        return self._train_graph

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if getattr(self, "_train_graph", None) is not None:
            self._train_graph.rebind()  # .to() / .cuda() rebind buffers: no stale pointers in the cached graph
        return out

    def forward(self, texts, text_lengths, alignment, pitch, energy, voiced, style,
                denormal_pitch, *, source_draws=None, prior=None, taps=None):
        wants_grad = torch.is_grad_enabled() and (
            any(p.requires_grad for p in self.parameters())
            or any(torch.is_tensor(t) and t.requires_grad for t in (pitch, energy, style)))
        if wants_grad:  # differentiable path: autograd primitives backed by the backward kernels
            audio = self.train_graph().forward(texts, text_lengths, alignment, pitch, energy, voiced,
                                               style, denormal_pitch, source_draws=source_draws,
                                               prior=prior)
            return DecoderPrediction(audio=audio, magnitude=None, phase=None)
        audio = self.engine().forward(texts, text_lengths, alignment, pitch, energy, voiced,
                                      style, denormal_pitch, source_draws=source_draws,
                                      prior=prior, taps=taps)
        return DecoderPrediction(audio=audio, magnitude=None, phase=None)


def adaptive_convnext_tree(dim, style_dim) -> Node:
    # AdaptiveConvNeXtBlock (conv_next.py:96-123): like the generator block but GELU, no snake
    inter = 4 * dim
    return Node(
        dwconv=conv(dim, dim, 7, groups=dim),
        norm=ada_fc(style_dim, dim),
        pwconv1=nn.Linear(dim, inter),
        grn=Node(gamma=zeros(1, 1, inter), beta=zeros(1, 1, inter)),
        pwconv2=nn.Linear(inter, dim),
    )


class _EngineModule(nn.Module):
    """Shared plumbing of the shells: lazily built inference engine (no_grad) and differentiable graph."""
    engine_cls_name = ""
    graph_cls_name = ""
    regularisers = True  # train() mode: dropout / Dropout1d / DropPath sites live, like the reference
    regulariser_seed = 0

    def engine(self):
        from . import engine as E

        if self._engine is None:
            self._engine = getattr(E, self.engine_cls_name)(self)
        return self._engine

    def train_graph(self):
        from . import train_engine as TE

        if getattr(self, "_train_graph", None) is None:
            self._train_graph = getattr(TE, self.graph_cls_name)(self)
        return self._train_graph

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if getattr(self, "_train_graph", None) is not None:
            self._train_graph.rebind()
        return out

    def _wants_grad(self, *inputs):
        return torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                            any(torch.is_tensor(t) and t.requires_grad for t in inputs))


class DurationPredictor(_EngineModule):
    """Drop-in for reference DurationPredictor (duration_predictor.py:15-87):
    forward(texts, text_lengths, style) -> (B, T, duration_classes)."""
    engine_cls_name = "DurationEngine"
    graph_cls_name = "DurationTrainGraph"

    def __init__(self, model_config):
        super().__init__()
        mc = model_config
        self.model_config = mc
        c = mc.inter_dim
        self.text_encoder = text_encoder_tree(c, mc.text_encoder)
        self.conv_next = seq([adaptive_convnext_tree(c, mc.style_dim)
                              for _ in range(mc.duration_predictor.n_layer)])
        lin = nn.Linear(c, mc.duration_predictor.duration_classes)
        nn.init.xavier_uniform_(lin.weight)
        self.duration_proj = Node(linear_layer=lin)
        self.query_norm = ada_fc(mc.style_dim, c)
        self.key_norm = ada_fc(mc.style_dim, c)
        att = Node(conv_q=conv(c, c, 1), conv_k=conv(c, c, 1), conv_v=conv(c, c, 1), conv_o=conv(c, c, 1))
        for n in ("conv_q", "conv_k", "conv_v"):
            nn.init.xavier_uniform_(getattr(att, n).weight)
        self.cross_attention = att
        post = Node()
        post.put("0", conv(c, c, 5, wn=True, groups=c))
        post.put("2", conv(c, c, 1, wn=True))
        self.cross_post = post
        self._engine = None

    def forward(self, texts, text_lengths, style, *, taps=None):
        if self._wants_grad(style):
            return self.train_graph().forward(texts, text_lengths, style)
        return self.engine().forward(texts, text_lengths, style, taps=taps)


class PitchEnergyPredictor(_EngineModule):
    """Drop-in for reference PitchEnergyPredictor (pitch_energy_predictor.py:8-82):
    forward(texts, text_lengths, alignment, style) -> (pitch (B,F), energy (B,F))."""
    engine_cls_name = "PitchEnergyEngine"
    graph_cls_name = "PitchEnergyTrainGraph"

    def __init__(self, model_config):
        super().__init__()
        mc = model_config
        self.model_config = mc
        d, sdim = mc.pitch_energy_predictor.inter_dim, mc.style_dim
        hc = d + sdim
        self.text_encoder = text_encoder_tree(d, mc.text_encoder)

        def mha():
            m = Node(conv_q=conv(hc, hc, 1), conv_k=conv(hc, hc, 1), conv_v=conv(hc, hc, 1),
                     conv_o=conv(hc, hc, 1))
            for n in ("conv_q", "conv_k", "conv_v"):
                nn.init.xavier_uniform_(getattr(m, n).weight)
            return m

        self.prosody_encoder = Node(
            attn_layers=[mha() for _ in range(3)],
            norm_layers_1=[ada_fc(sdim, hc) for _ in range(3)],
            ffn_layers=[Node(conv_1=conv(hc, 2 * hc, 1), conv_2=conv(2 * hc, hc, 1)) for _ in range(3)],
            norm_layers_2=[ada_fc(sdim, hc) for _ in range(3)],
            proj_layers=[conv(hc, d, 1) for _ in range(3)],
        )
        dims = [(hc, d), (d, d // 2), (d // 2, d // 2), (d // 2, d // 2)]
        self.F0 = seq([decoder_block_tree(a, b, sdim) for a, b in dims])
        self.N = seq([decoder_block_tree(a, b, sdim) for a, b in dims])
        self.F0_proj = conv(d // 2, 1, 1)
        self.N_proj = conv(d // 2, 1, 1)
        self._engine = None

    def forward(self, texts, text_lengths, alignment, style, *, taps=None):
        if self._wants_grad(style):
            return self.train_graph().forward(texts, text_lengths, alignment, style)
        return self.engine().forward(texts, text_lengths, alignment, style, taps=taps)


class DurationProcessor(nn.Module):
    """Drop-in for the reference DurationProcessor (utils.py:656-807).

    forward(pred, text_length) -> soft alignment (B,T,F) (prediction_to_duration + duration_to_alignment on the
    CUDA kernels, any integer coarse multiplier); the class <-> duration index maps (``dur_to_class``, ``class_to_dur_hard``,
    ``align_to_class``: table lookups, bit-exact, device-agnostic indexing) and ``class_to_dur_soft`` /
    ``prediction_to_duration`` that the duration stage uses for its targets and losses (stage_type.py:507-522)."""

    # durations (frames) represented by the 16 classes, and how many consecutive durations 0..50 map to each class
    CLASS_DURATIONS = (1, 2, 3, 4, 5, 6, 7, 9, 12, 15, 18, 22, 27, 32, 38, 46)
    CLASS_RUNS = (2, 1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 5, 5, 5, 7, 9)

    def __init__(self, class_count=16, max_dur=50):
        super().__init__()
        self.class_count, self.max_dur = class_count, max_dur
        self.register_buffer("class_to_dur_table", torch.tensor(self.CLASS_DURATIONS, dtype=torch.float32))
        self.register_buffer("dur_to_class_table", torch.repeat_interleave(
            torch.arange(len(self.CLASS_RUNS), dtype=torch.float32), torch.tensor(self.CLASS_RUNS)))

    def class_to_dur_soft(self, softdur):
        """expected duration of a class distribution (utils.py:726-730)"""
        return (softdur * self.class_to_dur_table).sum(dim=-1) / (softdur.sum(dim=-1) + 1e-9)

    def class_to_dur_hard(self, classes):
        return self.class_to_dur_table[classes.clamp(min=0, max=self.class_count)]

    def dur_to_class(self, durs):
        return self.dur_to_class_table[durs.clamp(min=1, max=self.max_dur).long()]

    def align_to_class(self, alignment):
        return self.dur_to_class(alignment.sum(dim=-1).clamp(min=1, max=50))

    def prediction_to_duration(self, pred, text_length):
        """softmax -> expected duration -> * sequence mask (utils.py:745-750); the CUDA kernel when pred is on the
        device (same result as forward()'s first half), the torch formula otherwise"""
        if pred.is_cuda and not pred.requires_grad:
            from .engine import soft_durations
            return soft_durations(pred, text_length)
        soft = self.class_to_dur_soft(torch.softmax(pred, dim=-1))
        mask = torch.arange(pred.shape[1], device=pred.device)[None, :] < text_length.to(pred.device)[:, None]
        return soft * mask

    def forward(self, pred, text_length, multiplier=1):
        from .engine import duration_to_alignment

        return duration_to_alignment(pred, text_length, multiplier)[0]


class Synthesizer(nn.Module):
    """Batched text -> wav graph, the reference's ExportModel.forward (export_model.py:40-63):
    duration predictor -> alignment -> pitch/energy predictor -> speech predictor."""

    def __init__(self, *, speech_predictor, pitch_energy_predictor, duration_predictor,
                 class_count=16, max_dur=50, coarse_multiplier=1):
        super().__init__()
        self.coarse_multiplier = coarse_multiplier
        self.speech_predictor = speech_predictor
        self.pitch_energy_predictor = pitch_energy_predictor
        self.duration_predictor = duration_predictor
        self.duration_processor = DurationProcessor(class_count, max_dur)

    @torch.no_grad()
    def forward(self, texts, text_lengths, speech_style, pe_style, duration_style, *,
                source_draws=None, prior=None, return_aux=False):
        dur_pred = self.duration_predictor(texts, text_lengths, duration_style)
        alignment = self.duration_processor(dur_pred, text_lengths)
        # export_model.py:42-45: the speech predictor gets the alignment at the fine frame rate
        fine = alignment if self.coarse_multiplier == 1 else self.duration_processor(
            dur_pred, text_lengths, multiplier=self.coarse_multiplier)
        pitch, energy = self.pitch_energy_predictor(texts, text_lengths, alignment, pe_style)
        voiced = (pitch > 20).float()
        pred = self.speech_predictor(texts, text_lengths, fine, pitch, energy, voiced,
                                     speech_style, pitch, source_draws=source_draws, prior=prior)
        if return_aux:
            return pred.audio, dict(dur_pred=dur_pred, alignment=alignment, pitch=pitch, energy=energy)
        return pred.audio


HOT_PATH_KEYS = ("speech_predictor", "duration_predictor", "pitch_energy_predictor", "speech_style_encoder",
                 "pe_style_encoder", "duration_style_encoder", "disc", "mrd0", "mrd1", "mrd2", "pitch_disc", "dur_disc")
ALL_KEYS = ("text_aligner", "duration_predictor", "pitch_energy_predictor", "speech_predictor",
            "disc", "mrd0", "mrd1", "mrd2", "speech_style_encoder", "pe_style_encoder",
            "duration_style_encoder", "pitch_disc", "dur_disc")


class ModelSet(dict):
    """dict with attribute access (the reference returns a ``Munch``)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def build_model(model_config, *, extra: Dict[str, nn.Module] | None = None) -> ModelSet:
    """Mirror of reference ``build_model`` (models.py:29-85).

    Returns the B200-native modules for the keys on the hot path.  Keys that
    are out of this engine's scope (aligner, discriminators, ...) can be
    supplied by the caller through ``extra`` (e.g. the reference's own modules
    when dropped into its train.py); they are passed through untouched.
    """
    nets = ModelSet()
    nets["duration_predictor"] = DurationPredictor(model_config)
    nets["pitch_energy_predictor"] = PitchEnergyPredictor(model_config)
    nets["speech_predictor"] = SpeechPredictor(model_config)
    from .style_encoder import MelStyleEncoder, PitchStyleEncoder

    se = model_config.style_encoder
    for key in ("speech_style_encoder", "duration_style_encoder"):  # models.py:49-54,62-67
        nets[key] = MelStyleEncoder(se.n_mels, model_config.style_dim, se.max_channels, se.skip_downsample)
    nets["pe_style_encoder"] = PitchStyleEncoder(se.n_mels, model_config.style_dim, se.max_channels,
                                                 se.skip_downsample,
                                                 coarse_multiplier=model_config.coarse_multiplier)
    # discriminators of the adversarial terms (models.py:43-46,78-83).  `text_aligner` is not re-implemented: pass it
    # through `extra` when needed.
    from .discriminator import ContextFreeDiscriminator, PitchDiscriminator, SpecDiscriminator

    nets["disc"] = ContextFreeDiscriminator()
    for key in ("mrd0", "mrd1", "mrd2"):
        nets[key] = SpecDiscriminator()
    nets["pitch_disc"] = PitchDiscriminator(dim_in=2, dim_hidden=64, kernel=21)
    nets["dur_disc"] = PitchDiscriminator(dim_in=1, dim_hidden=64, kernel=5)
    if extra:
        for k, v in extra.items():
            if k not in nets:
                nets[k] = v
    return nets
