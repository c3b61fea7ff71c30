"""Run the UNMODIFIED reference modules on given inputs (build container only).

TEST INFRASTRUCTURE.  Feeds the harmonic source's random draws as inputs by
substituting ``torch.rand/randn/randn_like`` inside the reference's generator
module for the duration of one forward, and collects tap
tensors with forward hooks.
"""
from __future__ import annotations

import contextlib

import torch

from . import ref_loader


class _TorchProxy:
    def __init__(self, real, draws):
        self._real = real
        self._draws = draws

    def __getattr__(self, k):
        return getattr(self._real, k)

    def rand(self, *shape, **kw):
        r = self._draws["rand_ini"]
        assert tuple(shape) == tuple(r.shape), (shape, r.shape)
        return r.clone()

    def randn(self, size, **kw):
        n = self._draws["noise"]
        assert tuple(size) == tuple(n.shape), (size, n.shape)
        return n.clone().to(kw.get("dtype", n.dtype))

    def randn_like(self, t, **kw):
        return self._real.zeros_like(t)


@contextlib.contextmanager
def injected_draws(draws):
    ref_loader.load()
    import stylish_tts.train.models.generator as gen_mod

    real = gen_mod.torch
    gen_mod.torch = _TorchProxy(real, draws)
    try:
        yield
    finally:
        gen_mod.torch = real


def speech_predictor_forward(ref_sp, inp, taps=None):
    """ref_sp: reference SpeechPredictor (eval).  Returns audio (B,1,L)."""
    hooks = []
    if taps is not None:
        def tap(name, mod, fn=lambda o: o):
            hooks.append(mod.register_forward_hook(
                lambda m, i, o, name=name, fn=fn: taps.__setitem__(name, fn(o).detach())))
        tap("text_encoding", ref_sp.text_encoder, lambda o: o[0])
        tap("prenet", ref_sp.text_encoder.prenet)
        tap("dec_encode", ref_sp.decoder.encode)
        tap("decoder", ref_sp.decoder, lambda o: o[0])
        tap("conformer", ref_sp.generator.amp_conformer, lambda o: o.transpose(1, 2))
        bg = ref_sp.generator.basegen
        tap("logamp_prior", bg.amp_prior_block)
        tap("phase_prior", bg.phase_prior_block)
        tap("amp_convnext", bg.amp_convnext[-1])
        tap("upsampled", bg.upblocks[-1])
        tap("logamp", bg.amp_output_conv)
        tap("real", bg.phase_output_real_conv)
        tap("imag", bg.phase_output_imag_conv)
        tap("prior_wave", bg.m_source, lambda o: o[0].squeeze(2))
        hooks.append(bg.amp_prior_conv.register_forward_hook(
            lambda m, i, o: taps.__setitem__("har_spec", i[0].detach())))
        hooks.append(bg.phase_prior_conv.register_forward_hook(
            lambda m, i, o: taps.__setitem__("har_phase", i[0].detach())))
    try:
        with torch.no_grad(), injected_draws(inp["draws"]):
            out = ref_sp(inp["texts"], inp["text_lengths"], inp["alignment"], inp["pitch"],
                         inp["energy"], inp["voiced"], inp["style"], inp["denormal_pitch"])
    finally:
        for h in hooks:
            h.remove()
    return out.audio
