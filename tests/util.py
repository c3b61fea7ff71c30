"""Shared helpers for the tests (golden loading, seeded model/inputs)."""
import os

import numpy as np
import torch

import stylish_tts_b200 as st
from stylish_tts_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def golden_case(name):
    """-> (module with the fixture's seeded weights, inputs, golden dict, tap stride)"""
    gold = load_golden(name)
    b, t, iseed, wseed, ragged, stride = [int(v) for v in gold["meta"]]
    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, wseed)
    inp = synth.speech_inputs(b, t, seed=iseed, ragged=bool(ragged))
    return sp, inp, gold, stride


def state_dict_of(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def decimate(t, stride):
    return t[..., ::stride] if t.shape[-1] > 2000 else t


def wrap_aware_phase_err(a, b):
    """mean |e^{ia} - e^{ib}| — phase comparison that ignores 2*pi wraps (F7)."""
    return float((torch.polar(torch.ones_like(a), a) - torch.polar(torch.ones_like(b), b))
                 .abs().mean())
