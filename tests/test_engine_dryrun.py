"""Host-logic test of the engine without a GPU: every C-ABI call is replaced by a
recorder, tensors stay on the CPU, so shapes / strides / buffer plumbing / call order of
the whole speech_predictor forward are exercised (no arithmetic is performed or checked)."""
import collections

import pytest
import torch

import stylish_tts_b200 as st
from stylish_tts_b200 import _lib as L
from stylish_tts_b200 import engine as E
from stylish_tts_b200 import synth


@pytest.fixture()
def recorder(monkeypatch):
    calls = []

    def fake_call(name, *args):
        assert name in L._SIGNATURES, name
        assert len(args) == len(L._SIGNATURES[name]), (name, len(args))
        calls.append(name)

    monkeypatch.setattr(L, "call", fake_call)
    monkeypatch.setattr(L, "_req", lambda *a, **k: None)
    monkeypatch.setattr(L, "stream_ptr", lambda: 0)
    monkeypatch.setattr(L, "load", lambda: None)
    return calls


def test_forward_plumbing(recorder, monkeypatch):
    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, 0)
    inp = synth.speech_inputs(2, 16, seed=1, ragged=True)
    eng = E.SpeechEngine(sp)
    P = eng.packed(torch.device("cpu"))
    assert P.fc_rows == sum(v.shape[0] for k, v in sp.state_dict().items() if k.endswith(".fc.weight"))

    # run the body of forward() with CPU tensors (bypassing the device check)
    B = 2
    h = torch.empty((B, P.fc_rows))
    taps = {}
    mu, _, _ = eng.text_encoder(P, inp["texts"], inp["text_lengths"], taps)
    assert mu.shape == (2, 128, 16)
    mel = eng.decoder(P, mu, inp["alignment"], inp["pitch"], inp["energy"], inp["voiced"], h, taps)
    Fr = inp["alignment"].shape[2]
    assert mel.shape == (2, 128, Fr)
    audio = eng.generator(P, mel, h, inp["denormal_pitch"], inp["voiced"], inp["draws"]["noise"],
                          taps=taps)
    assert audio.shape == (2, 1, Fr * 300)
    cnt = collections.Counter(recorder)
    # 8 encoder layers x (qkv, o, ffn1, ffn2) + prenet 3 + proj + proj_m ...
    assert cnt["sty_attention_fwd"] == 8 + 1
    assert cnt["sty_source_fwd"] == 1 and cnt["sty_stft_pitched_fwd"] == 1 and cnt["sty_istft_head_pitched_fwd"] == 1
    # C<=64 ConvNeXt fronts are fused into the pointwise conv (tensor-core path) when T >= 128
    assert cnt["sty_dwconv_ln_fwd"] == 5 + 1
    # the 9 output-rate ConvNeXt blocks are one fused call each (pass 1 + GRN scale + pass 2 inside the library)
    assert cnt["sty_convnext_fused_fwd"] == 9
    assert cnt["sty_grn_scale_fwd"] == 16 - 9
    # decoder AdaINs keep the two-pass statistics kernel; the 12 S-rate AdaINs of the two generator blocks
    # take their statistics from the producing conv's epilogue (out_sum / out_sumsq -> moments_affine)
    assert cnt["sty_instnorm_affine_fwd"] == 2 * 5
    assert cnt["sty_moments_affine_fwd"] == 2 * 6
    for k in ("prenet", "dec_encode", "conformer", "logamp_prior", "upsampled", "real", "imag"):
        assert k in taps


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        L.load()


def test_cpu_inputs_rejected():
    sp = st.build_model(st.default_model_config()).speech_predictor
    inp = synth.speech_inputs(1, 12, seed=1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        sp(inp["texts"], inp["text_lengths"], inp["alignment"], inp["pitch"], inp["energy"],
           inp["voiced"], inp["style"], inp["denormal_pitch"])
