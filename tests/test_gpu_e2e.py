"""End-to-end parity of the CUDA speech_predictor forward.

* against the committed golden fixtures made by the UNMODIFIED reference
  (tests/golden/*.npz), audio and every tap;
* against the CPU oracle at a BASELINE-sized utterance (T=258 tokens, ~10 s);
* the harmonic prior (RNG/phase-wrap sensitive) is checked on its own
  against the fp64 oracle, and injected (from the oracle) for everything downstream;
* size-independent properties at the full BASELINE config-2 batch (B=16).

Tolerance: the north-star bound is 1e-3 relative (fp32).  The vocoder / encoder convs run on
the tensor cores with a bf16 hi/lo operand split ("bf16x3", ~2^-17 relative per operand), so we
assert 5e-4 on the audio and 2e-4 on the intermediate taps (observed 1e-4 / 3e-5).
"""
import pytest
import torch

from oracle import speech_oracle as so
from stylish_tts_b200 import synth
import stylish_tts_b200 as st
from tests import util
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
AUDIO_TOL = 5e-4
TAP_TOL = 2e-4


def dev():
    return torch.device("cuda:0")


def run_oracle(sp, inp, taps=None, dtype=torch.float32, prior=None):
    sd = so.to_dtype(util.state_dict_of(sp), dtype)
    f = lambda t: t.to(dtype) if t.is_floating_point() else t
    draws = {k: f(v) for k, v in inp["draws"].items()}
    return so.speech_predictor(sd, inp["texts"], inp["text_lengths"], f(inp["alignment"]),
                               f(inp["pitch"]), f(inp["energy"]), f(inp["voiced"]),
                               f(inp["style"]), f(inp["denormal_pitch"]), draws, prior=prior,
                               taps=taps)


def run_gpu(sp, inp, *, prior=None, taps=None, draws=True):
    d = dev()
    sp = sp.to(d)
    c = lambda t: t.to(d)
    with torch.no_grad():
        out = sp(c(inp["texts"]), c(inp["text_lengths"]), c(inp["alignment"]), c(inp["pitch"]),
                 c(inp["energy"]), c(inp["voiced"]), c(inp["style"]), c(inp["denormal_pitch"]),
                 source_draws={k: c(v) for k, v in inp["draws"].items()} if draws else None,
                 prior=None if prior is None else (c(prior[0]), c(prior[1])), taps=taps)
    torch.cuda.synchronize()
    return out.audio.cpu()


@pytest.mark.parametrize("name", ["sp_b2_t16_ragged", "sp_b1_t12"])
def test_golden_reference_parity(name):
    sp, inp, gold, stride = util.golden_case(name)
    otaps = {}
    run_oracle(sp, inp, otaps)  # full-resolution prior to inject (the fixture is decimated)
    prior = (otaps["har_spec"], otaps["har_phase"])
    # the oracle's prior itself is pinned to the reference's by tests/test_oracle_golden.py
    taps = {}
    audio = run_gpu(sp, inp, prior=prior, taps=taps)
    assert audio.shape == gold["audio"].shape
    errs = {}
    for k, gv in gold.items():
        if not k.startswith("tap_") or k[4:] in ("prior_wave", "har_spec", "har_phase"):
            continue
        t = util.decimate(taps[k[4:]].cpu(), stride)
        assert t.shape == gv.shape, k
        errs[k] = rel_l2(t, gv)
    errs["audio"] = rel_l2(audio, gold["audio"])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < (AUDIO_TOL if k == "audio" else TAP_TOL), (k, v)


def test_harmonic_prior_vs_fp64_oracle():
    """Our excitation must be at least as close to the exact (fp64) value as the reference's own
    fp32 evaluation is (its phase accumulator loses ~0.01-0.1 rad, see csrc/source_stft.cu)."""
    sp, inp, gold, stride = util.golden_case("sp_b2_t16_ragged")
    t32, t64 = {}, {}
    run_oracle(sp, inp, t32)
    run_oracle(sp, inp, t64, dtype=torch.float64)
    taps = {}
    run_gpu(sp, inp, taps=taps)
    w = taps["prior_wave"].cpu().double()
    e_ref = rel_l2(t32["prior_wave"].double(), t64["prior_wave"])
    e_gpu = rel_l2(w, t64["prior_wave"])
    print(f"prior wave: fp32-oracle vs fp64 {e_ref:.2e}; gpu vs fp64 {e_gpu:.2e}")
    assert e_gpu <= max(e_ref, 1e-5) * 1.5
    # spectrum of the prior: magnitude tight, phase wrap-aware
    e_spec = rel_l2(taps["har_spec"].cpu().double(), t64["har_spec"])
    e_spec_ref = rel_l2(t32["har_spec"].double(), t64["har_spec"])
    assert e_spec <= max(e_spec_ref, 1e-5) * 1.5 + 1e-5
    e_ph = util.wrap_aware_phase_err(taps["har_phase"].cpu().double(), t64["har_phase"])
    e_ph_ref = util.wrap_aware_phase_err(t32["har_phase"].double(), t64["har_phase"])
    print(f"har_spec {e_spec:.2e} (ref {e_spec_ref:.2e}); har_phase wrap-aware {e_ph:.2e} (ref {e_ph_ref:.2e})")
    assert e_ph <= max(e_ph_ref, 1e-5) * 1.5 + 1e-5


def test_stft_istft_kernels_on_oracle_inputs():
    """stft / istft-head kernels alone, fed the oracle's exact inputs."""
    from stylish_tts_b200 import _lib as L
    sp, inp, gold, stride = util.golden_case("sp_b1_t12")
    taps = {}
    ref_audio = run_oracle(sp, inp, taps)
    d = dev()
    eng = sp.engine()
    P = eng.packed(d)
    wave = taps["prior_wave"].to(d).contiguous()
    B, Lw = wave.shape
    S = Lw // 4
    spec = torch.empty(B, 32, S, device=d)
    ph = torch.empty_like(spec)
    L.call("sty_stft_fwd", wave.data_ptr(), P.stft_f_re.data_ptr(), P.stft_f_im.data_ptr(),
           spec.data_ptr(), ph.data_ptr(), B, Lw, 64, 4, 32, L.stream_ptr())
    assert rel_l2(spec, taps["har_spec"]) < 2e-5
    assert util.wrap_aware_phase_err(ph.cpu(), taps["har_phase"]) < 1e-4
    la, re, im = (taps[k].to(d).contiguous() for k in ("logamp", "real", "imag"))
    audio = torch.empty(B, 1, S * 4, device=d)
    L.call("sty_istft_head_fwd", la.data_ptr(), la.stride(0), re.data_ptr(), im.data_ptr(),
           re.stride(0), P.stft_b_re.data_ptr(), P.stft_b_im.data_ptr(), audio.data_ptr(), B, S, 32,
           64, 4, L.stream_ptr())
    assert rel_l2(audio, ref_audio) < 2e-5


def test_baseline_sized_utterance_vs_oracle():
    """T=258 tokens (~10 s, S~60k steps): config-2 shapes at B=2, ragged lengths."""
    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, 21)
    inp = synth.speech_inputs(2, 258, seed=9, ragged=True)
    otaps = {}
    ref = run_oracle(sp, inp, otaps)
    taps = {}
    audio = run_gpu(sp, inp, prior=(otaps["har_spec"], otaps["har_phase"]), taps=taps)
    errs = {k: rel_l2(taps[k], otaps[k]) for k in ("text_encoding", "decoder", "conformer",
                                                   "logamp_prior", "phase_prior", "upsampled",
                                                   "logamp", "real", "imag")}
    errs["audio"] = rel_l2(audio, ref)
    print({k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < (AUDIO_TOL if k == "audio" else TAP_TOL), (k, v)


def test_config2_full_batch_audio_vs_oracle():
    """BASELINE config 2 itself (B=16, T=258, F=803): audio of every utterance against the CPU oracle."""
    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, 21)
    inp = synth.speech_inputs(16, 258, seed=3, ragged=True)
    otaps = {}
    ref = run_oracle(sp, inp, otaps)
    audio = run_gpu(sp, inp, prior=(otaps["har_spec"], otaps["har_phase"]))
    assert audio.shape == ref.shape == (16, 1, inp["alignment"].shape[2] * 300)
    per_row = [rel_l2(audio[b], ref[b]) for b in range(16)]
    print("config 2 audio rel-L2 per utterance:", [f"{e:.1e}" for e in per_row])
    assert max(per_row) < AUDIO_TOL, per_row
    assert rel_l2(audio, ref) < AUDIO_TOL


def test_full_batch_properties():
    """BASELINE config 2 (B=16, T=258): size-independent properties.
    (1) utterances are independent: row b of the batch == the same utterance run alone;
    (2) padded tokens and padded positions do not influence the audio;
    (3) the output is finite, bounded by tanh, and of the expected length."""
    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, 5)
    inp = synth.speech_inputs(16, 258, seed=3, ragged=True)
    full = run_gpu(sp, inp)
    Fr = inp["alignment"].shape[2]
    assert full.shape == (16, 1, Fr * 300)
    assert torch.isfinite(full).all() and float(full.abs().max()) <= 1.0
    for b in (0, 7, 15):
        one = {k: (v[b:b + 1] if torch.is_tensor(v) else {kk: vv[b:b + 1] for kk, vv in v.items()})
               for k, v in inp.items()}
        alone = run_gpu(sp, one)
        # not bit-exact: GRN statistics are accumulated with atomics (order varies with the grid)
        assert rel_l2(full[b:b + 1], alone) < 1e-4, b
    inp2 = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in inp.items()}
    for b in range(16):
        n = int(inp["text_lengths"][b])
        inp2["texts"][b, n:] = 77  # garbage in the padded region
    again = run_gpu(sp, inp2)
    assert rel_l2(again, full) < 1e-4
