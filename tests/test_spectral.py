"""Mel front-end + STFT / phase losses.

CPU: the oracle restatement and the product's host-side constants against the golden fixture made
by the UNMODIFIED reference (tests/golden/make_spectral_golden.py).
GPU: the fused FFT kernels, through the C ABI, against the golden fixture and the oracle, forward
and backward (gradient w.r.t. the predicted audio), plus size-independent properties at the full
config-3 size (B=32, 10 s).
"""
import math

import pytest
import torch

from oracle import spectral_oracle as so
from stylish_tts_b200 import spectral as sp
from tests.golden.make_spectral_golden import MEL_MEAN, MEL_STD, SAMPLE_RATE, W_MEL, W_PHASE, spectral_inputs
from tests.util import load_golden, rel_l2


@pytest.fixture(scope="module")
def gold():
    return load_golden("spectral")


# ----------------------------------------------------------------------------- CPU
def test_oracle_filterbank_matches_torchaudio_golden(gold):
    assert torch.equal(so.mel_fbank(257, 80, SAMPLE_RATE), gold["fb_257_80"])
    assert torch.equal(so.mel_fbank(1025, 128, SAMPLE_RATE), gold["fb_1025_128"])
    # the product's host-side table (same arithmetic, no torchaudio import)
    assert torch.equal(sp.melscale_fbanks(257, 0.0, 12000.0, 80, SAMPLE_RATE), gold["fb_257_80"])
    assert torch.equal(sp.melscale_fbanks(1025, 0.0, 12000.0, 128, SAMPLE_RATE), gold["fb_1025_128"])


def test_sparse_filterbank_roundtrip():
    plan = sp.SpectrogramPlan(n_fft=512, hop=128, win_length=512, n_mels=128, sample_rate=SAMPLE_RATE, power=1)
    h = plan._host
    fb = torch.zeros_like(plan.fb)
    for m in range(128):
        s, n, o = int(h["fb_start"][m]), int(h["fb_len"][m]), int(h["fb_off"][m])
        fb[s:s + n, m] = h["fb_w"][o:o + n]
    assert torch.equal(fb, plan.fb)
    fbt = torch.zeros_like(plan.fb)
    for k in range(257):
        for e in range(int(h["fbt_ptr"][k]), int(h["fbt_ptr"][k + 1])):
            fbt[k, int(h["fbt_mel"][e])] = h["fbt_w"][e]
    assert torch.equal(fbt, plan.fb)


def test_centered_window_matches_torch_stft_padding():
    w = sp.centered_window(1200, 2048)
    assert w.shape == (2048,) and float(w[:424].abs().sum()) == 0.0 and float(w[424 + 1200:].abs().sum()) == 0.0
    assert torch.equal(w[424:424 + 1200], torch.hann_window(1200))


def test_oracle_front_end_matches_reference_golden(gold):
    target, pred = spectral_inputs()
    for name, (n_fft, win) in {"mel": (512, 512), "style_mel": (2048, 1200)}.items():
        raw = so.mel_spectrogram(target, n_fft=n_fft, win=win, hop=300, n_mels=80, sample_rate=SAMPLE_RATE)
        assert rel_l2(raw, gold[name + "_raw"]) < 1e-6
        mel = so.calculate_mel(target, n_fft=n_fft, win=win, hop=300, n_mels=80, sample_rate=SAMPLE_RATE,
                               mean=MEL_MEAN, std=MEL_STD)
        assert mel.shape == gold[name].shape
        assert rel_l2(mel, gold[name]) < 1e-6
    e = so.log_energy(gold["mel"], MEL_MEAN, MEL_STD)
    assert rel_l2(e, gold["energy"]) < 1e-6


def test_oracle_losses_and_gradient_match_reference_golden(gold):
    target, pred = spectral_inputs()
    pred.requires_grad_(True)
    for r, res in enumerate(so.RESOLUTIONS):
        tm, tp, tf = so.multi_spectrogram_single(target, res, SAMPLE_RATE)
        assert rel_l2(tm, gold[f"t_spec{r}"]) < 1e-6
        assert rel_l2(tf, gold[f"t_fft{r}"]) < 1e-6
        assert torch.equal(tp, gold[f"t_phase{r}"])
    ls = so.acoustic_spectral_losses(target, pred, SAMPLE_RATE)
    assert abs(float(ls["mel"].detach()) - float(gold["mel_loss"])) < 1e-5 * float(gold["mel_loss"])
    assert abs(float(ls["multi_phase"].detach()) - float(gold["phase_loss"])) < 1e-5 * float(gold["phase_loss"])
    total = so.backwards_total(ls, dict(mel=W_MEL, multi_phase=W_PHASE))
    total.backward()
    assert rel_l2(pred.grad, gold["d_pred"]) < 1e-4


def test_cpu_tensors_are_rejected_loudly():
    to_mel = sp.MelSpectrogram(n_mels=80, n_fft=512, win_length=512, hop_length=300, sample_rate=SAMPLE_RATE)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        to_mel(torch.zeros(1, 4000))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sp.MultiResolutionSTFTLoss()(target_list=[torch.zeros(4)], pred_list=[torch.zeros(4)])


# ----------------------------------------------------------------------------- GPU
def _phase_err(a, b, mag, floor=2e-3):
    """wrap-aware phase error on bins safely above the reference's 1e-3 magnitude gate"""
    keep = mag > floor
    d = (torch.polar(torch.ones_like(a), a) - torch.polar(torch.ones_like(b), b)).abs()
    return float((d * keep).sum() / keep.sum())


@pytest.mark.gpu
def test_mel_front_end_vs_reference_golden(gold):
    target, _ = spectral_inputs()
    x = target.cuda()
    for name, (n_fft, win) in {"mel": (512, 512), "style_mel": (2048, 1200)}.items():
        to_mel = sp.MelSpectrogram(n_mels=80, n_fft=n_fft, win_length=win, hop_length=300, sample_rate=SAMPLE_RATE)
        raw = to_mel(x)
        assert raw.shape == gold[name + "_raw"].shape
        assert rel_l2(raw, gold[name + "_raw"]) < 1e-5
        mel, length = sp.calculate_mel(x, to_mel, MEL_MEAN, MEL_STD)
        assert mel.shape == gold[name].shape and int(length[0]) == mel.shape[2]
        assert rel_l2(mel, gold[name]) < 1e-5
        if name == "mel":
            e = sp.mel_energy(mel, MEL_MEAN, MEL_STD)
            assert rel_l2(e, gold["energy"]) < 1e-5
            ln = sp.log_norm(mel.unsqueeze(1), MEL_MEAN, MEL_STD)
            assert rel_l2(ln.squeeze(1), torch.exp(gold["energy"]) - 1e-9) < 1e-5


@pytest.mark.gpu
def test_multi_spectrogram_and_losses_vs_reference_golden(gold):
    target, pred = spectral_inputs()
    t, p = target.cuda(), pred.cuda().requires_grad_(True)
    ms = sp.MultiSpectrogram(sample_rate=SAMPLE_RATE)
    t_spec, p_spec, t_ph, p_ph, t_fft, p_fft = ms(target=t, pred=p)
    for r in range(3):
        assert t_spec[r].shape == gold[f"t_spec{r}"].shape and not t_spec[r].requires_grad
        assert rel_l2(t_spec[r], gold[f"t_spec{r}"]) < 1e-5
        assert rel_l2(p_spec[r], gold[f"p_spec{r}"]) < 1e-5
        assert rel_l2(t_fft[r], gold[f"t_fft{r}"]) < 1e-5
        assert t_ph[r].shape == gold[f"t_phase{r}"].shape
        assert _phase_err(t_ph[r].cpu(), gold[f"t_phase{r}"], gold[f"t_fft{r}"][:, 0]) < 1e-4

    class Log:
        def add_loss(self, k, v):
            self.k, self.v = k, v

    log = Log()
    mel_loss = sp.MultiResolutionSTFTLoss(sample_rate=SAMPLE_RATE)(target_list=t_spec, pred_list=p_spec, log=log)
    assert log.k == "mel" and log.v is mel_loss
    ph_loss = sp.multi_phase_loss(p_ph, t_ph, 512)
    assert abs(float(mel_loss) - float(gold["mel_loss"])) < 1e-5 * float(gold["mel_loss"])
    assert abs(float(ph_loss) - float(gold["phase_loss"])) < 1e-3 * float(gold["phase_loss"])
    total = W_MEL * mel_loss / (mel_loss.detach() + 1e-9) + W_PHASE * ph_loss / (ph_loss.detach() + 1e-9)
    total.backward()
    # the phase-loss gradient is a sum of sign() terms: a phase within float rounding of a wrap /
    # the 1e-3 gate flips single terms, so the bound is looser than for the smooth mel part
    assert rel_l2(p.grad, gold["d_pred"]) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["mel", "mag", "phase"])
def test_spectrogram_backward_vs_oracle_autograd(which):
    """smooth heads only (no sign() terms): d/d audio of <random cotangent, output>."""
    torch.manual_seed(3)
    _, pred = spectral_inputs(B=2, L=5000, seed=5)
    res = so.RESOLUTIONS[1]
    p_ref = pred.double().requires_grad_(True)
    mel, ph, mag = so.multi_spectrogram_single(p_ref, res, SAMPLE_RATE)
    out_ref = dict(mel=mel[:, 0], mag=mag[:, 0], phase=ph)[which]
    ct = torch.randn(out_ref.shape, dtype=torch.float64)
    if which == "phase":  # keep away from the magnitude gate and tiny bins (1/|X|^2 conditioning)
        ct = ct * (mag[:, 0].detach() > 0.05)
    (out_ref * ct).sum().backward()
    plan = sp.SpectrogramPlan(n_fft=res[0], hop=res[1], win_length=res[2], n_mels=128, sample_rate=SAMPLE_RATE,
                              power=1, mel_mode=sp.MEL_LOG1P)
    p = pred.cuda().requires_grad_(True)
    g_mag, g_ph, g_mel = sp._SpectrogramFn.apply(p, plan, True, True, True)
    out = dict(mel=g_mel, mag=g_mag, phase=g_ph)[which]
    (out * ct.float().cuda()).sum().backward()
    assert rel_l2(p.grad, p_ref.grad) < 2e-4


@pytest.mark.gpu
def test_full_size_properties():
    """config 3 size (B=32, 10 s): linearity of |X| in the signal gain, zero loss and zero gradient
    on identical inputs, Parseval energy check of the power spectrum."""
    g = torch.Generator().manual_seed(2)
    B, Ls = 32, 240000
    x = (0.1 * torch.randn(B, Ls, generator=g)).cuda()
    plan = sp.SpectrogramPlan(n_fft=1024, hop=256, win_length=1024, n_mels=0, sample_rate=SAMPLE_RATE, power=2)
    p1 = plan.forward(x, want_mag=True, want_mel=False)[0]
    p2 = plan.forward(2.0 * x, want_mag=True, want_mel=False)[0]
    assert rel_l2(p2, 4.0 * p1) < 1e-6
    # Parseval per frame: sum_k c_k |X_k|^2 = N * sum_n (w x)^2 with c = 1 at DC/Nyquist, 2 elsewhere
    c = torch.full((513,), 2.0, device="cuda")
    c[0] = c[-1] = 1.0
    lhs = (p1[:, :, 10] * c[None]).sum(1)
    w = torch.hann_window(1024, device="cuda")
    seg = x[:, 10 * 256 - 512:10 * 256 + 512] * w
    assert rel_l2(lhs, 1024.0 * (seg * seg).sum(1)) < 1e-4
    ms = sp.MultiSpectrogram(sample_rate=SAMPLE_RATE)
    xp = x.clone().requires_grad_(True)
    t_spec, p_spec, t_ph, p_ph, _, _ = ms(target=x, pred=xp)
    assert [tuple(s.shape) for s in t_spec] == [(B, 1, 128, 1876), (B, 1, 128, 938), (B, 1, 128, 469)]
    loss = sp.MultiResolutionSTFTLoss()(target_list=t_spec, pred_list=p_spec) + sp.multi_phase_loss(p_ph, t_ph)
    assert float(loss) == 0.0
    loss.backward()
    assert float(xp.grad.abs().max()) == 0.0
