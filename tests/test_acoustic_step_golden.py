"""The composed acoustic training step against a golden made by the UNMODIFIED reference's
own ``AcousticStep`` + ``LossLog.backwards_loss`` (tests/golden/make_acoustic_step_golden.py;
stage_type.py:61-193,346-366, loss_log.py:82-94).

CPU : the oracle pieces (mel front-end, energy, alignment, style encoder in train(), speech predictor with
      batch-stat BN, MultiSpectrogram, mel + multi-phase loss, backwards_loss normalisation) composed the way
      AcousticStep wires them reproduce the reference's losses, d(total)/d(audio) and the gradients of BOTH
      trained modules.
GPU : ``train_step.acoustic_step`` on the CUDA kernels (through the C ABI) reproduces the same numbers — loss
      values, energy, style vector, audio, the gradient w.r.t. the audio, and per-parameter gradient norms and
      probe dots of ``speech_predictor`` and ``speech_style_encoder`` through the WHOLE step.
The harmonic prior is phase-chaotic in fp32, so the GPU arm gets the fp32 oracle's prior injected.
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import spectral_oracle as sp_o
from oracle import speech_oracle as so
from oracle import style_oracle as st_o
import stylish_tts_b200 as st
from tests import util
from tests.golden.make_acoustic_step_golden import make_batch, probe, seeded_nets
from tests.util import rel_l2

MEAN, STD = -4.0, 4.0  # NormalizationStats defaults (train_context.py:47-53)


def load_gold():
    z = np.load(util.GOLDEN_DIR + "/acoustic_step.npz")
    return {k: z[k] for k in z.files}


def front_end_oracle(batch, mc):
    mel = sp_o.calculate_mel(batch["audio_gt"], n_fft=mc.n_fft, win=mc.win_length, hop=mc.hop_length,
                             n_mels=mc.n_mels, sample_rate=mc.sample_rate, mean=MEAN, std=STD)
    se = mc.style_encoder
    style_mel = sp_o.calculate_mel(batch["audio_gt"], n_fft=se.n_fft, win=se.win_length, hop=se.hop_length,
                                   n_mels=se.n_mels, sample_rate=mc.sample_rate, mean=MEAN, std=STD)
    energy = sp_o.log_energy(mel, MEAN, STD)
    alignment = so.duration_to_alignment(batch["alignment"][:, 0, :].float())
    return mel, style_mel, energy, alignment


def check_grads(gold, key, grads, t_norm, t_dot):
    names = [str(n) for n in gold[key + "_names"]]
    scale = float(np.sqrt((gold[key + "_norms"] ** 2).sum()))
    worst = 0.0
    for n, norm, dot in zip(names, gold[key + "_norms"], gold[key + "_dots"]):
        assert n in grads and grads[n] is not None, f"no gradient for {key}.{n}"
        g = grads[n].detach().float().cpu()
        e_n = abs(float(g.norm()) - norm)
        e_d = abs(float((g * probe(n, g.shape)).sum()) - dot)
        assert e_n <= t_norm * norm + 1e-5 * scale, (key, n, float(g.norm()), norm)
        assert e_d <= t_dot * norm + 1e-5 * scale, (key, n, e_d, norm)
        worst = max(worst, e_d / (norm + 1e-5 * scale))
    return worst


def test_oracle_composition_matches_reference_acoustic_step():
    gold = load_gold()
    mc = st.default_model_config()
    sp, se = seeded_nets()
    batch, draws = make_batch()
    sd_sp = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v.clone())
             for k, v in sp.state_dict().items()}
    sd_se = {k: (v.detach().clone().requires_grad_(True) if k.endswith(("weight_orig", "bias", "unshared.weight"))
                 else v.detach().clone()) for k, v in se.state_dict().items()}
    with torch.no_grad():
        mel, style_mel, energy, alignment = front_end_oracle(batch, mc)
    assert rel_l2(mel, torch.from_numpy(gold["mel_target"])) < 1e-5
    assert rel_l2(energy, torch.from_numpy(gold["energy"])) < 1e-5
    pitch = batch["pitch"]
    voiced = (pitch > 20).float()
    style = st_o.mel_style_encoder(sd_se, style_mel.unsqueeze(1), training=True)
    style.retain_grad()
    audio = so.speech_predictor(sd_sp, batch["text"], batch["text_length"], alignment, pitch, energy, voiced,
                                style, pitch, draws, bn_training=True)
    audio.retain_grad()
    losses = sp_o.acoustic_spectral_losses(batch["audio_gt"], audio.squeeze(1), mc.sample_rate)
    total = sp_o.backwards_total(losses, dict(mel=gold["weights"][0], multi_phase=gold["weights"][1]))
    total.backward()
    assert rel_l2(style, torch.from_numpy(gold["style"])) < 1e-5
    assert rel_l2(audio, torch.from_numpy(gold["audio"])) < 1e-5
    assert abs(float(losses["mel"]) - gold["mel_loss"]) < 1e-5 * gold["mel_loss"]
    assert abs(float(losses["multi_phase"]) - gold["phase_loss"]) < 1e-5 * gold["phase_loss"]
    assert abs(float(total) - gold["backward_scalar"]) < 1e-5
    # the phase term is piecewise (anti-wrapping round(), |X| > 1e-3 mask): an audio difference of 1e-6 moves a few
    # bins across a discontinuity, so two fp32 evaluations of the SAME formula agree to ~1.5e-4, not 1e-6
    assert rel_l2(audio.grad, torch.from_numpy(gold["d_audio"])) < 5e-4
    assert rel_l2(style.grad, torch.from_numpy(gold["d_style"])) < 2e-3
    g_sp = {k: v.grad for k, v in sd_sp.items() if v.is_floating_point() and v.grad is not None}
    g_se = {k: v.grad for k, v in sd_se.items() if v.requires_grad}
    # two fp32 evaluations of one graph in different summation orders
    print("sp worst", check_grads(gold, "sp", g_sp, 2e-3, 4e-3))
    print("se worst", check_grads(gold, "se", g_se, 2e-3, 4e-3))


@pytest.mark.gpu
@pytest.mark.parametrize("tensor_cores", [True, False])
def test_gpu_acoustic_step_matches_reference_golden(tensor_cores, monkeypatch):
    from stylish_tts_b200 import engine as E
    from stylish_tts_b200 import train_step as ts

    monkeypatch.setattr(E, "USE_UMMA", tensor_cores)
    gold = load_gold()
    mc = st.default_model_config()
    sp, se = seeded_nets()
    batch, draws = make_batch()
    # the fp32 oracle's harmonic prior (bit-compatible with the reference's), injected
    with torch.no_grad():
        sd32 = util.state_dict_of(sp)
        pitch = batch["pitch"]
        har_spec, har_phase, _ = so.harmonic_prior(sd32, "generator.basegen", pitch, (pitch > 20).float(), draws)
    dev = torch.device("cuda:0")
    sp, se = sp.to(dev).train(), se.to(dev).train()
    sp.regularisers = False  # the golden's pinned configuration: batch-stat BN, regularisers off
    nets = SimpleNamespace(speech_predictor=sp, speech_style_encoder=se)
    fe = ts.FrontEnd(mc, MEAN, STD)
    b = SimpleNamespace(**{k: v.to(dev) for k, v in batch.items()})
    kw = dict(w_mel=float(gold["weights"][0]), w_phase=float(gold["weights"][1]),
              prior=(har_spec.to(dev), har_phase.to(dev)))
    out = ts.acoustic_step(b, nets, fe, **kw)
    errs = dict(
        mel_target=rel_l2(out.mel_target, torch.from_numpy(gold["mel_target"])),
        energy=rel_l2(out.energy, torch.from_numpy(gold["energy"])),
        style=rel_l2(out.style, torch.from_numpy(gold["style"])),
        audio=rel_l2(out.pred.audio, torch.from_numpy(gold["audio"])),
        mel_loss=abs(float(out.mel) - gold["mel_loss"]) / gold["mel_loss"],
        phase_loss=abs(float(out.multi_phase) - gold["phase_loss"]) / gold["phase_loss"],
        total=abs(float(out.total) - gold["backward_scalar"]) / gold["backward_scalar"])
    print("acoustic step forward vs reference golden:", errs)
    assert errs["mel_target"] < 1e-4 and errs["energy"] < 1e-4
    assert errs["style"] < 5e-4 and errs["audio"] < 5e-4
    assert errs["mel_loss"] < 1e-3 and errs["phase_loss"] < 1e-3 and errs["total"] < 1e-5

    # ---- backward, factored by the chain rule so that each factor is compared on IDENTICAL inputs.
    # The multi-phase term is piecewise linear (|wrap(d)|: kinks at d = 0 and +-pi, and the |X| > 1e-3 mask), so
    # d(total)/d(audio) changes discontinuously with the audio: two evaluations whose audio differs by 2.6e-5
    # (ours vs the reference's) flip a few 1e-5 of the bins and differ by ~sqrt(that) = 3e-3..7e-3 in the
    # gradient; two fp32 CPU evaluations 1e-6 apart differ by 1.5e-4 (CPU test above).
    # (1) losses: d(total)/d(audio) evaluated AT the reference's audio
    from stylish_tts_b200.optim import acoustic_losses
    a_ref = torch.from_numpy(gold["audio"]).to(dev).squeeze(1).requires_grad_(True)
    tot, _, _ = acoustic_losses(a_ref, b.audio_gt, fe.multi_spectrogram, fe.stft_loss, w_mel=kw["w_mel"],
                                w_phase=kw["w_phase"])
    tot.backward()
    e_loss = rel_l2(a_ref.grad.unsqueeze(1), torch.from_numpy(gold["d_audio"]))
    print("d(total)/d(audio) at the reference's audio:", e_loss)
    assert e_loss < 1e-3, e_loss  # own FFT vs pocketfft: spectra ~1e-6 apart, same kink mechanism
    # (2) both modules: the reference's d(total)/d(audio) pulled back through speech_predictor AND
    # speech_style_encoder by the step's own graph
    out.pred.audio.backward(torch.from_numpy(gold["d_audio"]).to(dev), retain_graph=True, inputs=(
        [p for p in sp.parameters()] + [p for p in se.parameters()] + [out.style]))
    torch.cuda.synchronize()
    e_style = rel_l2(out.style.grad, torch.from_numpy(gold["d_style"]))
    print("d(total)/d(style) with the reference's audio gradient:", e_style)
    assert e_style < 1e-3, e_style
    # same bounds as the CPU test above (two fp32 evaluations of one graph: worst probe dot 5e-4 of the norm there,
    # in the text encoder, whose gradient the decoder's InstanceNorms amplify — tests/test_train_step.py)
    t = (2e-3, 4e-3)
    print("sp worst", check_grads(gold, "sp", {n: p.grad for n, p in sp.named_parameters()}, *t))
    print("se worst", check_grads(gold, "se", {n: p.grad for n, p in se.named_parameters()}, *t))
    # (3) the composed backward of the step itself: same numbers up to the conditioning discussed above
    for p in list(sp.parameters()) + list(se.parameters()):
        p.grad = None
    out.pred.audio.retain_grad()
    out.total.backward()
    torch.cuda.synchronize()
    e_comp = rel_l2(out.pred.audio.grad, torch.from_numpy(gold["d_audio"]))
    print("composed d(total)/d(audio):", e_comp)
    assert e_comp < 2e-2, e_comp
    print("composed sp worst", check_grads(gold, "sp", {n: p.grad for n, p in sp.named_parameters()}, 2e-2, 4e-2))
    print("composed se worst", check_grads(gold, "se", {n: p.grad for n, p in se.named_parameters()}, 2e-2, 4e-2))
