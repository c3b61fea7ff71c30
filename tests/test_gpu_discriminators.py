"""Spectrogram discriminators and adversarial losses on the device through the C ABI:

* ``SpecDiscriminator`` (drop-in for mrd0-2, discriminator.py:13-69) against the golden outputs of the UNMODIFIED
  reference (tests/golden/discriminators.npz) and, at a size that takes the tensor-core path (frames >= 64 after
  three stride-2 layers), against the fp64 oracle — scores, input gradient and every parameter gradient;
* LSGAN + TPRLS generator / discriminator losses (losses.py:251-278,339-363) incl. the radix-select median, against
  the oracle formulas (which tests/test_disc_oracle.py pins to the reference);
* the device-resident moving average / learning-rate multiplier.
"""
import numpy as np
import pytest
import torch

from oracle import disc_oracle as do
from stylish_tts_b200 import _lib as L
from stylish_tts_b200 import discriminator as D
from tests import util
from tests.golden.make_disc_golden import inputs, state_dict_from_table
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def gold():
    z = np.load(util.GOLDEN_DIR + "/discriminators.npz")
    return {k: z[k] for k in z.files}


def test_spec_discriminator_vs_reference_golden():
    g = gold()
    tf, _, _, _ = inputs()
    for i in range(3):
        m = D.SpecDiscriminator()
        m.load_state_dict(state_dict_from_table(g[f"mrd{i}_names"], g[f"mrd{i}_shapes"]), strict=True)
        m = m.to(dev())
        with torch.no_grad():
            outs, fmaps = m(tf[i].to(dev()))
        assert fmaps == [] and len(outs) == 5
        for j, o in enumerate(outs):
            ref = torch.from_numpy(g[f"mrd{i}_out{j}"])
            assert o.shape == ref.shape, (i, j, o.shape, ref.shape)
            assert rel_l2(o, ref) < 3e-5, (i, j, rel_l2(o, ref))


def _seeded_disc(seed):
    torch.manual_seed(seed)
    m = D.SpecDiscriminator()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() > 1 and p.shape[1:] != (1, 1, 1):
                p.copy_(torch.randn(p.shape) / np.sqrt(p[0].numel()))
            elif p.dim() == 1:
                p.copy_(0.1 * torch.randn(p.shape))
    return m


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("bins,frames", [(65, 523), (33, 1030), (20, 333)])
def test_spec_discriminator_tensor_core_path_vs_fp64_oracle(bins, frames, tensor_cores, fused, monkeypatch):
    """fused = the production path (first-layer / tail kernels of csrc/disc_ops.cu + fold-free data gradients);
    fused = False = every layer through the generic conv kernels.  (20, 333): 333 -> 167 -> 84 -> 42 frames, the
    last two layers run below UMMA_MIN_T and take the `wide` tensor-core rule (threshold lowered for the test)."""
    from stylish_tts_b200 import engine as E

    if not fused and frames == 333:
        pytest.skip("wide rule belongs to the fused path")
    monkeypatch.setattr(E, "USE_UMMA", tensor_cores)
    monkeypatch.setattr(E, "WIDE_MIN_ELEMS", 0)
    monkeypatch.setattr(D.SpecDiscriminator, "fused", fused)
    m = _seeded_disc(bins)
    sd64 = {k: v.detach().double().requires_grad_(True) for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(frames)
    y = torch.rand(2, 1, bins, frames, generator=gen) * 3
    y64 = y.double().requires_grad_(True)
    outs_ref = do.spec_discriminator(sd64, y64)
    cots = [torch.randn(o.shape, generator=gen).double() for o in outs_ref]
    sum((o * c).sum() for o, c in zip(outs_ref, cots)).backward()
    md = m.to(dev())
    yd = y.to(dev()).requires_grad_(True)
    calls = []
    orig = L.call
    L.call = lambda name, *a: (calls.append((name, a)), orig(name, *a))[1]
    try:
        outs, _ = md(yd)
    finally:
        L.call = orig
    umma = [a[0]._obj for n, a in calls if n == "sty_conv1d_fwd" and a[0]._obj.w_split]
    if tensor_cores:  # the space-to-depth convs and the last 3x3 conv run on tcgen05
        assert len(umma) >= 4 and {(c.CI, c.K) for c in umma} >= {(192, 5), (96, 3)}
    else:
        assert not umma
    sum((o * c.float().to(dev())).sum() for o, c in zip(outs, cots)).backward()
    torch.cuda.synchronize()
    for j, (o, r) in enumerate(zip(outs, outs_ref)):
        assert o.shape == r.shape and rel_l2(o, r) < 1e-4, (j, rel_l2(o, r))
    # LeakyReLU(0.1) has a kink at 0: a forward difference of 1e-5 (bf16x3) / 1e-6 (fp32 FMA) flips the slope of the
    # pre-activations that close to zero, and the gradient error goes like the square root of that fraction
    tol = 8e-3 if tensor_cores else 1e-3  # measured 3.9e-3 / 5.1e-4; the conv primitives themselves hold 6e-6 (test_style_encoder)
    print("d(input)", rel_l2(yd.grad, y64.grad))
    assert rel_l2(yd.grad, y64.grad) < tol, rel_l2(yd.grad, y64.grad)
    params = dict(md.named_parameters())
    worst = 0.0
    for k, v in sd64.items():
        assert params[k].grad is not None, k
        e = rel_l2(params[k].grad, v.grad)
        worst = max(worst, e)
        assert e < tol, (k, e)
    print("worst parameter gradient", worst)


@pytest.mark.parametrize("n", [1, 2, 7, 1000, 65537, 300001])
def test_tprls_kernels_vs_torch(n):
    gen = torch.Generator().manual_seed(n)
    a = torch.randn(n, generator=gen)
    b = torch.randn(n, generator=gen) * 0.7 + 0.1
    for eps, fn in ((1e-9, lambda r, g_: do.tprls_discriminator([r], [g_])), (0.0, None)):
        a64, b64 = a.double().requires_grad_(True), b.double().requires_grad_(True)
        if fn is None:  # generator form on (gen=a, real=b): do.tprls_generator(real, gen)
            ref = do.tprls_generator([b64], [a64])
        else:
            ref = fn(a64, b64)
        ad, bd = a.to(dev()).requires_grad_(True), b.to(dev()).requires_grad_(True)
        out = D._TprlsFn.apply(ad, bd, eps)
        if not torch.isfinite(ref):
            assert not torch.isfinite(out)  # empty selection: NaN in the reference formula too
            continue
        assert float(out) == pytest.approx(float(ref), rel=2e-5, abs=1e-7), (n, eps)
        ref.backward()
        out.backward()
        if float(a64.grad.abs().sum()) > 0:
            assert rel_l2(ad.grad, a64.grad) < 1e-4 and rel_l2(bd.grad, b64.grad) < 1e-4
        else:
            assert float(ad.grad.abs().sum()) == 0.0


def test_adversarial_losses_vs_oracle_and_lr_control():
    g = gold()
    tf, pf, _, _ = inputs()
    # frames >= 8 so that every scale has several elements; the golden's own sizes
    mods = []
    for i in range(3):
        m = D.SpecDiscriminator()
        m.load_state_dict(state_dict_from_table(g[f"mrd{i}_names"], g[f"mrd{i}_shapes"]), strict=True)
        mods.append(m.to(dev()))
    sds = {f"mrd{i}": {k: v.detach().cpu().double().requires_grad_(True) for k, v in mods[i].state_dict().items()}
           for i in range(3)}
    tfd, pfd = [t.to(dev()) for t in tf], [p.to(dev()).requires_grad_(True) for p in pf]
    pf64 = [p.double().requires_grad_(True) for p in pf]
    tf64 = [t.double() for t in tf]
    # generator side
    gl = D.GeneratorLoss(mrd0=mods[0], mrd1=mods[1], mrd2=mods[2])
    loss = gl(target_list=tfd, pred_list=pfd)
    ref = sum(do.helper_generator(lambda y, i=i: do.spec_discriminator(sds[f"mrd{i}"], y), tf64[i], pf64[i])
              for i in range(3))
    assert float(loss) == pytest.approx(float(ref), rel=1e-4)
    loss.backward()
    ref.backward()
    for i in range(3):
        assert rel_l2(pfd[i].grad, pf64[i].grad) < 1e-3, (i, rel_l2(pfd[i].grad, pf64[i].grad))
    assert all(p.grad is None for m in mods for p in m.parameters())  # constants of the generator step
    assert all(p.requires_grad for m in mods for p in m.parameters())
    # discriminator side
    for sd in sds.values():
        for v in sd.values():
            v.grad = None
    dl = D.DiscriminatorLoss(mrd0=mods[0], mrd1=mods[1], mrd2=mods[2], device=dev())
    dloss = dl(target_list=tfd, pred_list=[p.detach() for p in pfd])
    parts = [do.helper_discriminator(lambda y, i=i: do.spec_discriminator(sds[f"mrd{i}"], y), tf64[i],
                                     pf64[i].detach()) for i in range(3)]
    dref = sum(p[0] for p in parts)
    assert float(dloss) == pytest.approx(float(dref), rel=1e-4)
    dloss.backward()
    dref.backward()
    for i in range(3):
        for k, p in mods[i].named_parameters():
            e = rel_l2(p.grad, sds[f"mrd{i}"][k].grad)
            assert e < 2e-3, (i, k, e)
    # moving average of the plain LSGAN part and the resulting learning-rate multiplier (losses.py:237-249,287)
    last = 0.5 * 5 * 0.95 + float(parts[0][1]) * 0.05
    assert float(dl.lr_control["mrd0"].last_loss) == pytest.approx(last, rel=1e-5)
    assert float(dl.lr_control["mrd0"].last_loss) == pytest.approx(float(g["last_loss_mrd0"]), rel=1e-4)
    assert float(dl.lr_control["mrd0"].multiplier()) == pytest.approx(float(g["lr_mult_mrd0"]), rel=1e-3)


def test_acoustic_step_with_adversarial_terms_and_discriminator_step():
    """configs[2] "full train step": AcousticStep forward + mel / multi-phase / generator (mrd0-2) terms, backward,
    AdamW on speech_predictor + speech_style_encoder, then the discriminator half of Stage.train_batch
    (stage.py:125-146): detached spectrograms, d_loss * sqrt(B), one mrd optimizer stepped at lr_gen x multiplier."""
    from types import SimpleNamespace
    import stylish_tts_b200 as st
    from stylish_tts_b200 import optim, synth, train_step as ts

    d = dev()
    mc = st.default_model_config()
    nets = st.build_model(mc)
    assert {"mrd0", "mrd1", "mrd2", "pitch_disc", "dur_disc"} <= set(nets)
    synth.randomize_(nets.speech_predictor, 0)
    synth.converge_spectral_(nets.speech_style_encoder)
    for k in list(nets):
        nets[k] = nets[k].to(d)
    sp, se = nets.speech_predictor.train(), nets.speech_style_encoder.train()
    sp.regularisers = False
    B, Tn = 2, 18
    inp = synth.speech_inputs(B, Tn, seed=4)
    dur = torch.full((B, Tn), 3.0)
    dur[:, ::9] += 1.0
    frames = int(dur[0].sum())
    g = torch.Generator().manual_seed(8)
    batch = SimpleNamespace(audio_gt=(0.1 * torch.randn(B, frames * 300, generator=g)).to(d),
                            text=inp["texts"].to(d), text_length=inp["text_lengths"].to(d),
                            pitch=inp["pitch"].to(d), alignment=dur.unsqueeze(1).to(d))
    fe = ts.FrontEnd(mc)
    gen_opt = optim.FlatAdamW(list(sp.parameters()) + list(se.parameters()), lr=1e-4, betas=(0.85, 0.99), eps=1e-9,
                              weight_decay=1e-4, world_size=1)
    disc_opts = {f"mrd{i}": optim.FlatAdamW(nets[f"mrd{i}"].parameters(), lr=1e-4, betas=(0.85, 0.99), eps=1e-9,
                                            weight_decay=1e-4, world_size=1) for i in range(3)}
    gl = D.GeneratorLoss(mrd0=nets.mrd0, mrd1=nets.mrd1, mrd2=nets.mrd2)
    dl = D.DiscriminatorLoss(mrd0=nets.mrd0, mrd1=nets.mrd1, mrd2=nets.mrd2, device=d)
    draws = {"noise": inp["draws"]["noise"].to(d)}
    before = {k: o.flat.clone() for k, o in disc_opts.items()}
    out = ts.acoustic_step(batch, nets, fe, source_draws=draws, generator_loss=gl)
    assert out.generator is not None and torch.isfinite(out.generator)
    plain = ts.acoustic_step(batch, nets, fe, source_draws=draws)
    assert float(out.total) == pytest.approx(float(plain.total) + float(out.generator), rel=1e-5)
    out.total.backward()
    for n, p in list(sp.named_parameters()) + list(se.named_parameters()):
        if "m_source" not in n:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert all(p.grad is None for i in range(3) for p in nets[f"mrd{i}"].parameters())
    gen_opt.step()
    gen_opt.zero_grad()
    d_loss = ts.discriminator_step(out, batch, dl, disc_opts, disc_index=1, lr_source=gen_opt)
    torch.cuda.synchronize()
    assert torch.isfinite(d_loss)
    assert float((disc_opts["mrd1"].flat - before["mrd1"]).abs().max()) > 0      # stepped
    assert float((disc_opts["mrd0"].flat - before["mrd0"]).abs().max()) == 0     # not this index
    mult = float(dl.lr_control["mrd1"].multiplier())
    assert float(disc_opts["mrd1"].hyper[0]) == pytest.approx(1e-4 * mult, rel=1e-5)
    assert all(p.grad is None for i in range(3) for p in nets[f"mrd{i}"].parameters())


def test_shared_evaluation_matches_the_two_evaluation_losses():
    """AdversarialTerms (one discriminator evaluation per batch) against GeneratorLoss + DiscriminatorLoss (the
    reference's two evaluations, pinned to the oracle above): generator term and its gradient w.r.t. the predicted
    spectrograms, discriminator loss value, moving averages, and the stepped discriminator's parameter gradients
    (x sqrt(B)); the discriminators that are not stepped get no gradient at all."""
    import math

    tf, pf, _, _ = inputs()
    mods = [_seeded_disc(10 + i).to(dev()) for i in range(3)]
    tfd = [t.to(dev()) for t in tf]
    scale = math.sqrt(tf[0].shape[0])
    # reference schedule
    pa = [p.to(dev()).requires_grad_(True) for p in pf]
    gl = D.GeneratorLoss(mrd0=mods[0], mrd1=mods[1], mrd2=mods[2])
    la = gl(target_list=tfd, pred_list=pa)
    (2.5 * la).backward()
    dl = D.DiscriminatorLoss(mrd0=mods[0], mrd1=mods[1], mrd2=mods[2], device=dev())
    da = dl(target_list=tfd, pred_list=[p.detach() for p in pa])
    (da * scale).backward()
    grads_a = [{k: p.grad.clone() for k, p in m.named_parameters()} for m in mods]
    for m in mods:
        m.zero_grad(set_to_none=True)
    # shared evaluation
    for index in (0, 2):
        pb = [p.to(dev()).requires_grad_(True) for p in pf]
        adv = D.AdversarialTerms(mrd0=mods[0], mrd1=mods[1], mrd2=mods[2], device=dev())
        lb = adv(target_list=tfd, pred_list=pb)
        assert float(lb) == pytest.approx(float(la), rel=1e-6)
        (2.5 * lb).backward()
        assert all(p.grad is None for m in mods for p in m.parameters())  # constants of the generator step
        for i in range(3):
            assert rel_l2(pb[i].grad, pa[i].grad) < 1e-6, (i, rel_l2(pb[i].grad, pa[i].grad))
        before = [p.grad.clone() for p in pb]
        db = adv.discriminator_backward(index, scale)
        assert float(db) == pytest.approx(float(da), rel=1e-6)
        for i in range(3):
            assert float(adv.lr_control[f"mrd{i}"].last_loss) == pytest.approx(float(dl.lr_control[f"mrd{i}"].last_loss),
                                                                               rel=1e-6)
            for k, p in mods[i].named_parameters():
                if i == index:
                    assert rel_l2(p.grad, grads_a[i][k]) < 1e-4, (i, k, rel_l2(p.grad, grads_a[i][k]))  # fp32 sums in another order
                else:
                    assert p.grad is None, (i, k)
        assert all(torch.equal(p.grad, g0) for p, g0 in zip(pb, before))  # untouched by the second walk
        with pytest.raises(RuntimeError):
            adv.discriminator_backward(index, scale)  # the tape is released after the discriminator half
        for m in mods:
            m.zero_grad(set_to_none=True)


def _conv2d_ref(x, w, b, stride=(1, 1), pad=(1, 4)):
    return torch.nn.functional.conv2d(x, w, b, stride=stride, padding=pad)


@pytest.mark.parametrize("B,bins,W", [(2, 17, 600), (1, 5, 37), (3, 9, 513)])
def test_first_layer_kernels_vs_torch_fp64(B, bins, W):
    """sty_disc_first_{fwd,dgrad,wgrad}: Conv2d(1 -> 32, 3x9, pad (1,4)) and both gradients against torch fp64"""
    g = torch.Generator().manual_seed(bins * W)
    y = torch.randn(B, bins, W, generator=g)
    w = torch.randn(32, 1, 3, 9, generator=g) / 5
    b = torch.randn(32, generator=g)
    cot = torch.randn(B, 32, bins, W, generator=g)
    y64, w64, b64 = (t.double().requires_grad_(True) for t in (y, w, b))
    ref = _conv2d_ref(y64[:, None], w64, b64)
    (ref * cot.double()).sum().backward()
    yd, wd, bd = (t.to(dev()).requires_grad_(True) for t in (y, w, b))
    h = D.FirstConvFn.apply(yd, wd, bd, None)                         # (B, bins+2, 32, W)
    assert float(h[:, 0].abs().max()) == 0 and float(h[:, -1].abs().max()) == 0
    assert rel_l2(h[:, 1:-1].permute(0, 2, 1, 3), ref) < 2e-6
    coth = torch.zeros_like(h)
    coth[:, 1:-1] = cot.to(dev()).permute(0, 2, 1, 3)
    (h * coth).sum().backward()
    assert rel_l2(yd.grad, y64.grad) < 2e-6, rel_l2(yd.grad, y64.grad)
    assert rel_l2(wd.grad, w64.grad) < 5e-6, rel_l2(wd.grad, w64.grad)
    assert rel_l2(bd.grad, b64.grad) < 5e-6, rel_l2(bd.grad, b64.grad)


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("B,bins,W", [(2, 19, 300), (1, 6, 131), (2, 8, 58)])
def test_tail_kernels_vs_torch_fp64(B, bins, W, kind):
    """sty_disc_tail_{fwd,bwd} + sty_disc_score_wgrad: LeakyReLU(0.1) -> (score conv 32 -> 1 3x3 | identity |
    space-to-depth) and the single-pass backward, against torch fp64"""
    g = torch.Generator().manual_seed(bins * W + kind)
    Hp = bins + 2
    h = torch.zeros(B, Hp, 32, W)
    h[:, 1:-1] = torch.randn(B, bins, 32, W, generator=g)
    ws = torch.randn(1, 32, 3, 3, generator=g) / 8
    bs = torch.randn(1, generator=g)
    h64, w64, b64 = (t.double().requires_grad_(True) for t in (h, ws, bs))
    a64 = torch.nn.functional.leaky_relu(h64, 0.1)                   # (B,Hp,32,W), zero border rows stay zero
    score64 = _conv2d_ref(a64[:, 1:-1].permute(0, 2, 1, 3), w64, b64, pad=(1, 1))[:, 0]   # (B,bins,W)
    W2 = (W + 1) // 2
    if kind == 1:
        next64 = a64
    elif kind == 2:
        ap = torch.nn.functional.pad(a64, (0, 2 * W2 - W))
        next64 = ap.reshape(B, Hp, 32, W2, 2).permute(0, 1, 2, 4, 3).reshape(B, Hp, 64, W2)
    cs = torch.randn(score64.shape, generator=g)
    loss = (score64 * cs.double()).sum()
    cn = None
    if kind:
        cn = torch.randn(next64.shape, generator=g)
        cn[:, 0] = 0
        cn[:, -1] = 0          # consumers never send gradient into the border rows
        loss = loss + (next64 * cn.double()).sum()
    loss.backward()
    hd, wd, bd = (t.to(dev()).requires_grad_(True) for t in (h, ws, bs))
    score, nxt = D.TailFn.apply(hd, wd, bd, kind, None)
    assert rel_l2(score, score64) < 2e-6, rel_l2(score, score64)
    lossd = (score * cs.to(dev())).sum()
    if kind:
        assert nxt.shape == next64.shape and rel_l2(nxt, next64) < 1e-7
        lossd = lossd + (nxt * cn.to(dev())).sum()
    lossd.backward()
    assert float(hd.grad[:, 0].abs().max()) == 0 and float(hd.grad[:, -1].abs().max()) == 0
    assert rel_l2(hd.grad[:, 1:-1], h64.grad[:, 1:-1]) < 2e-6, rel_l2(hd.grad[:, 1:-1], h64.grad[:, 1:-1])
    assert rel_l2(wd.grad, w64.grad) < 5e-6, rel_l2(wd.grad, w64.grad)
    assert rel_l2(bd.grad, b64.grad) < 5e-6, rel_l2(bd.grad, b64.grad)


def _cfd(seed=0):
    g = gold()
    m = D.ContextFreeDiscriminator()
    m.load_state_dict(state_dict_from_table(g["disc_names"], g["disc_shapes"]), strict=True)
    return m


def test_context_free_discriminator_vs_reference_golden():
    """`disc` (ContextFreeDiscriminator, discriminator.py:119-175) against the output of the UNMODIFIED reference in
    train() mode (batch-statistics BatchNorm), same state dict"""
    g = gold()
    _, _, ta, _ = inputs()
    m = _cfd().to(dev()).train()
    with torch.no_grad():
        outs, fmaps = m(ta.to(dev()))
    assert fmaps == [] and len(outs) == 1
    ref = torch.from_numpy(g["disc_out"])
    assert outs[0].shape == ref.shape and rel_l2(outs[0], ref) < 2e-4, rel_l2(outs[0], ref)


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("training", [True, False])
def test_context_free_discriminator_vs_fp64_oracle(tensor_cores, training, monkeypatch):
    """scores, input gradient and every parameter gradient against oracle/disc_oracle.py in fp64; train() mode
    (batch statistics, running buffers updated like nn.BatchNorm1d) and eval() mode (running statistics)"""
    from stylish_tts_b200 import engine as E

    monkeypatch.setattr(E, "USE_UMMA", tensor_cores)
    pass
    m = _cfd()
    with torch.no_grad():  # non-trivial running statistics for the eval() case
        for k, b in m.named_buffers():
            if k.endswith("running_mean"):
                b.copy_(0.05 * torch.randn(b.shape, generator=torch.Generator().manual_seed(len(k))))
            if k.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=torch.Generator().manual_seed(len(k) + 1)))
    gen = torch.Generator().manual_seed(3)
    x = 0.2 * torch.randn(3, 1024 + 512 * 6, generator=gen)
    names = {k for k, _ in m.named_parameters()}
    sd64 = {k: (v.detach().double().requires_grad_(True) if k in names else v.detach().double().clone()
                if v.is_floating_point() else v.clone()) for k, v in m.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    ref = do.context_free_discriminator(sd64, x64, bn_training=training)[0]
    cot = torch.randn(ref.shape, generator=gen).double()
    (ref * cot).sum().backward()
    md = m.to(dev())
    md.train(training)
    before = {k: b.clone() for k, b in md.named_buffers()}
    xd = x.to(dev()).requires_grad_(True)
    out = md(xd)[0][0]
    assert out.shape == ref.shape and rel_l2(out, ref) < 2e-4, rel_l2(out, ref)
    (out * cot.float().to(dev())).sum().backward()
    torch.cuda.synchronize()
    # GELU / ReLU kinks + 9 BatchNorm layers amplify the bf16x3 forward difference in a handful of elements
    tol = 5e-3 if tensor_cores else 1e-3
    assert rel_l2(xd.grad, x64.grad) < tol, rel_l2(xd.grad, x64.grad)
    worst = 0.0
    for k, p in md.named_parameters():
        assert p.grad is not None, k
        r = sd64[k].grad
        if training and float(r.norm()) < 1e-10 * max(float(sd64[k].detach().norm()), 1.0):
            # a bias straight in front of a batch-statistics BatchNorm has NO gradient (the mean removes it): the fp64
            # value is rounding noise, ours must be noise of fp32 size
            assert float(p.grad.norm()) < 1e-5, (k, float(p.grad.norm()))
            continue
        e = rel_l2(p.grad, r)
        worst = max(worst, e)
        assert e < tol, (k, e)
    print("worst parameter gradient", worst)
    after = dict(md.named_buffers())
    if not training:
        assert all(torch.equal(before[k], after[k]) for k in before)
    else:
        # first BatchNorm: running statistics follow nn.BatchNorm1d (momentum 0.1, unbiased variance)
        h0 = torch.nn.functional.conv1d(x.double().unfold(1, 1024, 512).reshape(-1, 1, 1024),
                                        sd64["conv.0.net.0.weight"].detach(), stride=4, padding=5)
        mean, var = h0.mean((0, 2)), h0.var((0, 2), unbiased=True)
        rm = 0.9 * before["conv.0.net.1.running_mean"].double().cpu() + 0.1 * mean
        rv = 0.9 * before["conv.0.net.1.running_var"].double().cpu() + 0.1 * var
        assert rel_l2(after["conv.0.net.1.running_mean"], rm) < 1e-5
        assert rel_l2(after["conv.0.net.1.running_var"], rv) < 1e-5
        assert int(after["conv.0.net.1.num_batches_tracked"]) == 1


def test_adversarial_terms_with_waveform_discriminator():
    """AdversarialTerms with `disc`: generator term = sum over mrd0-2 + disc_weight x disc (losses.py:316-327), the
    discriminator half steps mrd{index} AND disc (stage.py:141-143); against the two-evaluation classes"""
    import math

    tf, pf, ta, pa = inputs()
    mods = [_seeded_disc(20 + i).to(dev()) for i in range(3)]
    disc = _cfd().to(dev()).train()
    tfd, tad = [t.to(dev()) for t in tf], ta.to(dev())
    scale = math.sqrt(2)
    pfa = [p.to(dev()).requires_grad_(True) for p in pf]
    paa = pa.to(dev()).requires_grad_(True)
    gl = D.GeneratorLoss(mrd0=mods[0], mrd1=mods[1], mrd2=mods[2], disc=disc)
    la = gl(target_list=tfd, pred_list=pfa, target_audio=tad, pred_audio=paa)
    la.backward()
    dl = D.DiscriminatorLoss(mrd0=mods[0], mrd1=mods[1], mrd2=mods[2], disc=disc, device=dev())
    da = dl(target_list=tfd, pred_list=[p.detach() for p in pfa], target_audio=tad, pred_audio=paa.detach())
    (da * scale).backward()
    grads_disc = {k: p.grad.clone() for k, p in disc.named_parameters()}
    grads_m1 = {k: p.grad.clone() for k, p in mods[1].named_parameters()}
    for m in mods + [disc]:
        m.zero_grad(set_to_none=True)
    pfb = [p.to(dev()).requires_grad_(True) for p in pf]
    pab = pa.to(dev()).requires_grad_(True)
    adv = D.AdversarialTerms(mrd0=mods[0], mrd1=mods[1], mrd2=mods[2], disc=disc, device=dev())
    lb = adv(target_list=tfd, pred_list=pfb, target_audio=tad, pred_audio=pab)
    assert float(lb) == pytest.approx(float(la), rel=1e-5)
    lb.backward()
    assert rel_l2(pab.grad, paa.grad) < 1e-4, rel_l2(pab.grad, paa.grad)
    db = adv.discriminator_backward(1, scale)
    assert float(db) == pytest.approx(float(da), rel=1e-5)
    for k, p in disc.named_parameters():
        if float(grads_disc[k].norm()) < 1e-6 and float(p.grad.norm()) < 1e-6:
            continue  # bias in front of a batch-statistics BatchNorm: zero gradient, both sides are rounding noise
        assert rel_l2(p.grad, grads_disc[k]) < 1e-3, (k, rel_l2(p.grad, grads_disc[k]))
    for k, p in mods[1].named_parameters():
        assert rel_l2(p.grad, grads_m1[k]) < 1e-4, (k, rel_l2(p.grad, grads_m1[k]))
    assert all(p.grad is None for i in (0, 2) for p in mods[i].parameters())


def _reference_set():
    """all six discriminators with the golden's state dicts (the reference's own values)"""
    g = gold()
    mods = {}
    for key, cls in (("mrd0", D.SpecDiscriminator), ("mrd1", D.SpecDiscriminator), ("mrd2", D.SpecDiscriminator),
                     ("disc", D.ContextFreeDiscriminator)):
        m = cls()
        m.load_state_dict(state_dict_from_table(g[f"{key}_names"], g[f"{key}_shapes"]), strict=True)
        mods[key] = m.to(dev()).train()
    for key, kw in (("pitch_disc", dict(dim_in=2, dim_hidden=64, kernel=21)), ("dur_disc", dict(dim_in=1, dim_hidden=64, kernel=5))):
        m = D.PitchDiscriminator(**kw)
        m.load_state_dict(state_dict_from_table(g[f"{key}_names"], g[f"{key}_shapes"]), strict=True)
        mods[key] = m.to(dev()).train()
    return g, mods


def test_loss_classes_vs_reference_golden_full_acoustic_set():
    """GeneratorLoss / DiscriminatorLoss with the reference's constructor and call keywords (mrd0-2 + disc + pitch +
    duration; used=, index=) against the values the UNMODIFIED reference classes produced (tests/golden/
    make_disc_golden.py): generator loss of the full acoustic set, its gradient w.r.t. the predicted audio and the
    first predicted spectrogram, the pitch / duration routes, the discriminator loss, gradients of `disc.last.2` and of
    an mrd1 weight-norm direction, the moving average and the learning-rate multiplier"""
    from tests.golden.make_disc_golden import curve_inputs

    g, mods = _reference_set()
    kw = dict(mrd0=mods["mrd0"], mrd1=mods["mrd1"], mrd2=mods["mrd2"], disc=mods["disc"], pitch=mods["pitch_disc"],
              duration=mods["dur_disc"])
    gl, dl = D.GeneratorLoss(**kw), D.DiscriminatorLoss(**kw, device=dev())
    tf, pf, ta, pa = inputs()
    tfd, tad = [t.to(dev()) for t in tf], ta.to(dev())
    pfd = [p.to(dev()).requires_grad_(True) for p in pf]
    pad = pa.to(dev()).requires_grad_(True)
    args = dict(target_list=tfd, target_audio=tad, used=["mrd0", "mrd1", "mrd2", "disc"], index=0)
    loss = gl(pred_list=pfd, pred_audio=pad, **args).mean()
    assert float(loss) == pytest.approx(float(g["gen_loss"]), rel=2e-4)
    loss.backward()
    assert rel_l2(pad.grad, torch.from_numpy(g["gen_d_pred_audio"])) < 2e-3
    assert rel_l2(pfd[0].grad, torch.from_numpy(g["gen_d_pred_fft0"])) < 2e-3
    assert all(p.grad is None for m in mods.values() for p in m.parameters())
    pc, du = curve_inputs()
    pc, du = pc.to(dev()), du.to(dev())
    pg = gl(target_list=[pc], pred_list=[pc * 0.9 + 0.1], target_audio=None, pred_audio=None, used=["pitch_disc"], index=0)
    assert float(pg) == pytest.approx(float(g["pitch_gen_loss"]), rel=2e-4)
    dd = dl(target_list=[du], pred_list=[du * 1.1 - 0.2], target_audio=None, pred_audio=None, used=["dur_disc"], index=0)
    assert float(dd) == pytest.approx(float(g["dur_disc_loss"]), rel=2e-4)
    d = dl(pred_list=[p.detach() for p in pfd], pred_audio=pad.detach(), **args).mean()
    assert float(d) == pytest.approx(float(g["disc_loss"]), rel=2e-4)
    for m in mods.values():
        m.zero_grad(set_to_none=True)
    d.backward()
    assert rel_l2(mods["disc"].last[2].weight.grad, torch.from_numpy(g["disc_d_last2_w"])) < 2e-3
    w = mods["mrd1"].discriminators[2].parametrizations.weight.original1
    assert float(w.grad.norm()) == pytest.approx(float(g["disc_d_mrd1_conv2_v_norm"]), rel=5e-3)
    assert float(dl.lr_control["mrd0"].last_loss) == pytest.approx(float(g["last_loss_mrd0"]), rel=1e-4)
    assert float(dl.get_disc_lr_multiplier("mrd0")) == pytest.approx(float(g["lr_mult_mrd0"]), rel=1e-3)
    sd = dl.state_dict()
    assert sd["discriminators.mrd0.last_loss"] == pytest.approx(float(g["last_loss_mrd0"]), rel=1e-4)
    dl.load_state_dict({"discriminators.disc.last_loss": 0.25})
    assert float(dl.lr_control["disc"].last_loss) == 0.25
