"""Mel style encoder.

CPU: state-dict compatibility with the reference's key list; the oracle restatement against golden
outputs / gradients of the UNMODIFIED reference (tests/golden/make_style_golden.py).
GPU: the CUDA module (row-channel images, Conv2d as Conv1d over stacked rows, through the C ABI) against the
oracle and the golden, forward (eval + train) and every parameter gradient; image-op kernels one by one.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_loader
from oracle import style_oracle as sto
from tests import util
from tests.golden.make_style_golden import build, probe, style_inputs
from tests.util import rel_l2


def gold():
    z = np.load(util.GOLDEN_DIR + "/style_encoder.npz")
    return {k: z[k] for k in z.files}


def sd_of(m, dtype=torch.float32, grad=False):
    out = {}
    for k, v in m.state_dict().items():
        t = v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()
        if grad and (k.endswith("weight_orig") or k.endswith("bias") or k.startswith("unshared")):
            t.requires_grad_(True)
        out[k] = t
    return out


def test_state_dict_keys_match_reference():
    g = gold()
    m = build()
    params = sorted(n for n, _ in m.named_parameters())
    assert params == sorted(str(n) for n in g["names"])
    keys = set(m.state_dict())
    for n in params:
        if n.endswith("weight_orig"):
            assert n[:-5] + "_u" in keys and n[:-5] + "_v" in keys
    assert sum(p.numel() for p in m.parameters()) == 9452624  # 9.45 M


def test_oracle_matches_reference_golden():
    g = gold()
    m = build()
    x, ct = style_inputs()
    with torch.no_grad():
        out_eval = sto.mel_style_encoder(sd_of(m), x, training=False)
    assert rel_l2(out_eval, torch.from_numpy(g["out_eval"])) < 1e-5
    sd = sd_of(m, grad=True)
    out = sto.mel_style_encoder(sd, x, training=True)
    assert rel_l2(out, torch.from_numpy(g["out_train"])) < 1e-5
    (out * ct).sum().backward()
    for n, norm, dot in zip([str(s) for s in g["names"]], g["norms"], g["dots"]):
        gr = sd[n].grad
        assert abs(float(gr.norm()) - norm) <= 2e-4 * norm + 1e-9, n
        assert abs(float((gr * probe(n, gr.shape)).sum()) - dot) <= 1e-3 * norm + 1e-9, n
    assert rel_l2(sd["shared.0.weight_u"], torch.from_numpy(g["u0"])) < 1e-5
    assert rel_l2(sd["shared.3.conv2.weight_v"], torch.from_numpy(g["v3"])) < 1e-5


def test_cpu_input_is_rejected():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        build()(torch.zeros(1, 1, 80, 64))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 1, 80, 52), (2, 1, 80, 176)])
def test_gpu_forward_and_gradients(shape):
    """(.., 52): odd widths after pooling, W < 128 everywhere (fp32 FMA convs); (.., 176): tensor-core convs"""
    g = gold()
    m = build()
    gen = torch.Generator().manual_seed(31)
    x = torch.randn(shape, generator=gen)
    ct = torch.randn(shape[0], 64, generator=gen)
    golden_case = shape == (2, 1, 80, 52)
    with torch.no_grad():
        ref_eval = sto.mel_style_encoder(sd_of(m, torch.float64), x.double(), training=False)
    sd = sd_of(m, torch.float64, grad=True)
    ref = sto.mel_style_encoder(sd, x.double(), training=True)
    (ref * ct.double()).sum().backward()

    mc = build().cuda()
    mc.eval()
    with torch.no_grad():
        out_eval = mc(x.cuda())
    assert out_eval.shape == (shape[0], 64)
    assert rel_l2(out_eval, ref_eval) < 2e-4
    mc.train()
    out = mc(x.cuda())
    (out * ct.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 2e-4
    if golden_case:
        assert rel_l2(out_eval, torch.from_numpy(g["out_eval"])) < 2e-4
        assert rel_l2(out, torch.from_numpy(g["out_train"])) < 2e-4
    worst = 0.0
    for n, p in mc.named_parameters():
        assert p.grad is not None, n
        e = rel_l2(p.grad, sd[n].grad)
        worst = max(worst, e)
        assert e < 1e-3, (n, e)
    print("style encoder: worst parameter-gradient error vs fp64 oracle", worst)
    # spectral-norm state advanced by one power iteration, like the reference in train mode
    assert rel_l2(mc.state_dict()["shared.0.weight_u"], sd["shared.0.weight_u"]) < 1e-5


@pytest.mark.gpu
def test_image_op_kernels():
    from stylish_tts_b200 import style_encoder as se

    gen = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 6, 8, 13
    x = torch.randn(B, C, H, W, generator=gen)

    def to_rc(t):  # (B,C,H,W) -> row-channel with zero border rows
        Bq, Cq, Hq, Wq = t.shape
        out = torch.zeros(Bq, Hq + 2, Cq, Wq, dtype=t.dtype)
        out[:, 1:Hq + 1] = t.permute(0, 2, 1, 3)
        return out

    def from_rc(t):
        return t[:, 1:-1].permute(0, 2, 1, 3)

    # learned stride-2 depthwise conv
    w, b = torch.randn(C, 1, 3, 3, generator=gen), torch.randn(C, generator=gen)
    xr, wr, br = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, stride=2, padding=1, groups=C)
    xc = to_rc(x).cuda().requires_grad_(True)
    wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = se.DwDownFn.apply(xc, wc, bc)
    assert float(out[:, 0].abs().max()) == 0.0 and float(out[:, -1].abs().max()) == 0.0
    ct = torch.randn(ref.shape, generator=gen)
    (ref * ct.double()).sum().backward()
    (out * to_rc(ct).cuda()).sum().backward()
    assert rel_l2(from_rc(out), ref) < 1e-5
    assert rel_l2(from_rc(xc.grad), xr.grad) < 1e-5 and rel_l2(wc.grad, wr.grad) < 1e-5 and rel_l2(bc.grad, br.grad) < 1e-5

    # average pool with odd width
    xr = x.double().requires_grad_(True)
    ref = F.avg_pool2d(torch.cat([xr, xr[..., -1:]], -1), 2)
    xc = to_rc(x).cuda().requires_grad_(True)
    out = se.AvgPool2Fn.apply(xc)
    ct = torch.randn(ref.shape, generator=gen)
    (ref * ct.double()).sum().backward()
    (out * to_rc(ct).cuda()).sum().backward()
    assert rel_l2(from_rc(out), ref) < 1e-6 and rel_l2(from_rc(xc.grad), xr.grad) < 1e-6

    # 3x3 conv as a row conv (with LeakyReLU prologue, residual, scale)
    Co = 10
    w4, b4 = torch.randn(Co, C, 3, 3, generator=gen) * 0.2, torch.randn(Co, generator=gen)
    res = torch.randn(B, Co, H, W, generator=gen)
    xr, wr, br, rr = (t.double().requires_grad_(True) for t in (x, w4, b4, res))
    ref = 0.7 * F.conv2d(F.leaky_relu(xr, 0.2), wr, br, padding=1) + 0.7 * rr
    xc, wc, bc, rc = to_rc(x).cuda().requires_grad_(True), w4.cuda().requires_grad_(True), \
        b4.cuda().requires_grad_(True), to_rc(res).cuda().requires_grad_(True)
    enc = build()
    mask = enc._row_mask(B, H + 2, W, 3, torch.device("cuda:0"))
    out = se.RowConvFn.apply(xc, wc, bc, rc, dict(in_act=2, row_mask=mask, out_scale=0.7, res_scale=0.7))
    assert float(out[:, 0].abs().max()) == 0.0 and float(out[:, -1].abs().max()) == 0.0
    ct = torch.randn(ref.shape, generator=gen)
    (ref * ct.double()).sum().backward()
    (out * to_rc(ct).cuda()).sum().backward()
    assert rel_l2(from_rc(out), ref) < 1e-5
    for a, r in ((from_rc(xc.grad), xr.grad), (wc.grad, wr.grad), (bc.grad, br.grad), (from_rc(rc.grad), rr.grad)):
        assert rel_l2(a, r) < 2e-5

    # region mean
    xr = x.double().requires_grad_(True)
    ref = xr[:, :, 1:6, 2:11].mean(dim=(2, 3))
    xc = to_rc(x).cuda().requires_grad_(True)
    out = se.RegionMeanFn.apply(xc, 2, 5, 2, 9)
    ct = torch.randn(ref.shape, generator=gen)
    (ref * ct.double()).sum().backward()
    (out * ct.cuda()).sum().backward()
    assert rel_l2(out, ref) < 1e-6 and rel_l2(from_rc(xc.grad), xr.grad) < 1e-6


def test_pitch_style_encoder_oracle_matches_reference_golden():
    from tests.golden.make_style_golden import build_pitch, pitch_inputs

    m = build_pitch()
    with torch.no_grad():
        out = sto.pitch_style_encoder(sd_of(m), *pitch_inputs())
    assert rel_l2(out, torch.from_numpy(gold()["pe_out_eval"])) < 1e-5


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_pitch_style_encoder_coarse_multiplier_vs_live_reference():
    """coarse_multiplier = 2: pitch / energy at the fine rate, mel at the coarse one (mel_style_encoder.py:188-206)"""
    from tests.golden.make_style_golden import build_pitch, pitch_inputs

    ref_loader.load()
    from stylish_tts.train.models.mel_style_encoder import PitchStyleEncoder as RefPSE

    m = build_pitch()
    ref = RefPSE(80, 64, 384, True, coarse_multiplier=2).eval()
    ref.load_state_dict(m.state_dict(), strict=True)
    x, pitch, energy = pitch_inputs()
    fine_p, fine_e = pitch.repeat_interleave(2, dim=1) * 1.01, energy.repeat_interleave(2, dim=1) + 0.1
    with torch.no_grad():
        want = ref(x, fine_p, fine_e)
        got = sto.pitch_style_encoder(sd_of(m), x, fine_p, fine_e, coarse_multiplier=2)
    assert rel_l2(got, want) < 1e-5


@pytest.mark.gpu
def test_pitch_style_encoder_coarse_multiplier_gpu():
    from stylish_tts_b200.style_encoder import PitchStyleEncoder
    from tests.golden.make_style_golden import build_pitch, pitch_inputs

    m = build_pitch()
    x, pitch, energy = pitch_inputs()
    fine_p, fine_e = pitch.repeat_interleave(2, dim=1) * 1.01, energy.repeat_interleave(2, dim=1) + 0.1
    with torch.no_grad():
        want = sto.pitch_style_encoder(sd_of(m, torch.float64), x.double(), fine_p.double(), fine_e.double(),
                                       coarse_multiplier=2)
    mc = PitchStyleEncoder(80, 64, 384, True, coarse_multiplier=2)
    mc.load_state_dict(m.state_dict(), strict=True)
    mc = mc.cuda().eval()
    with torch.no_grad():
        got = mc(x.cuda(), fine_p.cuda(), fine_e.cuda())
    assert rel_l2(got, want) < 2e-4


@pytest.mark.gpu
def test_pitch_style_encoder_gpu():
    from tests.golden.make_style_golden import build_pitch, pitch_inputs

    m = build_pitch()
    x, pitch, energy = pitch_inputs()
    sd = sd_of(m, torch.float64, grad=True)
    for k in list(sd):
        if "parametrizations" in k:
            sd[k].requires_grad_(True)
    ref = sto.pitch_style_encoder(sd, x.double(), pitch.double(), energy.double(), training=True)
    ref.square().sum().backward()
    mc = build_pitch().cuda().train()
    out = mc(x.cuda(), pitch.cuda(), energy.cuda())
    out.square().sum().backward()
    assert rel_l2(out, ref) < 2e-4
    for n, p in mc.named_parameters():
        assert p.grad is not None and rel_l2(p.grad, sd[n].grad) < 1e-3, n
    mc.eval()
    with torch.no_grad():
        assert rel_l2(build_pitch().cuda().eval()(x.cuda(), pitch.cuda(), energy.cuda()),
                      torch.from_numpy(gold()["pe_out_eval"])) < 2e-4


def test_build_model_has_style_encoders():
    import stylish_tts_b200 as st

    nets = st.build_model(st.default_model_config())
    for k in ("speech_style_encoder", "pe_style_encoder", "duration_style_encoder"):
        assert k in nets
    assert "preconv.parametrizations.weight.original0" in nets.pe_style_encoder.state_dict()
