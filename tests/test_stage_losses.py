"""Loss terms of the textual / duration stages (stylish_tts_b200/stage_losses.py) against the reference's own code
(live when /root/reference is mounted) and closed forms."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_loader
from stylish_tts_b200 import stage_losses as sl
from stylish_tts_b200.modules import DurationProcessor


def case():
    g = torch.Generator().manual_seed(3)
    raw = torch.randn(3, 9, 16, generator=g, requires_grad=True)
    lens = torch.tensor([9, 5, 7])
    target = torch.randint(1, 40, (3, 9), generator=g)
    weight = sl.duration_class_weights(torch.randint(1, 500, (16,), generator=g))
    return raw, lens, target, weight


def test_curve_loss_closed_form():
    t = torch.tensor([[0.0, 1.0, 3.0]])
    p = torch.tensor([[0.5, 1.0, 6.0]])
    want = (0.125 + 0 + 2.5) / 3 + (0.125 + 2.5) / 2  # smooth-L1 (beta 1) of the curve + of its difference
    assert float(sl.curve_loss(t, p)) == pytest.approx(want, rel=1e-6)


def test_duration_losses_match_reference():
    raw, lens, target, weight = case()
    proc = DurationProcessor(16, 50)
    dur = proc.prediction_to_duration(raw, lens)
    cls = proc.dur_to_class(target)
    l1, ce = sl.duration_losses(raw, dur, target, cls, lens, weight)
    (l1 + ce).backward()
    assert raw.grad is not None and torch.isfinite(raw.grad).all()
    assert float(raw.grad[1, 5:].abs().max()) == 0.0  # padded tokens carry no gradient
    if not ref_loader.available():
        pytest.skip("/root/reference not mounted")
    ref_loader.load()
    from stylish_tts.train.losses import DurationLoss
    from stylish_tts.train.utils import DurationProcessor as RefDP

    rp = RefDP(16, 50)
    raw2 = raw.detach().clone().requires_grad_(True)
    dur2 = rp.prediction_to_duration(raw2, lens)
    want_l1 = sum(F.smooth_l1_loss(dur2[i, :lens[i]], target[i, :lens[i]]) for i in range(3)) / 3  # stage_type.py:514-518
    want_ce, _ = DurationLoss(class_count=16, weight=weight)(raw2, rp.dur_to_class(target), lens)
    (want_l1 + want_ce).backward()
    assert float(l1.detach()) == pytest.approx(float(want_l1.detach()), rel=1e-6)
    assert float(ce.detach()) == pytest.approx(float(want_ce.detach()), rel=1e-6)
    assert torch.allclose(raw.grad, raw2.grad, atol=1e-7)
