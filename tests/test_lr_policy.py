"""Learning-rate policy: cosine schedule on the 10 000-logical-step clock and the gap-aware
discriminator multiplier, against the live reference objects when mounted (optimizers.py:96-103 with
transformers' cosine schedule; losses.py:229-250) and against the pinned oracle formula otherwise."""
import math

import pytest
import torch

from oracle import disc_oracle, ref_loader
from stylish_tts_b200 import optim


def test_cosine_schedule_matches_transformers_scheduler():
    import transformers

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-4)
    sched = transformers.get_cosine_schedule_with_warmup(opt, num_warmup_steps=0, num_training_steps=10000)
    for step, limit in [(0, 1000), (1, 1000), (137, 1000), (500, 777), (899, 1000), (900, 1000), (999, 1000), (5, 7)]:
        logical = step * 10000 // limit  # MultiOptimizer.scheduler, optimizers.py:96-103
        logical = min(logical, 10000 * 0.9)
        sched.last_epoch = logical
        sched.step()
        ref = opt.param_groups[0]["lr"]
        assert math.isclose(optim.cosine_lr(1e-4, step, limit), ref, rel_tol=1e-12, abs_tol=1e-18), (step, limit)
    # plateau: nothing changes after 90 % of the stage
    assert optim.cosine_lr(1e-4, 950, 1000) == optim.cosine_lr(1e-4, 900, 1000)


@pytest.mark.parametrize("sub_count", [1, 5])
def test_discriminator_lr_multiplier(sub_count):
    d = optim.DiscriminatorLR(sub_count)
    helper = None
    if ref_loader.available():
        ref_loader.load()
        from stylish_tts.train.losses import DiscriminatorLossHelper
        helper = DiscriminatorLossHelper(torch.nn.Identity(), sub_count)
    g = torch.Generator().manual_seed(sub_count)
    for _ in range(200):
        loss = torch.rand((), generator=g) * 1.2 * sub_count
        d.update(loss)
        want = disc_oracle.disc_lr_multiplier(float(d.last_loss), sub_count)
        assert math.isclose(float(d.multiplier()), want, rel_tol=2e-6), (float(d.last_loss), want)
        if helper is not None:
            helper.last_loss = helper.last_loss * 0.95 + loss.item() * 0.05  # losses.py:287
            assert math.isclose(float(d.last_loss), helper.last_loss, rel_tol=1e-5)
            assert math.isclose(float(d.multiplier()), helper.get_disc_lr_multiplier(), rel_tol=1e-4)
