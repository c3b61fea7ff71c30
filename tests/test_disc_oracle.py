"""oracle/disc_oracle.py (discriminators + adversarial losses, SURVEY §8f rank 1 — the next row; no CUDA path
yet) against golden vectors of the UNMODIFIED reference (tests/golden/make_disc_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import disc_oracle as do
from tests import util
from tests.golden.make_disc_golden import curve_inputs, inputs, state_dict_from_table
from tests.util import rel_l2


@pytest.fixture(scope="module")
def gold():
    z = np.load(util.GOLDEN_DIR + "/discriminators.npz")
    return {k: z[k] for k in z.files}


def sds_of(gold, grad=False):
    out = {}
    for key in ("mrd0", "mrd1", "mrd2", "disc", "pitch_disc", "dur_disc"):
        sd = state_dict_from_table(gold[key + "_names"], gold[key + "_shapes"])
        if grad:
            sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                  for k, v in sd.items()}
        out[key] = sd
    return out


def test_discriminator_outputs(gold):
    sds = sds_of(gold)
    tf, _, ta, _ = inputs()
    with torch.no_grad():
        for i in range(3):
            outs = do.spec_discriminator(sds[f"mrd{i}"], tf[i])
            assert len(outs) == 5
            for j, o in enumerate(outs):
                assert rel_l2(o, torch.from_numpy(gold[f"mrd{i}_out{j}"])) < 2e-6, (i, j)
        d = do.context_free_discriminator(sds["disc"], ta, bn_training=True)
        assert len(d) == 1 and d[0].shape == gold["disc_out"].shape
        assert rel_l2(d[0], torch.from_numpy(gold["disc_out"])) < 1e-5


def test_adversarial_losses_and_gradients(gold):
    tf, pf, ta, pa = inputs()
    sds = sds_of(gold)
    pa_g = pa.clone().requires_grad_(True)
    pf_g = [p.clone().requires_grad_(True) for p in pf]
    g = do.acoustic_generator_loss(sds, tf, pf_g, ta, pa_g)
    g.backward()
    assert float(g.detach()) == pytest.approx(float(gold["gen_loss"]), rel=2e-6)
    assert rel_l2(pa_g.grad, torch.from_numpy(gold["gen_d_pred_audio"])) < 1e-4
    assert rel_l2(pf_g[0].grad, torch.from_numpy(gold["gen_d_pred_fft0"])) < 1e-4
    sds = sds_of(gold, grad=True)
    d = do.acoustic_discriminator_loss(sds, tf, pf, ta, pa)
    d.backward()
    assert float(d.detach()) == pytest.approx(float(gold["disc_loss"]), rel=2e-6)
    v = sds["mrd1"]["discriminators.2.parametrizations.weight.original1"]
    assert float(v.grad.norm()) == pytest.approx(float(gold["disc_d_mrd1_conv2_v_norm"]), rel=1e-4)
    assert rel_l2(sds["disc"]["last.2.weight"].grad, torch.from_numpy(gold["disc_d_last2_w"])) < 1e-4


def test_pitch_and_duration_discriminators(gold):
    sds = sds_of(gold)
    pc, du = curve_inputs()
    with torch.no_grad():
        for key, x, k in (("pitch_disc", pc, 21), ("dur_disc", du, 5)):
            outs = do.pitch_discriminator(sds[key], x, k)
            for j, o in enumerate(outs):
                assert rel_l2(o, torch.from_numpy(gold[f"{key}_out{j}"])) < 2e-6, (key, j)
        g = do.helper_generator(lambda y: do.pitch_discriminator(sds["pitch_disc"], y, 21), pc, pc * 0.9 + 0.1)
        assert float(g) == pytest.approx(float(gold["pitch_gen_loss"]), rel=2e-6)
        d, _ = do.helper_discriminator(lambda y: do.pitch_discriminator(sds["dur_disc"], y, 5), du, du * 1.1 - 0.2)
        assert float(d) == pytest.approx(float(gold["dur_disc_loss"]), rel=2e-6)


def test_lr_controller(gold):
    tf, pf, _, _ = inputs()
    sds = sds_of(gold)
    with torch.no_grad():
        _, plain = do.helper_discriminator(lambda y: do.spec_discriminator(sds["mrd0"], y), tf[0], pf[0])
    last = 0.5 * 5 * 0.95 + float(plain) * 0.05  # DiscriminatorLossHelper.forward, losses.py:287
    assert last == pytest.approx(float(gold["last_loss_mrd0"]), rel=1e-6)
    assert do.disc_lr_multiplier(last, 5) == pytest.approx(float(gold["lr_mult_mrd0"]), rel=1e-5)
    assert do.disc_lr_multiplier(2.5 + 0.3, 5) == 4.0 and do.disc_lr_multiplier(2.5 - 0.3, 5) == 0.01
    assert do.disc_lr_multiplier(2.5, 5) == 1.0
