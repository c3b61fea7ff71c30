"""Golden vectors for oracle/disc_oracle.py from the UNMODIFIED reference discriminators and adversarial losses
(build container only):   python tests/golden/make_disc_golden.py

The modules' parameters are overwritten with values derived from their NAMES (``det_tensor``), so the test can
rebuild the same state dicts from the (name, shape) table stored in the fixture without the reference and without
storing 1.5 M weights."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SEED = 5


def det_tensor(name, shape):
    """deterministic parameter value from its name: N(0,1)/sqrt(fan_in) for weights, small for biases / gains"""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    t = torch.randn(tuple(shape), generator=g)
    if name.endswith("original0"):          # weight-norm gain
        return 1.0 + 0.1 * t
    if len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        return t / np.sqrt(fan_in)
    if name.endswith(".net.1.weight"):      # BatchNorm gamma
        return 1.0 + 0.2 * t
    return 0.1 * t


def state_dict_from_table(names, shapes):
    sd = {}
    for n, sh in zip(names, shapes):
        n = str(n)
        sh = tuple(int(v) for v in sh if v >= 0)
        if n.endswith("running_mean"):
            sd[n] = torch.zeros(sh)
        elif n.endswith("running_var"):
            sd[n] = torch.ones(sh)
        elif n.endswith("num_batches_tracked"):
            sd[n] = torch.zeros((), dtype=torch.long)
        else:
            sd[n] = det_tensor(n, sh)
    return sd


def inputs():
    g = torch.Generator().manual_seed(17)
    tf = [torch.rand(2, 1, bins, frames, generator=g) * 3 for bins, frames in ((257, 24), (513, 12), (1025, 6))]
    pf = [t * (0.7 + 0.6 * torch.rand(t.shape, generator=g)) for t in tf]
    ta = 0.1 * torch.randn(2, 3072, generator=g)
    pa = ta + 0.03 * torch.randn(2, 3072, generator=g)
    return tf, pf, ta, pa


def curve_inputs():
    g = torch.Generator().manual_seed(23)
    return torch.randn(2, 2, 50, generator=g), 5 * torch.rand(2, 1, 18, generator=g)


def build_reference():
    from oracle import ref_loader
    ref_loader.load()
    from stylish_tts.train.models.discriminator import SpecDiscriminator, ContextFreeDiscriminator
    from stylish_tts.train.models.pitch_discriminator import PitchDiscriminator
    torch.manual_seed(SEED)
    mods = dict(mrd0=SpecDiscriminator(), mrd1=SpecDiscriminator(), mrd2=SpecDiscriminator(),
                disc=ContextFreeDiscriminator(),
                pitch_disc=PitchDiscriminator(dim_in=2, dim_hidden=64, kernel=21),   # models.py:81-82
                dur_disc=PitchDiscriminator(dim_in=1, dim_hidden=64, kernel=5))
    tables = {}
    for key, m in mods.items():
        sd = m.state_dict()
        names = list(sd)
        shapes = [list(sd[n].shape) for n in names]
        m.load_state_dict(state_dict_from_table([f"{n}" for n in names], shapes), strict=True)
        tables[key] = (names, shapes)
    return mods, tables


def main():
    from oracle import ref_loader
    ref_loader.load()
    from stylish_tts.train.losses import DiscriminatorLoss, GeneratorLoss
    mods, tables = build_reference()
    for m in mods.values():
        m.train()
    tf, pf, ta, pa = inputs()
    blob = {}
    for key, (names, shapes) in tables.items():
        blob[key + "_names"] = np.array(names)
        pad = max(len(s_) for s_ in shapes)
        blob[key + "_shapes"] = np.array([list(s_) + [-1] * (pad - len(s_)) for s_ in shapes])
    with torch.no_grad():
        for i in range(3):
            outs, _ = mods[f"mrd{i}"](tf[i])
            for j, o in enumerate(outs):
                blob[f"mrd{i}_out{j}"] = o.numpy()
        blob["disc_out"] = mods["disc"](ta)[0][0].numpy()
        pc, du = curve_inputs()
        for j, o in enumerate(mods["pitch_disc"](pc)[0]):
            blob[f"pitch_disc_out{j}"] = o.numpy()
        for j, o in enumerate(mods["dur_disc"](du)[0]):
            blob[f"dur_disc_out{j}"] = o.numpy()
    kw = dict(mrd0=mods["mrd0"], mrd1=mods["mrd1"], mrd2=mods["mrd2"], disc=mods["disc"], pitch=mods["pitch_disc"],
              duration=mods["dur_disc"])
    gl, dl = GeneratorLoss(**kw), DiscriminatorLoss(**kw)
    args = dict(target_list=tf, pred_list=pf, target_audio=ta, pred_audio=pa, used=["mrd0", "mrd1", "mrd2", "disc"],
                index=0)
    pa_g = pa.clone().requires_grad_(True)
    pf_g = [p.clone().requires_grad_(True) for p in pf]
    g = gl(**{**args, "pred_list": pf_g, "pred_audio": pa_g})
    g.backward()
    blob["gen_loss"] = np.array(float(g.detach()))
    blob["gen_d_pred_audio"] = pa_g.grad.numpy()
    blob["gen_d_pred_fft0"] = pf_g[0].grad.numpy()
    pc, du = curve_inputs()
    blob["pitch_gen_loss"] = np.array(float(gl(target_list=[pc], pred_list=[pc * 0.9 + 0.1], target_audio=None,
                                               pred_audio=None, used=["pitch_disc"], index=0).detach()))
    blob["dur_disc_loss"] = np.array(float(dl(target_list=[du], pred_list=[du * 1.1 - 0.2], target_audio=None,
                                              pred_audio=None, used=["dur_disc"], index=0).detach()))
    d = dl(**args)
    for m in mods.values():
        m.zero_grad()
    d.backward()
    blob["disc_loss"] = np.array(float(d.detach()))
    w = mods["mrd1"].discriminators[2].parametrizations.weight.original1
    blob["disc_d_mrd1_conv2_v_norm"] = np.array(float(w.grad.norm()))
    blob["disc_d_last2_w"] = mods["disc"].last[2].weight.grad.numpy()
    blob["last_loss_mrd0"] = np.array(dl.discriminators["mrd0"].last_loss)
    blob["lr_mult_mrd0"] = np.array(dl.get_disc_lr_multiplier("mrd0"))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "discriminators.npz")
    np.savez_compressed(path, **blob)
    print(sorted(blob), os.path.getsize(path) // 1024, "KiB", float(g), float(d))


if __name__ == "__main__":
    main()
