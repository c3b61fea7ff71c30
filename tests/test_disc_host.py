"""Host-side pieces of the discriminators that need no GPU: the weight transforms that turn strided / grouped
convolutions into stride-1 dense ones (checked against torch's own strided / grouped conv1d / conv2d on the CPU) and
state-dict compatibility of the drop-in modules with the reference's classes."""
import os

import pytest
import torch
import torch.nn.functional as F

from stylish_tts_b200 import discriminator as D

REF = "/root/reference"


@pytest.mark.parametrize("K,s", [(11, 4), (7, 2), (5, 2), (3, 1), (9, 3)])
def test_strided_weight_equals_strided_conv(K, s):
    """Conv1d(stride s, padding K//2) == stride-1 'same' conv with strided_weight() on the space-to-depth input
    (ContextFreeDiscriminator's four strided layers, discriminator.py:124-132)"""
    g = torch.Generator().manual_seed(K * 10 + s)
    Cc, Co, T = 3, 5, 24 * s
    x = torch.randn(2, Cc, T, generator=g, dtype=torch.float64)
    w = torch.randn(Co, Cc, K, generator=g, dtype=torch.float64)
    ref = F.conv1d(x, w, stride=s, padding=K // 2)
    xs = x.reshape(2, Cc, T // s, s).permute(0, 1, 3, 2).reshape(2, Cc * s, T // s)  # channel c*s + p = phase p of c
    ws = D.strided_weight(w, s)
    out = F.conv1d(xs, ws, padding=ws.shape[2] // 2)
    assert out.shape == ref.shape and torch.allclose(out, ref, atol=1e-12)
    w.requires_grad_(True)  # plain tensor ops: the gradient reaches the original kernel
    D.strided_weight(w, s).square().sum().backward()
    assert torch.allclose(w.grad, 2 * w.detach())


def test_stride2_weight_equals_strided_conv2d():
    """SpecDiscriminator's stride-(1,2) 3x9 layers on the space-to-depth input (discriminator.py:24-40)"""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 7, 30, generator=g, dtype=torch.float64)
    w = torch.randn(6, 4, 3, 9, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=(1, 2), padding=(1, 4))
    xs = x.reshape(2, 4, 7, 15, 2).permute(0, 1, 4, 2, 3).reshape(2, 8, 7, 15)
    out = F.conv2d(xs, D.stride2_weight(w), padding=(1, 2))
    assert out.shape == ref.shape and torch.allclose(out, ref, atol=1e-12)


def test_grouped_as_dense_equals_grouped_conv():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 32, 10, generator=g, dtype=torch.float64)
    for co, k in ((16, 3), (96, 1), (32, 7)):
        w = torch.randn(co, 4, k, generator=g, dtype=torch.float64)
        dense = D.grouped_as_dense(w, 8)
        assert dense.shape == (co, 32, k)
        assert torch.allclose(F.conv1d(x, dense, padding=k // 2), F.conv1d(x, w, padding=k // 2, groups=8), atol=1e-12)


def test_gap_layout_constants():
    """the end-to-end window layout of ContextFreeDiscriminator: every level's pitch divides by the next stride, the
    data fraction is the same at every level (one bn_frac), the gap covers every kernel's reach"""
    P, Tw = D.ContextFreeDiscriminator.PITCH, D.ContextFreeDiscriminator.DATA
    strides, reach = (4, 4, 2, 2), (2, 2, 2, 1)  # J = ceil((K//2)/s) of the strided kernels at the OUTPUT rate
    for lvl, (s, j) in enumerate(zip(strides, reach)):
        assert P[lvl] == s * P[lvl + 1] and Tw[lvl] == s * Tw[lvl + 1]
        assert P[lvl + 1] - Tw[lvl + 1] >= j
    assert len({Tw[i] / P[i] for i in range(5)}) == 1
    assert P[4] - Tw[4] >= 3 and P[4] % 4 == 0   # k7 temporal conv; 16-byte rows for the segment kernels


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container)")
def test_state_dicts_match_the_reference_classes():
    """every discriminator drop-in loads the reference module's state dict with strict=True and vice versa"""
    from oracle import ref_loader

    ref_loader.load()
    from stylish_tts.train.models.discriminator import ContextFreeDiscriminator, SpecDiscriminator
    from stylish_tts.train.models.pitch_discriminator import PitchDiscriminator

    pairs = [(SpecDiscriminator(), D.SpecDiscriminator()), (ContextFreeDiscriminator(), D.ContextFreeDiscriminator()),
             (PitchDiscriminator(dim_in=2, dim_hidden=64, kernel=21), D.PitchDiscriminator(dim_in=2, dim_hidden=64, kernel=21))]
    for ref, ours in pairs:
        rs, os_ = ref.state_dict(), ours.state_dict()
        assert list(rs) == list(os_), (type(ref).__name__, set(rs) ^ set(os_))
        assert all(rs[k].shape == os_[k].shape and rs[k].dtype == os_[k].dtype for k in rs)
        ours.load_state_dict(rs, strict=True)
        ref.load_state_dict(ours.state_dict(), strict=True)
