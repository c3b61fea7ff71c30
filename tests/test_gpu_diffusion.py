"""Diffusion style sampler (BASELINE configs[3]) — **restatement, parity unpinned** (no implementation exists in
the reference: SURVEY F2).  The CUDA path (TMA-fed tcgen05 GEMMs with bf16 hi|lo planes, token-major flash attention,
LayerNorm with fused time-embedding add) against the fp64 evaluation of oracle/diffusion_oracle.py — the only parity
available for this component is self-consistency; tolerance = the north-star's 1e-3 relative."""
import math

import pytest
import torch

from oracle import diffusion_oracle as do
from stylish_tts_b200 import _lib as L
from stylish_tts_b200 import diffusion as DF
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def planes(x):
    return DF._planes(x)


@pytest.mark.parametrize("M,N,K,act,res", [(128, 256, 64, 0, False), (256, 768, 192, 5, False), (384, 256, 1024, 0, True),
                                           (16512, 1536, 1024, 0, False)])
def test_gemm_split_vs_fp64(M, N, K, act, res):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    bias = torch.randn(N, generator=g) * 0.1
    r = torch.randn(M, N, generator=g) if res else None
    ref = a.double() @ w.double().t() + bias.double()
    if act == 5:
        ref = torch.nn.functional.gelu(ref)
    if res:
        ref = ref + r.double()
    d = dev()
    out = torch.empty(M, N, device=d)
    outp = torch.empty(2, M, N, device=d, dtype=torch.bfloat16)
    ap, wp, bd = planes(a.to(d)), planes(w.to(d)), bias.to(d)  # keep the operands alive across the call
    rd = None if r is None else r.to(d)
    L.call("sty_gemm_split_fwd", ap.data_ptr(), wp.data_ptr(), bd.data_ptr(), L.ptr(rd), out.data_ptr(),
           outp.data_ptr(), M, N, K, act, L.stream_ptr())
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 2e-5, rel_l2(out, ref)
    recon = outp[0].float() + outp[1].float()
    assert rel_l2(recon, out) < 1e-5  # the output planes carry the fp32 result to ~2^-17


def seeded_denoiser(seed=0):
    torch.manual_seed(seed)
    m = DF.StyleDenoiser()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("norm.weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape))
            elif p.dim() == 1 and "fourier" not in n:
                p.copy_(0.05 * torch.randn(p.shape))
    return m


@pytest.mark.parametrize("B,T", [(3, 50), (2, 258)])
def test_denoiser_network_vs_fp64_oracle(B, T):
    m = seeded_denoiser(B)
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, 256, generator=g)
    emb = torch.randn(B, T, 768, generator=g)
    c_noise = torch.randn(B, generator=g) * 0.5
    ref = do.network(sd, x.double(), c_noise.double(), emb.double())
    md = m.to(dev())
    out = md(x.to(dev()), c_noise.to(dev()), emb.to(dev()))
    torch.cuda.synchronize()
    assert out.shape == (B, 256)
    assert rel_l2(out, ref) < 1e-3, rel_l2(out, ref)
    print("denoiser network vs fp64 oracle:", rel_l2(out, ref))


def test_sampler_vs_fp64_oracle():
    """ADPM2 over the Karras schedule, 5 steps (8 denoiser evaluations), identical noise in both arms"""
    B, T, steps = 4, 40, 5
    m = seeded_denoiser(7)
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(B, 256, generator=g)
    emb = torch.randn(B, T, 768, generator=g)
    step_noise = [torch.randn(B, 256, generator=g) for _ in range(steps - 1)]
    ref = do.adpm2_sample(sd, noise.double(), emb.double(), steps, [s.double() for s in step_noise])
    sig = DF.karras_sigmas(steps)
    ref_sig = do.karras_sigmas(steps)
    assert all(abs(a - float(b)) <= 1e-12 * max(1.0, abs(a)) for a, b in zip(sig, ref_sig))
    d = dev()
    sampler = DF.DiffusionSampler(m.to(d))
    out = sampler(noise.to(d), embedding=emb.to(d), num_steps=steps, step_noise=[s.to(d) for s in step_noise])
    torch.cuda.synchronize()
    assert out.shape == (B, 256) and torch.isfinite(out).all()
    print("sampler vs fp64 oracle:", rel_l2(out, ref))
    assert rel_l2(out, ref) < 1e-3, rel_l2(out, ref)
