"""Time the loss front-end at config-3 size (B=32, 10 s): fwd + bwd of the three STFT resolutions
and the mel front-ends; prints ms and effective GB/s.  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stylish_tts_b200 import spectral as sp


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B, Ls = 32, 240000
    x = 0.1 * torch.randn(B, Ls, device="cuda")
    y = (0.9 * x + 0.01 * torch.randn_like(x)).requires_grad_(True)
    ms = sp.MultiSpectrogram(sample_rate=24000)
    stft_loss = sp.MultiResolutionSTFTLoss()
    for i, plan in enumerate(ms.plans):
        nf = plan.n_frames(Ls)
        t = timeit(lambda: plan.forward(x, want_mag=True, want_phase=True, want_mel=True))
        out_bytes = B * nf * (2 * plan.K + plan.n_mels) * 4
        print(f"spectrogram fwd n_fft={plan.n_fft}: {t:.3f} ms  out {out_bytes/1e6:.1f} MB -> {out_bytes/t/1e6:.0f} GB/s")
        dm = torch.randn(B, plan.n_mels, nf, device="cuda")
        dp = torch.randn(B, plan.K, nf, device="cuda")
        t = timeit(lambda: plan.backward(x, nf, None, dp, dm))
        print(f"spectrogram bwd n_fft={plan.n_fft}: {t:.3f} ms")

    def step():
        y.grad = None
        t_spec, p_spec, t_ph, p_ph, _, _ = ms(target=x, pred=y)
        loss = stft_loss(target_list=t_spec, pred_list=p_spec) + sp.multi_phase_loss(p_ph, t_ph)
        loss.backward()
    print(f"full loss front-end fwd+bwd (3 resolutions, target+pred): {timeit(step):.3f} ms")
    to_mel = sp.MelSpectrogram(n_mels=80, n_fft=2048, win_length=1200, hop_length=300, sample_rate=24000)
    print(f"calculate_mel style (2048/1200/300): {timeit(lambda: sp.calculate_mel(x, to_mel, -4.0, 4.0)):.3f} ms")
    to_mel = sp.MelSpectrogram(n_mels=80, n_fft=512, win_length=512, hop_length=300, sample_rate=24000)
    print(f"calculate_mel (512/512/300): {timeit(lambda: sp.calculate_mel(x, to_mel, -4.0, 4.0)):.3f} ms")
    # cuFFT-based torch path for comparison (library baseline)
    w = torch.hann_window(1024, device="cuda")
    print(f"torch.stft n_fft=1024 alone (cuFFT, library baseline): {timeit(lambda: torch.stft(x, 1024, 256, 1024, w, return_complex=True)):.3f} ms")


if __name__ == "__main__":
    main()
