"""Golden fixtures for the mel front-end and the STFT / phase losses, made by the UNMODIFIED
reference objects (run in the build container; needs /root/reference and torchaudio):

    python tests/golden/make_spectral_golden.py

Inputs are regenerated from seeds by ``spectral_inputs`` below (also imported by the tests), so
only outputs are stored: the reference's MultiSpectrogram lists, calculate_mel / energy, the two
losses, the normalised total (LossLog.backwards_loss weights mel 5, multi_phase 8 from
config/config.yml) and its gradient w.r.t. the predicted audio.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 24000
MEL_MEAN, MEL_STD = -4.0, 4.0
W_MEL, W_PHASE = 5.0, 8.0


def spectral_inputs(B=2, L=7200, seed=11):
    """target: noisy harmonic signal; pred: a perturbed copy (so phases are comparable)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L, dtype=torch.float64) / SAMPLE_RATE
    f0 = 110.0 + 40.0 * torch.rand(B, 1, generator=g, dtype=torch.float64)
    target = sum((0.3 / h) * torch.sin(2 * torch.pi * h * f0 * t) for h in range(1, 6))
    target = (target + 0.05 * torch.randn(B, L, generator=g, dtype=torch.float64)).float()
    pred = (0.9 * target + 0.05 * torch.randn(B, L, generator=g)).float()
    return target.contiguous(), pred.contiguous()


def main():
    from oracle import ref_loader

    ref_loader.load()
    import torchaudio
    from stylish_tts.train.multi_spectrogram import MultiSpectrogram
    from stylish_tts.train.losses import MultiResolutionSTFTLoss, multi_phase_loss
    from stylish_tts.train.utils import calculate_mel, log_norm

    class Log:
        def __init__(self):
            self.metrics = {}

        def add_loss(self, k, v):
            self.metrics[k] = v

    target, pred = spectral_inputs()
    pred.requires_grad_(True)
    ms = MultiSpectrogram(sample_rate=SAMPLE_RATE)
    t_spec, p_spec, t_ph, p_ph, t_fft, p_fft = ms(target=target, pred=pred)
    log = Log()
    mel_loss = MultiResolutionSTFTLoss(sample_rate=SAMPLE_RATE)(target_list=t_spec, pred_list=p_spec, log=log)
    ph_loss = multi_phase_loss(p_ph, t_ph, 512)
    total = W_MEL * mel_loss / (mel_loss.detach() + 1e-9) + W_PHASE * ph_loss / (ph_loss.detach() + 1e-9)
    total.backward()
    blob = dict(mel_loss=mel_loss.detach().numpy(), phase_loss=ph_loss.detach().numpy(),
                total=total.detach().numpy(), d_pred=pred.grad.numpy())
    for r in range(3):
        blob[f"t_spec{r}"] = t_spec[r].numpy()
        blob[f"p_spec{r}"] = p_spec[r].detach().numpy()
        blob[f"t_phase{r}"] = t_ph[r].numpy()
        blob[f"t_fft{r}"] = t_fft[r].numpy()
    # mel front-ends of train_context.py:155-169 (model.yml: 512/512/300 and 2048/1200/300, 80 mels)
    for name, (n_fft, win) in {"mel": (512, 512), "style_mel": (2048, 1200)}.items():
        to_mel = torchaudio.transforms.MelSpectrogram(n_mels=80, n_fft=n_fft, win_length=win, hop_length=300,
                                                      sample_rate=SAMPLE_RATE)
        mel, length = calculate_mel(target, to_mel, MEL_MEAN, MEL_STD)
        blob[name] = mel.numpy()
        blob[name + "_raw"] = to_mel(target).numpy()
        if name == "mel":
            e = log_norm(mel.unsqueeze(1), MEL_MEAN, MEL_STD).squeeze(1)
            blob["energy"] = torch.log(e + 1e-9).numpy()
    # the torchaudio filterbanks themselves (pins the restated fbank arithmetic)
    blob["fb_257_80"] = torchaudio.functional.melscale_fbanks(257, 0.0, 12000.0, 80, SAMPLE_RATE).numpy()
    blob["fb_1025_128"] = torchaudio.functional.melscale_fbanks(1025, 0.0, 12000.0, 128, SAMPLE_RATE).numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spectral.npz")
    np.savez_compressed(path, **blob)
    print({k: v.shape for k, v in blob.items()}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
