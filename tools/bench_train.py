"""Time one training step of speech_predictor (config 3: B=32, 258 tokens, 803 frames):
forward + multi-resolution STFT/phase losses + backward, eager.  Prints ms and a per-kernel table."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stylish_tts_b200 as st
from stylish_tts_b200 import synth, spectral, _lib as L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--warm", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, 0)
    sp = sp.to(dev).train()
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synth.speech_inputs(a.batch, 258, seed=1).items()}
    noise = inp["draws"]["noise"].to(dev)
    ms = spectral.MultiSpectrogram(sample_rate=24000)
    stft_loss = spectral.MultiResolutionSTFTLoss()
    target = 0.1 * torch.randn(a.batch, inp["pitch"].shape[1] * 300, device=dev)

    def step():
        for p in sp.parameters():
            p.grad = None
        out = sp(inp["texts"], inp["text_lengths"], inp["alignment"], inp["pitch"], inp["energy"], inp["voiced"],
                 inp["style"], inp["denormal_pitch"], source_draws={"noise": noise})
        audio = out.audio.squeeze(1)
        t_spec, p_spec, t_ph, p_ph, _, _ = ms(target=target, pred=audio)
        mel = stft_loss(target_list=t_spec, pred_list=p_spec)
        ph = spectral.multi_phase_loss(p_ph, t_ph)
        total = 5.0 * mel / (mel.detach() + 1e-9) + 8.0 * ph / (ph.detach() + 1e-9)
        total.backward()
        return total

    for _ in range(a.warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / a.steps
    secs = a.batch * inp["pitch"].shape[1] * 300 / 24000
    print(f"train step B={a.batch}: {ms_step:.1f} ms  ({1000/ms_step:.2f} steps/s, {secs/ms_step*1000:.0f} audio-s/s trained), "
          f"peak mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB")
    if a.profile:
        L.profile_log = []
        step()
        torch.cuda.synchronize()
        agg = {}
        for sig, s0, s1, info in L.profile_log:
            t = s0.elapsed_time(s1)
            n, tt = agg.get(sig, (0, 0.0))
            agg[sig] = (n + 1, tt + t)
        L.profile_log = None
        tot = sum(v[1] for v in agg.values())
        print(f"sum of our kernel time: {tot:.1f} ms over {sum(v[0] for v in agg.values())} C-ABI calls")
        for sig, (n, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"  {100*tt/tot:6.2f}%  n={n:4d}  {tt:8.3f} ms  {sig}")
    # torch profiler view: how much is ATen glue
    if a.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))


if __name__ == "__main__":
    main()
