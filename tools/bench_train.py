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
    from types import SimpleNamespace
    from stylish_tts_b200 import train_step as ts, optim
    mc = st.default_model_config()
    nets = st.build_model(mc)
    sp, se = nets.speech_predictor, nets.speech_style_encoder
    synth.randomize_(sp, 0)
    synth.converge_spectral_(se)
    sp, se = sp.to(dev).train(), se.to(dev).train()
    opt = optim.FlatAdamW(list(sp.parameters()) + list(se.parameters()), world_size=1)
    fe = ts.FrontEnd(mc)
    inp = synth.speech_inputs(a.batch, 258, seed=1)
    dur = torch.full((a.batch, 258), 3.0)
    dur[:, ::9] += 1.0
    dur[:, -1] += 1.0
    frames = int(dur[0].sum())
    pitch = torch.cat([inp["pitch"], inp["pitch"][:, -1:]], 1)
    batch = SimpleNamespace(audio_gt=(0.1 * torch.randn(a.batch, frames * 300)).to(dev), text=inp["texts"].to(dev),
                            text_length=inp["text_lengths"].to(dev), pitch=pitch.to(dev),
                            alignment=dur.unsqueeze(1).to(dev))

    def step():
        out = ts.acoustic_step(batch, nets, fe)
        out.total.backward()
        opt.step()
        opt.zero_grad()
        return out.total

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
    secs = a.batch * frames * 300 / 24000
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
