"""Small forward + acoustic training step for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import stylish_tts_b200 as st
from stylish_tts_b200 import synth, optim, train_step as ts

dev = torch.device("cuda:0")
mc = st.default_model_config()
nets = st.build_model(mc)
synth.randomize_(nets.speech_predictor, 0)
synth.converge_spectral_(nets.speech_style_encoder)
sp, se = nets.speech_predictor.to(dev), nets.speech_style_encoder.to(dev)
B, Tn = 2, 45   # 45 tokens -> 140 frames -> S = 10500 (several 128-row tiles incl. a partial one)
inp = synth.speech_inputs(B, Tn, seed=4, ragged=True)
c = lambda t: t.to(dev)
sp.eval()
with torch.no_grad():
    a = sp(c(inp["texts"]), c(inp["text_lengths"]), c(inp["alignment"]), c(inp["pitch"]), c(inp["energy"]),
           c(inp["voiced"]), c(inp["style"]), c(inp["denormal_pitch"])).audio
torch.cuda.synchronize()
print("fwd ok", tuple(a.shape), float(a.abs().mean()))
sp.train(); se.train()
dur = torch.full((B, Tn), 3.0); dur[:, ::9] += 1.0
frames = int(dur[0].sum())
if frames % 2: dur[:, -1] += 1; frames += 1
pitch = torch.nn.functional.pad(inp["pitch"], (0, frames - inp["pitch"].shape[1]), mode="replicate")
batch = SimpleNamespace(audio_gt=(0.1 * torch.randn(B, frames * 300)).to(dev), text=c(inp["texts"]),
                        text_length=c(inp["text_lengths"]), pitch=c(pitch), alignment=c(dur.unsqueeze(1)))
fe = ts.FrontEnd(mc)
opt = optim.FlatAdamW(list(sp.parameters()) + list(se.parameters()), world_size=1)
out = ts.acoustic_step(batch, nets, fe)
out.total.backward()
opt.step()
torch.cuda.synchronize()
print("train step ok", float(out.mel), float(out.multi_phase))
