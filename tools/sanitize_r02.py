"""Round-2 kernels at small sizes for compute-sanitizer (memcheck): forward (TMA convs, fused ConvNeXt block,
second-generation attention, register-tiled iSTFT head, side-stream prior branch), one acoustic training step with ALL
adversarial terms (mrd0-2 + the waveform discriminator through AdversarialTerms, tcgen05 attention backward,
fold-free data gradients, gapped-layout convs) and the discriminator half-step, and a 3-step diffusion sampling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import stylish_tts_b200 as st
from stylish_tts_b200 import synth, optim, train_step as ts, discriminator as D, diffusion as DF

dev = torch.device("cuda:0")
mc = st.default_model_config()
nets = st.build_model(mc)
synth.randomize_(nets.speech_predictor, 0)
synth.converge_spectral_(nets.speech_style_encoder)
for k in list(nets):
    nets[k] = nets[k].to(dev)
sp, se = nets.speech_predictor, nets.speech_style_encoder
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
keys = ["mrd0", "mrd1", "mrd2", "disc"]
for k in keys:
    nets[k].train()
dopts = {k: optim.FlatAdamW(nets[k].parameters(), world_size=1) for k in keys}
adv = D.AdversarialTerms(mrd0=nets.mrd0, mrd1=nets.mrd1, mrd2=nets.mrd2, disc=nets.disc, device=dev)
out = ts.acoustic_step(batch, nets, fe, generator_loss=adv)
out.total.backward()
opt.step(); opt.zero_grad()
d_loss = ts.discriminator_step(out, batch, adv, dopts, disc_index=1, lr_source=opt)
torch.cuda.synchronize()
print("adversarial train step ok", float(out.mel), float(out.multi_phase), float(out.generator), float(d_loss))
torch.manual_seed(0)
m = DF.StyleDenoiser().to(dev)
sampler = DF.DiffusionSampler(m)
g = torch.Generator(device=dev).manual_seed(1)
noise = torch.randn(3, 256, device=dev, generator=g)
emb = torch.randn(3, 70, 768, device=dev, generator=g)
sn = [torch.randn(3, 256, device=dev, generator=g) for _ in range(2)]
s = sampler(noise, embedding=emb, num_steps=3, step_noise=sn)
torch.cuda.synchronize()
print("diffusion ok", tuple(s.shape), bool(torch.isfinite(s).all()))
