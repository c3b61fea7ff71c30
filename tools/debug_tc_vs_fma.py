"""GPU-box diagnostic: per-parameter gradient difference tensor-core path vs fp32-FMA path vs fp64 oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stylish_tts_b200 as st
from stylish_tts_b200 import synth, engine as E
from oracle import speech_oracle as so
from tests import util
from tests.test_train_step import oracle_grads, cotangent, CASE
from tests.util import rel_l2

dev = torch.device("cuda:0")
tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 258
sp = st.build_model(st.default_model_config()).speech_predictor
synth.randomize_(sp, CASE["wseed"]); synth.condition_phase_head_(sp)
inp = synth.speech_inputs(2, tokens, seed=9, ragged=True)
taps = {}
with torch.no_grad():
    so.speech_predictor(util.state_dict_of(sp), inp["texts"], inp["text_lengths"], inp["alignment"], inp["pitch"], inp["energy"], inp["voiced"], inp["style"], inp["denormal_pitch"], inp["draws"], taps=taps)
prior = (taps["har_spec"], taps["har_phase"])
a64, g64, d64, _ = oracle_grads(sp, inp, torch.float64, prior=prior)
a32, g32, d32, _ = oracle_grads(sp, inp, torch.float32, prior=prior)
res = {}
for tc in (False, True):
    E.USE_UMMA = tc
    m = st.build_model(st.default_model_config()).speech_predictor
    m.load_state_dict(sp.state_dict()); m = m.to(dev).train(); m.regularisers = False
    c = lambda t: t.to(dev)
    style, pitch, energy = (c(inp[k]).clone().requires_grad_(True) for k in ("style", "pitch", "energy"))
    out = m(c(inp["texts"]), c(inp["text_lengths"]), c(inp["alignment"]), pitch, energy, c(inp["voiced"]), style, c(inp["denormal_pitch"]), prior=(c(prior[0]), c(prior[1])))
    (out.audio * c(cotangent(out.audio.shape))).sum().backward()
    torch.cuda.synchronize()
    res[tc] = {n: p.grad.double().cpu() for n, p in m.named_parameters() if p.grad is not None}
print(f"{'parameter':72s} cpu32  fma    tc     (rel-L2 vs fp64 oracle)")
for n in g64:
    if not n.startswith("generator"): continue
    if g64[n].numel() < 64: continue
    e = [rel_l2(x, g64[n]) for x in (g32[n], res[False][n], res[True][n])]
    flag = " <<<" if e[2] > 3 * max(e[0], e[1], 1e-4) else ""
    print(f"{n:72s} {e[0]:.1e} {e[1]:.1e} {e[2]:.1e}{flag}")
