"""GPU-box diagnostic: gradient error of speech_predictor vs the fp64 oracle as a function of utterance length."""
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
for tokens in [int(a) for a in sys.argv[1:]] or [40, 130, 258]:
    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, CASE["wseed"]); synth.condition_phase_head_(sp)
    inp = synth.speech_inputs(2, tokens, seed=9, ragged=True)
    taps = {}
    with torch.no_grad():
        so.speech_predictor(util.state_dict_of(sp), inp["texts"], inp["text_lengths"], inp["alignment"], inp["pitch"], inp["energy"], inp["voiced"], inp["style"], inp["denormal_pitch"], inp["draws"], taps=taps)
    prior = (taps["har_spec"], taps["har_phase"])
    t = time.time()
    a64, g64, d64, _ = oracle_grads(sp, inp, torch.float64, prior=prior)
    a32, g32, d32, _ = oracle_grads(sp, inp, torch.float32, prior=prior)
    print(f"== tokens {tokens} S={prior[0].shape[2]} oracle time {time.time()-t:.1f}s")
    tot64 = torch.cat([g64[n].flatten() for n in g64])
    tot32 = torch.cat([g32[n].flatten().double() for n in g64])
    print("  cpu fp32 oracle vs fp64: params", rel_l2(tot32, tot64), {k: rel_l2(d32[k], d64[k]) for k in d64})
    for tc in (False, True):
        E.USE_UMMA = tc
        m = st.build_model(st.default_model_config()).speech_predictor
        m.load_state_dict(sp.state_dict()); m = m.to(dev).train(); m.regularisers = False
        c = lambda t: t.to(dev)
        style, pitch, energy = (c(inp[k]).clone().requires_grad_(True) for k in ("style", "pitch", "energy"))
        out = m(c(inp["texts"]), c(inp["text_lengths"]), c(inp["alignment"]), pitch, energy, c(inp["voiced"]), style, c(inp["denormal_pitch"]), prior=(c(prior[0]), c(prior[1])))
        (out.audio * c(cotangent(out.audio.shape))).sum().backward()
        torch.cuda.synchronize()
        params = dict(m.named_parameters())
        tot = torch.cat([params[n].grad.flatten().double().cpu() for n in g64])
        print(f"  gpu tc={tc}: audio {rel_l2(out.audio, a64):.2e} params {rel_l2(tot, tot64):.2e}", {k: f"{rel_l2(t.grad, d64[k]):.2e}" for k, t in (("style", style), ("pitch", pitch), ("energy", energy))})
        scale = float(tot64.norm())
        errs = sorted(((float((params[n].grad.double().cpu() - g64[n]).norm()) / scale, n, float(g64[n].norm()) / scale) for n in g64), reverse=True)[:12]
        for e, n, s in errs:
            print(f"      {n:70s} err/|G| {e:.2e}  |g|/|G| {s:.2e}")
