"""One eager forward of BASELINE configs[1] (B=16, 258 tokens, 803 frames) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python tools/ncu_fwd_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stylish_tts_b200 as st
from stylish_tts_b200 import synth

dev = torch.device("cuda:0")
sp = st.build_model(st.default_model_config()).speech_predictor
synth.randomize_(sp, 0)
sp = sp.to(dev).eval()
inp = synth.speech_inputs(16, 258, seed=1)
args = [inp[k].to(dev) for k in ("texts", "text_lengths", "alignment", "pitch", "energy", "voiced", "style", "denormal_pitch")]
with torch.no_grad():
    for _ in range(2):
        sp(*args)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    sp(*args)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
