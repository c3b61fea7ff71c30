#!/usr/bin/env python
"""Benchmark of the speech_predictor forward hot path (BASELINE.json configs[1]:
single speaker, 256-char utterances (258 tokens with pads), batch 16, fwd-only).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one forward of a batch of 16 ten-second utterances (text -> wav).
Prints ONE JSON line (rank 0).  metric = audio-seconds synthesized per second.

ours:       `value` = device-resident inputs, CUDA-event timed graph replays;
            `e2e`   = pinned host inputs -> H2D -> forward -> D2H audio, per step;
            `roofline` = dominant kernel (by device time) of one event-profiled step;
            `cpu_baseline` = the CPU oracle (torch CPU port of the reference) on the
            box's host cores, bounded sample (rank 0, N=1 only).
reference:  the reference's CPU implementation of the path (oracle port: /root/reference
            cannot travel to the GPU box) on all host threads, one utterance per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-seconds synthesized/sec (fwd)"
UNIT = "audio-s/s"
SAMPLE_RATE = 24000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception as e:  # nvidia-smi missing: record nothing
            log("clock sampler unavailable:", e)
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if not (t0 - 0.05 <= ts <= t1 + 0.25):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_forward_fn(batch, tokens, seed=1):
    """CPU oracle (torch CPU port of the reference forward) closure + audio seconds per call."""
    import torch
    import stylish_tts_b200 as st
    from stylish_tts_b200 import synth
    from oracle import speech_oracle as so

    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, 0)
    sd = {k: v.detach().clone() for k, v in sp.state_dict().items()}
    inp = synth.speech_inputs(batch, tokens, seed=seed)
    secs = batch * inp["alignment"].shape[2] * 300 / SAMPLE_RATE

    def run():
        with torch.no_grad():
            return so.speech_predictor(sd, inp["texts"], inp["text_lengths"], inp["alignment"],
                                       inp["pitch"], inp["energy"], inp["voiced"], inp["style"],
                                       inp["denormal_pitch"], inp["draws"])
    return run, secs


def cpu_baseline(tokens, budget_s=20.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, secs = oracle_forward_fn(1, tokens)
    run()  # warm-up
    best, spent, n = None, 0.0, 0
    while n < 3 and spent < budget_s:
        t = time.perf_counter()
        run()
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
        spent += dt
        n += 1
    return {"value": secs / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 utterance ({secs:.2f} audio-s, B=1 T={tokens}: the faster CPU configuration, "
                      f"SURVEY.md §6), best of {n} runs, torch CPU {torch.get_num_threads()} threads"}


def eager_gpu_baseline(batch, tokens, dev, iters=3):
    """The GPU-library bar: the SAME torch port of the reference forward, moved to the GPU
    (eager PyTorch: cuDNN / cuBLAS / ATen kernels), batch `batch`.  Bench-side only — the product path never
    touches it."""
    import torch
    import stylish_tts_b200 as st
    from stylish_tts_b200 import synth
    from oracle import speech_oracle as so

    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, 0)
    sd = {k: v.detach().clone().to(dev) for k, v in sp.state_dict().items()}
    inp = synth.speech_inputs(batch, tokens, seed=1)
    d = lambda t: t.to(dev)
    draws = {k: d(v) for k, v in inp["draws"].items()}
    args = [d(inp[k]) for k in ("texts", "text_lengths", "alignment", "pitch", "energy", "voiced", "style",
                                "denormal_pitch")]
    secs = batch * inp["alignment"].shape[2] * 300 / SAMPLE_RATE

    def run():
        with torch.no_grad(), torch.device(dev):
            return so.speech_predictor(sd, *args, draws)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"value": secs / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": batch,
            "what": "oracle/speech_oracle.py (torch port of the reference forward) on cuda:0 in eager PyTorch "
                    f"{torch.__version__}, default settings (TF32 off)"}


def config0_text_to_wav(dev, cpu=True):
    """BASELINE configs[0]: ExportModel.forward (export_model.py:40-63) on the sample_dataset's 10 phoneme strings
    plus one 258-token utterance, batch 1 each (the reference's inference plumbing), random-init weights, styles
    randn(1,64)*0.1 — ours (Synthesizer on the GPU, host tokens in / host audio out) next to the CPU port."""
    import torch
    import stylish_tts_b200 as st
    from stylish_tts_b200 import synth, text as T
    from oracle import speech_oracle as so

    mc = st.default_model_config()
    nets = st.build_model(mc)
    for i, k in enumerate(("duration_predictor", "pitch_energy_predictor", "speech_predictor")):
        synth.randomize_(nets[k], 30 + i)
    tc = T.TextCleaner(mc.symbol)
    g = torch.Generator().manual_seed(1)
    utts = [torch.tensor(tc(s)) for s in T.SAMPLE_PHONEMES]
    long_u = torch.randint(1, 178, (258,), generator=g)
    long_u[0] = long_u[-1] = 0
    utts.append(long_u)
    styles = [torch.randn(1, 64, generator=g) * 0.1 for _ in range(3)]
    sds = {k: {n: v.detach().clone() for n, v in nets[k].state_dict().items()}
           for k in ("duration_predictor", "pitch_energy_predictor", "speech_predictor")}
    syn = st.Synthesizer(speech_predictor=nets.speech_predictor.to(dev).eval(),
                         pitch_energy_predictor=nets.pitch_energy_predictor.to(dev).eval(),
                         duration_predictor=nets.duration_predictor.to(dev).eval())
    sty_d = [x.to(dev) for x in styles]

    def ours():
        secs = 0.0
        for u in utts:
            texts = u.unsqueeze(0).to(dev, non_blocking=True)
            lengths = torch.tensor([u.numel()], device=dev)
            audio = syn(texts, lengths, *sty_d).cpu()
            secs += audio.shape[-1] / SAMPLE_RATE
        return secs
    ours()
    torch.cuda.synchronize()
    t = time.perf_counter()
    secs = ours()
    dt = time.perf_counter() - t
    out = {"utterances": len(utts), "tokens": [int(u.numel()) for u in utts], "audio_s": round(secs, 2),
           "ours_audio_s_per_s": secs / dt, "ours_wall_s": dt}
    if cpu:
        def draws_fn(frames):
            return {"rand_ini": torch.rand(1, 9, generator=g), "noise": torch.randn(1, frames * 300, 9, generator=g)}
        t = time.perf_counter()
        secs_c = 0.0
        with torch.no_grad():
            for u in utts:
                a, _ = so.synthesize(sds, u.unsqueeze(0), torch.tensor([u.numel()]), styles[0], styles[1], styles[2],
                                     draws_fn)
                secs_c += a.shape[-1] / SAMPLE_RATE
        dtc = time.perf_counter() - t
        out.update(cpu_audio_s_per_s=secs_c / dtc, cpu_wall_s=dtc, cpu_cores=os.cpu_count())
    return out


def diffusion_sampler_bench(dev, iters=3):
    """BASELINE configs[3] (kernel isolation): 10-step diffusion style sampler, batch 64 styles, 258-token context.
    Restatement, parity unpinned (the reference has no implementation: SURVEY F2); CUDA-graph replay, CUDA events."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_diffusion", os.path.join(ROOT, "tools", "bench_diffusion.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.measure(B=64, T=258, steps=10, iters=iters, graph=True, dev=dev)
    r["tensor_pipe_note"] = ("sm__pipe_tensor_cycles_active of the GEMM launches: 72-74 % (q|kv, FF1, FF2), 28 % "
                             "(out-proj, K=512 with residual + two outputs): profiles/r02_ncu_full_gemm_split.txt")
    return r


def run_reference(args):
    """--impl reference: the reference's CPU implementation (oracle port) on host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if getattr(args, "mode", "fwd") == "train":
        return run_reference_train(args, cores)
    run, secs = oracle_forward_fn(1, args.tokens)
    for _ in range(max(1, min(args.warmup, 2))):
        run()
    steps = max(1, min(args.steps, 20))
    t = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t
    val = steps * secs / dt
    sample = (f"each step = 1 of the batch's {args.batch} utterances (B=1, T={args.tokens}, "
              f"{secs:.2f} audio-s) through the CPU port of the reference forward "
              f"(oracle/speech_oracle.py, pinned to the reference by tests/golden)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "configs[1]: single-speaker 24 kHz, 258-token utterances, fwd-only "
                               "(speech_predictor text->wav)", "batch_per_step": 1,
                   "tokens": args.tokens},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


TRAIN_METRIC = "train steps/sec (acoustic step: fwd+bwd + multi-res STFT/phase loss + AdamW)"


def run_reference_train(args, cores):
    """--impl reference --mode train: the acoustic step through the CPU oracles; each timed step is ONE utterance
    (a bounded sample), a batch-`train_batch` step is that time x the batch size"""
    batch = getattr(args, "train_batch", 32)
    run, secs = oracle_train_fn(args.tokens)
    for _ in range(max(1, min(args.warmup, 1))):
        run()
    steps = max(1, min(args.steps, 5))
    t = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t) / steps
    val = 1.0 / (dt * batch)
    sample = (f"each step = 1 utterance ({secs:.2f} audio-s) fwd + losses + bwd through the CPU oracles "
              f"({dt:.2f} s); a batch-{batch} step is extrapolated as {batch}x that")
    print(json.dumps({
        "impl": "reference", "metric": TRAIN_METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt * batch, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[2]: full train step, batch {batch}", "batch_per_step": 1,
                   "tokens": args.tokens},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def oracle_train_fn(tokens, seed=1):
    """CPU arm of the training step: oracle forward (batch-stat BN) + spectral losses + autograd backward
    for ONE utterance; returns (closure, audio seconds per call)."""
    import torch
    import stylish_tts_b200 as st
    from stylish_tts_b200 import synth
    from oracle import dropout_oracle as do, speech_oracle as so, spectral_oracle as spo, style_oracle as sto

    nets = st.build_model(st.default_model_config())
    sp, se = nets.speech_predictor, nets.speech_style_encoder
    synth.randomize_(sp, 0)
    synth.converge_spectral_(se)
    grad = lambda m: {k: (v.detach().clone().requires_grad_(k.split(".")[-1] not in ("weight_u", "weight_v")
                                                            and "running" not in k)
                          if v.is_floating_point() else v.clone()) for k, v in m.state_dict().items()}
    sd, sde = grad(sp), grad(se)
    inp = synth.speech_inputs(1, tokens, seed=seed)
    frames = inp["alignment"].shape[2]
    target = 0.1 * torch.randn(1, frames * 300, generator=torch.Generator().manual_seed(3))
    secs = frames * 300 / SAMPLE_RATE

    def run():
        for v in list(sd.values()) + list(sde.values()):
            v.grad = None
        with torch.no_grad():  # calculate_mel x2 + energy (stage_type.py:75-97)
            mel = spo.calculate_mel(target, n_fft=512, win=512, hop=300, n_mels=80, sample_rate=SAMPLE_RATE,
                                    mean=-4.0, std=4.0)
            style_mel = spo.calculate_mel(target, n_fft=2048, win=1200, hop=300, n_mels=80,
                                          sample_rate=SAMPLE_RATE, mean=-4.0, std=4.0)
            spo.log_energy(mel, -4.0, 4.0)
        style = sto.mel_style_encoder(sde, style_mel.unsqueeze(1), training=True)
        so.MASKS = do.Masks(17)  # train() mode like the GPU arm: dropout sites + decoder box smoothing live
        try:
            audio = so.speech_predictor(sd, inp["texts"], inp["text_lengths"], inp["alignment"], inp["pitch"],
                                        inp["energy"], inp["voiced"], style, inp["denormal_pitch"],
                                        inp["draws"], bn_training=True, smoothing=(7, 15))
        finally:
            so.MASKS = None
        ls = spo.acoustic_spectral_losses(target, audio.squeeze(1), SAMPLE_RATE)
        spo.backwards_total(ls, dict(mel=5.0, multi_phase=8.0)).backward()
    return run, secs


def measure_train(args, dev, world, rank, barrier, steps, warmup, batch):
    """config 3 / 5: one optimizer step of speech_predictor on `batch` 10-s utterances per GPU."""
    import torch
    import torch.distributed as dist
    import stylish_tts_b200 as st
    from stylish_tts_b200 import _lib, synth, spectral, optim

    from types import SimpleNamespace
    from stylish_tts_b200 import train_step as ts

    mc = st.default_model_config()
    nets = st.build_model(mc)
    sp, se = nets.speech_predictor, nets.speech_style_encoder
    synth.randomize_(sp, 0)
    synth.converge_spectral_(se)
    sp, se = sp.to(dev).train(), se.to(dev).train()
    sp.regulariser_seed = rank  # train() mode: dropout masks differ between the data-parallel ranks
    # acoustic stage: speech_predictor + speech_style_encoder are trained (stage_type.py:393-410)
    opt = optim.FlatAdamW(list(sp.parameters()) + list(se.parameters()), lr=1e-4, betas=(0.85, 0.99), eps=1e-9,
                          weight_decay=1e-4, world_size=world)
    fe = ts.FrontEnd(mc)
    adversarial = None
    if getattr(args, "adversarial", False):
        # the adversarial half of configs[2]: generator term on mrd0-2 + one discriminator stepped per batch
        from stylish_tts_b200 import discriminator as D
        mrd = [nets[f"mrd{i}"].to(dev).train() for i in range(3)]
        disc_opts = {f"mrd{i}": optim.FlatAdamW(mrd[i].parameters(), lr=1e-4, betas=(0.85, 0.99), eps=1e-9,
                                                weight_decay=1e-4, world_size=world) for i in range(3)}
        wave_disc = None
        if not getattr(args, "no_wave_disc", False):  # `disc`: ContextFreeDiscriminator on 1024-sample windows
            wave_disc = nets["disc"].to(dev).train()
            disc_opts["disc"] = optim.FlatAdamW(wave_disc.parameters(), lr=1e-4, betas=(0.85, 0.99), eps=1e-9,
                                                weight_decay=1e-4, world_size=world)
        if getattr(args, "adversarial_two_pass", False):  # the reference's literal schedule: 2 evaluations per batch
            adversarial = (D.GeneratorLoss(mrd0=mrd[0], mrd1=mrd[1], mrd2=mrd[2], disc=wave_disc),
                           D.DiscriminatorLoss(mrd0=mrd[0], mrd1=mrd[1], mrd2=mrd[2], disc=wave_disc, device=dev),
                           disc_opts)
        else:  # one evaluation of mrd0-2 per batch feeds both halves (same numbers: tests/test_gpu_discriminators.py)
            adv = D.AdversarialTerms(mrd0=mrd[0], mrd1=mrd[1], mrd2=mrd[2], disc=wave_disc, device=dev)
            adversarial = (adv, adv, disc_opts)
    host = synth.speech_inputs(batch, args.tokens, seed=11 + rank)
    dur = torch.full((batch, args.tokens), 3.0)
    dur[:, ::9] += 1.0  # the durations synth.speech_inputs builds its alignment from
    frames = int(dur[0].sum())
    pitch = host["pitch"]
    if frames % 2:  # calculate_mel keeps an even number of frames (utils.py:829-830): the data path's audio is
        dur[:, -1] += 1.0  # binned accordingly; make the synthetic utterance one frame longer
        frames += 1
        pitch = torch.cat([pitch, pitch[:, -1:]], 1)
    hb = dict(audio_gt=0.1 * torch.randn(batch, frames * 300, generator=torch.Generator().manual_seed(5 + rank)),
              text=host["texts"], text_length=host["text_lengths"], pitch=pitch, alignment=dur.unsqueeze(1))
    assert pitch.shape[1] == frames
    pinned = {k: v.pin_memory() for k, v in hb.items()}
    resident = {k: v.to(dev) for k, v in hb.items()}
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def train_step(b):
        import random
        bb = SimpleNamespace(**b)
        out = ts.acoustic_step(bb, nets, fe, generator_loss=adversarial[0] if adversarial else None)
        out.total.backward()  # source noise drawn on the device (reference too)
        opt.step()
        opt.zero_grad()
        if adversarial:
            ts.discriminator_step(out, bb, adversarial[1], adversarial[2], disc_index=random.randrange(3),
                                  lr_source=opt)
        return torch.stack([out.total.detach(), out.mel.detach(), out.multi_phase.detach()])

    graphed = None
    if not args.no_graph:
        try:  # the whole iteration (fwd, losses, bwd, all-reduce, AdamW) as one CUDA graph
            from stylish_tts_b200.runtime import GraphedAcousticStep
            graphed = GraphedAcousticStep(nets, fe, opt, SimpleNamespace(**resident), adversarial=adversarial)
        except Exception as e:  # still our kernels, launched eagerly
            log(f"[bench] train-step graph capture failed ({type(e).__name__}: {e}); eager launches")
            graphed = None
            torch.cuda.synchronize()
            opt.zero_grad()

    def step_resident():
        return graphed() if graphed is not None else train_step(resident)

    def step_e2e():
        if graphed is not None:
            loss_host.copy_(graphed(SimpleNamespace(**pinned))[:3], non_blocking=True)
        else:
            loss_host.copy_(train_step({k: v.to(dev, non_blocking=True) for k, v in pinned.items()}),
                            non_blocking=True)

    def timed(fn):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = _lib.launches
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.launches - before

    ms, launched = timed(step_resident)
    ms_e2e, _ = timed(step_e2e)
    torch.cuda.synchronize()
    if graphed is not None:
        launched = graphed.launches_per_replay * steps
    audio_s = batch * frames * 300 / SAMPLE_RATE
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    res = {
        "steps_per_s": steps / (ms / 1e3), "ms_per_step": ms / steps,
        "trained_audio_s_per_s": world * audio_s * steps / (ms / 1e3),
        "batch_per_gpu": batch, "global_batch": batch * world, "frames": frames,
        "e2e": {"steps_per_s": steps / (ms_e2e / 1e3), "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12},
        "gpu_launches": launched, "cuda_graph": graphed is not None,
        "grad_allreduce_bytes": opt.numel * 4 if world > 1 else 0,
        "params": opt.numel, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1),
        "loss": [round(float(x), 5) for x in loss_host.tolist()],
        "adversarial": ("generator term on mrd0-2 (LSGAN + TPRLS) + discriminator half-step (loss of all three, "
                        "mrd{random index} stepped, sqrt(B) scaling, gap-aware lr) from ONE evaluation of the "
                        "discriminators per batch (discriminator.AdversarialTerms); waveform discriminator `disc` "
                        + ("included (disc_weight 3, stepped every batch)" if "disc" in adversarial[2] else "not included"))
        if adversarial else "off",
        "scope": "AcousticStep(use_predicted_pe=False, predict_audio=True): calculate_mel x2 + energy + alignment "
                 "+ speech_style_encoder + speech_predictor (train() mode: batch-stat BN, dropout sites and decoder box smoothing live) + "
                 "MultiSpectrogram x3 + mel & multi-phase losses (backwards_loss normalisation) + backward of "
                 "both modules + fused AdamW" + ("; + adversarial terms (stage.py:104-146)" if adversarial
                                                 else "; adversarial terms measured separately (train_adversarial)"),
    }
    if not adversarial:
        # one event-profiled EAGER step: per-kernel device time of the training step and the roofline of its
        # dominant convolution call (the graph replays above are what `ms_per_step` measures).  EVERY rank runs the
        # two steps (they contain the gradient all-reduce); only rank 0 records events.
        try:
            train_step(resident)
            torch.cuda.synchronize()
            if rank == 0:
                _lib.profile_log = []
            train_step(resident)
            torch.cuda.synchronize()
            agg = {}
            for sig, e0, e1, info in (_lib.profile_log or []):
                d = agg.setdefault(sig, {"ms": 0.0, "n": 0, "info": info})
                d["ms"] += e0.elapsed_time(e1)
                d["n"] += 1
            _lib.profile_log = None
            total = sum(d["ms"] for d in agg.values()) or 1.0
            ranked = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
            res["top_kernels"] = [{"kernel": k, "share": round(v["ms"] / total, 4), "launches_per_step": v["n"],
                                   "avg_ms": round(v["ms"] / v["n"], 4)} for k, v in ranked[:12]]
            peak, how = measured_peaks()
            with_info = [(k, v) for k, v in ranked if v["info"] is not None]
            if with_info:
                k, v = with_info[0]
                avg_s = v["ms"] / v["n"] / 1e3
                byts = conv_alg_bytes(v["info"])
                res["roofline"] = {
                    "bound": "hbm", "kernel": k, "achieved": round(byts / avg_s / 1e9, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(byts / avg_s / 1e9 / peak, 4), "traffic": None, "peak_source": how,
                    "alg_bytes_per_launch": byts, "avg_launch_ms": round(avg_s * 1e3, 4),
                    "share_of_step": round(v["ms"] / total, 4),
                    "note": "dominant convolution call of one eager training step (input + output (+ residual) bytes / "
                            "CUDA-event time); the step's device time is spread over ~940 C-ABI calls (top_kernels)"}
        except Exception as e:
            _lib.profile_log = None
            log(f"[bench] train-step kernel profile failed: {type(e).__name__}: {e}")
            torch.cuda.synchronize()
    del graphed, opt, sp, se, nets
    torch.cuda.empty_cache()
    return res


def ncu_traffic(sig):
    """DRAM bytes per launch of the kernel from the committed `ncu --set full` captures (profiles/), if one
    matches this kernel signature; None otherwise."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        for key, val in j["traffic_bytes_per_launch"].items():
            if key in sig:
                return val
    except Exception:
        pass
    return None


def conv_alg_bytes(info):
    if info.get("kind") == "convnext_fused":  # read x twice (GRN recompute) + write y = 3 u
        return 3 * info["B"] * info["C"] * info["T"] * 4
    n = info["B"] * info["T"] * (info["CI"] + info["CO"]) * 4
    if info["res"]:
        n += info["B"] * info["T"] * info["CO"] * 4
    return n


def kernel_flops(info):
    if info.get("kind") == "convnext_fused":  # two pointwise GEMMs (the first one twice) + depthwise k7 (twice)
        return 2.0 * info["B"] * info["T"] * (3 * info["C"] * info["J"] + 2 * 7 * info["C"])
    return 2.0 * info["B"] * info["T"] * info["CI"] * info["CO"] * info["K"]


def ncu_traffic_r02(sig):
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        for key, val in j["traffic_bytes_per_launch"].items():
            if key in sig:
                return val
    except Exception:
        pass
    return ncu_traffic(sig)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import stylish_tts_b200 as st
    from stylish_tts_b200 import _lib, synth
    from stylish_tts_b200.runtime import GraphedSpeech, INPUT_KEYS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # a rank that dies or skips a collective must abort the job in minutes, not after the 10-minute default
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, T = args.batch, args.tokens
    sp = st.build_model(st.default_model_config()).speech_predictor
    synth.randomize_(sp, 0)
    sp = sp.to(dev).eval()
    host = synth.speech_inputs(B, T, seed=1 + rank)
    frames = host["alignment"].shape[2]
    audio_s = B * frames * 300 / SAMPLE_RATE
    pinned = {k: host[k].pin_memory() for k in INPUT_KEYS}
    resident = {k: host[k].to(dev) for k in INPUT_KEYS}

    graph = None
    if not args.no_graph:
        try:
            graph = GraphedSpeech(sp, resident)
        except Exception as e:  # still our kernels, launched eagerly
            log(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); eager launches")
            graph = None
            torch.cuda.synchronize()

    def step_resident():
        if graph is not None:
            return graph.replay()
        with torch.no_grad():
            return sp(*[resident[k] for k in INPUT_KEYS]).audio

    out_host = torch.empty((B, 1, frames * 300), dtype=torch.float32).pin_memory()
    h2d = sum(pinned[k].numel() * pinned[k].element_size() for k in INPUT_KEYS)
    d2h = out_host.numel() * 4

    # e2e: EVERY step copies its inputs from pinned host memory and reads its audio back to pinned host memory, all
    # inside the timed region.  The copies are software-pipelined like a serving loop would: the H2D of step i+1
    # (copy-in stream, into device staging buffers) and the D2H of step i (copy-out stream) run on the DMA engines
    # while step i / i+1 computes; the hand-over to / from the graph's static buffers is a device-to-device copy on the
    # compute stream.  `--e2e-serial` keeps everything on one stream (copy in -> forward -> copy out).
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    stage_in = {k: torch.empty_like(resident[k]) for k in INPUT_KEYS}
    stage_out = torch.empty((B, 1, frames * 300), device=dev, dtype=torch.float32)
    ev = {k: torch.cuda.Event() for k in ("h2d", "taken", "ready", "copied")}
    pipe = {"started": False}

    def step_e2e_serial():
        if graph is not None:
            audio = graph(pinned)
        else:
            dv = [pinned[k].to(dev, non_blocking=True) for k in INPUT_KEYS]
            with torch.no_grad():
                audio = sp(*dv).audio
        out_host.copy_(audio, non_blocking=True)

    def step_e2e():
        main = torch.cuda.current_stream()
        if pipe["started"]:
            s_in.wait_event(ev["taken"])       # the previous step has read the staging inputs
        with torch.cuda.stream(s_in):
            for k in INPUT_KEYS:
                stage_in[k].copy_(pinned[k], non_blocking=True)
            ev["h2d"].record(s_in)
        main.wait_event(ev["h2d"])
        if graph is not None:
            for k in INPUT_KEYS:
                graph.static[k].copy_(stage_in[k], non_blocking=True)
            ev["taken"].record(main)
            audio = graph.replay()
        else:
            with torch.no_grad():
                audio = sp(*[stage_in[k] for k in INPUT_KEYS]).audio
            ev["taken"].record(main)
        if pipe["started"]:
            main.wait_event(ev["copied"])      # the previous step's D2H has read the staging output
        stage_out.copy_(audio, non_blocking=True)
        ev["ready"].record(main)
        s_out.wait_event(ev["ready"])
        with torch.cuda.stream(s_out):
            out_host.copy_(stage_out, non_blocking=True)
            ev["copied"].record(s_out)
        pipe["started"] = True

    def drain_e2e():  # the timed region ends when the last step's audio is in host memory
        if pipe["started"]:
            torch.cuda.current_stream().wait_event(ev["copied"])

    def timed(fn, steps, warmup, drain=None):
        for _ in range(warmup):
            fn()
        if drain is not None:
            drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = _lib.launches
        e0.record()
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launched = _lib.launches - before
        return float(t.item()), launched

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t_start = time.time()
    ms, launched = timed(step_resident, args.steps, max(args.warmup, 3))
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    if getattr(args, "e2e_serial", False):
        ms_e2e, _ = timed(step_e2e_serial, args.steps, max(args.warmup, 3))
    else:
        ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 3), drain=drain_e2e)
    if graph is not None:
        launched = graph.launches_per_replay * args.steps
    value = world * audio_s * args.steps / (ms / 1e3)
    e2e_value = world * audio_s * args.steps / (ms_e2e / 1e3)

    # ---- one event-profiled eager step: per-kernel device time -> roofline of the dominant one
    roofline, top = None, []
    if rank == 0:
        with torch.no_grad():
            sp(*[resident[k] for k in INPUT_KEYS])  # warm
            torch.cuda.synchronize()
            _lib.profile_log = []
            for _ in range(2):
                sp(*[resident[k] for k in INPUT_KEYS])
            torch.cuda.synchronize()
        agg = {}
        for sig, e0, e1, info in _lib.profile_log:
            d = agg.setdefault(sig, {"ms": 0.0, "n": 0, "info": info})
            d["ms"] += e0.elapsed_time(e1)
            d["n"] += 1
        _lib.profile_log = None
        total = sum(d["ms"] for d in agg.values())
        ranked = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
        top = [{"kernel": k, "share": round(v["ms"] / total, 4), "launches_per_step": v["n"] // 2,
                "avg_ms": round(v["ms"] / v["n"], 4)} for k, v in ranked[:12]]
        peak, how = measured_peaks()

        def roof(k, v):
            avg_s = v["ms"] / v["n"] / 1e3
            byts = conv_alg_bytes(v["info"])
            ach = byts / avg_s / 1e9
            return {"bound": "hbm", "kernel": k, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": ncu_traffic_r02(k), "peak_source": how,
                    "alg_bytes_per_launch": byts, "avg_launch_ms": round(avg_s * 1e3, 4),
                    "share_of_step": round(v["ms"] / total, 4),
                    "fp32_tflops": round(kernel_flops(v["info"]) / avg_s / 1e12, 2)}

        with_info = [(k, v) for k, v in ranked if v["info"] is not None]
        if with_info:
            roofline = roof(*with_info[0])
            roofline["note"] = (
                "dominant C-ABI call of the step.  convnext_fused_fwd = the whole GeneratorConvNeXtBlock (pass 1 + GRN "
                "scale + pass 2, tcgen05 bf16x3, TMA-fed) charged with SURVEY 8(d)'s 3 u (x read twice, y written once; "
                "the 4C intermediate never leaves the SM) — it is bound by instruction issue / MUFU in the Snake "
                "epilogue, not by HBM; conv1d[...+umma] = tcgen05 bf16x3 conv charged with its own input + output "
                "(+ residual) bytes.  achieved = algorithmic bytes / CUDA-event time of the call")
            roofline["others"] = [roof(k, v) for k, v in with_info[1:6]]
        log(f"[bench] per-kernel device time of one step (event-timed, eager; total {total / 2:.2f} ms):")
        for k, v in ranked:
            log(f"    {v['ms'] / total * 100:6.2f}%  n={v['n'] // 2:3d}  avg {v['ms'] / v['n']:8.4f} ms  {k}")

    cpu, config0 = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline(T)
        except Exception as e:
            log("[bench] cpu baseline failed:", e)
        try:
            if cpu is not None:
                cpu["eager_gpu"] = eager_gpu_baseline(B, T, dev)
        except Exception as e:
            log(f"[bench] eager-GPU baseline failed: {type(e).__name__}: {e}")
            torch.cuda.synchronize()
        try:
            config0 = config0_text_to_wav(dev)
        except Exception as e:
            log(f"[bench] configs[0] failed: {type(e).__name__}: {e}")
        torch.cuda.empty_cache()
    diffusion = None
    if rank == 0 and world == 1:
        try:
            diffusion = diffusion_sampler_bench(dev)
        except Exception as e:
            log(f"[bench] configs[3] failed: {type(e).__name__}: {e}")
        torch.cuda.empty_cache()

    train = train_adv = None
    had_graph = graph is not None
    if not args.no_train:
        had_graph, graph = graph is not None, None  # free the captured graph's memory pool first
        torch.cuda.empty_cache()
        try:
            train = measure_train(args, dev, world, rank, barrier, max(3, min(args.steps, 5)), 3,
                                  args.train_batch)
        except Exception as e:
            log(f"[bench] train measurement failed: {type(e).__name__}: {e}")
        try:  # the same step with the adversarial terms of the acoustic stage (mrd0-2)
            import copy
            adv_args = copy.copy(args)
            adv_args.adversarial = True
            torch.cuda.empty_cache()
            train_adv = measure_train(adv_args, dev, world, rank, barrier, 3, 3, args.train_batch)
        except Exception as e:
            log(f"[bench] adversarial train measurement failed: {type(e).__name__}: {e}")

    if rank == 0:
        peak_hbm, _ = measured_peaks()
        extra_cfg = {
            # 58 MB of algorithmic HBM traffic per audio-second of the whole forward path
            "path_hbm_frac_survey_8d": round(value / world * 58e6 / (peak_hbm * 1e9), 4),
            "host_cores": os.cpu_count(),
        }
        if train is not None:  # configs[2] / configs[4] figures where the driver keeps them
            extra_cfg.update(train_ms_per_step=round(train["ms_per_step"], 3),
                             train_steps_per_s=round(train["steps_per_s"], 4),
                             train_audio_s_per_s=round(train["trained_audio_s_per_s"], 1),
                             train_global_batch=train["global_batch"],
                             train_grad_allreduce_bytes=train["grad_allreduce_bytes"],
                             train_e2e_ms_per_step=round(train["e2e"]["ms_per_step"], 3))
        if train_adv is not None:
            extra_cfg.update(train_adversarial_ms_per_step=round(train_adv["ms_per_step"], 3),
                             train_adversarial_steps_per_s=round(train_adv["steps_per_s"], 4))
        if config0 is not None:
            extra_cfg.update(config0_ours_audio_s_per_s=round(config0["ours_audio_s_per_s"], 1),
                             config0_cpu_audio_s_per_s=round(config0.get("cpu_audio_s_per_s", 0.0), 2))
        if cpu is not None and "eager_gpu" in cpu:
            extra_cfg["eager_gpu_audio_s_per_s"] = round(cpu["eager_gpu"]["value"], 1)
        if diffusion is not None:
            extra_cfg.update(config3_diffusion_ms_per_64_styles=round(diffusion["ms_per_sample_batch"], 2),
                             config3_diffusion_mma_tflops=round(diffusion["mma_tflops_bf16x3"], 1))
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: single-speaker 24 kHz (model.yml), 258-token "
                                   "utterances, batch 16, fwd-only speech_predictor text->wav",
                       "batch_per_gpu": B, "tokens": T, "frames": frames,
                       "audio_s_per_step_per_gpu": audio_s, "parallelism": f"dp{world} (no collective)",
                       "weights": "random-init (seeded), reference architecture",
                       "cuda_graph": had_graph,
                       "l2": "no flush needed: per-step working set (~6 GB of activations) >> 126 MB L2",
                       **extra_cfg},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "how": ("one stream: copy in -> forward -> copy out" if getattr(args, "e2e_serial", False) else
                            "three-stream pipeline: every step's H2D (pinned -> staging) and D2H (staging -> pinned) "
                            "inside the timed region, overlapped with the neighbouring steps' compute; the region "
                            "ends when the last step's audio is in host memory")},
            "gpu_launches": launched,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "top_kernels": top,
            "train": train,
            "train_adversarial": train_adv,
            "config0": config0,
            "diffusion": diffusion,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """--mode train: BASELINE configs[2] (1 GPU) / configs[4] (DDP, batch 32 per GPU, gradient all-reduce)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # a rank that dies or skips a collective must abort the job in minutes, not after the 10-minute default
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t0 = time.time()
    tr = measure_train(args, dev, world, rank, barrier, args.steps, max(args.warmup, 3), args.train_batch)
    clocks = sampler.stop(t0, time.time()) if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            run, secs = oracle_train_fn(args.tokens)
            t = time.perf_counter()
            run()
            dt = time.perf_counter() - t
            cpu = {"value": 1.0 / (dt * args.train_batch), "unit": "steps/s", "cores": cores, "kind": "port",
                   "sample": f"1 utterance ({secs:.2f} audio-s) fwd+loss+bwd through the CPU oracle took {dt:.2f} s; "
                             f"a batch-{args.train_batch} step is extrapolated as {args.train_batch}x that"}
        except Exception as e:
            log("[bench] cpu train baseline failed:", e)
    if rank == 0:
        e2e = tr.pop("e2e")
        print(json.dumps({
            "metric": TRAIN_METRIC, "value": tr["steps_per_s"], "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": tr["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("configs[2]: full train step, batch 32, 1xB200" if world == 1 else
                                    "configs[4]: DDP train, batch 32/GPU, 10 s utterances, NCCL gradient all-reduce"),
                       "batch_per_gpu": tr["batch_per_gpu"], "global_batch": tr["global_batch"],
                       "tokens": args.tokens, "frames": tr["frames"], "parallelism": f"dp{world}",
                       "l2": "no flush needed: per-step working set (~28 GB) >> 126 MB L2"},
            "e2e": {"value": e2e["steps_per_s"], "unit": "steps/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": e2e["d2h_bytes_per_step"], "ms_per_step": e2e["ms_per_step"]},
            "gpu_launches": tr["gpu_launches"], "clocks": clocks, "cpu_baseline": cpu, "train": tr,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--tokens", type=int, default=258)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true",
                    help="e2e on one stream (copy in -> forward -> copy out) instead of the three-stream pipeline")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mode", default="fwd", choices=["fwd", "train"],
                    help="fwd: configs[1] (the headline line, plus a short `train` sub-measurement); "
                         "train: configs[2]/[4] as the line itself")
    ap.add_argument("--train-batch", type=int, default=32)
    ap.add_argument("--no-train", action="store_true", help="skip the train sub-measurement of --mode fwd")
    ap.add_argument("--no-wave-disc", action="store_true",
                    help="with --adversarial: leave the waveform discriminator `disc` out (mrd0-2 only)")
    ap.add_argument("--adversarial-two-pass", action="store_true",
                    help="with --adversarial: evaluate the discriminators twice per batch like the reference's schedule")
    ap.add_argument("--adversarial", action="store_true",
                    help="train measurement with the adversarial terms (spectrogram discriminators mrd0-2)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
