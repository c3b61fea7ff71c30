"""GPU-box: BASELINE configs[3] — diffusion style sampler, 10 steps (18 denoiser evaluations), batch 64 styles,
258-token context; CUDA-graph replay timed with CUDA events.  python tools/bench_diffusion.py [--iters N] [--eager]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stylish_tts_b200 import diffusion as DF, _lib as L


def measure(B=64, T=258, steps=10, iters=5, graph=True, dev=None):
    dev = dev or torch.device("cuda:0")
    torch.manual_seed(0)
    m = DF.StyleDenoiser().to(dev)
    sampler = DF.DiffusionSampler(m)
    g = torch.Generator(device=dev).manual_seed(1)
    noise = torch.randn(B, 256, device=dev, generator=g)
    emb = torch.randn(B, T, 768, device=dev, generator=g)
    step_noise = [torch.randn(B, 256, device=dev, generator=g) for _ in range(steps - 1)]
    run = lambda: sampler(noise, embedding=emb, num_steps=steps, step_noise=step_noise)
    before = L.launches
    out = run()
    launches = L.launches - before
    torch.cuda.synchronize()
    gr = None
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = run()
    fn = gr.replay if gr is not None else run
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    evals = 2 * (steps - 1)
    M = B * T
    flops_eval = 3 * 2.0 * M * (1024 * 1536 + 512 * 1024 + 2 * 1024 * 2048) + 3 * 4.0 * B * 8 * T * T * 64
    return {"workload": f"configs[3]: diffusion style sampler, {steps} steps ({evals} denoiser evaluations), batch {B} styles, "
                        f"{T}-token context; restatement, parity unpinned (SURVEY F2)",
            "ms_per_sample_batch": ms, "styles_per_s": B / (ms / 1e3), "ms_per_eval": ms / evals,
            "logical_tflops": flops_eval * evals / (ms / 1e3) / 1e12,
            "mma_tflops_bf16x3": 3 * flops_eval * evals / (ms / 1e3) / 1e12,
            "c_abi_calls_per_sample": launches, "cuda_graph": gr is not None, "finite": bool(torch.isfinite(out).all())}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    print(json.dumps(measure(B=a.batch, iters=a.iters, graph=not a.eager)))
