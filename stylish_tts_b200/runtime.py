"""CUDA-graph replay of the speech_predictor forward (launch-bound host loop ->
one graph launch per batch).  Shapes are static per graph; inputs are copied into
the graph's static buffers on the launching stream."""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib as L

INPUT_KEYS = ("texts", "text_lengths", "alignment", "pitch", "energy", "voiced", "style",
              "denormal_pitch")


class GraphedSpeech:
    """Capture ``speech_predictor(**inputs)`` once, then replay.

    >>> g = GraphedSpeech(sp, example_inputs)   # tensors already on the GPU
    >>> audio = g(inputs)                       # (B,1,L) view of the static output
    """

    def __init__(self, sp, example: Dict[str, torch.Tensor], warmup: int = 2):
        self.sp = sp
        self.static = {k: example[k].clone() for k in INPUT_KEYS}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # packs weights, builds caches, sets func attributes
                sp(*[self.static[k] for k in INPUT_KEYS])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        from . import _lib
        before = _lib.launches
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = sp(*[self.static[k] for k in INPUT_KEYS]).audio
        self.launches_per_replay = _lib.launches - before

    def __call__(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        for k in INPUT_KEYS:
            self.static[k].copy_(inputs[k], non_blocking=True)
        self.graph.replay()
        return self.out

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.out


class GraphedAcousticStep:
    """One whole acoustic training iteration — forward graph, losses, backward, gradient all-reduce and the fused
    AdamW update — captured in ONE CUDA graph (≈4 400 kernel launches per replay).  Inputs are copied into
    static buffers; the learning rate / step count live in device memory (`FlatAdamW.hyper`), so replays keep
    advancing the optimizer.  Shapes are fixed at capture time."""

    def __init__(self, nets, frontend, optimizer, example_batch, *, warmup: int = 3, source_draws=None,
                 adversarial=None):
        """adversarial = (GeneratorLoss, DiscriminatorLoss, {key: FlatAdamW of mrd0..2}): the iteration then also
        carries the generator's adversarial term and the discriminator half-step (stage.py:116-146).  The reference
        draws the stepped discriminator with random.randrange(3) per batch, so one graph per index is captured
        (shared memory pool) and picked at replay."""
        import random
        from types import SimpleNamespace
        from .train_step import acoustic_step, discriminator_step

        self.static = {k: v.clone() for k, v in vars(example_batch).items()}
        self.opt = optimizer
        batch = SimpleNamespace(**self.static)
        # stochastic regularisers: the mask seed and the decoder's smoothing filters live in device memory and
        # are renewed by begin_step() OUTSIDE the graph, before every replay
        self.train_graphs = [nets.speech_predictor.train_graph()]
        for g in self.train_graphs:
            g.auto_step = False

        self.adversarial = adversarial
        self._random = random

        def iteration(disc_index=0):
            if adversarial is None:
                out = acoustic_step(batch, nets, frontend, source_draws=source_draws)
            else:
                out = acoustic_step(batch, nets, frontend, source_draws=source_draws, generator_loss=adversarial[0])
            out.total.backward()
            optimizer.step()
            optimizer.zero_grad()
            terms = [out.total.detach(), out.mel.detach(), out.multi_phase.detach()]
            if adversarial is not None:
                terms.append(out.generator.detach())
                terms.append(discriminator_step(out, batch, adversarial[1], adversarial[2], disc_index=disc_index,
                                                lr_source=optimizer))
            return torch.stack(terms)

        # the warm-up iterations are real launches (they fill the device-constant caches and size the allocator), but
        # constructing the graph must not train: parameters, Adam moments, step / lr cells, module buffers (BatchNorm
        # running statistics, spectral-norm u / v) and the discriminators' moving averages are restored afterwards
        opts = [optimizer] + (list(adversarial[2].values()) if adversarial is not None else [])
        snaps = [o.state_snapshot() for o in opts]
        mods = [m for m in nets.values() if isinstance(m, torch.nn.Module)] if hasattr(nets, "values") else []
        bufs = [(b, b.detach().clone()) for m in mods for b in m.buffers()]
        ctl = getattr(adversarial[1], "lr_control", {}) if adversarial is not None else {}
        ctl_snap = {k: c.last_loss.clone() for k, c in ctl.items()}
        py_state = random.getstate()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # also fills the device-constant caches: no host copies while capturing
            for _ in range(max(warmup, 1)):
                self.begin_step()
                iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for o, sn in zip(opts, snaps):
                o.state_restore(sn)
            for b, c in bufs:
                b.copy_(c)
            for k, c in ctl.items():
                c.last_loss.copy_(ctl_snap[k])
        random.setstate(py_state)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()  # the warm-up's activations go back to the driver before the graph pool grows
        self.begin_step()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = L.launches
        with torch.cuda.graph(self.graph):
            self.losses = iteration(0)
        self.launches_per_replay = L.launches - before
        optimizer.step_count -= 1  # capturing launches nothing: only the warm-up iterations were real updates
        self.graphs, self.loss_bufs = [self.graph], [self.losses]
        if adversarial is not None:
            always = [adversarial[2]["disc"]] if "disc" in adversarial[2] else []  # `disc` is stepped with every index
            adversarial[2]["mrd0"].step_count -= 1
            for o in always:
                o.step_count -= 1
            for idx in (1, 2):  # one graph per stepped discriminator, same memory pool (replayed one at a time);
                # nothing index-specific is allocated lazily (all three discriminators run in every iteration)
                self.begin_step()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self.graph.pool()):
                    buf = iteration(idx)
                optimizer.step_count -= 1
                adversarial[2][f"mrd{idx}"].step_count -= 1
                for o in always:
                    o.step_count -= 1
                self.graphs.append(g)
                self.loss_bufs.append(buf)

    def begin_step(self):
        for g in self.train_graphs:
            if g.stochastic():
                g.begin_step()

    def __call__(self, batch=None):
        self.begin_step()
        if batch is not None:
            for k, v in vars(batch).items():
                self.static[k].copy_(v, non_blocking=True)
        idx = self._random.randrange(3) if self.adversarial is not None else 0  # stage.py:119
        self.graphs[idx].replay()
        self.opt.step_count += 1
        if self.adversarial is not None:
            self.adversarial[2][f"mrd{idx}"].step_count += 1
            if "disc" in self.adversarial[2]:
                self.adversarial[2]["disc"].step_count += 1
        L.param_epoch += 1
        return self.loss_bufs[idx]
