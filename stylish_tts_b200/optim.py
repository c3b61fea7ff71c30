"""Flat-arena AdamW + data-parallel gradient exchange for the training path.

The reference builds one ``torch.optim.AdamW`` per model key (optimizers.py:106-117: lr 1e-4,
weight_decay 1e-4, betas (0.85, 0.99), eps 1e-9) and lets HuggingFace accelerate wrap the modules in
DDP, i.e. a NCCL sum all-reduce of the gradient buckets followed by a division by the world size
(train_context.py:94-104).  Here every parameter of the wrapped modules lives in ONE flat fp32 arena
(``p.data`` are views into it), the gradients of a step are packed into a second arena, exchanged with a
single ``all_reduce`` over NVLink (the only collective of the data-parallel step) and applied
by one fused kernel launch (``sty_adamw_step``) that also folds in the 1/world averaging.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import _lib as L


class FlatAdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], *, lr=1e-4, betas=(0.85, 0.99), eps=1e-9,
                 weight_decay=1e-4, process_group=None, world_size: Optional[int] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdamW: no trainable parameters")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.group = process_group
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.world = world_size
        dev = self.params[0].device
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):  # re-home the parameters inside the arena
                self.flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + p.numel()].view(p.shape)
        if self.world > 1:
            # DDP broadcasts rank 0's parameters when it wraps a module (train_context.py:94-104 via accelerate);
            # here the arena IS the parameters, so one broadcast keeps the replicas identical from step 0
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.broadcast(self.flat, src=0 if process_group is None else dist.get_global_rank(process_group, 0),
                               group=process_group)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.step_count = 0
        # [lr, step] on the device: read by the update kernel, so a captured step keeps advancing (CUDA graphs)
        self.hyper = torch.tensor([lr, 0.0], device=dev, dtype=torch.float32)

    def state_snapshot(self):
        """copies of everything a step changes (parameters, moments, step / lr cell) — used to undo warm-up steps"""
        return (self.flat.clone(), self.m.clone(), self.v.clone(), self.hyper.clone(), self.step_count)

    def state_restore(self, snap):
        flat, m, v, hyper, self.step_count = snap
        self.flat.copy_(flat)
        self.m.copy_(m)
        self.v.copy_(v)
        self.hyper.copy_(hyper)
        L.param_epoch += 1

    # -- checkpointing: the layout of torch.optim.AdamW.state_dict(), which is what the reference's per-key optimizers
    #    write through accelerate.save_state (train.py:453-469) — a checkpoint of either side resumes on the other
    def state_dict(self) -> dict:
        state = {}
        if self.step_count > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.m[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.v[o:o + n].view(p.shape).clone()}
        group = {"lr": float(self.hyper[0]), "betas": tuple(self.betas), "eps": self.eps,
                 "weight_decay": self.weight_decay, "amsgrad": False, "maximize": False, "foreach": None,
                 "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": True,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError(f"FlatAdamW.load_state_dict: {len(order)} parameters in the checkpoint, {len(self.params)} here")
        g0 = groups[0]
        self.betas, self.eps, self.weight_decay = tuple(g0["betas"]), g0["eps"], g0["weight_decay"]
        step = 0
        with torch.no_grad():
            self.m.zero_()
            self.v.zero_()
            for slot, (p, o) in zip(order, zip(self.params, self.offsets)):
                st = sd["state"].get(slot)
                if st is None:
                    continue
                n = p.numel()
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"FlatAdamW.load_state_dict: moment shape {tuple(st['exp_avg'].shape)} != {tuple(p.shape)}")
                self.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                self.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                step = max(step, int(float(st["step"])))
            self.step_count = step
            self.hyper[1] = float(step)
        self.set_lr(float(g0["lr"]))

    def set_lr(self, lr: float):
        self.lr = lr
        self.hyper[0] = lr

    def schedule(self, step: int, step_limit: int, base_lr: Optional[float] = None) -> float:
        """the reference's per-batch scheduler call (batch_manager.py:239 -> optimizers.py:96-103): cosine on the
        10 000-logical-step clock with the 90 % plateau, from the stage's base rate (the constructor's `lr` unless
        given); the new rate is written into the device cell the update kernel reads"""
        if base_lr is not None:
            self.base_lr = base_lr
        elif not hasattr(self, "base_lr"):
            self.base_lr = self.lr
        lr = cosine_lr(self.base_lr, step, step_limit)
        self.set_lr(lr)
        return lr

    # -- gradient plumbing (works on any device; the gloo tests exercise it on CPU) --------------
    def pack_gradients(self) -> torch.Tensor:
        """copy the per-parameter ``.grad`` tensors into the flat arena (missing gradients = 0, like
        DDP's find_unused_parameters for ``m_source.l_linear``)"""
        views = []
        for p in self.params:
            g = p.grad
            views.append(torch.zeros(p.numel(), device=self.grad.device, dtype=torch.float32) if g is None
                         else g.detach().reshape(-1))
        torch.cat(views, out=self.grad)
        return self.grad

    def reduce_gradients(self) -> torch.Tensor:
        """sum over ranks; the 1/world factor is applied inside the update kernel"""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)
        return self.grad

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    # -- the update (CUDA only) -------------------------------------------------------------------
    def step(self):
        if not self.flat.is_cuda:
            raise RuntimeError("stylish_tts_b200: FlatAdamW.step needs CUDA parameters (no CPU fallback)")
        self.pack_gradients()
        self.reduce_gradients()
        self.step_count += 1
        self.hyper[1:2].add_(1.0)
        L.call("sty_adamw_step_dev", self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(),
               self.v.data_ptr(), self.numel, self.hyper.data_ptr(), self.betas[0], self.betas[1], self.eps,
               self.weight_decay, 1.0 / self.world, L.stream_ptr())
        L.param_epoch += 1  # the arena changed under the views: invalidates the engines' packed weights


def acoustic_losses(audio_pred, audio_target, multi_spectrogram, stft_loss, *, w_mel=5.0, w_phase=8.0,
                    return_fft=False):
    """mel + multi_phase terms of the acoustic stage with LossLog.backwards_loss normalisation
    (stage_type.py:170-193, loss_log.py:82-94, weights config.yml:73-107).  -> (total, mel, phase)
    [+ (target |X| list, predicted |X| list) — the discriminators' inputs, stage_type.py:208-219]"""
    from .spectral import multi_phase_loss

    t_spec, p_spec, t_ph, p_ph, t_fft, p_fft = multi_spectrogram(target=audio_target, pred=audio_pred)
    mel = stft_loss(target_list=t_spec, pred_list=p_spec)
    ph = multi_phase_loss(p_ph, t_ph)
    total = w_mel * mel / (mel.detach() + 1e-9) + w_phase * ph / (ph.detach() + 1e-9)
    if return_fft:
        return total, mel, ph, t_fft, p_fft
    return total, mel, ph


# ---------------------------------------------------------------------------------------------------------
# learning-rate policy of the reference (optimizers.py:54-65,96-103; losses.py:229-250)
# ---------------------------------------------------------------------------------------------------------
LOGICAL_STEP_LIMIT = 10000  # optimizers.py:10: every stage is mapped onto a 10 000-"logical-step" cosine
PLATEAU = 0.9               # optimizers.py:98: the cosine stops decaying after 90 % of the stage


def cosine_lr(base_lr: float, step: int, step_limit: int) -> float:
    """MultiOptimizer.scheduler (optimizers.py:96-103) + transformers.get_cosine_schedule_with_warmup(0 warm-up,
    10 000 steps, half a cycle): the stage's `step` of `step_limit` is mapped to a logical step, capped at the
    plateau, and the scheduler is evaluated one past it (``scheduler.step()`` increments before it evaluates)."""
    import math

    logical = step * LOGICAL_STEP_LIMIT // step_limit
    logical = min(logical, LOGICAL_STEP_LIMIT * PLATEAU)
    progress = float(logical + 1) / float(max(1, LOGICAL_STEP_LIMIT))
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))


class DiscriminatorLR:
    """Gap-aware discriminator learning rate (DiscriminatorLossHelper, losses.py:229-250; applied by
    MultiOptimizer.step_discriminator_schedulers, optimizers.py:54-65): lr_disc = lr_gen * f(last_loss) with
    ``last_loss`` an exponential moving average of the LSGAN discriminator loss.  Everything lives in DEVICE
    tensors (no ``.item()`` like losses.py:287), so it can sit inside a captured training iteration:
    ``update(loss)`` folds a new loss in, ``apply(gen_opt, disc_opt)`` writes the discriminator's learning rate
    into its ``FlatAdamW.hyper`` cell."""

    def __init__(self, sub_count: int, device="cpu"):
        self.ideal = 0.5 * sub_count
        self.f_max, self.h_min = 4.0, 0.01
        self.x_max = self.x_min = 0.05 * sub_count
        self.last_loss = torch.full((), self.ideal, device=device, dtype=torch.float32)

    def update(self, disc_loss: torch.Tensor) -> None:
        self.last_loss.mul_(0.95).add_(disc_loss.detach().to(self.last_loss.dtype), alpha=0.05)

    def multiplier(self) -> torch.Tensor:
        last, ideal = self.last_loss, self.ideal
        x = (last - ideal).abs()
        up = torch.clamp(torch.pow(torch.full_like(x, self.f_max), x / self.x_max), max=self.f_max)
        down = torch.clamp(torch.pow(torch.full_like(x, self.h_min), x / self.x_min), min=self.h_min)
        mid = torch.where(last > ideal, up, down)
        return torch.where(last > ideal + self.x_max, torch.full_like(x, self.f_max),
                           torch.where(last < ideal - self.x_min, torch.full_like(x, self.h_min), mid))

    def apply(self, gen_opt: "FlatAdamW", disc_opt: "FlatAdamW") -> None:
        disc_opt.hyper[0:1].copy_(gen_opt.hyper[0:1] * self.multiplier())
