"""Flat-arena AdamW and the data-parallel gradient exchange.

CPU (gloo, world_size 2): rank-0 parameter broadcast at construction, gradient packing + sum all-reduce of the flat
arena, unused parameters as zeros, identical result on both ranks.   GPU: the fused AdamW kernel against torch.optim.AdamW with the
reference's hyper-parameters (optimizers.py:106-117), including the 1/world gradient scaling."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stylish_tts_b200 import optim


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_params(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(shape, generator=g)) for shape in ((7, 3), (5,), (2, 3, 4), (1,))]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = _make_params(seed=rank)  # every rank starts from its OWN initialisation ...
        before = [p.detach().clone() for p in _make_params(seed=0)]
        opt = optim.FlatAdamW(params, world_size=world)
        for p, b in zip(params, before):  # ... re-homed into the arena and overwritten with rank 0's values (DDP's
            # parameter broadcast at wrap time, train_context.py:94-104)
            assert torch.equal(p.detach(), b) and p.data_ptr() >= opt.flat.data_ptr()
        for i, p in enumerate(params):
            if i == 3:
                continue  # unused parameter: no gradient on any rank
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        opt.pack_gradients()
        opt.reduce_gradients()
        expect = torch.cat([torch.full((p.numel(),), 3.0 * (i + 1) if i != 3 else 0.0) for i, p in enumerate(params)])
        assert torch.equal(opt.grad, expect), (opt.grad, expect)
        gathered = [torch.empty_like(opt.grad) for _ in range(world)]
        dist.all_gather(gathered, opt.grad)
        assert torch.equal(gathered[0], gathered[1])
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            opt.step()
        out.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(30)
    assert res == {0: "ok", 1: "ok"}, res


def test_state_dict_is_torch_adamw_compatible():
    """FlatAdamW.state_dict / load_state_dict speak torch.optim.AdamW's layout (what the reference's per-key
    optimizers write through accelerate.save_state, train.py:453-469): a reference checkpoint resumes here and one
    written here loads into torch.optim.AdamW (CPU: no update is launched)"""
    kw = dict(lr=1e-4, betas=(0.85, 0.99), eps=1e-9, weight_decay=1e-4)
    ref_params = _make_params(seed=3)
    ref = torch.optim.AdamW(ref_params, **kw)
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        for p in ref_params:
            p.grad = torch.randn(p.shape, generator=g)
        ref.step()
    sd = ref.state_dict()
    ours = optim.FlatAdamW(_make_params(seed=3), world_size=1, **kw)
    assert ours.state_dict()["state"] == {}                      # like a fresh torch optimizer
    ours.load_state_dict(sd)
    assert ours.step_count == 3 and float(ours.hyper[1]) == 3.0 and float(ours.hyper[0]) == pytest.approx(1e-4)
    for i, (p, o) in enumerate(zip(ours.params, ours.offsets)):
        n = p.numel()
        assert torch.equal(ours.m[o:o + n].view(p.shape), sd["state"][i]["exp_avg"])
        assert torch.equal(ours.v[o:o + n].view(p.shape), sd["state"][i]["exp_avg_sq"])
    back = ours.state_dict()
    fresh = torch.optim.AdamW(_make_params(seed=3), **kw)
    fresh.load_state_dict(back)                                   # torch accepts our layout
    for i in range(len(ref_params)):
        assert torch.equal(fresh.state_dict()["state"][i]["exp_avg_sq"], sd["state"][i]["exp_avg_sq"])
        assert float(fresh.state_dict()["state"][i]["step"]) == 3.0
    with pytest.raises(ValueError):
        optim.FlatAdamW(_make_params()[:2], world_size=1).load_state_dict(sd)


@pytest.mark.gpu
def test_resume_from_torch_adamw_checkpoint_continues_identically():
    """two steps with torch.optim.AdamW, checkpoint, then a third step here == the third step of torch"""
    d = torch.device("cuda:0")
    kw = dict(lr=1e-3, betas=(0.85, 0.99), eps=1e-9, weight_decay=1e-4)
    ref_params = [torch.nn.Parameter(p.detach().to(d)) for p in _make_params(seed=5)]
    ref = torch.optim.AdamW(ref_params, **kw)
    g = torch.Generator().manual_seed(2)
    grads = [[torch.randn(p.shape, generator=g).to(d) for p in ref_params] for _ in range(3)]
    for it in range(2):
        for p, gr in zip(ref_params, grads[it]):
            p.grad = gr.clone()
        ref.step()
    ours_params = [torch.nn.Parameter(p.detach().clone()) for p in ref_params]
    ours = optim.FlatAdamW(ours_params, world_size=1, **kw)
    ours.load_state_dict(ref.state_dict())
    for p, q, gr in zip(ref_params, ours_params, grads[2]):
        p.grad, q.grad = gr.clone(), gr.clone()
    ref.step()
    ours.step()
    torch.cuda.synchronize()
    for p, q in zip(ref_params, ours_params):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 4])
def test_fused_adamw_matches_torch(world):
    dev = torch.device("cuda:0")
    ref_p = [torch.nn.Parameter(p.detach().clone().double()) for p in _make_params(3)]
    ours = [torch.nn.Parameter(p.detach().clone().to(dev)) for p in _make_params(3)]
    ref = torch.optim.AdamW(ref_p, lr=1e-4, betas=(0.85, 0.99), eps=1e-9, weight_decay=1e-4)
    opt = optim.FlatAdamW(ours, lr=1e-4, betas=(0.85, 0.99), eps=1e-9, weight_decay=1e-4, world_size=1)
    opt.world = world  # exercise the 1/world scaling without a process group
    opt.reduce_gradients = lambda: opt.grad
    g = torch.Generator().manual_seed(4)
    for step in range(5):
        for r, o in zip(ref_p, ours):
            gr = torch.randn(r.shape, generator=g)
            r.grad = gr.double()
            o.grad = (gr * world).to(dev)  # what a sum all-reduce over `world` equal ranks would hold
        ref.step()
        opt.step()
        opt.zero_grad()
    torch.cuda.synchronize()
    for r, o in zip(ref_p, ours):
        assert o.grad is None
        assert float((o.detach().cpu().double() - r.detach()).abs().max()) < 2e-6
