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
