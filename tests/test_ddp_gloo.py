"""world_size-2 gloo test (CPU) of the data-parallel host logic: the flat gradient arena is
what gets all-reduced, parameters/gradients stay views of the arenas, the average over ranks
equals the single-process gradient of the concatenated batch."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from one_to_many_gan_b200.optim import GradArena

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    arena = GradArena(net.parameters())
    assert arena.data_parallel and arena.world == world
    for p, off in zip(arena.params, arena.offsets):
        assert p.data_ptr() == arena.param_arena.data_ptr() + 4 * off
        assert p.grad.data_ptr() == arena.grad_arena.data_ptr() + 4 * off
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 5, generator=g)
    y = torch.randn(8, 3, generator=g)
    shard = slice(rank * 4, rank * 4 + 4)
    arena.zero_grad()
    torch.nn.functional.mse_loss(net(x[shard]), y[shard]).backward()
    # autograd accumulated in place: grads are still arena views
    for p, off in zip(arena.params, arena.offsets):
        assert p.grad.data_ptr() == arena.grad_arena.data_ptr() + 4 * off
    arena.all_reduce_async()
    arena.wait_all_reduce()
    avg = arena.grad_arena / world
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
        ref.load_state_dict(net.state_dict())
        torch.nn.functional.mse_loss(ref(x), y).backward()
        for p, off, q in zip(arena.params, arena.offsets, ref.parameters()):
            torch.testing.assert_close(avg[off : off + p.numel()].view_as(p), q.grad, rtol=1e-5, atol=1e-6)
        out.put("ok")
    arena.zero_grad()
    assert float(arena.grad_arena.abs().sum()) == 0.0
    # coalesced exchange over two arenas (engine: G decoder slice + S, then G encoder slice + M):
    # one collective, bucket bookkeeping per arena, the second call fills exactly the gaps
    from one_to_many_gan_b200.optim import all_reduce_buckets

    other = GradArena(torch.nn.Linear(4, 4).parameters())
    arena.grad_arena.fill_(float(rank + 1))
    other.grad_arena.fill_(10.0 * (rank + 1))
    cut = arena.offsets[2]
    all_reduce_buckets([(arena, cut, None), (other, 0, None)])
    all_reduce_buckets([(a, lo, hi) for a in (arena, other) for lo, hi in a._missing()])
    for a in (arena, other):
        a.wait_all_reduce()
    assert torch.all(arena.grad_arena == 3.0) and torch.all(other.grad_arena == 30.0)
    assert arena._issued == [] and other._pending == []
    dist.barrier()
    dist.destroy_process_group()


def test_grad_arena_all_reduce_two_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
