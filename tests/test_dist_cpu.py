"""Multi-process host logic of the data-parallel step on CPU (gloo, world size 2): ray sharding by rank and the single
all-reduce of the flat gradient buffer whose views are the parameters' .grad (bench.py `flat_grad_views`)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Softplus(beta=100.0), torch.nn.Linear(16, 3))
    flat, params = bench.flat_grad_views(model)
    assert flat.numel() == sum(p.numel() for p in params) and all(p.grad.data_ptr() >= flat.data_ptr() for p in params)
    # every rank renders its own view (weak scaling): different rays per rank, same weights
    o, d = bench.pinhole_rays(12, 16, bench.orbit(rank, world))
    x = torch.cat([o, d], -1)
    flat.zero_()
    model(x).square().mean().backward()                 # accumulates straight into the flat buffer
    local = flat.clone()
    dist.all_reduce(flat)                               # the one collective of a step
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(flat, sum(gathered), rtol=1e-6, atol=1e-7)
    assert torch.allclose(params[0].grad.flatten(), flat[:params[0].numel()])       # .grad are still views of the reduced buffer
    # strong scaling: ONE frame dealt to the ranks by rows (round-robin) -- the shards partition the frame, equal sizes, rows stay whole
    fo, fd = bench.pinhole_rays(bench.H, bench.W, bench.orbit(0, 8))
    so, sd = bench.strong_shard(fo, fd, rank, world, bench.H * bench.W, bench.H * bench.W // world)
    assert so.shape[0] == bench.H * bench.W // world
    assert torch.equal(so.view(-1, bench.W, 3), fo.view(bench.H, bench.W, 3)[rank::world])
    key = (sd * torch.tensor([1.0, 1e3, 1e6])).sum(-1).double()             # a fingerprint per ray direction
    keys = [torch.zeros_like(key) for _ in range(world)]
    dist.all_gather(keys, key)
    allk = torch.cat(keys).sort().values
    ref = (fd * torch.tensor([1.0, 1e3, 1e6])).sum(-1).double().sort().values
    assert torch.equal(allk, ref)                                            # every ray of the frame exactly once over the ranks
    ro_, rd_ = bench.strong_shard(fo, fd, rank, world, 4096, 4096 // world, seed=5)
    assert ro_.shape[0] == 4096 // world
    if rank == 0:
        out.put((float(flat.abs().sum()), [float(g.abs().sum()) for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_ray_shard_and_single_allreduce_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    total, parts = out.get(timeout=10)
    assert total > 0 and parts[0] != parts[1]           # the two ranks really saw different rays


def test_reference_arm_runs_on_rank0_only(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import bench
    class A: steps, warmup, ref_rays, gpus = 1, 0, 8, 2
    bench.run_reference(A, rank=1)                       # other ranks exit without work
    assert capsys.readouterr().out == ""
