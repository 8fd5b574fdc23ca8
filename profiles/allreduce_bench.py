"""Time of the step's one collective on this box: all-reduce (sum) of the flat fp32 gradient (12.15 M floats = 48.6 MB), CUDA events, max over ranks.
    NCCL_ALGO=... torchrun --nproc-per-node N profiles/allreduce_bench.py"""
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 12_147_270
for numel in (n, n // 8, 4 * n):
    x = torch.randn(numel, device=dev)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dist.all_reduce(x); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if dist.get_rank() == 0:
        w = dist.get_world_size()
        print(f"world {w}  {numel * 4 / 1e6:8.1f} MB  median {float(t):.3f} ms  busbw {2 * (w - 1) / w * numel * 4 / float(t) / 1e6:8.1f} GB/s  "
              f"NCCL_ALGO={os.environ.get('NCCL_ALGO', '-')} NCCL_PROTO={os.environ.get('NCCL_PROTO', '-')}", flush=True)
dist.destroy_process_group()
