"""Kernel-time table of one bench step with torch.profiler (CUPTI); cheaper than an ncu launch list for iterating.
    python profiles/torch_profile_step.py > gpurun_out/torch_prof.txt"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neuralsim_b200.renderer import SingleVolumeRenderer  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model(dev).train()
r = SingleVolumeRenderer(dict(near=0.01)).train()
flat, params = bench.flat_grad_views(model)
o, d = bench.pinhole_rays(bench.H, bench.W, bench.orbit(0, 8))
o, d = o.to(dev), d.to(dev)
CH = int(os.environ.get("CHUNK", 480000))
ha = torch.zeros(CH, 4, device=dev)


def step():
    flat.zero_()
    for s in range(0, o.shape[0], CH):
        e = min(s + CH, o.shape[0])
        out = r.render(model, o[s:e], d[s:e], rays_h_appear=ha[:e - s])["rendered"]
        loss = bench.loss_of(out) * ((e - s) / o.shape[0])
        if loss.requires_grad:
            loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=70))
