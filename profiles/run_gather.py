"""Micro-driver for ncu: runs the hash-gather kernels on the boundary points of one real 65 536-ray chunk of the bench
scene (so the access pattern is the one the step sees).  Usage (on the GPU box):
    ncu --set full --clock-control none --import-source on -k regex:k_fused_sdf_tc -s 2 -c 1 -o gpurun_out/prof_gather python profiles/run_gather.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model(dev).train()
o, d = bench.pinhole_rays(bench.H, bench.W, bench.orbit(0, 8))
sel = slice(200 * 800, 200 * 800 + 65536)                      # rows 200..281: through the object
o, d = o[sel].to(dev).contiguous(), d[sel].to(dev).contiguous()
rt = model.ray_test(o, d, near=0.01)
R = rt["num_rays"]
t = rt["near"][:, None] + (rt["far"] - rt["near"])[:, None] * torch.linspace(0, 1, 116, device=dev)[None, :]
ridx = torch.arange(R, device=dev)
print("rays", R, "points", t.numel())
surf = model.implicit_surface
reps = int(os.environ.get("REPS", 5))
with torch.no_grad():
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sdf = surf.fused_sdf_rays(ridx, t, rt["rays_o"], rt["rays_d"])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"fused_sdf_tc: {t.numel()/dt/1e9:.2f} Gpts/s  {t.numel()*512/dt/1e9:.0f} GB/s algorithmic")
x = torch.addcmul(rt["rays_o"][:, None, :], rt["rays_d"][:, None, :], t[..., None]).reshape(-1, 3)
w = torch.randn(x.shape[0], device=dev)
w[torch.rand_like(w) < 0.8] = 0
for i in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = surf.fused_sdf_autograd(x)
    (s * w).sum().backward()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"fused fwd+bwd: {x.shape[0]/dt/1e9:.2f} Gpts/s")
