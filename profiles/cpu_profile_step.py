"""Host-side cost of a TRAINING-BATCH step (4096 random rays, launch / sync bound): cProfile of N steps.
    python profiles/cpu_profile_step.py > gpurun_out/cpu_prof.txt"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neuralsim_b200.renderer import SingleVolumeRenderer  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model(dev).train()
r = SingleVolumeRenderer(dict(near=0.01)).train()
flat, params = bench.flat_grad_views(model)
o, d = bench.pinhole_rays(bench.H, bench.W, bench.orbit(0, 8))
N = int(os.environ.get("RAYS", 4096))
sel = torch.randperm(o.shape[0], generator=torch.Generator().manual_seed(7))[:N]
o, d = o[sel].to(dev).contiguous(), d[sel].to(dev).contiguous()
ha = torch.zeros(N, 4, device=dev)


def step():
    flat.zero_()
    out = r.render(model, o, d, rays_h_appear=ha)["rendered"]
    loss = bench.loss_of(out)
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(100):
    step()
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) * 10:.3f} ms / step wall ({N} rays)")
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats("tottime").print_stats(45)
