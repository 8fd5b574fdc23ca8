"""A/B of the SDF query on the boundary points of one bench frame, ray-tiled order only: default (two levels per trip) vs four levels per
trip (NSB_SDF_VARIANT=5) at 4..6 CTAs / SM.  python profiles/ab_gather2.py > gpurun_out/ab2.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neuralsim_b200.graphics.raysample import batch_sample_step_linear  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model(dev).train()
o, d = bench.pinhole_rays(bench.H, bench.W, bench.orbit(0, 8))
rt = model.ray_test(o.to(dev), d.to(dev), near=0.01)
ro, rd, near, far = rt["rays_o"].contiguous(), rt["rays_d"].contiguous(), rt["near"].contiguous(), rt["far"].contiguous()
R = ro.shape[0]
coarse = batch_sample_step_linear(near, far, 65, prefix_shape=[R]).contiguous()
t = coarse.reshape(-1).contiguous()
pinfo = torch.stack([torch.arange(R, device=dev) * 65, torch.full((R,), 65, device=dev)], 1).contiguous()
surf = model.implicit_surface
fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for variant, ctas in ((1, 6), (5, 6), (5, 5), (5, 4), (1, 5), (1, 8)):
    os.environ["NSB_SDF_VARIANT"], os.environ["NSB_SDF_CTAS"] = str(variant), str(ctas)
    ts = []
    for i in range(6):
        fl.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        with torch.no_grad():
            s = surf.fused_sdf_rays(None, t, ro, rd, packs=(pinfo, None))
        b.record()
        torch.cuda.synchronize()
        if i:
            ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    if ref is None:
        ref = s
    print(f"variant {variant} ctas {ctas}: {ms:7.3f} ms  {t.numel() * 512 / ms / 1e6:7.0f} GB/s algorithmic   bit-equal to the default: {bool(torch.equal(s, ref))}", flush=True)
