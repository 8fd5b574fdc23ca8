"""A/B timing of the fused SDF query variants on the REAL boundary points of one bench frame (the 28.6 M coarse + fine
samples of the autograd query, in the order the step issues them).  python profiles/ab_gather.py > gpurun_out/ab.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neuralsim_b200.graphics import neus_fused as NF  # noqa: E402
from neuralsim_b200.graphics.raysample import batch_sample_step_linear  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model(dev).train()
o, d = bench.pinhole_rays(bench.H, bench.W, bench.orbit(0, 8))
rt = model.ray_test(o.to(dev), d.to(dev), near=0.01)
ro, rd, near, far = rt["rays_o"].contiguous(), rt["rays_d"].contiguous(), rt["near"].contiguous(), rt["far"].contiguous()
R = ro.shape[0]
coarse = batch_sample_step_linear(near, far, 65, prefix_shape=[R]).contiguous()
ridx = torch.arange(R, device=dev).unsqueeze(-1).expand(R, 65).reshape(-1).contiguous()
t = coarse.reshape(-1).contiguous()
print("rays", R, "points", t.numel(), flush=True)
surf = model.implicit_surface
fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(tag, reps=4):
    ts = []
    for i in range(reps + 1):
        fl.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        with torch.no_grad():
            s = surf.fused_sdf_rays(ridx, t, ro, rd, packs=PACKS)
        b.record()
        torch.cuda.synchronize()
        if i:
            ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    print(f"{tag:42s} {ms:8.3f} ms  {t.numel() / ms / 1e6:6.2f} Gpts/s  {t.numel() * 512 / ms / 1e6:7.0f} GB/s algorithmic", flush=True)
    return s


pinfo = torch.stack([torch.arange(R, device=dev) * 65, torch.full((R,), 65, device=dev)], 1).contiguous()
PACKS = None


def run_packs(tag):
    global PACKS
    PACKS = (pinfo, None)
    s = run(tag)
    PACKS = None
    return s


ref = None
for variant, name in ((1, "sfu softplus, 2 levels/trip (default)"), (4, "sfu, 2 levels/trip, PAIRED corner loads (experiment)"),
                      (0, "libm softplus, 2 levels/trip")):
    for ctas in (6, 7):
        os.environ["NSB_SDF_VARIANT"], os.environ["NSB_SDF_CTAS"] = str(variant), str(ctas)
        s = run(f"{name}, {ctas} CTA/SM")
        s2 = run_packs(f"{name}, {ctas} CTA/SM, RAY-TILED")
        print("    tiled == ray-major:", bool(torch.equal(s, s2)), flush=True)
        if ref is None:
            ref = s
        else:
            dd = (s - ref).abs()
            print(f"    vs variant 0: max |d| {float(dd.max()):.3e}, differing {float((dd > 0).float().mean()) * 100:.3f} %", flush=True)
# point order: the same points, pixel-patch-major (8x4 pixel tiles x 65 depths) instead of ray-major
os.environ["NSB_SDF_VARIANT"], os.environ["NSB_SDF_CTAS"] = "2", "6"
perm = torch.randperm(t.numel(), device=dev)
ridx_r, t_r = ridx[perm].contiguous(), t[perm].contiguous()
ridx, t = ridx_r, t_r
run("variant 0, random point order")
ridx2 = torch.arange(R, device=dev).unsqueeze(0).expand(65, R).reshape(-1).contiguous()         # depth-major: lanes = adjacent rays
t2 = coarse.t().reshape(-1).contiguous()
ridx, t = ridx2, t2
run("variant 0, depth-major (lanes = adjacent rays)")
