"""PSNR of OUR 800x600 render against the render of the REFERENCE'S OWN KERNELS (oracle/_ref, op-by-op chain) for the same weights and rays
(BASELINE.json: "PSNR vs ref"; graphics/utils.py:89-105 of the reference: -10 log10(mean((x - y)^2))).  Not collected by pytest: run on a GPU box
    python profiles/psnr_vs_reference.py > gpurun_out/psnr.txt
Both renders happen in child processes of their own (the reference arm patches module-level back ends)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %(root)r)
import bench
if %(ref)r:
    assert bench.use_reference_cuda_kernels(), "oracle/_ref is not built"
from neuralsim_b200.renderer import SingleVolumeRenderer
dev = torch.device("cuda:0")
model = bench.build_model(dev).eval()
o, d = bench.pinhole_rays(bench.H, bench.W, bench.orbit(%(view)d, 8))
with torch.no_grad():
    out = SingleVolumeRenderer(dict(near=0.01)).eval().render(model, o.to(dev), d.to(dev), rays_h_appear=torch.zeros(o.shape[0], 4, device=dev))["rendered"]
torch.save({k: v.cpu() for k, v in out.items()}, %(path)r)
'''


def render(ref, view, path):
    code = CHILD % dict(root=ROOT, ref=ref, view=view, path=path)
    subprocess.run([sys.executable, "-c", code], check=True)


if __name__ == "__main__":
    import torch
    res = {}
    for view in (0, 3):
        a, b = f"/tmp/psnr_ours_{view}.pt", f"/tmp/psnr_ref_{view}.pt"
        render(False, view, a)
        render(True, view, b)
        x, y = torch.load(a), torch.load(b)
        for k in ("rgb_volume", "depth_volume", "normals_volume", "mask_volume"):
            mse = float((x[k].double() - y[k].double()).square().mean())
            rel = float((x[k].double() - y[k].double()).norm() / y[k].double().norm().clamp_min(1e-30))
            res[f"view{view}.{k}"] = dict(psnr_db=(float("inf") if mse == 0 else -10.0 * float(torch.log10(torch.tensor(mse)))), rel_l2=rel)
    print(json.dumps(res, indent=1))
