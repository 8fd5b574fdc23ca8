"""How fast can the table-gradient scatter go?  The real colour samples of one bench frame (2.8 M points near the surface):
  (a) the stand-alone scatter kernel k_lotd_bwd_grid (thread = (point, level), 64 warps / SM, nothing else in flight)
  (b) the fused backward k_sdf_bwd_tc (gather recompute + 3 MMAs + scatter, 16 warps / SM)
python profiles/ab_scatter.py > gpurun_out/ab_scatter.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neuralsim_b200 import _lib as L  # noqa: E402
from neuralsim_b200.bindings import _lotd  # noqa: E402
from neuralsim_b200.renderer import SingleVolumeRenderer  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_model(dev).train()
o, d = bench.pinhole_rays(bench.H, bench.W, bench.orbit(0, 8))
ren = SingleVolumeRenderer(dict(near=0.01)).train()
with torch.no_grad():
    vb = ren.ray_query(model, o.to(dev), d.to(dev), rays_h_appear=torch.zeros(o.shape[0], 4, device=dev))["volume_buffer"]
x = vb["net_x"].detach().contiguous()
K = x.shape[0]
print("colour samples", K, flush=True)
enc = model.implicit_surface.encoding
meta, params = enc.meta, enc.flattened_params.detach()
xt = (x * 0.5 + 0.5).clamp(1e-6, 1 - 1e-6).contiguous()
g = torch.randn(K, 32, device=dev)
fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, tag, reps=4):
    ts = []
    for i in range(reps + 1):
        fl.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i:
            ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    print(f"{tag:70s} {ms:8.3f} ms  {K / ms / 1e6:6.2f} Gpts/s  {K * 128 / ms / 1e6:7.1f} G corner-updates/s", flush=True)


acc = torch.zeros(params.shape[0], device=dev)


def scatter_only():
    L.check(L.lib().nsb_lotd_bwd_grid(meta.c_ref, L.ptr(g), 0, L.ptr(xt, "f32"), L.c_i64(K), L.c_i32(meta.n_levels), L.c_f32(1.0), L.ptr(acc),
                                      L.stream_ptr()), "bwd_grid")


timeit(scatter_only, "(a) k_lotd_bwd_grid: scatter only, full occupancy")
perm = torch.randperm(K, device=dev)
xt_r, g_r = xt[perm].contiguous(), g[perm].contiguous()
xt, g = xt_r, g_r
timeit(scatter_only, "(a') the same points in random order")
surf = model.implicit_surface
w = torch.randn(K, device=dev)


def fused_fwd_bwd():
    model.zero_grad(set_to_none=True)
    s = surf.fused_sdf_autograd(x)
    (s * w).sum().backward()


L.KERNEL_TIMER.enable()
for _ in range(5):
    fused_fwd_bwd()
kt = L.KERNEL_TIMER.summary()
L.KERNEL_TIMER.disable()
for k, v in kt.items():
    print(f"(b) {k}: {v['ms'] / v['launches']:.3f} ms per launch, {v['units'] / v['launches'] / 1e6:.2f} M points", flush=True)
