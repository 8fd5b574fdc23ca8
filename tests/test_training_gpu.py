"""A few training iterations of the shipped recipe on the synthetic sphere (GPU): random pixel batches, perturb=True, eikonal term through the
second-order LoTD path, Adam on the fp32 masters (-> the fp16 images of table / decoder / radiance net are rebuilt every step), the accel's
sample collection inside the query kernels and its EMA update every 16 iterations.  No oracle here: the checks are the ones a training run
relies on -- finite gradients on every parameter, a loss that goes down, an occupancy grid that stays alive and tracks the surface."""
import math
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_short_training_run(cuda):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from neuralsim_b200.renderer import SingleVolumeRenderer
    torch.manual_seed(0)
    model = bench.build_model(cuda, collect_samples=True).train()
    occ = model.accel.occ
    occ.n_steps_warmup = 0                                   # exercise the occupied / empty voxel sampling branch of the EMA update
    occ.update_from_net_cfg = dict(num_steps=2, num_pts=2 ** 17)
    ren = SingleVolumeRenderer(dict(near=0.01, perturb=True, depth_use_normalized_vw=False)).train()
    # Adam moves every touched parameter by ~lr per step: keep the geometry (table + decoder) on a small rate so that 40 steps do not wreck
    # the sphere, and let the radiance net move fast enough to learn the constant colour
    opt = torch.optim.Adam([dict(params=model.implicit_surface.parameters(), lr=1e-4), dict(params=model.radiance_net.parameters(), lr=5e-3),
                            dict(params=model.ctrl_var.parameters(), lr=1e-3)], eps=1e-15)
    frames = [bench.pinhole_rays(bench.H, bench.W, bench.orbit(k, 8)) for k in range(4)]
    target = torch.tensor([0.8, 0.3, 0.1], device=cuda)
    n_occ0 = int(occ.occ_grid.sum())
    losses = []
    for it in range(1, 41):
        model.training_before_per_step(it)
        o, d = frames[it % 4]
        # pixels inside the sphere's silhouette +- margin, so that most rays of the batch see the surface
        sel = torch.randperm(o.shape[0])[:4096]
        o_b, d_b = o[sel].to(cuda), d[sel].to(cuda)
        out = ren.render(model, o_b, d_b, rays_h_appear=torch.zeros(4096, 4, device=cuda))["rendered"]
        m = out["mask_volume"].detach() > 0.5
        rgb_loss = ((out["rgb_volume"] - target).abs() * m.unsqueeze(-1)).sum() / m.sum().clamp_min(1)
        eik_pts = model.sample_pts_uniform(2048)
        eik = ((eik_pts["nablas"].norm(dim=-1) - 1.0) ** 2).mean()
        loss = rgb_loss + 0.01 * eik
        opt.zero_grad(set_to_none=True)
        loss.backward()
        for name, p in model.named_parameters():
            assert p.grad is None or torch.isfinite(p.grad).all(), name
        assert model.implicit_surface.encoding.flattened_params.grad is not None
        assert model.radiance_net.blocks.layers[0].weight.grad is not None and model.ctrl_var.ln_inv_s.grad is not None
        opt.step()
        losses.append(float(rgb_loss))
    assert all(math.isfinite(x) for x in losses)
    assert sum(losses[-5:]) / 5 < 0.6 * sum(losses[:5]) / 5, losses        # the radiance net learns the constant colour
    n_occ = int(occ.occ_grid.sum())
    assert 0.3 * n_occ0 < n_occ < 3.0 * n_occ0, (n_occ0, n_occ)             # two EMA updates later the grid still hugs the sphere
    assert float(occ.occ_val_grid.max()) > 0.9
