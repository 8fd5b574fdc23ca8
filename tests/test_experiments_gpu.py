"""Experiments that are in the tree but NOT on the default path (round-2 groundwork).  Collected only with NSB_TEST_EXPERIMENTS=1:
    NSB_TEST_EXPERIMENTS=1 python -m pytest tests/test_experiments_gpu.py -q"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("NSB_TEST_EXPERIMENTS") != "1", reason="experiments are opt-in (NSB_TEST_EXPERIMENTS=1)")]


def test_persistent_upsampling_equals_the_stage_kernels(cuda):
    """csrc/ray_upsample.cu: fine samples of every hit ray == cat of the stages the multi-kernel path produces (bit for bit)."""
    from util import make_pair
    from neuralsim_b200.graphics import neus_fused as NF
    from oracle import scene as oscene
    P, model = make_pair(cuda)
    ro, rd = oscene.pinhole_rays(60, 80, oscene.orbit_camera(1, 8))
    rt = model.ray_test(ro.to(cuda), rd.to(cuda), near=0.01)
    o, d = rt["rays_o"].contiguous(), rt["rays_d"].contiguous()
    ridx_hit, pinfo, t0, ridx = NF.march_lean(model.accel.occ.occ_grid, o, d, rt["near"].contiguous(), rt["far"].contiguous(), step_size=0.005, max_steps=4096)
    num_fine, factors, inv_s = [9, 9, 33], [1, 4, 16], 64.0
    surf = model.implicit_surface
    with torch.no_grad():
        sdf = surf.fused_sdf_rays(ridx, t0, o, d)
        depth, pi, stages = t0, pinfo, []
        for i, f in enumerate(factors):
            cdf = NF.upsample_cdf(sdf, depth, pi, inv_s * f, True)
            fine = NF.sample_cdf_uniform(depth, cdf, pi, num_fine[i])
            stages.append(fine)
            if i < 2:
                sdf_f = surf.fused_sdf_rays(ridx_hit, fine, o, d).contiguous()
                depth, sdf, pi = NF.merge_sorted_vals(depth, sdf, pi, fine, sdf_f)
        ref = torch.cat(stages, -1)
        got, overflow = NF.upsample_persistent(surf, ridx_hit, pinfo, t0, o, d, [inv_s * f for f in factors], num_fine, use_estimate_alpha=True)
    ok = overflow == 0
    assert bool(ok.any()) and int((~ok).sum()) <= 0.05 * ok.numel()
    assert torch.equal(got[ok], ref[ok])
