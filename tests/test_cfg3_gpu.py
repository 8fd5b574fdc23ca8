"""Workload cfg3 (BASELINE.json configs[2], bench_cfg3.py): the StreetSurf close-range model -- cuboid aabb, cuboid LoTD from the `ngp` auto
config, per-axis occupancy grid, step 0.2 / 128 coarse samples -- and LiDAR-style rays (with_rgb=False, with_normal=True)."""
import pytest
import torch

import bench_cfg3 as C
from util import rel_l2

pytestmark = pytest.mark.gpu


def _small(cuda, levels):
    # the same construction with a 2^16 table and ~2 Mi parameters: seconds instead of a 32 Mi-parameter model
    return C.build_model(cuda, max_num_levels=levels, log2_hashmap_size=16, target_num_params=(levels + 2) * 2 ** 17).train()


def test_auto_ngp_cfg_matches_the_reference_formula():
    from neuralsim_b200.fields.encoding import auto_ngp_cfg
    c = auto_ngp_cfg([40., 150., 15.], 32 * 2 ** 20, dim=3, n_feats=2, log2_hashmap_size=20, min_res=16, max_num_levels=None)
    assert len(c["lod_res"]) == 17 and c["lod_types"].count("Dense") == 2            # lotd_cfg.py:59-133 evaluated by hand: 2 dense + 15 hashed levels
    assert c["lod_res"][1] == [66, 250, 25] and c["hashmap_size"] == 2 ** 20
    c16 = auto_ngp_cfg([40., 150., 15.], 32 * 2 ** 20, dim=3, n_feats=2, log2_hashmap_size=20, min_res=16, max_num_levels=16)
    assert c16["lod_res"] == c["lod_res"][:16]


def test_plane_scene_renders_the_road(cuda):
    from neuralsim_b200.renderer import SingleVolumeRenderer
    model = _small(cuda, 16)
    assert model.implicit_surface._fusable() and list(model.accel.occ.occ_grid.shape) == [40, 150, 15]
    (co, cd), (lo, ld) = C.make_views(2)
    with torch.no_grad():
        out = SingleVolumeRenderer(dict(near=C.NEAR, far=C.FAR)).eval().render(model.eval(), co.to(cuda), cd.to(cuda), rays_h_appear=torch.zeros(co.shape[0], 4, device=cuda))["rendered"]
    down = cd[:, 2] < -0.05                                            # rays that look at the road within the box
    t_exp = (C.ROAD_Z - co[:, 2]) / cd[:, 2]
    hitp = co + cd * t_exp.unsqueeze(-1)
    inside = down & (t_exp < 60) & (hitp[:, 0].abs() < 19.0) & (hitp[:, 1].abs() < 74.0)
    hit = out["mask_volume"].cpu() > 0.9
    assert float((hit & inside).sum()) > 0.9 * float(inside.sum())
    err = (out["depth_volume"].cpu()[hit & inside] - t_exp[hit & inside]).abs()
    assert float(err.median()) < 0.1                                   # metres: the plane is where it was put
    n = out["normals_volume"].cpu()[hit & inside]
    assert float(n[:, 2].mean()) > 0.9                                 # normals point up


@pytest.mark.parametrize("levels", [16, 17])
def test_lidar_rays_fused_equals_op_by_op(cuda, levels):
    """with_rgb=False, with_normal=True: the fused colour op (16 levels) / the generic path (17 levels, as the shipped config) against the chain with
    every fused path off; depth, normals, mask and the gradients of the table / decoder"""
    import neuralsim_b200.graphics.neus as GN
    import neuralsim_b200.fields.space as SP
    from neuralsim_b200.fields.networks import LoTDSDF
    from neuralsim_b200.renderer import SingleVolumeRenderer
    model = _small(cuda, levels)
    assert model.implicit_surface._fusable() == (levels == 16)
    _, (lo, ld) = C.make_views(1)
    lo, ld = lo[:2048].to(cuda), ld[:2048].to(cuda)
    r = SingleVolumeRenderer(dict(near=C.NEAR, far=C.FAR, with_rgb=False, with_normal=True)).train()
    outs = []
    for fused in (True, False):
        model.zero_grad(set_to_none=True)
        saved = (GN.FUSED_STAGES, SP.FUSED_RAY_TEST, LoTDSDF._fusable)
        if not fused:
            GN.FUSED_STAGES, SP.FUSED_RAY_TEST, LoTDSDF._fusable = False, False, (lambda self: False)
        try:
            out = r.render(model, lo, ld)["rendered"]
            assert "rgb_volume" not in out
            C.loss_lidar(out).backward()
        finally:
            GN.FUSED_STAGES, SP.FUSED_RAY_TEST, LoTDSDF._fusable = saved
        g = model.implicit_surface.encoding.flattened_params.grad
        outs.append(({k: v.detach().clone() for k, v in out.items()}, g.clone(), model.implicit_surface.decoder.layers[0].weight.grad.clone()))
        assert all(p.grad is None or float(p.grad.abs().sum()) == 0.0 for p in model.radiance_net.parameters())      # nothing reaches the radiance net
    (a, ga, wa), (b, gb, wb) = outs
    assert float(a["mask_volume"].sum()) > 100
    for k in ("depth_volume", "normals_volume", "mask_volume"):
        assert rel_l2(a[k], b[k]) <= 1e-4, (k, rel_l2(a[k], b[k]))
    assert rel_l2(ga, gb) <= 2e-2 and rel_l2(wa, wb) <= 2e-2, (rel_l2(ga, gb), rel_l2(wa, wb))


def test_static_step_on_cfg3(cuda):
    """camera + LiDAR rays through the static (graph) step == the host-sized path, bit for bit"""
    from neuralsim_b200.graphics.neus_static import StaticFrame
    from neuralsim_b200.renderer import SingleVolumeRenderer
    model = _small(cuda, 16)
    (co, cd), (lo, ld) = C.make_views(3)
    co, cd, lo, ld = (t[:4096].to(cuda) for t in (co, cd, lo, ld))
    ha = torch.zeros(4096, 4, device=cuda)
    with torch.no_grad():
        ref_c = SingleVolumeRenderer(dict(near=C.NEAR, far=C.FAR)).train().render(model, co, cd, rays_h_appear=ha)["rendered"]
        ref_l = SingleVolumeRenderer(dict(near=C.NEAR, far=C.FAR, with_rgb=False)).train().render(model, lo, ld)["rendered"]
        fc = StaticFrame(model, 4096, near=C.NEAR, far=C.FAR, slack=2.0)
        fl = StaticFrame(model, 4096, near=C.NEAR, far=C.FAR, with_rgb=False, slack=2.0)
        fc.step(co, cd, ha)
        fl.step(lo, ld, None)
    assert fc.counts()["overflow"] == 0 and fl.counts()["overflow"] == 0
    for k, v in ref_c.items():
        assert torch.equal(fc.rendered[k], v), k
    for k, v in ref_l.items():
        assert torch.equal(fl.rendered[k], v), k
    assert "rgb_volume" not in fl.rendered
