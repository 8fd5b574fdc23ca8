"""Distant-view (NeRF++) background model: the `ngp4d` auto config and the shell sampler against hand-evaluated values (CPU),
the model + its merge with the close-range buffer against the oracle restatement (GPU)."""
import math

import pytest
import torch

from oracle import distant as odist
from oracle import render as orender
from oracle import scene as oscene
from util import make_pair, rel_l2


def test_auto_ngp4d_cfg_known_answer():
    from neuralsim_b200.fields.distant import auto_ngp4d_cfg
    # the shipped BMVS Distant block (lotd_neus.bmvs.230814.yaml:205-215): 8 Mi params, min_res_xyz 8, min_res_w 4, 2^19 table, cubic space
    c = auto_ngp4d_cfg(dim=4, stretch=[2., 2., 2.], target_num_params=8 * 2 ** 20, min_res_xyz=8, min_res_w=4, n_feats=2, log2_hashmap_size=19, per_level_scale=1.382)
    assert c["lod_res"][0] == [8, 8, 8, 4] and c["lod_res"][1] == [12, 12, 12, 6]            # ceil(8 * 1.382) = 12, ceil(4 * 1.382) = 6
    assert c["lod_types"][:5] == ["Dense"] * 5 and c["lod_types"][5] == "Hash"              # 30^3 * 15 = 405 000 < 2^19 < 41^3 * 21
    n = sum((math.prod(r) if t == "Dense" else 2 ** 19) * 2 for r, t in zip(c["lod_res"], c["lod_types"]))
    assert n <= 8 * 2 ** 20 < n + 2 ** 20                                                   # levels are added while the budget lasts


def test_shell_sampler_against_the_oracle_and_by_hand():
    """CPU: product `_ray_marching` == oracle `march_shells`; one ray checked by hand"""
    from neuralsim_b200.fields.distant import ray_box_intersect, shell_radii
    r = shell_radii(1.0, 1000.0, 64)
    assert r.shape[0] == 64 and float(r[0]) == 1.0 and abs(float(r[-1]) - (1.0 - 63 * 0.999 / 64)) < 1e-6
    o = torch.tensor([[0.2, 0.0, 0.0]])
    d = torch.tensor([[1.0, 0.0, 0.0]])
    t = ray_box_intersect(o, d, torch.tensor([[1.0, 2.0, 4.0]]))
    assert torch.allclose(t, torch.tensor([[0.8, 1.8, 3.8]]))
    t = ray_box_intersect(torch.tensor([[0.0, 5.0, 0.0]]), d, torch.tensor([[1.0]]))          # passes beside the box
    assert bool(torch.isnan(t).all())


def _cfg():
    return dict(lod_res=[[6, 6, 6, 4], [9, 9, 9, 6], [13, 13, 13, 8], [20, 20, 20, 12]], lod_n_feats=[2] * 4, lod_types=["Dense", "Dense", "Hash", "Hash"],
                hashmap_size=2 ** 12)


def _make_distant(cuda, P):
    from neuralsim_b200.fields.distant import LoTDNeRFDistant
    m = LoTDNeRFDistant(encoding_cfg=dict(input_ch=4, lotd_cfg=P.lotd_cfg), radiance_decoder_cfg=dict(n_appear_embedding=P.n_appear),
                        radius_scale_min=1.0, radius_scale_max=1000.0, include_inf_distance=True, device=cuda,
                        ray_query_cfg=dict(query_mode="march", query_param=dict(march_cfg=dict(sample_mode="box", max_steps=32))))
    with torch.no_grad():
        m.encoding.flattened_params.copy_(P.grid.to(cuda))
        d, r = m.density_decoder.layers, m.rgb_decoder.blocks.layers
        d[0].weight.copy_(P.den_W1.to(cuda)); d[0].bias.copy_(P.den_b1.to(cuda)); d[1].weight.copy_(P.den_W2.to(cuda)); d[1].bias.copy_(P.den_b2.to(cuda))
        r[0].weight.copy_(P.rad_W1.to(cuda)); r[0].bias.copy_(P.rad_b1.to(cuda)); r[1].weight.copy_(P.rad_W2.to(cuda)); r[1].bias.copy_(P.rad_b2.to(cuda))
        r[2].weight.copy_(P.rad_W3.to(cuda)); r[2].bias.copy_(P.rad_b3.to(cuda))
    return m


@pytest.mark.gpu
def test_distant_model_against_the_oracle(cuda):
    """4-D LoTD + density / radiance MLPs + shells + compression: buffer structure equal, values within the fp16 rounding of the MLPs; gradients"""
    P = odist.DistantParams(_cfg())
    m = _make_distant(cuda, P)
    ro, rd = oscene.pinhole_rays(12, 16, oscene.orbit_camera(3, 8, radius=3.0))
    near = torch.full([ro.shape[0]], 0.01)
    ha = torch.zeros(ro.shape[0], P.n_appear)
    P.requires_grad_(True)
    ref = odist.ray_query(P, ro, rd, near, ha, radius_scale_min=1.0, radius_scale_max=1000.0, max_steps=32)
    rt = dict(rays_o=ro.to(cuda), rays_d=rd.to(cuda), near=near.to(cuda), far=None, num_rays=ro.shape[0], rays_inds=torch.arange(ro.shape[0], device=cuda),
              rays_h_appear=ha.to(cuda))
    got = m.ray_query(ray_tested=rt, config=dict(with_rgb=True))["volume_buffer"]
    assert got["type"] == "packed" and ref["type"] == "packed"
    assert torch.equal(got["rays_inds_hit"].cpu(), ref["rays_inds_hit"]) and torch.equal(got["pack_infos_hit"].cpu(), ref["pack_infos_hit"])
    assert torch.allclose(got["t"].cpu(), ref["t"], rtol=1e-5, atol=1e-5)
    assert rel_l2(got["opacity_alpha"], ref["opacity_alpha"]) <= 2e-3 and rel_l2(got["rgb"], ref["rgb"]) <= 2e-3
    (ref["opacity_alpha"].sum() + ref["rgb"].sum()).backward()
    (got["opacity_alpha"].sum() + got["rgb"].sum()).backward()
    assert rel_l2(m.encoding.flattened_params.grad, P.grid.grad) <= 3e-2
    assert rel_l2(m.density_decoder.layers[0].weight.grad, P.den_W1.grad) <= 3e-2
    assert rel_l2(m.rgb_decoder.blocks.layers[2].weight.grad, P.rad_W3.grad) <= 3e-2


@pytest.mark.gpu
def test_close_range_plus_distant_render_against_the_oracle(cuda):
    """SingleVolumeRenderer with the distant model: dv sampled behind the close-range box, buffers merged per ray, one integration"""
    from neuralsim_b200.renderer import SingleVolumeRenderer
    Pn, model = make_pair(cuda)
    Pd = odist.DistantParams(_cfg())
    dv = _make_distant(cuda, Pd)
    ro, rd = oscene.pinhole_rays(18, 24, oscene.orbit_camera(1, 8, radius=3.0, elev_deg=25.0))
    n = ro.shape[0]
    ha = torch.zeros(n, 4)
    with torch.no_grad():
        rt = orender.ray_test(ro, rd, near=0.01)
        vb_cr, _ = orender.neus_ray_query(Pn, oscene.make_occ_grid(), rt, rays_h_appear=torch.zeros(rt["num_rays"], Pn.n_appear))
        near_dv = torch.full([n], 0.01)
        near_dv[rt["rays_inds"]] = rt["far"]
        vb_dv = odist.ray_query(Pd, ro, rd, near_dv, ha, radius_scale_min=1.0, radius_scale_max=1000.0, max_steps=32)
        ref = orender.volume_integration(odist.merge_buffers(vb_cr, vb_dv, n), n, training=True)
        out = SingleVolumeRenderer(dict(near=0.01)).train().render(model.train(), ro.to(cuda), rd.to(cuda), rays_h_appear=ha.to(cuda), distant_model=dv,
                                                                 return_buffer=True)
    got = out["rendered"]
    assert float(got["mask_volume"].mean()) > 0.9          # the shell at 1e10 closes every ray
    for k in ("rgb_volume", "depth_volume", "mask_volume", "normals_volume"):
        assert rel_l2(got[k], ref[k]) <= 2e-3, (k, rel_l2(got[k], ref[k]))
    # rays that miss the close-range box are rendered by the distant model alone
    miss = torch.ones(n, dtype=torch.bool)
    miss[rt["rays_inds"]] = False
    only_dv = orender.volume_integration(vb_dv, n)
    assert rel_l2(got["rgb_volume"].cpu()[miss], only_dv["rgb_volume"][miss]) <= 2e-3
