"""N-object composition (neuralsim_b200/compose.py; reference app/renderers/buffer_compose_renderer.py:644-714): collect + per-ray sort +
integration of several packed volume buffers, and the batched ray test."""
import numpy as np
import pytest
import torch


def _random_buffer(rng, n_rays, p_hit, device, lo=1, hi=6):
    hit = np.nonzero(rng.uniform(0, 1, n_rays) < p_hit)[0]
    lens = rng.integers(lo, hi, hit.shape[0])
    first = np.cumsum(lens) - lens
    S = int(lens.sum())
    t = np.concatenate([np.sort(rng.uniform(0.5, 4.0, l)) for l in lens]) if S else np.zeros(0)
    mk = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(device)
    return dict(type="packed", rays_inds_hit=mk(hit, torch.int64), pack_infos_hit=mk(np.stack([first, lens], 1), torch.int64), t=mk(t),
                opacity_alpha=mk(rng.uniform(0.0, 0.6, S)).requires_grad_(True), rgb=mk(rng.uniform(0, 1, (S, 3))).requires_grad_(True),
                nablas_in_world=mk(rng.normal(0, 1, (S, 3))).requires_grad_(True))


def _oracle_compose(buffers, n_rays):
    """per ray: concatenate the objects' samples, stable sort by depth, alpha compositing (plain python / float64)"""
    out = dict(mask=np.zeros(n_rays), depth=np.zeros(n_rays), rgb=np.zeros((n_rays, 3)), nab=np.zeros((n_rays, 3)))
    per_ray = [[] for _ in range(n_rays)]
    for b in buffers:
        for r, (f, l) in zip(b["rays_inds_hit"].tolist(), b["pack_infos_hit"].tolist()):
            for k in range(f, f + l):
                per_ray[r].append((float(b["t"][k]), float(b["opacity_alpha"][k]), b["rgb"][k].detach().cpu().double().numpy(), b["nablas_in_world"][k].detach().cpu().double().numpy()))
    for r, samples in enumerate(per_ray):
        samples.sort(key=lambda s: s[0])
        T = 1.0
        for t, a, c, nb in samples:
            if T < 1e-4:
                break
            w = a * T
            out["mask"][r] += w; out["depth"][r] += w * t; out["rgb"][r] += w * c; out["nab"][r] += w * nb
            T *= 1.0 - a
        out["depth"][r] /= out["mask"][r] + 1e-10
    return out


@pytest.mark.gpu
def test_compose_three_buffers_against_a_per_ray_oracle(cuda):
    from neuralsim_b200.compose import compose_render
    rng = np.random.default_rng(0)
    n_rays = 500
    bufs = [_random_buffer(rng, n_rays, p, cuda) for p in (0.6, 0.3, 0.05)] + [dict(type="empty", rays_inds_hit=[])]
    rendered, total = compose_render(bufs, n_rays, training=True)
    ref = _oracle_compose(bufs[:3], n_rays)
    assert np.allclose(rendered["mask_volume"].detach().cpu().numpy(), ref["mask"], atol=1e-5)
    assert np.allclose(rendered["depth_volume"].detach().cpu().numpy(), ref["depth"], atol=1e-4)
    assert np.allclose(rendered["rgb_volume"].detach().cpu().numpy(), ref["rgb"], atol=1e-5)
    assert np.allclose(rendered["normals_volume"].detach().cpu().numpy(), ref["nab"], atol=1e-4)
    # sorted per ray, every sample exactly once
    t, pi = total["t"].cpu(), total["pack_infos_hit"].cpu()
    for f, l in pi.tolist():
        assert bool((t[f:f + l].diff() >= 0).all())
    assert sorted(total["src_index"].tolist()) == list(range(sum(b["t"].numel() for b in bufs[:3])))
    # gradients reach every object's buffers
    (rendered["rgb_volume"].sum() + rendered["mask_volume"].sum()).backward()
    for b in bufs[:3]:
        assert b["opacity_alpha"].grad is not None and float(b["opacity_alpha"].grad.abs().sum()) > 0
        assert float(b["rgb"].grad.abs().sum()) > 0


@pytest.mark.gpu
def test_two_spheres_in_one_scene(cuda):
    """two copies of the sphere model at different places / scales: rays through both see the nearer one first; a ray through one object
    renders exactly what that object renders alone"""
    from util import make_pair
    from oracle import scene as oscene
    from neuralsim_b200.compose import BufferComposeRenderer, ObjectPose
    from neuralsim_b200.renderer import SingleVolumeRenderer
    _, model = make_pair(cuda)
    model.train()
    ro, rd = oscene.pinhole_rays(30, 40, oscene.orbit_camera(0, 8, radius=6.0, elev_deg=10.0))
    ro, rd = ro.to(cuda), rd.to(cuda)
    ha = torch.zeros(ro.shape[0], 4, device=cuda)
    near_obj = ObjectPose(translation=[1.5, 0.0, 0.0], scale=1.0, device=cuda)        # between the camera (x ~ +5.9) and the origin
    far_obj = ObjectPose(translation=[-1.5, 0.0, 0.0], scale=1.5, device=cuda)
    r = BufferComposeRenderer(dict(near=0.01)).train()
    both = r.render([(model, near_obj), (model, far_obj)], ro, rd, ha, return_buffer=True)
    only_near = r.render([(model, near_obj)], ro, rd, ha)["rendered"]
    only_far = r.render([(model, far_obj)], ro, rd, ha)["rendered"]
    m_near, m_far = only_near["mask_volume"] > 0.99, only_far["mask_volume"] > 0.99
    assert int(m_near.sum()) > 20 and int(m_far.sum()) > 20 and int((m_near & m_far).sum()) > 5
    out = both["rendered"]
    sel = m_near & m_far                                    # rays through both: the nearer object occludes
    assert torch.allclose(out["depth_volume"][sel], only_near["depth_volume"][sel], atol=5e-2)     # <= 1 % residual transmittance x 3 units
    sel = m_far & ~(only_near["mask_volume"] > 0)           # rays that only meet the far one
    assert torch.allclose(out["rgb_volume"][sel], only_far["rgb_volume"][sel], atol=1e-6)
    # one object posed at the identity == the single-object renderer
    one = r.render([(model, ObjectPose(device=cuda))], ro, rd, ha)["rendered"]
    ref = SingleVolumeRenderer(dict(near=0.01)).train().render(model, ro, rd, rays_h_appear=ha)["rendered"]
    for k in ref:
        assert torch.allclose(one[k], ref[k], atol=1e-6), k
    sum(v.mean() for v in out.values()).backward()
    assert float(model.implicit_surface.encoding.flattened_params.grad.abs().sum()) > 0


def test_batched_ray_test_orders_pairs_by_ray():
    """cur_batch__ray_test (batched.py:95-145) on CPU: pairs (ray, object) consecutive in the ray index; compact_batch drops unused objects"""
    from neuralsim_b200.compose import BatchedBlockSpace
    sp = BatchedBlockSpace(2.0)
    g = torch.Generator().manual_seed(0)
    B, N = 5, 64
    o = torch.randn(B, N, 3, generator=g) * 0.3 + torch.tensor([0., 0., -4.])
    d = torch.nn.functional.normalize(torch.randn(B, N, 3, generator=g) * 0.2 + torch.tensor([0., 0., 1.]), dim=-1)
    o[3] += 100.0                                            # object 3 is far away: nothing hits it
    for compact in (False, True):
        rt = sp.cur_batch__ray_test(o, d, near=0.1, far=20.0, compact_batch=compact)
        assert rt["num_rays"] == rt["rays_inds"].numel() > 0
        assert bool((rt["rays_inds"].diff() >= 0).all())
        assert 3 not in rt["rays_full_bidx"].tolist()
        assert torch.equal(rt["full_bidx_map"][rt["rays_bidx"]], rt["rays_full_bidx"])
        if compact:
            assert rt["full_bidx_map"].tolist() == [0, 1, 2, 4]
        x_in = rt["rays_o"] + rt["rays_d"] * ((rt["near"] + rt["far"]) / 2).unsqueeze(-1)
        assert bool((x_in.abs() <= 1.0 + 1e-4).all())
