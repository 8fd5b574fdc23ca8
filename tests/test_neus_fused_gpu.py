"""Fused per-ray NeuS stage kernels (csrc/neus_fused.cu) against (1) the CPU oracle and (2) the unfused chain of
pack_ops / elementwise calls they replace (the reference's own formulation, graphics/neus.py with FUSED_STAGES off)."""
import numpy as np
import pytest
import torch

from util import random_packs, rel_l2

pytestmark = pytest.mark.gpu


def _packs(seed, n_packs=700, lo=1, hi=150):
    rng = np.random.default_rng(seed)
    pi = random_packs(rng, n_packs, lo, hi, "cuda")
    S = int(pi[-1].sum())
    # a ray crossing a surface: sdf decreasing through zero with noise, depth increasing
    g = torch.Generator().manual_seed(seed)
    depth, sdf = torch.empty(S), torch.empty(S)
    for b, n in pi.cpu().tolist():
        t = torch.rand(n, generator=g).sort().values * 2 + 0.5
        depth[b:b + n] = t
        sdf[b:b + n] = (1.4 - t) * (0.5 + torch.rand(1, generator=g)) + 0.02 * torch.randn(n, generator=g)
    return pi, sdf.cuda(), depth.cuda()


@pytest.mark.parametrize("estimate", [False, True])
@pytest.mark.parametrize("inv_s", [64.0, 1024.0])
def test_upsample_cdf_and_sampling(estimate, inv_s):
    from neuralsim_b200.graphics import neus as G, neus_fused as NF
    from neuralsim_b200.graphics.pack_ops import packed_cumsum, packed_div
    from neuralsim_b200.graphics.nerf import packed_alpha_to_vw
    from neuralsim_b200.graphics.raysample import packed_sample_cdf
    from oracle import render as orender
    pi, sdf, depth = _packs(3)
    alpha = (G.neus_packed_sdf_to_upsample_alpha(sdf, depth, inv_s, pi) if estimate else G.neus_packed_sdf_to_alpha(sdf, inv_s, pi))
    vw = packed_alpha_to_vw(alpha, pi)
    cdf = packed_cumsum(vw, pi, exclusive=True)
    cdf = packed_div(cdf, cdf[pi[:, 0] + pi[:, 1] - 1].clamp_min(1e-5), pi)
    got = NF.upsample_cdf(sdf, depth, pi, inv_s, estimate)
    assert torch.allclose(got, cdf, atol=2e-6, rtol=1e-5), float((got - cdf).abs().max())
    # oracle (CPU) alpha -> same cdf
    o_alpha = (orender.neus_packed_sdf_to_upsample_alpha(sdf.cpu(), depth.cpu(), inv_s, pi.cpu()) if estimate
               else orender.neus_packed_sdf_to_alpha(sdf.cpu(), inv_s, pi.cpu()))
    o_vw = orender.packed_alpha_to_vw(o_alpha, pi.cpu())
    o_cdf = orender.packed_cumsum_exclusive(o_vw, pi.cpu())
    o_cdf = orender.packed_div(o_cdf, o_cdf[pi.cpu()[:, 0] + pi.cpu()[:, 1] - 1].clamp_min(1e-5), pi.cpu())
    assert torch.allclose(got.cpu(), o_cdf, atol=5e-6, rtol=1e-5), float((got.cpu() - o_cdf).abs().max())
    for n in (9, 33):
        ref = packed_sample_cdf(depth, cdf, pi, n)[0]
        fine = NF.sample_cdf_uniform(depth, cdf, pi, n)
        assert torch.equal(fine, ref)            # same cdf in -> bit-identical samples


def test_neus_alpha_compress_forward_backward():
    from neuralsim_b200.graphics import neus as G, neus_fused as NF
    from neuralsim_b200.graphics.nerf import packed_volume_render_compression
    from oracle import render as orender
    pi, sdf, depth = _packs(5)
    for ln in (3.0, 5.5):
        ln_inv_s = torch.tensor(ln, device="cuda", requires_grad=True)
        s1 = sdf.clone().requires_grad_(True)
        a_ref = G.neus_packed_sdf_to_alpha(s1, ln_inv_s.exp(), pi)
        nidx_r, pinf_r, pidx_r = packed_volume_render_compression(a_ref.detach(), pi)
        s2 = sdf.clone().requires_grad_(True)
        ln2 = ln_inv_s.detach().clone().requires_grad_(True)
        a, nidx, pinf, pidx = NF.neus_alpha_compress(s2, ln2.exp(), pi)
        assert torch.allclose(a, a_ref, atol=1e-7, rtol=1e-6)
        o_alpha = orender.neus_packed_sdf_to_alpha(sdf.cpu(), float(np.exp(np.float32(ln))), pi.cpu())
        assert torch.allclose(a.detach().cpu(), o_alpha, atol=1e-6, rtol=1e-5)
        assert torch.equal(nidx, nidx_r) and torch.equal(pinf, pinf_r) and torch.equal(pidx, pidx_r)
        w = torch.randn(pidx.numel(), device="cuda", generator=torch.Generator("cuda").manual_seed(1))
        (a_ref[pidx_r] * w).sum().backward()
        (a[pidx] * w).sum().backward()
        assert rel_l2(s2.grad, s1.grad) < 1e-4
        assert abs(float(ln2.grad) - float(ln_inv_s.grad)) <= 1e-4 * abs(float(ln_inv_s.grad)) + 1e-6


@pytest.mark.parametrize("normalize_depth", [True, False])
def test_composite_forward_backward(normalize_depth):
    from neuralsim_b200.graphics import neus_fused as NF
    from neuralsim_b200.graphics.nerf import packed_alpha_to_vw
    from neuralsim_b200.graphics.pack_ops import packed_div, packed_sum
    pi, sdf, depth = _packs(7, n_packs=500, hi=90)
    K = sdf.numel()
    g = torch.Generator("cuda").manual_seed(2)
    alpha0 = (torch.rand(K, device="cuda", generator=g) ** 3).clamp(0, 0.95)
    alpha0[torch.rand(K, device="cuda", generator=g) < 0.2] = 0
    rgb0, nab0 = torch.rand(K, 3, device="cuda", generator=g), torch.randn(K, 3, device="cuda", generator=g)
    cot = [torch.randn(s, device="cuda", generator=g) for s in ((pi.shape[0],), (pi.shape[0],), (pi.shape[0], 3), (pi.shape[0], 3), (K,))]

    def run(fused):
        a, r, nb = (t.clone().requires_grad_(True) for t in (alpha0, rgb0, nab0))
        if fused:
            vw, m, d, c, n_ = NF.composite(a, depth, pi, rgb=r, nablas=nb, normalize_depth=normalize_depth)
        else:
            vw = packed_alpha_to_vw(a, pi)
            m = packed_sum(vw, pi)
            dw = packed_div(vw, m + 1e-10, pi) if normalize_depth else vw
            d = packed_sum(dw * depth, pi)
            c, n_ = packed_sum(vw.view(-1, 1) * r, pi), packed_sum(vw.view(-1, 1) * nb, pi)
        loss = (m * cot[0]).sum() + (d * cot[1]).sum() + (c * cot[2]).sum() + (n_ * cot[3]).sum() + (vw * cot[4]).sum()
        loss.backward()
        return (vw, m, d, c, n_), (a.grad, r.grad, nb.grad)

    (vw, m, d, c, n_), grads = run(True)
    (vw_r, m_r, d_r, c_r, n_r), grads_r = run(False)
    assert torch.equal(vw, vw_r)                              # serial recurrence: bit-exact
    for x, y in ((m, m_r), (d, d_r), (c, c_r), (n_, n_r)):
        assert torch.allclose(x, y, atol=2e-6, rtol=1e-5)
    for x, y in zip(grads, grads_r):
        assert rel_l2(x, y) < 1e-5


@pytest.mark.parametrize("perturb,stages", [(False, 3), (True, 3), (False, 1), (False, 2)])
def test_render_fused_equals_unfused_chain(perturb, stages):
    """The whole query + integration with the fused stages on and off: same samples, same image, same gradients
    (with `perturb`, from the same seed: the fused path consumes the random stream exactly as the op-by-op chain does)."""
    from neuralsim_b200.graphics import neus as G
    from neuralsim_b200.renderer import SingleVolumeRenderer
    from oracle import scene as oscene
    from util import make_pair, product_grads
    P, model = make_pair("cuda")
    if stages != 3:                                        # fewer up-sampling stages: no merge at all (1) / one merge (2)
        qp = dict(model.ray_query_cfg["query_param"])
        qp.update(num_fine=[8, 8, 32][:stages] if stages == 2 else 16, upsample_inv_s_factors=[1, 4, 16][:stages])
        model.ray_query_cfg["query_param"] = qp
    rays_o, rays_d = oscene.pinhole_rays(36, 48, oscene.orbit_camera(1, 8, radius=3.0, elev_deg=25.0))
    rays_o, rays_d = rays_o.cuda(), rays_d.cuda()
    ren = SingleVolumeRenderer(dict(near=0.01, far=None, perturb=perturb))
    h_appear = torch.linspace(-0.5, 0.5, rays_o.shape[0] * 4, device="cuda").view(-1, 4).contiguous()
    outs = {}
    for fused in (True, False):
        G.FUSED_STAGES = fused
        try:
            model.zero_grad(set_to_none=True)
            torch.manual_seed(123)
            ret = ren.render(model, rays_o, rays_d, rays_h_appear=h_appear, return_buffer=True)
            r = ret["rendered"]
            (r["rgb_volume"].square().sum() + r["depth_volume"].sum() * 0.1 + r["mask_volume"].sum() * 0.3
             + r["normals_volume"].square().sum() * 0.05).backward()
            outs[fused] = ({k: v.detach().clone() for k, v in r.items()}, product_grads(model), ret["volume_buffer"]["pack_infos_hit"].clone())
        finally:
            G.FUSED_STAGES = True
    assert torch.equal(outs[True][2], outs[False][2])
    for k in outs[True][0]:
        assert rel_l2(outs[True][0][k], outs[False][0][k]) < 1e-5, k
    for k, g in outs[True][1].items():
        if g is not None:
            assert rel_l2(g, outs[False][1][k]) < 2e-3, k


def test_conditioned_query_kwargs_reach_every_network_call(cuda):
    """rays_ts / rays_fidx of `ray_tested` (neus_ray_query.py:776-790, 846-866): a model that declares `use_ts` / `use_fidx` receives, with EVERY
    sdf / colour query, the value of the ray each sample belongs to; the result equals the unconditioned query (the stand-in ignores the values)"""
    from oracle import scene as oscene
    from util import make_pair
    import neuralsim_b200.graphics.neus as N
    _, model = make_pair(cuda)
    ro, rd = oscene.pinhole_rays(20, 24, oscene.orbit_camera(1, 8))
    ro, rd = ro.to(cuda), rd.to(cuda)
    n = ro.shape[0]
    ts = torch.arange(n, device=cuda, dtype=torch.float32) * 0.5
    fidx = torch.arange(n, device=cuda) % 7
    rt0 = model.ray_test(ro, rd, near=0.01, rays_h_appear=torch.zeros(n, 4, device=cuda))
    rt = model.ray_test(ro, rd, near=0.01, rays_h_appear=torch.zeros(n, 4, device=cuda), rays_ts=ts, rays_fidx=fidx)
    assert torch.equal(rt["rays_ts"], ts[rt["rays_inds"]])
    qp = dict(model.ray_query_cfg["query_param"])
    model.eval()
    with torch.no_grad():
        want, _ = N.neus_ray_query_march_occ_multi_upsample_compressed(model, rt0, **qp)
    seen = []

    class Conditioned:
        use_ts, use_fidx = True, True

        def __getattr__(self, k):
            return getattr(model, k)

        def forward_sdf(self, x, ts=None, fidx=None, **kw):
            assert ts is not None and fidx is not None and ts.shape[0] == x.reshape(-1, 3).shape[0] == fidx.shape[0]
            assert torch.equal((ts * 2).round().long() % 7, fidx)          # both belong to the same ray
            seen.append(ts.shape[0])
            return model.forward_sdf(x, **kw)

        def forward(self, x, ts=None, fidx=None, **kw):
            assert ts is not None and ts.shape[0] == x.shape[0] and torch.equal((ts * 2).round().long() % 7, fidx)
            seen.append(-x.shape[0])
            return model.forward(x, **kw)

    with torch.no_grad():
        got, _ = N.neus_ray_query_march_occ_multi_upsample_compressed(Conditioned(), rt, **qp)
    assert len([s for s in seen if s > 0]) >= 4 and len([s for s in seen if s < 0]) == 1       # marched + 2 fine stages + boundary; one colour query
    assert torch.equal(got["rays_inds_hit"], want["rays_inds_hit"]) and torch.equal(got["pack_infos_hit"], want["pack_infos_hit"])
    assert torch.allclose(got["t"], want["t"]) and torch.allclose(got["rgb"], want["rgb"], atol=2e-3) and torch.allclose(got["opacity_alpha"], want["opacity_alpha"], atol=2e-3)
