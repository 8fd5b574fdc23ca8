"""GPU: neuralsim_b200's host layer + kernels against the vectors produced by the reference's own Python."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python.npz"))


def test_product_matches_reference_python_vectors(cuda):
    from neuralsim_b200.graphics import neus as N, nerf as NF, pack_ops as P, raysample as RS
    T = lambda k: torch.from_numpy(G[k]).to(cuda)
    pi, sdf, depth = T("pi"), T("neus.sdf"), T("neus.depth")
    assert torch.equal(P.get_pack_infos_from_n(pi[:, 1].contiguous()), pi) and torch.equal(P.get_pack_infos_from_batch(5, 7, device=cuda), T("pi_from_batch"))
    ids = torch.repeat_interleave(torch.arange(pi.shape[0], device=cuda), pi[:, 1])
    assert torch.equal(P.get_pack_infos_from_boundary(P.mark_pack_boundaries(ids)), pi)
    for inv_s in (20, 64, 256, 2000):
        assert torch.allclose(N.neus_packed_sdf_to_alpha(sdf, float(inv_s), pi), T(f"neus.packed_alpha.{inv_s}"), rtol=1e-5, atol=1e-6)
        assert torch.allclose(N.neus_packed_sdf_to_upsample_alpha(sdf, depth, float(inv_s), pi), T(f"neus.upsample_alpha.{inv_s}"), rtol=1e-5, atol=1e-6)
    assert torch.allclose(N.neus_ray_sdf_to_alpha(T("neus.sdf_b"), 64.0), T("neus.ray_alpha"), rtol=1e-5, atol=1e-6)
    assert torch.allclose(N.neus_ray_sdf_to_alpha(T("neus.sdf_b"), 64.0, True), T("neus.ray_alpha_app1"), rtol=1e-5, atol=1e-6)
    assert torch.allclose(N.neus_ray_sdf_to_upsample_alpha(T("neus.sdf_b"), T("neus.t_b"), 64.0), T("neus.ray_upsample_alpha"), rtol=1e-5, atol=1e-6)
    assert torch.allclose(N.neus_ray_sdf_to_vw(T("neus.sdf_b"), 64.0), T("neus.ray_vw"), rtol=1e-5, atol=1e-6)
    a = T("vw.alpha_b")
    pib = P.get_pack_infos_from_batch(*a.shape, device=cuda)
    assert torch.allclose(NF.ray_alpha_to_vw(a), T("vw.ray_alpha_to_vw"), rtol=1e-5, atol=1e-7)
    assert torch.allclose(P.packed_alpha_to_vw(a.flatten(), pib), T("vw.packed_default"), rtol=1e-6, atol=1e-8)
    nidx, cpi, pidx = P.packed_volume_render_compression(a.flatten(), pib)
    assert torch.equal(nidx, T("vw.compress.nidx")) and torch.equal(cpi, T("vw.compress.pack_infos")) and torch.equal(pidx, T("vw.compress.pidx"))
    t, dt = RS.batch_sample_step_linear(T("rs.near"), T("rs.far"), 65, return_dt=True)
    assert torch.allclose(t, T("rs.linear_t"), rtol=1e-6, atol=1e-6) and torch.allclose(dt, T("rs.linear_dt"))
    s, i = P.packed_invert_cdf(T("rs.kat_bins"), T("rs.kat_cdfs"), torch.linspace(0., 1., 42, device=cuda)[1:-1].expand(3, 40).contiguous(), T("rs.kat_pi"))
    assert torch.equal(i, T("rs.kat_idx")) and torch.allclose(s, T("rs.kat_samples"), rtol=1e-6, atol=1e-7)
    assert torch.allclose(RS.packed_sample_cdf(T("rs.kat_bins"), T("rs.kat_cdfs"), T("rs.kat_pi"), 9)[0], T("rs.kat_sample_cdf9"), rtol=1e-6, atol=1e-7)
    assert torch.allclose(RS.batch_sample_pdf(T("rs.pdf_bins"), T("rs.pdf_w"), 12), T("rs.batch_sample_pdf"), rtol=1e-5, atol=1e-6)
    pa, pb, pinf = P.merge_two_packs_sorted(T("merge_sorted.va"), T("merge_sorted.pia"), T("merge_sorted.na"), T("merge_sorted.vb"),
                                            T("merge_sorted.pib"), T("merge_sorted.nb"))
    assert torch.equal(pa, T("merge_sorted.pa")) and torch.equal(pb, T("merge_sorted.pb")) and torch.equal(pinf, T("merge_sorted.pinf"))
    pa, pb, pinf = P.merge_two_batch_a_includes_b(T("merge_batch.A"), T("merge_batch.nA"), T("merge_batch.B"), T("merge_batch.nB"))
    assert torch.equal(pa, T("merge_batch.pa")) and torch.equal(pb, T("merge_batch.pb")) and torch.equal(pinf, T("merge_batch.pinf"))
    # a_includes_b known answer (unit_test.py:998-1011)
    va = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 11.1, 11.2, 0.2, 0.8], device=cuda); pia = P.get_pack_infos_from_n(torch.tensor([5, 2, 2], device=cuda))
    vb = torch.tensor([0.0, 0.25, 0.26, 0.6, 0.1, 0.2, 0.3, 0.4], device=cuda); pib2 = P.get_pack_infos_from_n(torch.tensor([4, 4], device=cuda))
    pa, pb, pinf = P.merge_two_packs_sorted_a_includes_b(va, pia, torch.tensor([11, 12, 13], device=cuda), vb, pib2, torch.tensor([11, 13], device=cuda))
    assert pa.tolist() == [1, 2, 5, 6, 7, 9, 10, 13, 16] and pb.tolist() == [0, 3, 4, 8, 11, 12, 14, 15] and pinf.tolist() == [[0, 9], [9, 2], [11, 6]]


def test_product_autograd_matches_reference_rules(cuda):
    from neuralsim_b200.graphics import pack_ops as P
    T = lambda k: torch.from_numpy(G[k]).to(cuda)
    pi, f, w = T("pi"), T("grad.f"), T("grad.w")
    x = f.clone().requires_grad_(True)
    for name, fn in (("sum", lambda v: P.packed_sum(v, pi)), ("cumsum_excl", lambda v: P.packed_cumsum(v, pi, exclusive=True))):
        y = fn(x)
        g, = torch.autograd.grad((y * w[:y.shape[0]]).sum(), x)
        assert torch.allclose(y, T(f"grad.{name}.y"), rtol=1e-5, atol=1e-5) and torch.allclose(g, T(f"grad.{name}.g"), rtol=1e-5, atol=1e-5), name
    x1 = f[:, 0].contiguous().clone().requires_grad_(True)
    for name, fn in (("diff", lambda v: P.packed_diff(v, pi)), ("bdiff", lambda v: P.packed_backward_diff(v, pi))):
        y = fn(x1)
        g, = torch.autograd.grad((y * w[:, 0]).sum(), x1)
        assert torch.allclose(y, T(f"grad.{name}.y")) and torch.allclose(g, T(f"grad.{name}.g"), atol=1e-6), name
    o = T("grad.div.other").clone().requires_grad_(True)
    x2 = f[:, 0].contiguous().clone().requires_grad_(True)
    y = P.packed_div(x2, o, pi)
    gi, go = torch.autograd.grad((y * w[:, 0]).sum(), [x2, o])
    assert torch.allclose(y, T("grad.div.y"), rtol=1e-6) and torch.allclose(gi, T("grad.div.gi"), rtol=1e-6) and torch.allclose(go, T("grad.div.go"), rtol=1e-4, atol=1e-5)
    al = T("grad.a2vw.alpha").clone().requires_grad_(True)
    vw = P.packed_alpha_to_vw(al, pi)
    g, = torch.autograd.grad((vw * w[:, 0]).sum(), al)
    assert torch.allclose(vw, T("grad.a2vw.vw")) and torch.allclose(g, T("grad.a2vw.g"), rtol=1e-5, atol=1e-6)
