"""GPU: our kernels AND the CPU oracle against the REFERENCE'S OWN CUDA kernels, compiled from /root/reference into
oracle/_ref by oracle/build_ref.py (the .so files travel with the snapshot; skipped when they are absent).
This is what pins the oracle for the parts the reference ships no fixtures for: LoTD, marching, alpha compositing."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))


def _ref(name):
    import build_ref
    try:
        mod = build_ref.load(name)
    except Exception as ex:  # pragma: no cover
        pytest.skip(f"oracle/_ref/{name} not loadable: {ex}")
    if mod is None:
        pytest.skip(f"oracle/_ref/{name}.so absent (built only where /root/reference exists)")
    return mod


def test_lotd_against_reference_kernels(cuda):
    from oracle import lotd as olotd
    from neuralsim_b200.bindings import _lotd as ours
    ref = _ref("_lotd")
    cfg = olotd.gen_ngp_cfg()
    rm = ref.LoDMeta(3, cfg["lod_res"], cfg["lod_n_feats"], cfg["lod_types"], cfg["hashmap_size"], False)
    om = ours.LoDMeta(3, cfg["lod_res"], cfg["lod_n_feats"], cfg["lod_types"], cfg["hashmap_size"])
    cm = olotd.LoDMeta(3, **cfg)
    assert rm.n_params == om.n_params and list(rm.level_offsets) == om.level_offsets and list(rm.level_sizes) == om.level_sizes
    rng = np.random.default_rng(0)
    p = torch.from_numpy(rng.uniform(-0.1, 0.1, rm.n_params).astype(np.float16)).to(cuda)
    x = torch.from_numpy(rng.uniform(1e-6, 1 - 1e-6, (60000, 3)).astype(np.float32)).to(cuda)
    y_r, d_r = ref.lod_fwd(rm, x, p, None, None, None, None, True)
    y_o, d_o = ours.lod_fwd(om, x, p, None, None, None, None, True)
    y_r = y_r.contiguous()
    assert torch.equal(y_r.view(torch.int16), y_o.view(torch.int16))                       # features: bit-exact
    assert torch.allclose(d_r.reshape(d_o.shape), d_o, rtol=1e-6, atol=1e-7)
    y_c, d_c = olotd.lod_fwd(cm, x[:5000].cpu().numpy(), p.cpu().numpy(), need_input_grad=True)   # the CPU oracle, same check
    assert np.array_equal(y_c.view(np.uint16), y_r[:5000].cpu().numpy().view(np.uint16))
    assert np.allclose(d_c.reshape(5000, -1), d_r.reshape(60000, -1)[:5000].cpu().numpy(), rtol=1e-6, atol=1e-7)
    for ml in (7, 0):
        a, _ = ref.lod_fwd(rm, x, p, None, None, None, ml, False)
        b, _ = ours.lod_fwd(om, x, p, None, None, None, ml, False)
        assert torch.equal(a.contiguous().view(torch.int16), b.view(torch.int16))
    # gradients: the reference accumulates with fp16 atomics (order dependent, saturating); ours in fp32 -> tolerance
    g = torch.from_numpy((rng.normal(size=(60000, 32)) * 0.05).astype(np.float16)).to(cuda)
    gx_r, gp_r = ref.lod_bwd(rm, g, x, p, d_r, None, None, None, None, True, True)
    gx_o, gp_o = ours.lod_bwd(om, g, x, p, d_o, None, None, None, None, True, True)
    assert torch.allclose(gx_r, gx_o, rtol=1e-4, atol=1e-4)
    err = float((gp_r.float() - gp_o.float()).norm() / gp_o.float().norm())
    assert err < 2e-2, err
    gin = torch.from_numpy(rng.normal(size=(60000, 3)).astype(np.float32)).to(cuda)
    a_r, b_r, _ = ref.lod_bwd_bwd_input(rm, gin, g, x, p, d_r.contiguous(), None, None, None, None, True, True, False)
    a_o, b_o, _ = ours.lod_bwd_bwd_input(om, gin, g, x, p, d_o, None, None, None, None, True, True, False)
    assert float((a_r.float() - a_o.float()).norm() / a_o.float().norm()) < 5e-3
    assert float((b_r.float() - b_o.float()).norm() / b_o.float().norm()) < 5e-2


@pytest.mark.parametrize("dt_gamma", [0.0, 0.01])
def test_marching_against_reference_kernels(cuda, dt_gamma):
    from oracle import march as omarch, render as orender, scene as oscene
    from neuralsim_b200.bindings import _occ_grid as ours
    ref = _ref("_occ_grid")
    os_, ds_ = [], []
    for k in range(4):
        o, d = oscene.pinhole_rays(60, 80, oscene.orbit_camera(k, 4, radius=2.5 + 0.3 * k, elev_deg=10 + 15 * k))
        os_.append(o); ds_.append(d)
    rt = orender.ray_test(torch.cat(os_), torch.cat(ds_), near=0.01)
    o, d, near, far = (rt[k].contiguous() for k in ("rays_o", "rays_d", "near", "far"))
    rng = np.random.default_rng(1)
    for grid in (oscene.make_occ_grid(64), torch.from_numpy(rng.random((48, 32, 40)) < 0.1)):
        roi = torch.tensor([-1., -1, -1, 1, 1, 1])
        args = (o.to(cuda), d.to(cuda), near.to(cuda), far.to(cuda), roi.to(cuda), grid.to(cuda))
        r = ref.ray_marching(*args, ref.ContractionType.AABB, 0.005, 0.1, dt_gamma, 1024, True)
        g = ours.ray_marching(*args, ours.ContractionType.AABB, 0.005, 0.1, dt_gamma, 1024, True)
        for a, b in zip(r, g):
            assert torch.equal(a, b)                                   # counts, t_starts, t_ends, ridx, gidx: bit-exact
        c = omarch.ray_marching(o, d, near, far, roi, grid, 0.005, 0.1, dt_gamma, 1024)            # and the C oracle
        assert torch.equal(c[0], r[0].cpu()) and torch.equal(c[1], r[1].squeeze(-1).cpu()) and torch.equal(c[4], r[4].cpu())
        assert int(r[0][:, 1].sum()) > 1000


def test_pack_ops_against_reference_kernels(cuda):
    from oracle import pack_ops as opk
    from neuralsim_b200.bindings import _pack_ops as ours
    from util import random_packs
    ref = _ref("_pack_ops")
    rng = np.random.default_rng(2)
    pi = random_packs(rng, 500, 1, 140).to(cuda)
    S = int(pi[:, 1].sum())
    f = torch.from_numpy(rng.normal(size=(S, 3)).astype(np.float32)).to(cuda)
    assert torch.allclose(ref.packed_sum(f, pi), ours.packed_sum(f, pi), rtol=1e-5, atol=1e-5)
    for ex in (False, True):
        for rev in (False, True):
            assert torch.allclose(ref.packed_cumsum(f, pi, ex, rev), ours.packed_cumsum(f, pi, ex, rev), rtol=1e-5, atol=2e-5)
    assert torch.equal(ref.packed_diff(f, pi, None, None), ours.packed_diff(f, pi, None, None))
    assert torch.equal(ref.packed_backward_diff(f, pi, None, None), ours.packed_backward_diff(f, pi, None, None))
    o = torch.from_numpy((rng.random((500, 3)) + 1).astype(np.float32)).to(cuda)
    assert torch.equal(ref.packed_div(f, o, pi), ours.packed_div(f, o, pi)) and torch.equal(ref.packed_add(f, o, pi), ours.packed_add(f, o, pi))
    a = torch.from_numpy((rng.random(S) ** 3).astype(np.float32)).to(cuda)
    a[torch.rand(S, device=cuda) < 0.3] = 0
    a[torch.rand(S, device=cuda) < 0.02] = 0.999
    w_r = ref.packed_alpha_to_vw_forward(a, pi, 1e-4, 0.0, False)[0]
    w_o = ours.packed_alpha_to_vw_forward(a, pi, 1e-4, 0.0, False)[0]
    assert torch.equal(w_r, w_o)                                        # the serial recurrence: bit-exact
    assert torch.equal(w_r.cpu(), opk.packed_alpha_to_vw_forward(a.cpu(), pi.cpu(), 1e-4, 0.0, False)[0])
    _, info_r, sel_r = ref.packed_alpha_to_vw_forward(a, pi, 1e-4, 0.0, True)
    _, info_o, sel_o = ours.packed_alpha_to_vw_forward(a, pi, 1e-4, 0.0, True)
    assert torch.equal(info_r.long(), info_o) and torch.equal(sel_r, sel_o)
    gw = torch.randn(S, device=cuda)
    ga_r = ref.packed_alpha_to_vw_backward(w_r, gw, a, pi, 1e-4, 0.0)
    ga_o = ours.packed_alpha_to_vw_backward(w_o, gw, a, pi, 1e-4, 0.0)
    assert float((ga_r - ga_o).norm() / ga_r.norm()) < 1e-4
    cdf = ours.packed_cumsum(a, pi, True, False)
    cdf = ours.packed_div(cdf, cdf[pi[:, 0] + pi[:, 1] - 1].clamp_min(1e-5).contiguous(), pi)
    bins = ours.packed_cumsum(torch.rand(S, device=cuda), pi, False, False)
    u = torch.linspace(0, 1, 35, device=cuda)[1:-1].expand(500, 33).contiguous()
    s_r, i_r = ref.packed_invert_cdf(bins, cdf, u, pi)
    s_o, i_o = ours.packed_invert_cdf(bins, cdf, u, pi)
    assert torch.equal(i_r, i_o) and torch.equal(s_r, s_o)
    assert torch.equal(ref.packed_searchsorted(cdf, u, pi), ours.packed_searchsorted(cdf, u, pi))
    pib = random_packs(rng, 500, 1, 40).to(cuda)
    vb = ours.packed_cumsum(torch.rand(int(pib[:, 1].sum()), device=cuda), pib, False, False)
    for b_sorted in (True, False):
        r = ref.try_merge_two_packs_sorted_aligned(bins, pi, vb, pib, b_sorted)
        g = ours.try_merge_two_packs_sorted_aligned(bins, pi, vb, pib, b_sorted)
        assert all(torch.equal(x, y) for x, y in zip(r, g))
    n = pi[:, 1].contiguous()
    assert all(torch.equal(x, y) for x, y in zip(ref.interleave_arange(n, True), ours.interleave_arange(n, True)))
    st, step = torch.randn(500, device=cuda), torch.rand(500, device=cuda)
    assert all(torch.equal(x, y) for x, y in zip(ref.interleave_linstep(st, n, step, True), ours.interleave_linstep(st, n, step, True)))
    v = torch.randn(S, device=cuda)
    v1, v2 = v.clone(), v.clone()
    i1 = ref.packed_sort_qsort(v1, pi, True); i2 = ours.packed_sort_qsort(v2, pi, True)
    assert torch.equal(v1, v2) and torch.equal(v[i1], v[i2])            # quicksort is unstable: compare values
    ids = torch.repeat_interleave(torch.arange(500, device=cuda), n)
    assert torch.equal(ref.mark_pack_boundaries_cuda(ids), ours.mark_pack_boundaries_cuda(ids))


def test_shencoder_against_reference_kernel(cuda):
    from neuralsim_b200.bindings import _shencoder as ours
    ref = _ref("_shencoder")
    v = torch.nn.functional.normalize(torch.randn(5000, 3, device=cuda), dim=-1)
    for C in (1, 2, 3, 4):
        out_r, out_o = torch.empty(5000, C * C, device=cuda), torch.empty(5000, C * C, device=cuda)
        j_r, j_o = torch.empty(5000, 3 * C * C, device=cuda), torch.empty(5000, 3 * C * C, device=cuda)
        ref.sh_encode_forward(v, out_r, 5000, 3, C, True, j_r)
        torch.cuda.synchronize()                                         # the reference launches on the default stream
        ours.sh_encode_forward(v, out_o, 5000, 3, C, True, j_o)
        assert torch.allclose(out_r, out_o, rtol=1e-6, atol=1e-7) and torch.allclose(j_r, j_o, rtol=1e-5, atol=1e-6)
        g = torch.randn(5000, C * C, device=cuda)
        gi_r, gi_o = torch.zeros(5000, 3, device=cuda), torch.zeros(5000, 3, device=cuda)
        ref.sh_encode_backward(g, v, 5000, 3, C, j_r, gi_r)
        torch.cuda.synchronize()
        ours.sh_encode_backward(g, v, 5000, 3, C, j_o, gi_o)
        assert torch.allclose(gi_r, gi_o, rtol=1e-4, atol=1e-5)
