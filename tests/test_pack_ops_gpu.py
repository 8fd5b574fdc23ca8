"""pack_ops kernels (csrc/pack_ops.cu) against the serial CPU oracle (oracle/pack_ops.py).
Index-valued results are compared bit-exactly, fp32 values within summation-order tolerance."""
import numpy as np
import pytest
import torch

from oracle import pack_ops as opk
from util import random_packs

pytestmark = pytest.mark.gpu


def _g(t, cuda):
    return None if t is None else t.to(cuda)


@pytest.fixture()
def data():
    rng = np.random.default_rng(5)
    pi = random_packs(rng, 700, 0, 150)          # includes empty packs and packs longer than 4 warps' chunk
    pi[0, 1] = 0
    n = pi[:, 1].clone(); pi = torch.stack([n.cumsum(0) - n, n], 1)
    S = int(n.sum())
    return rng, pi, S


@pytest.mark.parametrize("C", [1, 3])
def test_sum_cumsum_diff(cuda, data, C):
    from neuralsim_b200.bindings import _pack_ops as B
    rng, pi, S = data
    f = torch.from_numpy(rng.normal(size=(S,) if C == 1 else (S, C)).astype(np.float32))
    ref = opk.packed_sum(f, pi)
    got = B.packed_sum(f.to(cuda), pi.to(cuda)).cpu()
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)
    for ex in (False, True):
        for rev in (False, True):
            ref = opk.packed_cumsum(f, pi, ex, rev)
            got = B.packed_cumsum(f.to(cuda), pi.to(cuda), ex, rev).cpu()
            assert torch.allclose(got, ref, rtol=1e-5, atol=2e-5), (ex, rev)
    app = torch.from_numpy(rng.normal(size=(pi.shape[0],) if C == 1 else (pi.shape[0], C)).astype(np.float32))
    for kw in (dict(), dict(a=app), dict(b=app)):
        ref = opk.packed_diff(f, pi, kw.get("a"), kw.get("b"))
        got = B.packed_diff(f.to(cuda), pi.to(cuda), _g(kw.get("a"), cuda), _g(kw.get("b"), cuda)).cpu()
        assert torch.equal(got, ref)
        ref = opk.packed_backward_diff(f, pi, kw.get("a"), kw.get("b"))
        got = B.packed_backward_diff(f.to(cuda), pi.to(cuda), _g(kw.get("a"), cuda), _g(kw.get("b"), cuda)).cpu()
        assert torch.equal(got, ref)


def test_binary_ops(cuda, data):
    from neuralsim_b200.bindings import _pack_ops as B
    rng, pi, S = data
    for C in (1, 4):
        f = torch.from_numpy(rng.normal(size=(S,) if C == 1 else (S, C)).astype(np.float32))
        o = torch.from_numpy((rng.normal(size=(pi.shape[0],) if C == 1 else (pi.shape[0], C)) + 3).astype(np.float32))
        for name in ("add", "sub", "mul", "div", "gt", "geq", "lt", "leq", "eq", "neq"):
            ref = getattr(opk, f"packed_{name}")(f, o, pi)
            got = getattr(B, f"packed_{name}")(f.to(cuda), o.to(cuda), pi.to(cuda)).cpu()
            assert torch.equal(got, ref), name


def test_search_invert_cdf(cuda, data):
    from neuralsim_b200.bindings import _pack_ops as B
    rng, pi, S = data
    pi = pi[pi[:, 1] > 0]
    n = pi[:, 1].clone(); pi = torch.stack([n.cumsum(0) - n, n], 1); S = int(n.sum())
    w = torch.from_numpy(rng.random(S).astype(np.float32))
    w[rng.random(S) < 0.3] = 0                                   # flat cdf stretches -> pmf < eps branch
    cdf = opk.packed_cumsum(w, pi, True, False)
    last = cdf[pi[:, 0] + pi[:, 1] - 1].clamp_min(1e-5)
    cdf = opk.packed_div(cdf, last, pi)
    bins = opk.packed_cumsum(torch.from_numpy(rng.random(S).astype(np.float32)), pi, False, False)
    u = torch.linspace(0, 1, 11)[1:-1].expand(pi.shape[0], 9).contiguous()
    s_ref, i_ref = opk.packed_invert_cdf(bins, cdf, u, pi)
    s, i = B.packed_invert_cdf(bins.to(cuda), cdf.to(cuda), u.to(cuda), pi.to(cuda))
    assert torch.equal(i.cpu(), i_ref) and torch.equal(s.cpu(), s_ref)            # same fp32 op sequence -> bit-exact
    assert torch.equal(B.packed_searchsorted(cdf.to(cuda), u.to(cuda), pi.to(cuda)).cpu(), opk.packed_searchsorted(cdf, u, pi))


def test_merge_sorted_aligned(cuda, data):
    from neuralsim_b200.bindings import _pack_ops as B
    rng, pia, _ = data
    pib = random_packs(rng, pia.shape[0], 0, 40)
    va = opk.packed_cumsum(torch.from_numpy(rng.random(int(pia[:, 1].sum())).astype(np.float32)), pia, False, False)
    vb = opk.packed_cumsum(torch.from_numpy(rng.random(int(pib[:, 1].sum())).astype(np.float32)), pib, False, False)
    vb[::7] = va[rng.integers(0, va.shape[0], vb[::7].shape[0])] if va.shape[0] else vb[::7]   # cross ties (unsorted inside b -> fix below)
    # keep b sorted per pack after injecting ties
    for b, n in pib.tolist():
        vb[b:b + n] = vb[b:b + n].sort().values
    ra, rb, rp = opk.try_merge_two_packs_sorted_aligned(va, pia, vb, pib, True)
    ga, gb, gp = B.try_merge_two_packs_sorted_aligned(va.to(cuda), pia.to(cuda), vb.to(cuda), pib.to(cuda), True)
    assert torch.equal(gp.cpu(), rp) and torch.equal(ga.cpu(), ra) and torch.equal(gb.cpu(), rb)
    merged = torch.empty(va.shape[0] + vb.shape[0])
    merged[ga.cpu()], merged[gb.cpu()] = va, vb
    for b, n in rp.tolist():                                      # property: a permutation that sorts every pack
        assert torch.all(merged[b:b + n][1:] >= merged[b:b + n][:-1])
    # reference known-answer vector (pack_ops/unit_test.py:956-965)
    ka = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 0.2, 0.8]); kb = torch.tensor([0.0, 0.25, 0.26, 0.6, 0.1, 0.15, 0.3, 0.4])
    pa = torch.tensor([[0, 5], [5, 2]]); pb = torch.tensor([[0, 4], [4, 4]])
    ga, gb, _ = B.try_merge_two_packs_sorted_aligned(ka.to(cuda), pa.to(cuda), kb.to(cuda), pb.to(cuda), True)
    assert ga.tolist() == [1, 2, 5, 6, 7, 11, 14] and gb.tolist() == [0, 3, 4, 8, 9, 10, 12, 13]
    # b_sorted=False keeps the reference's serial bookkeeping
    ra, rb, _ = opk.try_merge_two_packs_sorted_aligned(ka, pa, kb, pb, False)
    ga, gb, _ = B.try_merge_two_packs_sorted_aligned(ka.to(cuda), pa.to(cuda), kb.to(cuda), pb.to(cuda), False)
    assert torch.equal(ga.cpu(), ra) and torch.equal(gb.cpu(), rb)


def test_alpha_to_vw_and_compression(cuda, data):
    from neuralsim_b200.bindings import _pack_ops as B
    rng, pi, S = data
    a = torch.from_numpy((rng.random(S) ** 3).astype(np.float32))
    a[rng.random(S) < 0.3] = 0.0
    a[rng.random(S) < 0.02] = 0.999                               # drives T below the early-stop threshold
    for thre in (0.0, 0.01):
        w_ref = opk.packed_alpha_to_vw_forward(a, pi, 1e-4, thre, False)[0]
        w = B.packed_alpha_to_vw_forward(a.to(cuda), pi.to(cuda), 1e-4, thre, False)[0]
        assert torch.equal(w.cpu(), w_ref)                        # serial recurrence replayed -> bit-exact
        _, info_ref, sel_ref = opk.packed_alpha_to_vw_forward(a, pi, 1e-4, thre, True)
        _, info, sel = B.packed_alpha_to_vw_forward(a.to(cuda), pi.to(cuda), 1e-4, thre, True)
        assert torch.equal(info.cpu(), info_ref) and torch.equal(sel.cpu(), sel_ref)
        gw = torch.from_numpy(rng.normal(size=S).astype(np.float32))
        ga_ref = opk.packed_alpha_to_vw_backward(w_ref, gw, a, pi, 1e-4, thre)
        ga = B.packed_alpha_to_vw_backward(w, gw.to(cuda), a.to(cuda), pi.to(cuda), 1e-4, thre)
        # cotangents are divided by (1 - alpha) ~ 1e-3 for the near-opaque samples: compare relative to the pack scale
        assert float((ga.cpu() - ga_ref).norm() / ga_ref.norm()) < 1e-4
        assert torch.allclose(ga.cpu(), ga_ref, rtol=1e-3, atol=1e-3 * float(ga_ref.abs().max()))


def test_producers_sort_boundaries(cuda, data):
    from neuralsim_b200.bindings import _pack_ops as B
    rng, pi, S = data
    n = pi[:, 1].contiguous()
    out, nidx = B.interleave_arange(n.to(cuda), True)
    ro, rn = opk.interleave_arange(n, True)
    assert torch.equal(out.cpu(), ro) and torch.equal(nidx.cpu(), rn)
    st = torch.from_numpy(rng.normal(size=n.shape[0]).astype(np.float32)); step = torch.from_numpy(rng.random(n.shape[0]).astype(np.float32))
    out, nidx = B.interleave_linstep(st.to(cuda), n.to(cuda), step.to(cuda), True)
    ro, rn = opk.interleave_linstep(st, n, step, True)
    assert torch.equal(out.cpu(), ro) and torch.equal(nidx.cpu(), rn)
    out, _ = B.interleave_linstep(st.to(cuda), n.to(cuda), 0.25, False)
    assert torch.equal(out.cpu(), opk.interleave_linstep(st, n, 0.25, False)[0])
    v = torch.from_numpy(rng.normal(size=S).astype(np.float32))
    k = min(v[::5].shape[0], v[1::5].shape[0]); v[::5][:k] = v[1::5][:k].clone()      # ties
    vg = v.clone().to(cuda)
    idx = B.packed_sort_qsort(vg, pi.to(cuda), True)
    vr = v.clone(); idx_ref = opk.packed_sort_qsort(vr, pi, True)
    assert torch.equal(vg.cpu(), vr) and torch.equal(idx.cpu(), idx_ref)
    ids = torch.repeat_interleave(torch.arange(pi.shape[0]), n)
    assert torch.equal(B.mark_pack_boundaries_cuda(ids.to(cuda)).cpu(), opk.mark_pack_boundaries_cuda(ids))


def test_autograd_wrappers(cuda, data):
    """graphics.pack_ops gradient rules against torch autograd of an index_add / gather formulation."""
    from neuralsim_b200.graphics import pack_ops as G
    rng, pi, S = data
    pi = pi[pi[:, 1] > 0]
    n = pi[:, 1].clone(); pi = torch.stack([n.cumsum(0) - n, n], 1).to(cuda); S = int(n.sum())
    pidx = torch.repeat_interleave(torch.arange(pi.shape[0], device=cuda), pi[:, 1])
    f = torch.randn(S, 3, device=cuda, requires_grad=True)
    G.packed_sum(f, pi).square().sum().backward()
    f2 = f.detach().clone().requires_grad_(True)
    torch.zeros(pi.shape[0], 3, device=cuda).index_add(0, pidx, f2).square().sum().backward()
    assert torch.allclose(f.grad, f2.grad, rtol=1e-4, atol=1e-4)
    a = torch.randn(S, device=cuda, requires_grad=True); o = (torch.rand(pi.shape[0], device=cuda) + 1).requires_grad_(True)
    (G.packed_div(a, o, pi) * torch.arange(S, device=cuda)).sum().backward()
    a2, o2 = a.detach().clone().requires_grad_(True), o.detach().clone().requires_grad_(True)
    ((a2 / o2[pidx]) * torch.arange(S, device=cuda)).sum().backward()
    assert torch.allclose(a.grad, a2.grad, rtol=1e-5) and torch.allclose(o.grad, o2.grad, rtol=1e-4, atol=1e-2)
    d = torch.randn(S, device=cuda, requires_grad=True)
    w = torch.randn(S, device=cuda)
    (G.packed_diff(d, pi) * w).sum().backward()
    d2 = d.detach().clone().requires_grad_(True)
    last = torch.zeros(S, dtype=torch.bool, device=cuda); last[pi[:, 0] + pi[:, 1] - 1] = True
    (torch.where(last, torch.zeros_like(d2), d2.roll(-1) - d2) * w).sum().backward()
    assert torch.allclose(d.grad, d2.grad, rtol=1e-5, atol=1e-6)
    c = torch.randn(S, device=cuda, requires_grad=True)
    (G.packed_cumsum(c, pi, exclusive=True) * w).sum().backward()
    ref = opk.packed_cumsum(w.cpu(), pi.cpu(), True, True)
    assert torch.allclose(c.grad.cpu(), ref, rtol=1e-4, atol=1e-4)
