"""LoTD kernels (csrc/lotd.cu through the C ABI / bindings._lotd) against the CPU oracle (oracle/lotd.py)."""
import numpy as np
import pytest
import torch

from oracle import lotd as olotd

pytestmark = pytest.mark.gpu


def _setup(cuda, cfg, n, seed=0, amp=0.1, dtype=np.float16):
    from neuralsim_b200.bindings import _lotd
    rng = np.random.default_rng(seed)
    om = olotd.LoDMeta(3, **cfg)
    gm = _lotd.LoDMeta(3, cfg["lod_res"], cfg["lod_n_feats"], cfg["lod_types"], cfg["hashmap_size"])
    assert gm.n_params == om.n_params and gm.level_offsets == om.level_offsets and gm.level_sizes == om.level_sizes
    p = rng.uniform(-amp, amp, om.n_params).astype(dtype)
    x = rng.uniform(1e-6, 1 - 1e-6, (n, 3)).astype(np.float32)
    x[:8] = np.array([[1e-6] * 3, [1 - 1e-6] * 3, [0.5] * 3, [1e-6, 0.5, 1 - 1e-6], [0.25, 0.75, 0.5], [0.999, 0.001, 0.3],
                      [0.3333, 0.6667, 0.1], [0.125, 0.125, 0.875]], dtype=np.float32)
    return _lotd, om, gm, p, x, torch.from_numpy(p).to(cuda), torch.from_numpy(x).to(cuda)


@pytest.mark.parametrize("full", [False, True])
def test_fwd_bit_exact_fp16(cuda, full):
    from util import small_lotd_cfg
    cfg = olotd.gen_ngp_cfg() if full else small_lotd_cfg()
    _lotd, om, gm, p, x, pg, xg = _setup(cuda, cfg, 20000)
    y_ref, d_ref = olotd.lod_fwd(om, x, p, need_input_grad=True)
    y, d = _lotd.lod_fwd(gm, xg, pg, None, None, None, None, True)
    assert y.dtype == torch.float16 and y.shape == (x.shape[0], om.n_encoded_dims)
    # fp16 features: bit-exact (same rounding sequence as the reference's <float,half,float> kernel)
    assert np.array_equal(y.cpu().numpy().view(np.uint16), y_ref.view(np.uint16))
    d = d.view(x.shape[0], om.n_encoded_dims, 3).cpu().numpy()
    assert np.allclose(d, d_ref, rtol=1e-6, atol=1e-7)
    y2, none = _lotd.lod_fwd(gm, xg, pg, None, None, None, None, False)
    assert none is None and torch.equal(y2, y)


def test_fwd_fp32_params_and_max_level(cuda):
    from util import small_lotd_cfg
    _lotd, om, gm, p, x, pg, xg = _setup(cuda, small_lotd_cfg(), 5000, dtype=np.float32)
    for ml in (None, 3, 0, -1):
        y_ref, _ = olotd.lod_fwd(om, x, p, max_level=ml)
        y, _ = _lotd.lod_fwd(gm, xg, pg, None, None, None, ml, False)
        assert np.allclose(y.cpu().numpy(), y_ref, rtol=1e-6, atol=1e-8), ml


def test_bwd_grid_and_input(cuda):
    from util import small_lotd_cfg
    _lotd, om, gm, p, x, pg, xg = _setup(cuda, small_lotd_cfg(), 30000)
    rng = np.random.default_rng(1)
    g = rng.normal(size=(x.shape[0], om.n_encoded_dims)).astype(np.float16)
    _, d_ref = olotd.lod_fwd(om, x, p, need_input_grad=True)
    gp_ref = olotd.lod_bwd_grid(om, g, x, om.n_params)
    gx_ref = olotd.lod_bwd_input(g, d_ref)
    _, dg = _lotd.lod_fwd(gm, xg, pg, None, None, None, None, True)
    gx, gp = _lotd.lod_bwd(gm, torch.from_numpy(g).to(cuda), xg, pg, dg, None, None, None, None, True, True)
    assert gp.dtype == torch.float16
    assert np.allclose(gx.cpu().numpy(), gx_ref, rtol=1e-5, atol=1e-4)
    # fp32 accumulation then one rounding to fp16: within fp16 resolution of the exact (fp64) sum
    got = gp.float().cpu().numpy().astype(np.float64)
    assert np.all(np.abs(got - gp_ref) <= 1e-3 * np.abs(gp_ref) + 1e-3)
    # max_level masks whole levels
    gx2, gp2 = _lotd.lod_bwd(gm, torch.from_numpy(g).to(cuda), xg, pg, dg, None, None, None, 2, False, True)
    assert gx2 is None and float(gp2[om.level_offsets[3]:].abs().max()) == 0.0


def test_bwd_bwd(cuda):
    from util import small_lotd_cfg
    _lotd, om, gm, p, x, pg, xg = _setup(cuda, small_lotd_cfg(), 20000)
    rng = np.random.default_rng(2)
    g = (rng.normal(size=(x.shape[0], om.n_encoded_dims)) * 0.1).astype(np.float16)
    gin = rng.normal(size=(x.shape[0], 3)).astype(np.float32)
    _, d_ref = olotd.lod_fwd(om, x, p, need_input_grad=True)
    a_ref, b_ref, _ = olotd.lod_bwd_bwd_input(om, gin, g, x, p, d_ref)
    _, dg = _lotd.lod_fwd(gm, xg, pg, None, None, None, None, True)
    a, b, c = _lotd.lod_bwd_bwd_input(gm, torch.from_numpy(gin).to(cuda), torch.from_numpy(g).to(cuda), xg, pg, dg, None, None, None, None,
                                      True, True, False)
    assert c is None
    assert np.allclose(a.float().cpu().numpy(), a_ref, rtol=2e-3, atol=2e-3)          # returned in dL_dy's dtype (fp16)
    got = b.float().cpu().numpy().astype(np.float64)
    assert np.all(np.abs(got - b_ref) <= 2e-3 * np.abs(b_ref) + 5e-2)


def test_errors(cuda):
    from neuralsim_b200.bindings import _lotd
    with pytest.raises(RuntimeError):
        _lotd.LoDMeta(3, [16, 2], [2, 2], ["Dense", "Dense"], None)           # res <= 2
    with pytest.raises(RuntimeError):
        _lotd.LoDMeta(3, [16], [2], ["Hash"], None)                            # hash without hashmap_size
    m = _lotd.LoDMeta(3, [16], [2], ["Dense"], None)
    with pytest.raises(RuntimeError):
        _lotd.lod_fwd(m, torch.zeros(4, 3, device=cuda), torch.zeros(7, device=cuda, dtype=torch.half))
    with pytest.raises(RuntimeError):
        _lotd.lod_fwd(m, torch.zeros(4, 3), torch.zeros(m.n_params, dtype=torch.half))   # CPU tensors: no CPU path
    y, _ = _lotd.lod_fwd(m, torch.zeros(0, 3, device=cuda), torch.zeros(m.n_params, device=cuda, dtype=torch.half))
    assert y.shape == (0, 2)
