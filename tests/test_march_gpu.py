"""Occupancy-grid marching (csrc/march.cu) against oracle/march.c: integer outputs and t values bit-exact."""
import numpy as np
import pytest
import torch

from oracle import march as omarch
from oracle import render as orender
from oracle import scene as oscene

pytestmark = pytest.mark.gpu


def _rays(n_views=3, H=48, W=64):
    os_, ds_ = [], []
    for k in range(n_views):
        o, d = oscene.pinhole_rays(H, W, oscene.orbit_camera(k, n_views, radius=3.0, elev_deg=15 + 20 * k))
        os_.append(o); ds_.append(d)
    o, d = torch.cat(os_), torch.cat(ds_)
    rt = orender.ray_test(o, d, near=0.01)
    return rt["rays_o"].contiguous(), rt["rays_d"].contiguous(), rt["near"].contiguous(), rt["far"].contiguous()


@pytest.mark.parametrize("grid_kind", ["sphere", "random", "empty", "full"])
@pytest.mark.parametrize("dt_gamma", [0.0, 0.01])
def test_single_grid_bit_exact(cuda, grid_kind, dt_gamma):
    from neuralsim_b200.bindings import _occ_grid
    o, d, near, far = _rays()
    rng = np.random.default_rng(3)
    if grid_kind == "sphere":
        grid = oscene.make_occ_grid(64)
    elif grid_kind == "random":
        grid = torch.from_numpy(rng.random((32, 24, 40)) < 0.15)
    elif grid_kind == "empty":
        grid = torch.zeros(16, 16, 16, dtype=torch.bool)
    else:
        grid = torch.ones(16, 16, 16, dtype=torch.bool)
    roi = torch.tensor([-1., -1, -1, 1, 1, 1])
    ref = omarch.ray_marching(o, d, near, far, roi, grid, 0.005, 0.05, dt_gamma, 256)
    got = _occ_grid.ray_marching(o.to(cuda), d.to(cuda), near.to(cuda), far.to(cuda), roi.to(cuda), grid.to(cuda),
                                 _occ_grid.ContractionType.AABB, 0.005, 0.05, dt_gamma, 256, True)
    assert torch.equal(got[0].cpu(), ref[0])                                   # packed_info
    assert torch.equal(got[3].cpu(), ref[3]) and torch.equal(got[4].cpu(), ref[4])   # ridx, gidx
    assert torch.equal(got[1].squeeze(-1).cpu(), ref[1]) and torch.equal(got[2].squeeze(-1).cpu(), ref[2])  # t bit-exact
    if grid_kind == "empty":
        assert got[1].shape[0] == 0
    if grid_kind == "full" and dt_gamma == 0.0:
        assert int(got[0][:, 1].max()) == 256                                  # max_steps cap


def test_batched_bit_exact(cuda):
    from neuralsim_b200.bindings import _occ_grid
    o, d, near, far = _rays(2, 32, 40)
    rng = np.random.default_rng(4)
    B = 5
    grid = torch.from_numpy(rng.random((B, 16, 16, 16)) < 0.2)
    roi = torch.tensor([-1., -1, -1, 1, 1, 1]).tile(B, 1).contiguous()
    bi = torch.from_numpy(rng.integers(-1, B, o.shape[0])).int()
    ref = omarch.ray_marching(o, d, near, far, roi, grid, 0.01, 1e10, 0.0, 128, batch_inds=bi)
    got = _occ_grid.batched_ray_marching(o.to(cuda), d.to(cuda), near.to(cuda), far.to(cuda), bi.to(cuda), 0, roi.to(cuda), grid.to(cuda),
                                         _occ_grid.ContractionType.AABB, 0.01, 1e10, 0.0, 128, True)
    info, t0, t1, ridx, bidx, gidx = got
    assert torch.equal(info.cpu(), ref[0]) and torch.equal(ridx.cpu(), ref[3]) and torch.equal(gidx.cpu(), ref[4])
    assert torch.equal(bidx.cpu(), ref[5]) and torch.equal(t0.squeeze(-1).cpu(), ref[1])
    assert int(info[bi.to(cuda) < 0][:, 1].sum()) == 0                          # rays with batch_ind < 0 produce nothing


@pytest.mark.parametrize("grid_kind", ["sphere", "random", "full"])
def test_recorded_march_equals_two_rounds(cuda, grid_kind):
    """nsb_ray_marching_record + nsb_march_compact (march once, copy) == the two-round march bit for bit, with the ray count on the device"""
    import ctypes
    from neuralsim_b200 import _lib as L
    from neuralsim_b200.bindings import _occ_grid
    o, d, near, far = (x.to(cuda) for x in _rays(2, 40, 56))
    rng = np.random.default_rng(5)
    grid = {"sphere": oscene.make_occ_grid(64), "random": torch.from_numpy(rng.random((32, 24, 40)) < 0.15),
            "full": torch.ones(16, 16, 16, dtype=torch.bool)}[grid_kind].to(cuda)
    roi = torch.tensor([-1., -1, -1, 1, 1, 1], device=cuda)
    R, max_steps, n_live = o.shape[0], 192, o.shape[0] - 37
    ref = _occ_grid.ray_marching(o[:n_live].contiguous(), d[:n_live].contiguous(), near[:n_live].contiguous(), far[:n_live].contiguous(), roi, grid,
                                 _occ_grid.ContractionType.AABB, 0.005, 0.05, 0.0, max_steps, False)
    lib, P = L.lib(), L.ptr
    cnt = torch.tensor([n_live, 0], dtype=torch.int64, device=cuda)
    num = torch.full((R,), -1, dtype=torch.int32, device=cuda)
    rec = torch.full((R * max_steps,), float("nan"), device=cuda)
    g8 = grid.contiguous().view(torch.uint8)
    lib.nsb_bind_device_counts(ctypes.c_void_p(cnt.data_ptr()), None)
    L.check(lib.nsb_ray_marching_record(L.c_i64(R), P(o.contiguous()), P(d.contiguous()), P(near), P(far), P(roi), L.c_i32(grid.shape[0]), L.c_i32(grid.shape[1]),
                                        L.c_i32(grid.shape[2]), P(g8), L.c_f32(0.005), L.c_f32(0.05), L.c_f32(0.0), ctypes.c_uint32(max_steps), P(num), P(rec),
                                        None, L.stream_ptr()), "record")
    assert torch.equal(num[:n_live], ref[0][:, 1]) and int(num[n_live:].abs().sum()) == 0
    info2 = torch.stack([num.cumsum(0, dtype=torch.int32) - num, num], -1).contiguous()
    hit = torch.nonzero(num).flatten().contiguous()
    M = int(num.sum())
    assert M == ref[1].shape[0] and M > 0
    t, ridx = torch.full((M + 8,), -7.0, device=cuda), torch.full((M + 8,), -7, dtype=torch.int32, device=cuda)
    cnt[1] = hit.numel()
    lib.nsb_bind_device_counts(ctypes.c_void_p(cnt.data_ptr() + 8), None)
    L.check(lib.nsb_march_compact(P(rec), ctypes.c_uint32(max_steps), P(info2), P(hit), L.c_i64(R), P(t), P(ridx), L.stream_ptr()), "compact")
    assert torch.equal(t[:M], ref[1].squeeze(-1)) and torch.equal(ridx[:M], ref[3])
    assert bool((t[M:] == -7.0).all()) and bool((ridx[M:] == -7).all())                     # nothing beyond the packed samples
    if grid_kind == "full":
        assert int(num.max()) == max_steps
