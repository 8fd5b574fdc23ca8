"""Occupancy-grid EMA maintenance: the oracle restatement of the reference's update (oracle/occgrid.py), the product's
torch chain (fields/accel.py:_update, runs on CPU) and -- on the GPU -- the three-launch device update (csrc/occ_ema.cu)."""
import numpy as np
import pytest
import torch

from oracle import occgrid as og


def _case(seed, n=20000, res=(16, 12, 20), touch_all=False):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1.02, 1.02, (n, 3)).astype(np.float32)
    if not touch_all:
        pts[:, 0] = np.abs(pts[:, 0])                      # half of the voxels stay untouched
    sdf = (np.linalg.norm(pts, axis=-1) - 0.5).astype(np.float16).astype(np.float32) * rng.choice([1.0, 0.02], n).astype(np.float32)
    grid = rng.uniform(0, 1, res).astype(np.float32)
    pcl = np.where(rng.uniform(0, 1, res) < 0.1, rng.uniform(0, 1, res), 0).astype(np.float32)
    return pts, sdf, grid, pcl


def test_oracle_semantics_known_answers():
    g = np.zeros((2, 2, 2), np.float32)
    g[0, 0, 0], g[1, 1, 1], g[0, 1, 0] = 0.8, 0.5, 0.9
    gidx = np.array([[0, 0, 0], [0, 0, 0], [1, 1, 1]])
    og.update_occ_val_grid_idx(g, gidx, np.array([0.1, 0.3, 0.9], np.float32), 0.5)
    assert g[0, 0, 0] == np.float32(0.4)          # max(0.5 * 0.8, 0.1, 0.3): the decayed old value takes part in the maximum
    assert g[1, 1, 1] == np.float32(0.9)
    assert g[0, 1, 0] == np.float32(0.9)          # untouched voxels do not decay
    assert og.normalized_logistic_density_half(np.array([0.0]), 256.0)[0] == np.float16(1.0)
    assert og.normalized_logistic_density_half(np.array([1.0]), 256.0)[0] == np.float16(0.0)       # 1 / cosh(20) underflows in fp16
    assert np.array_equal(og.voxel_index(np.array([[-1.5, 0.0, 0.999999]]), (4, 4, 4)), [[0, 2, 3]])


@pytest.mark.parametrize("seed", [0, 1])
def test_torch_chain_equals_oracle(seed):
    """the product's torch restatement (scatter_reduce_ 'amax' for torch_scatter's scatter_max), on CPU"""
    from neuralsim_b200.fields.accel import OccGridEma
    pts, sdf, grid, pcl = _case(seed)
    occ = OccGridEma(resolution=list(grid.shape), occ_thre=0.3, ema_decay=0.95, update_from_samples_cfg=dict())
    occ.occ_val_grid = torch.from_numpy(grid.copy())
    ref_grid, ref_occ, _ = og.step_update_occ(grid.copy(), pts, sdf, inv_s=256.0, ema_decay=0.95, occ_thre=0.3, pcl=None)
    occ._update(occ._gidx_of(torch.from_numpy(pts)), occ.occ_val_fn(torch.from_numpy(sdf)), 0.95)
    assert np.array_equal(occ.occ_val_grid.numpy(), ref_grid)
    assert np.array_equal(occ.occ_grid.numpy(), ref_occ)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,with_pcl", [(0, True), (1, False), (2, True)])
def test_device_update_equals_oracle(cuda, seed, with_pcl):
    from neuralsim_b200.fields.accel import OccGridEma
    pts, sdf, grid, pcl = _case(seed, touch_all=(seed == 2))
    occ = OccGridEma(resolution=list(grid.shape), occ_thre=0.3, ema_decay=0.95, update_from_samples_cfg=dict() if with_pcl else None, device=cuda).train()
    occ.occ_val_grid.copy_(torch.from_numpy(grid))
    if with_pcl:
        occ._occ_val_grid_pcl.copy_(torch.from_numpy(pcl))
    ref_grid, ref_occ, _ = og.step_update_occ(grid.copy(), pts, sdf, inv_s=256.0, ema_decay=0.95, occ_thre=0.3, pcl=pcl.copy() if with_pcl else None)
    occ._step_update_device(torch.from_numpy(pts).to(cuda), torch.from_numpy(sdf).to(cuda))
    assert np.array_equal(occ.occ_val_grid.cpu().numpy(), ref_grid)
    assert np.array_equal(occ.occ_grid.cpu().numpy(), ref_occ)
    if with_pcl:
        assert float(occ._occ_val_grid_pcl.abs().sum()) == 0.0


@pytest.mark.gpu
def test_step_uses_the_device_update_and_matches_the_torch_chain(cuda):
    """OccGridEma.step end to end (same random points through both paths)"""
    import neuralsim_b200.fields.accel as A
    outs = []
    for dev_path in (True, False):
        A.DEVICE_EMA = dev_path
        try:
            torch.manual_seed(5)
            occ = A.OccGridEma(resolution=[32, 32, 32], update_from_net_cfg=dict(num_steps=2, num_pts=2 ** 15), n_steps_warmup=0, device=cuda).train()
            c = (torch.arange(32, device=cuda) + 0.5) / 32 * 2 - 1
            gx, gy, gz = torch.meshgrid(c, c, c, indexing="ij")
            occ.set_occ_grid(((gx * gx + gy * gy + gz * gz).sqrt() - 0.5).abs() < 0.1)
            occ._occ_val_grid_pcl[3, 4, 5] = 0.7
            assert occ.step(16, lambda x: (x.norm(dim=-1) - 0.5).half().float())
            outs.append((occ.occ_val_grid.clone(), occ.occ_grid.clone()))
        finally:
            A.DEVICE_EMA = True
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_batched_update_equals_oracle():
    """OccGridEmaBatched.update (CPU torch chain) == the oracle's restatement of update_batched_occ_val_grid_idx_"""
    from neuralsim_b200.fields.accel import OccGridEmaBatched
    rng = np.random.default_rng(3)
    B, res, n = 4, (10, 8, 12), 6000
    pts = rng.uniform(-1.02, 1.02, (n, 3)).astype(np.float32)
    bidx = rng.integers(0, B - 1, n)                               # grid B-1 stays untouched
    sdf = (np.linalg.norm(pts, axis=-1) - 0.5).astype(np.float16).astype(np.float32) * 0.05
    grid = rng.uniform(0, 1, (B,) + res).astype(np.float32)
    occ = OccGridEmaBatched(B, resolution=list(res), update_from_samples_cfg=None)
    occ.occ_val_grid = torch.from_numpy(grid.copy())
    occ.update(torch.from_numpy(pts), torch.from_numpy(bidx), torch.from_numpy(sdf))
    ref = og.update_batched_occ_val_grid_idx(grid.copy(), bidx, og.voxel_index(pts, res), og.normalized_logistic_density_half(sdf, 256.0).astype(np.float32), 0.95)
    assert np.array_equal(occ.occ_val_grid.numpy(), ref)
    assert np.array_equal(occ.occ_val_grid.numpy()[B - 1], grid[B - 1])
    assert np.array_equal(occ.occ_grid.numpy(), ref > np.float32(0.3))


@pytest.mark.gpu
def test_batched_accel_marches_the_conditioned_grids(cuda):
    """set_condition selects the instances' grids; rays carry the batch index; an empty grid yields no sample, a full one marches the box"""
    from neuralsim_b200.fields.accel import OccGridAccelBatched
    acc = OccGridAccelBatched(6, resolution=[16, 16, 16], device=cuda)
    acc.occ.occ_grid[4] = True                                      # instance 4: everything occupied; instance 2: nothing
    acc.set_condition(2, ins_inds_per_batch=torch.tensor([4, 2], device=cuda))
    o = torch.tensor([[0., 0., -3.]] * 4, device=cuda)
    d = torch.tensor([[0., 0., 1.]] * 4, device=cuda)
    bidx = torch.tensor([0, 1, 0, 1], device=cuda)
    ret = acc.cur_batch__ray_march(o, d, bidx, near=torch.full([4], 2.0, device=cuda), far=torch.full([4], 4.0, device=cuda), step_size=0.1, max_steps=64)
    assert ret.ridx_hit.tolist() == [0, 2] and ret.pack_infos[:, 1].tolist() == [20, 20]
    assert bool((ret.bidx == 0).all())
    assert acc.cur_batch__query_occupancy(torch.zeros(2, 3, device=cuda), torch.tensor([0, 1], device=cuda)).tolist() == [True, False]
    pts, b = acc.cur_batch__sample_pts_in_occupied(100)
    assert bool((b == 0).all()) and pts.shape[0] >= 100
