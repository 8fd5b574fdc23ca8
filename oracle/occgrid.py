"""TEST INFRASTRUCTURE (oracle): CPU restatement of the occupancy-grid EMA maintenance of the reference.

Follows nr3d_lib/models/accelerations/occgrid/utils.py:63-68 (`sdf_to_occ_val` -> `normalized_logistic_density`,
nr3d_lib/maths/common.py:122-133), :84-87 (`binarize`), :89-109 (`update_occ_val_grid_idx_`, `update_occ_val_grid_`) and
occgrid/ema_single.py:176-190 (`_step_update_occ`).  numpy only; every function cites what it restates.

PARITY UNPINNED: the reference computes the per-voxel maximum with torch_scatter's `scatter_max(out=...)`, which is not installed
here and has no test in the reference; this file restates its documented semantics (`out` takes part in the maximum,
utils.py:96-99) with `np.maximum.at`.  Only tests/ may import this module.
"""
import numpy as np


def normalized_logistic_density_half(sdf, inv_s):
    """(1 / cosh(clamp(inv_s x / 2, -20, 20)))^2 evaluated on a HALF tensor as torch does: every op computes in fp32 and rounds its
    result to fp16 (maths/common.py:133; the sdf the model returns is half, lotd_sdf.py under autocast)."""
    x = np.asarray(sdf, dtype=np.float32).astype(np.float16)
    a = (x.astype(np.float32) * np.float32(inv_s)).astype(np.float16)
    a = (a.astype(np.float32) / np.float32(2.)).astype(np.float16)
    a = np.clip(a, np.float16(-20), np.float16(20))
    with np.errstate(over="ignore"):                       # cosh(20) = 2.4e8 is inf in fp16, as in torch
        c = np.cosh(a.astype(np.float32)).astype(np.float16)
    r = (np.float32(1.) / c.astype(np.float32)).astype(np.float16)
    return (r.astype(np.float32) * r.astype(np.float32)).astype(np.float16)


def voxel_index(pts, res):
    """((pts/2 + 0.5) * res).long().clamp(0, res-1)   (ema_single.py:179, utils.py:107) -- fp32 ops, truncation toward zero"""
    p = np.asarray(pts, dtype=np.float32)
    u = (p / np.float32(2.) + np.float32(0.5)).astype(np.float32) * np.asarray(res, dtype=np.float32)
    return np.clip(np.trunc(u).astype(np.int64), 0, np.asarray(res, dtype=np.int64) - 1)


def update_occ_val_grid_idx(grid, gidx, occ_val, ema_decay=1.0):
    """utils.py:89-101: new = scatter_max(occ_val, ravel(gidx), out = ema_decay * grid); grid[gidx] = new[ravel(gidx)].  In place; returns grid."""
    shape = grid.shape
    flat = (gidx * np.array([shape[1] * shape[2], shape[2], 1], dtype=np.int64)).sum(-1)
    new = (np.float32(ema_decay) * grid.reshape(-1).astype(np.float32)).astype(np.float32)
    np.maximum.at(new, flat, np.asarray(occ_val, dtype=np.float32).reshape(-1))
    grid.reshape(-1)[flat] = new[flat]
    return grid


def binarize(grid, thre):
    """utils.py:84-87 with consider_mean=False"""
    return grid > np.float32(thre)


def step_update_occ(occ_val_grid, pts, sdf, *, inv_s, ema_decay, occ_thre, pcl=None):
    """ema_single.py:176-190.  -> (occ_val_grid (updated in place), occ_grid bool, pcl (zeroed))"""
    res = occ_val_grid.shape
    occ_val = normalized_logistic_density_half(sdf, inv_s).astype(np.float32)
    gidx = voxel_index(np.asarray(pts).reshape(-1, 3), res)
    if pcl is not None:
        idx = np.argwhere(pcl != 0)
        if idx.size:
            gidx = np.concatenate([gidx, idx.astype(np.int64)], 0)
            occ_val = np.concatenate([occ_val.reshape(-1), pcl[tuple(idx.T)].astype(np.float32)], 0)
        pcl[...] = 0
    update_occ_val_grid_idx(occ_val_grid, gidx, occ_val, ema_decay)
    return occ_val_grid, binarize(occ_val_grid, occ_thre), pcl


def update_batched_occ_val_grid_idx(grid, bidx, gidx, occ_val, ema_decay=1.0):
    """utils.py:111-120 (per-point batch indices): the same update on a [B,X,Y,Z] grid, raveled with the batch index in front.  In place."""
    shape = grid.shape[1:]
    flat = np.asarray(bidx, dtype=np.int64) * int(np.prod(shape)) + (gidx * np.array([shape[1] * shape[2], shape[2], 1], dtype=np.int64)).sum(-1)
    new = (np.float32(ema_decay) * grid.reshape(-1).astype(np.float32)).astype(np.float32)
    np.maximum.at(new, flat, np.asarray(occ_val, dtype=np.float32).reshape(-1))
    grid.reshape(-1)[flat] = new[flat]
    return grid
