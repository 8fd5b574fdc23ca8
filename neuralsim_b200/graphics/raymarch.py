"""Occupancy-grid ray marching front-end -- API of `nr3d_lib.graphics.raymarch.occgrid_raymarch`
(reference: nr3d_lib/nr3d_lib/graphics/raymarch/occgrid_raymarch.py:25-222, raymarch/__init__.py)."""
from __future__ import annotations

from dataclasses import dataclass, fields

import torch

from ..bindings import _occ_grid as _backend
from .pack_ops import packed_diff


@dataclass
class RaymarchRetBase:
    num_hit_rays: int
    ridx_hit: torch.Tensor        # [num_hit]     indices (into the input rays) of rays that produced samples
    samples: torch.Tensor         # [M,3]
    depth_samples: torch.Tensor   # [M]
    deltas: torch.Tensor          # [M]
    ridx: torch.Tensor            # [M]           ray index of every sample
    pack_infos: torch.Tensor      # [num_hit,2]

    def __iter__(self):
        return iter(tuple(getattr(self, f.name) for f in fields(self)))

    def __getitem__(self, name):
        return getattr(self, name)


@dataclass
class RaymarchRetSingle(RaymarchRetBase):
    gidx: torch.Tensor
    gidx_pack_infos: torch.Tensor


@dataclass
class RaymarchRetBatched(RaymarchRetBase):
    bidx: torch.Tensor
    gidx: torch.Tensor
    gidx_pack_infos: torch.Tensor


def _full(v, like):
    return like.new_full(like.shape[:-1], v) if not isinstance(v, torch.Tensor) else v


def occgrid_raymarch(occ_grid, rays_o, rays_d, near, far, *, constraction="aabb", perturb=False, perturb_before_march=False,
                     roi=None, step_size=1e-3, max_step_size=1e10, dt_gamma=0.0, max_steps=512, step_size_factor=1.0,
                     generator=None) -> RaymarchRetSingle:
    if constraction.lower() != "aabb":
        raise RuntimeError(f"occgrid_raymarch: contraction {constraction!r} is not built (AABB only)")
    step_size, dt_gamma = step_size * step_size_factor, dt_gamma * step_size_factor
    near, far = _full(near, rays_o), _full(far, rays_o)
    if roi is None:
        roi = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=rays_o.dtype, device=rays_o.device)
    if perturb and perturb_before_march:
        near = near + step_size * torch.rand(near.shape, device=near.device, dtype=near.dtype, generator=generator)
    pack_infos, t_starts, t_ends, ridx, gidx = _backend.ray_marching(
        rays_o.contiguous(), rays_d.contiguous(), near.contiguous(), far.contiguous(), roi, occ_grid, _backend.ContractionType.AABB,
        step_size, max_step_size, dt_gamma, max_steps, True)
    ridx, gidx = ridx.long(), gidx.long()
    ridx_hit = pack_infos[..., 1].nonzero().long()[..., 0].contiguous()
    if ridx_hit.numel() == 0:
        return RaymarchRetSingle(0, None, None, None, None, None, None, None, None)
    pack_infos = pack_infos[ridx_hit].contiguous().long()
    t_starts, t_ends = t_starts.squeeze(-1), t_ends.squeeze(-1)
    deltas = t_ends - t_starts
    if perturb and not perturb_before_march:
        # the single-grid variant jitters only what `deltas` is measured from; depths stay at t_starts (:96-110)
        noise = torch.rand(deltas.shape, dtype=deltas.dtype, device=deltas.device, generator=generator)
        deltas = packed_diff(torch.addcmul(t_starts, noise, deltas), pack_infos)
    samples = torch.addcmul(rays_o.index_select(0, ridx), rays_d.index_select(0, ridx), t_starts.unsqueeze(-1))
    return RaymarchRetSingle(ridx_hit.numel(), ridx_hit, samples, t_starts, deltas, ridx, pack_infos, gidx, None)


def occgrid_raymarch_batched(occ_grid, rays_o, rays_d, near, far, bidx=None, *, constraction="aabb", perturb=False,
                             perturb_before_march=False, roi=None, step_size=1e-3, max_step_size=1e10, dt_gamma=0.0,
                             max_steps=512, step_size_factor=1.0, generator=None) -> RaymarchRetBatched:
    """occ_grid [B,X,Y,Z]; rays either carry `bidx` [R] or are laid out [B, R/B] (occgrid_raymarch.py:114-222)."""
    if constraction.lower() != "aabb":
        raise RuntimeError(f"occgrid_raymarch_batched: contraction {constraction!r} is not built (AABB only)")
    step_size, dt_gamma = step_size * step_size_factor, dt_gamma * step_size_factor
    B = occ_grid.shape[0]
    batch_data_size = 0
    if bidx is None:
        batch_data_size = rays_o.shape[1]
        rays_o, rays_d = rays_o.flatten(0, 1), rays_d.flatten(0, 1)
        near = near.flatten() if isinstance(near, torch.Tensor) else near
        far = far.flatten() if isinstance(far, torch.Tensor) else far
    near, far = _full(near, rays_o), _full(far, rays_o)
    if roi is None:
        roi = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=rays_o.dtype, device=rays_o.device).tile(B, 1)
    if perturb and perturb_before_march:
        near = near + step_size * torch.rand(near.shape, device=near.device, dtype=near.dtype, generator=generator)
    pack_infos, t_starts, t_ends, ridx, bidx_out, gidx = _backend.batched_ray_marching(
        rays_o.contiguous(), rays_d.contiguous(), near.contiguous(), far.contiguous(), None if bidx is None else bidx.int().contiguous(),
        batch_data_size, roi, occ_grid, _backend.ContractionType.AABB, step_size, max_step_size, dt_gamma, max_steps, True)
    ridx, gidx, bidx_out = ridx.long(), gidx.long(), bidx_out.long()
    ridx_hit = pack_infos[..., 1].nonzero().long()[..., 0].contiguous()
    if ridx_hit.numel() == 0:
        return RaymarchRetBatched(0, None, None, None, None, None, None, None, None, None)
    pack_infos = pack_infos[ridx_hit].contiguous().long()
    t_starts, t_ends = t_starts.squeeze(-1), t_ends.squeeze(-1)
    deltas = t_ends - t_starts
    t_samples = t_starts
    if perturb and not perturb_before_march:
        noise = torch.rand(deltas.shape, dtype=deltas.dtype, device=deltas.device, generator=generator)
        t_samples = torch.addcmul(t_starts, noise, deltas)
        deltas = packed_diff(t_samples, pack_infos)
    samples = torch.addcmul(rays_o.index_select(0, ridx), rays_d.index_select(0, ridx), t_samples.unsqueeze(-1))
    return RaymarchRetBatched(ridx_hit.numel(), ridx_hit, samples, t_samples, deltas, ridx, pack_infos, bidx_out, gidx, None)
