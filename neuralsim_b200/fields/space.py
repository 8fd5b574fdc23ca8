"""Axis-aligned bounding space -- API of `nr3d_lib.models.spatial.AABBSpace` (reference: models/spatial/aabb.py:20-99)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..graphics.raytest import ray_box_intersection_fast_float_nocheck


class AABBSpace(nn.Module):
    def __init__(self, bounding_size: float = 2.0, aabb=None, dtype=torch.float, device=None):
        super().__init__()
        if aabb is None:
            h = bounding_size / 2.
            aabb = [[-h, -h, -h], [h, h, h]]
        aabb = torch.as_tensor(aabb, dtype=dtype, device=device)
        self.register_buffer("aabb", aabb, persistent=True)
        self.register_buffer("radius3d_original", (aabb[1] - aabb[0]) / 2., persistent=True)

    @property
    def center(self):
        return (self.aabb[1] + self.aabb[0]) / 2.

    @property
    def radius3d(self):
        return (self.aabb[1] - self.aabb[0]) / 2.

    def normalize_coords(self, x):
        return (x - self.center) / self.radius3d

    def unnormalize_coords(self, x):
        return x * self.radius3d + self.center

    def normalize_rays(self, rays_o, rays_d):
        """So that o + d*t lands in [-1,1]^3 for the same depth t (|d| changes)."""
        return (rays_o - self.center) / self.radius3d, rays_d / self.radius3d

    def sample_pts_uniform(self, num_pts: int, generator=None):
        return torch.empty([num_pts, 3], dtype=self.aabb.dtype, device=self.aabb.device).uniform_(-1, 1, generator=generator)

    def ray_test(self, rays_o, rays_d, near=None, far=None, return_rays=True, normalized=False, **extra_ray_data):
        """Slab test against the unit cube -> dict(num_rays, rays_inds, near, far, rays_o, rays_d, **extras) of the hit rays."""
        if not normalized:
            rays_o, rays_d = self.normalize_rays(rays_o, rays_d)
        with torch.no_grad():
            near_, far_ = ray_box_intersection_fast_float_nocheck(rays_o, rays_d, -1., 1.)
            if near is not None:
                near_.clamp_min_(near)
            if far is not None:
                far_.clamp_max_(far)
            mask = (far_ > near_) & (far_ > (0 if near is None else near))
            if far is not None:
                mask &= near_ < far
            ridx = mask.nonzero().long()[..., 0]
        ret = dict(num_rays=ridx.shape[0], rays_inds=ridx, near=near_[ridx], far=far_[ridx])
        ret.update({k: (v[ridx] if isinstance(v, torch.Tensor) else v) for k, v in extra_ray_data.items()})
        if return_rays:
            ret.update(rays_o=rays_o[ridx], rays_d=rays_d[ridx])
        return ret
