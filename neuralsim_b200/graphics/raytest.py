"""Ray / box tests -- API of `nr3d_lib.graphics.raytest` (reference: nr3d_lib/nr3d_lib/graphics/raytest.py:150-175)."""
from __future__ import annotations

import torch


def ray_box_intersection_fast_float_nocheck(rays_o, rays_d, aabb_min: float, aabb_max: float):
    """Slab test against the axis-aligned cube [aabb_min, aabb_max]^3 -> (t_near, t_far), no validity mask."""
    t_a = (aabb_min - rays_o) / rays_d
    t_b = (aabb_max - rays_o) / rays_d
    return torch.minimum(t_a, t_b).max(dim=-1).values, torch.maximum(t_a, t_b).min(dim=-1).values


def ray_box_intersection_fast_float(rays_o, rays_d, aabb_min: float, aabb_max: float):
    t_near, t_far = ray_box_intersection_fast_float_nocheck(rays_o, rays_d, aabb_min, aabb_max)
    return t_near, t_far, (t_far > t_near) & (t_far > 0)
