"""Volume-graphics helpers -- API of `nr3d_lib.graphics.nerf.nerf_utils`
(reference: nr3d_lib/nr3d_lib/graphics/nerf/nerf_utils.py)."""
from __future__ import annotations

import torch

from .pack_ops import packed_alpha_to_vw, packed_cumsum, packed_volume_render_compression  # noqa: F401

__all__ = ["tau_to_alpha", "packed_alpha_to_vw", "packed_volume_render_compression", "packed_tau_to_vw", "ray_alpha_to_vw", "ray_tau_to_vw"]


def tau_to_alpha(tau):
    return 1 - torch.exp(-tau)


def packed_tau_to_vw(tau, pack_infos):
    transmittance = torch.exp(-packed_cumsum(tau.view(-1), pack_infos, exclusive=True))
    return (1 - torch.exp(-tau.view(-1))) * transmittance


def ray_alpha_to_vw(alpha):
    """w_i = alpha_i * prod_{k<i} (1 - alpha_k) along the last dim (nerf_utils.py:98-110)."""
    keep = torch.roll((1 + 1e-10) - alpha, 1, dims=-1)
    keep[..., 0] = 1
    return alpha * torch.cumprod(keep, dim=-1)


def ray_tau_to_vw(tau):
    trans = torch.exp(-torch.cumsum(tau, dim=-1))
    trans = torch.cat([torch.ones_like(trans[..., :1]), trans[..., :-1]], -1)
    return (1 - torch.exp(-tau)) * trans
