"""Ray sampling helpers -- API of `nr3d_lib.graphics.raysample` (reference: nr3d_lib/nr3d_lib/graphics/raysample.py)."""
from __future__ import annotations

import torch

from .pack_ops import packed_invert_cdf

__all__ = ["batch_sample_step_linear", "packed_sample_cdf", "batch_sample_cdf", "batch_sample_pdf"]


@torch.no_grad()
def batch_sample_step_linear(near, far, num_samples, prefix_shape=None, perturb=False, return_dt=False, generator=None):
    """Evenly spaced depths in [near, far] per ray; with `perturb`, one uniform jitter per stratum
    (raysample.py:285-310).  -> t [..., num_samples] (, deltas)."""
    if prefix_shape is None:
        prefix_shape = [1] if list(near.shape) == [1] else near.squeeze().shape
    near = near.squeeze().expand(prefix_shape).unsqueeze(-1)
    far = far.squeeze().expand(prefix_shape).unsqueeze(-1)
    device, dtype = near.device, near.dtype
    steps = torch.arange(num_samples, device=device)
    if not perturb:
        dt = (far - near) / (num_samples - 1)
        idx = steps.to(dtype)
    else:
        dt = (far - near) / num_samples
        idx = steps + torch.rand([*prefix_shape, num_samples], dtype=dtype, device=device, generator=generator)
    t = torch.addcmul(near, idx.to(dtype), dt)
    if not return_dt:
        return t
    if not perturb:
        return t, dt.expand((*prefix_shape, num_samples))
    deltas = torch.zeros_like(t)
    deltas[..., :-1], deltas[..., -1] = t.diff(dim=-1), (far[..., 0] - near[..., 0]) / num_samples
    return t, deltas


@torch.no_grad()
def packed_sample_cdf(bins, cdfs, pack_infos, num_to_sample, perturb=False, generator=None):
    """Inverse-CDF sampling inside every pack (raysample.py:38-61).  -> (t [P,n], global bin index [P,n])."""
    P, device, dtype = pack_infos.shape[0], bins.device, bins.dtype
    if not perturb:
        u = torch.linspace(0., 1., num_to_sample + 2, device=device, dtype=dtype)[1:-1].expand((P, num_to_sample))
    else:
        u = batch_sample_step_linear(bins.new_zeros(P), bins.new_ones(P), num_to_sample, perturb=True, generator=generator)
    return packed_invert_cdf(bins, cdfs.to(dtype), u.contiguous(), pack_infos)


@torch.no_grad()
def batch_sample_cdf(bins, cdf, num_to_sample, perturb=False, eps=1e-5, generator=None):
    """Batched inverse-CDF sampling (raysample.py:221-262)."""
    prefix, device, dtype = bins.shape[:-1], bins.device, bins.dtype
    if not perturb:
        u = torch.linspace(0., 1., num_to_sample + 2, device=device, dtype=dtype)[1:-1].expand((*prefix, num_to_sample))
    else:
        u = batch_sample_step_linear(bins.new_zeros(prefix), bins.new_ones(prefix), num_to_sample, perturb=True, generator=generator)
    u = u.contiguous()
    inds = torch.searchsorted(cdf.detach(), u, right=False)
    lo, hi = (inds - 1).clamp_min(0), inds.clamp_max(cdf.shape[-1] - 1)
    c0, c1 = cdf.gather(-1, lo), cdf.gather(-1, hi)
    b0, b1 = bins.gather(-1, lo), bins.gather(-1, hi)
    denom = c1 - c0
    denom[denom < eps] = 1
    return b0 + (u - c0) / denom * (b1 - b0)


@torch.no_grad()
def batch_sample_pdf(bins, weights, num_to_sample, perturb=False, eps=1e-5, generator=None):
    pdf = weights / weights.sum(-1, keepdim=True).clamp_min(eps)
    cdf = torch.cat([pdf.new_zeros([*pdf.shape[:-1], 1]), pdf.cumsum(-1)], -1)
    return batch_sample_cdf(bins, cdf, num_to_sample, perturb, eps, generator)
