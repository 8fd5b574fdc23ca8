"""Synthetic scenes for the oracle, the tests and bench.py (no dataset / checkpoint exists offline).
TEST INFRASTRUCTURE (bench.py builds the same scene through its own numpy code path; see bench.py).

SURVEY.md §8(d) "Synthetic inputs":
  cfg 1  analytic sphere SDF |x|-0.5, camera at (-4,0,0), 64x64 pinhole rays, 32 uniform samples,
         inv_s=64, rgb = 0.5+0.5*normalize(grad sdf)  -> `sphere_cfg1_forward_backward`
  cfg 2  CFG-sized LoTDNeuS whose SDF is a noisy sphere, 64^3 occupancy grid, 800x600 pinhole rays on a
         radius-3 orbit.
The real pipeline reaches a sphere-like SDF by `pretrain_sdf_sphere` (nr3d_lib/models/fields/sdf/utils.py:53-,
500 Adam steps).  Here the same state is *constructed*: the coarsest-but-one dense level stores the sampled
sphere SDF in its feature 0, two hidden units of the decoder pass it through exactly
(softplus_b(k s) - softplus_b(-k s) = k s), every other weight / table entry keeps its random init so that all
16 levels and all hidden units contribute to value and gradient.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import nets as onets


def pinhole_rays(H, W, cam_pos, look_at=(0., 0., 0.), up=(0., 0., 1.), focal=None, dtype=torch.float32):
    """OpenCV-style pinhole with half-pixel offset (app/resources/observers cameras/pinhole.py:176-178).
    -> rays_o [H*W,3], rays_d [H*W,3] (unit norm)."""
    focal = (H + W) / 2. if focal is None else focal
    cam_pos = torch.tensor(cam_pos, dtype=torch.float64)
    fwd = torch.tensor(look_at, dtype=torch.float64) - cam_pos
    fwd = fwd / fwd.norm()
    upv = torch.tensor(up, dtype=torch.float64)
    right = torch.linalg.cross(fwd, upv)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    x = (i + 0.5 - W / 2.) / focal
    y = (j + 0.5 - H / 2.) / focal
    d = x[..., None] * right + y[..., None] * down + fwd
    d = d / d.norm(dim=-1, keepdim=True)
    o = cam_pos.expand_as(d)
    return o.reshape(-1, 3).to(dtype).contiguous(), d.reshape(-1, 3).to(dtype).contiguous()


def orbit_camera(k, n, radius=3.0, elev_deg=20.0):
    a = 2 * math.pi * k / max(n, 1)
    e = math.radians(elev_deg)
    return (radius * math.cos(e) * math.cos(a), radius * math.cos(e) * math.sin(a), radius * math.sin(e))


def make_sphere_params(seed=42, radius=0.5, noise=2.0e-3, k_pass=8.0, lotd_cfg=None, sdf_level=5, ln_inv_s_init=0.5298):
    """LoTDNeuSParams whose decoded SDF ~ |x| - radius (+ random detail).  inv_s = exp(10*ln_inv_s_init) ~ 200."""
    P = onets.LoTDNeuSParams(seed=seed, lotd_cfg=lotd_cfg, lotd_bound=noise, ln_inv_s_init=ln_inv_s_init)
    meta = P.meta
    assert meta.level_types[sdf_level] == 0, "sdf_level must be a Dense level"
    res = meta.level_res_multidim[sdf_level]
    # vertex v of a level sits at x01 = (v - 0.5) / (res - 2)  (pos = x01*(res-2)+0.5), x = 2*x01 - 1
    ax = [(torch.arange(r, dtype=torch.float64) - 0.5) / (r - 2) * 2 - 1 for r in res]
    gx, gy, gz = torch.meshgrid(*ax, indexing="ij")
    s = torch.sqrt(gx * gx + gy * gy + gz * gz) - radius
    off = meta.level_offsets[sdf_level]
    nf = meta.level_n_feats[sdf_level]
    lvl = P.grid[off:off + meta.level_n_params[sdf_level]].view(*res, nf)
    lvl[..., 0] = s.to(torch.float32)
    f_idx = sum(meta.level_n_feats[:sdf_level])          # column of that feature in h
    W = P.dec_W1.shape[0]
    # shrink the random part so the sphere dominates, then wire the pass-through pair
    P.dec_W1[:, f_idx] *= 0.0
    P.dec_W2 *= 0.05
    P.dec_b2.zero_()
    P.dec_W1[0].zero_(); P.dec_W1[1].zero_()
    P.dec_W1[0, f_idx] = k_pass
    P.dec_W1[1, f_idx] = -k_pass
    P.dec_b1[0] = 0.; P.dec_b1[1] = 0.
    P.dec_W2[0, 0] = 1. / k_pass
    P.dec_W2[0, 1] = -1. / k_pass
    return P


def make_occ_grid(res=64, radius=0.5, band=0.012):
    """bool[res,res,res]: voxels whose cube can intersect the |sdf| < band shell of the sphere
    (stands in for OccGridEma.init(from_net), ema_single.py:133-190, thresholded at occ_thre)."""
    c = (torch.arange(res, dtype=torch.float64) + 0.5) / res * 2 - 1
    gx, gy, gz = torch.meshgrid(c, c, c, indexing="ij")
    half_diag = math.sqrt(3.) / res
    return ((torch.sqrt(gx * gx + gy * gy + gz * gz) - radius).abs() < band + half_diag).contiguous()


# ------------------------------------------------------------------------------------------------
# cfg 1: the reference's pure-PyTorch CPU path
# ------------------------------------------------------------------------------------------------
def sphere_cfg1_forward_backward(H=64, W=64, num_samples=32, inv_s=64.0, radius=0.5):
    """neus_ray_sdf_to_alpha (neus_utils.py:59-77) + ray_alpha_to_vw (nerf_utils.py:98-110) on an analytic
    sphere; returns (rendered dict, loss).  Pure torch, runs on the CPU threads torch is given."""
    rays_o, rays_d = pinhole_rays(H, W, (-4., 0., 0.))
    t0 = (-1. - rays_o) / rays_d
    t1 = (1. - rays_o) / rays_d
    near = torch.minimum(t0, t1).max(-1).values.clamp_min(0.01)
    far = torch.maximum(t0, t1).min(-1).values
    ok = far > near
    rays_o, rays_d, near, far = rays_o[ok], rays_d[ok], near[ok], far[ok]
    t = near[:, None] + (far - near)[:, None] * torch.linspace(0, 1, num_samples + 1)[None, :]
    x = (rays_o[:, None, :] + rays_d[:, None, :] * t[..., None]).requires_grad_(True)
    sdf = x.norm(dim=-1) - radius
    nablas = torch.autograd.grad(sdf.sum(), x, create_graph=True)[0]
    cdf = torch.sigmoid(sdf * inv_s)
    alpha = (-1 * cdf.diff(dim=-1) / (cdf[..., :-1] + 1e-5)).clamp_min(0)
    shifted = torch.roll((1 + 1e-10) - alpha, 1, dims=-1)
    shifted[..., 0] = 1
    vw = alpha * torch.cumprod(shifted, dim=-1)
    tm = 0.5 * (t[..., 1:] + t[..., :-1])
    nm = torch.nn.functional.normalize(nablas[:, :-1], dim=-1)
    out = dict(mask_volume=vw.sum(-1), depth_volume=(vw * tm).sum(-1) / (vw.sum(-1) + 1e-10),
               rgb_volume=(vw[..., None] * (0.5 + 0.5 * nm)).sum(-2), normals_volume=(vw[..., None] * nablas[:, :-1]).sum(-2))
    loss = sum(v.mean() for v in out.values())
    loss.backward()
    return out, loss, int(ok.sum())
