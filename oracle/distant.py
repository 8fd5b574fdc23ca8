"""TEST INFRASTRUCTURE (oracle): CPU restatement of the Distant-view (NeRF++) background model and of its merge with the close-range
buffer.  Follows nr3d_lib/models/fields_distant/nerf/renderer_mixin.py:53-85 (ray_box_intersect), :170-288 (_ray_marching, sample_mode box,
interval inverse_proportional), :290-381 (_ray_query_march with the visibility compression), models/fields/nerf/lotd_nerf.py:136-153 (forward:
encoding + identity embedding -> density MLP with softplus output; radiance MLP on SH(v), h, h_appear) and
app/renderers/single_volume_renderer.py:286-375 (distant near = close-range far, per-ray merge of the two sorted buffers).
The 4-D LoTD, the autocast linear layers and the SH basis are the oracle's own (oracle/lotd.py, oracle/nets.py).  torch on CPU, fp32.

PARITY: no golden vector of the reference exists for this model; its building blocks are pinned (LoTD kernels incl. the 4-D instantiation against
the reference's compiled `_lotd`, tests/test_ref_parity_gpu.py; pack ops; SH), the model-level composition here is a restatement.
Only tests/ may import this module."""
import math

import torch

from . import lotd as olotd
from . import nets as onets
from . import render as orender


class DistantParams:
    def __init__(self, lotd_cfg, seed=7, W=64, n_appear=4, lotd_bound=1.0e-1, aabb=None):
        g = torch.Generator().manual_seed(seed)
        self.lotd_cfg = lotd_cfg
        self.meta = olotd.LoDMeta(4, **lotd_cfg)
        nf = self.meta.n_encoded_dims
        self.grid = (torch.rand(self.meta.n_params, generator=g) * 2 - 1) * lotd_bound
        self.den_W1, self.den_b1 = onets.kaiming_linear(g, W, nf + 4)
        self.den_W2, self.den_b2 = onets.kaiming_linear(g, 1, W)
        self.n_appear = n_appear
        in_rad = 16 + nf + n_appear
        self.rad_W1, self.rad_b1 = onets.kaiming_linear(g, W, in_rad)
        self.rad_W2, self.rad_b2 = onets.kaiming_linear(g, W, W)
        self.rad_W3, self.rad_b3 = onets.kaiming_linear(g, 3, W)
        self.aabb = torch.tensor([[-1., -1., -1.], [1., 1., 1.]]) if aabb is None else torch.as_tensor(aabb, dtype=torch.float)

    def tensors(self):
        return dict(grid=self.grid, den_W1=self.den_W1, den_b1=self.den_b1, den_W2=self.den_W2, den_b2=self.den_b2, rad_W1=self.rad_W1,
                    rad_b1=self.rad_b1, rad_W2=self.rad_W2, rad_b2=self.rad_b2, rad_W3=self.rad_W3, rad_b3=self.rad_b3)

    def requires_grad_(self, flag=True):
        for t in self.tensors().values():
            t.requires_grad_(flag)
        return self


def encode(P, x):
    """LoTDEncoding.forward: x in [-1,1]^4 -> x/2+0.5 -> fp16 table lookup (lotd_encoding.py:150-170)"""
    return onets._LoTDFwd.apply(P.meta, (x / 2. + 0.5).detach(), onets._Grid16.apply(P.grid), None)


def density(P, x):
    h = encode(P, x)
    z = torch.relu(onets.linear16(torch.cat([h.float(), x], -1), P.den_W1, P.den_b1))
    out = onets.linear16(z, P.den_W2, P.den_b2)
    sigma = onets.r16(torch.nn.functional.softplus(out))[..., 0]            # nn.Softplus() as the output activation, under autocast
    return sigma, h


def forward(P, x, v, h_appear):
    sigma, h = density(P, x)
    parts = [onets.sh_encode(v, 4), h.float()]
    if P.n_appear:
        parts.append(h_appear)
    a = torch.relu(onets.linear16(torch.cat(parts, -1), P.rad_W1, P.rad_b1))
    a = torch.relu(onets.linear16(a, P.rad_W2, P.rad_b2))
    return sigma, onets.r16(torch.sigmoid(onets.linear16(a, P.rad_W3, P.rad_b3)))


def ray_box_intersect(o, d, r):
    o, d, r = o.unsqueeze(1), d.unsqueeze(1), r[..., None]
    t_min, t_max = (-r - o) / d, (r - o) / d
    t_near = torch.minimum(t_min, t_max).max(dim=-1).values
    t_far = torch.maximum(t_min, t_max).min(dim=-1).values
    t_far[~((t_far > t_near) & (t_far > 0))] = math.nan
    return t_far


def march_shells(P, rays_o, rays_d, t_min, *, radius_scale_min=1.0, radius_scale_max=1000.0, max_steps=64, include_inf_distance=True):
    n = rays_o.shape[0]
    a, b = 1. / radius_scale_min, 1. / radius_scale_max
    r_reci = torch.arange(a, b, (b - a) / max_steps, dtype=torch.float32).expand(n, -1)
    r = r_reci.reciprocal()
    r_ext = torch.cat([r, torch.full([n, 1], 1.0e10 if include_inf_distance else radius_scale_max)], -1)
    c, rad = (P.aabb[1] + P.aabb[0]) / 2., (P.aabb[1] - P.aabb[0]) / 2.
    o_n, d_n = (rays_o - c) / rad, rays_d / rad
    t_ext = ray_box_intersect(o_n, d_n, r_ext)
    deltas, t = t_ext.diff(dim=-1), t_ext[:, :-1]
    x = torch.addcmul(o_n.unsqueeze(-2), d_n.unsqueeze(-2), t.unsqueeze(-1))
    x4 = torch.cat([x * r_reci.unsqueeze(-1), r_reci.unsqueeze(-1) * 2. - 1], -1)
    valid = ~(torch.isnan(t) | (t < t_min[:, None]))
    ridx, pidx = valid.nonzero().t()
    if ridx.numel() == 0:
        return None
    pack_infos = orender.get_pack_infos_from_n(valid.sum(-1))
    ridx_hit = pack_infos[:, 1].nonzero()[:, 0]
    return ridx_hit, x4[ridx, pidx], t[ridx, pidx], deltas[ridx, pidx], ridx, pack_infos[ridx_hit]


def ray_query(P, rays_o, rays_d, near, rays_h_appear=None, **march_kw):
    """-> packed volume buffer (rays_inds_hit, pack_infos_hit, t, opacity_alpha, rgb) of ALL given rays, or an empty one"""
    m = march_shells(P, rays_o, rays_d, near, **march_kw)
    if m is None:
        return dict(type="empty", rays_inds_hit=[])
    ridx_hit, samples, depth, deltas, ridx, pack_infos = m
    with torch.no_grad():
        alphas = 1 - torch.exp(-(density(P, samples)[0].float() * deltas))
    nidx, pack_infos, pidx = orender.packed_volume_render_compression(alphas, pack_infos)
    if nidx.numel() == 0:
        return dict(type="empty", rays_inds_hit=[])
    ridx_hit, samples, depth, deltas, ridx = ridx_hit[nidx], samples[pidx], depth[pidx], deltas[pidx], ridx[pidx]
    v = rays_d / rays_d.norm(dim=-1, keepdim=True).clamp_min(1e-10)
    sigma, rgb = forward(P, samples, v[ridx], None if rays_h_appear is None else rays_h_appear[ridx])
    return dict(type="packed", rays_inds_hit=ridx_hit, pack_infos_hit=pack_infos, t=depth, opacity_alpha=1 - torch.exp(-(sigma.float() * deltas)), rgb=rgb.float())


def merge_buffers(vb_cr, vb_dv, num_rays):
    """the per-ray merge of two depth-sorted packed buffers (single_volume_renderer.py:337-375), written as a stable sort by (ray, depth)"""
    bufs = [b for b in (vb_cr, vb_dv) if b["type"] != "empty"]
    if not bufs:
        return dict(type="empty", rays_inds_hit=[])
    ray = torch.cat([torch.repeat_interleave(b["rays_inds_hit"], b["pack_infos_hit"][:, 1]) for b in bufs])
    t = torch.cat([b["t"] for b in bufs])
    order = torch.argsort(t, stable=True)
    order = order[torch.argsort(ray[order], stable=True)]
    counts = torch.bincount(ray, minlength=num_rays)
    hit = counts.nonzero()[:, 0]
    out = dict(type="packed", rays_inds_hit=hit, pack_infos_hit=orender.get_pack_infos_from_n(counts)[hit], t=t[order],
               opacity_alpha=torch.cat([b["opacity_alpha"] for b in bufs])[order])
    out["rgb"] = torch.cat([b["rgb"] if "rgb" in b else torch.zeros(b["t"].numel(), 3) for b in bufs])[order]
    out["nablas"] = torch.cat([b["nablas"] if "nablas" in b else torch.zeros(b["t"].numel(), 3) for b in bufs])[order]
    return out
