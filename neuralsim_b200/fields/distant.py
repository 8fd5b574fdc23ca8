"""Distant-view (NeRF++) background model -- API of `app.models.single.LoTDNeRFDistant` = `NeRFRendererMixinDistant` + `LoTDNeRF`
(reference: nr3d_lib/models/fields_distant/nerf/renderer_mixin.py:87-440, models/fields/nerf/lotd_nerf.py:32-190, lotd_cfg.py:136-194;
present in every shipped config, e.g. code_single/configs/object_centric/lotd_neus.bmvs.230814.yaml:200-262).

Everything outside the close-range box is sampled on `max_steps` cuboid shells r aabb with 1/r uniform in [1/radius_scale_max, 1/radius_scale_min]
(+ one shell at 1e10), encoded by a 4-D LoTD in (x / r, 2 / r - 1), turned into density and colour by two small MLPs, compressed with the
same visibility pass as the close-range buffer and merged with it per ray (`merge_two_packs_sorted`, single_volume_renderer.py:337-375).

Kernels: the 4-D LoTD instantiation of csrc/lotd.cu, the pack_ops kernels, the SH kernel; the MLPs are the reference's autocast layers (cuBLAS).
This model is NOT fused: it is the `next` row of SURVEY.md §8(f) built op by op on this library's kernels.
"""
from __future__ import annotations

import math
from operator import itemgetter

import numpy as np
import torch
import torch.nn as nn

from ..graphics.nerf import packed_volume_render_compression
from ..graphics.pack_ops import get_pack_infos_from_n
from .encoding import LoTDEncoding
from .networks import MLP, RadianceNet
from .space import AABBSpace

__all__ = ["auto_ngp4d_cfg", "LoTDNeRF", "LoTDNeRFDistant", "ray_box_intersect", "tau_to_alpha", "shell_radii"]


def auto_ngp4d_cfg(dim=4, n_feats=2, stretch=1.0, target_num_params=2 ** 32, max_layers=128, min_dense_layers=0, log2_hashmap_size=19, min_res_xyz=4,
                   min_res_w=4, per_level_scale=1.382):
    """[Dense -> Hash] ladder for the (xyz / r, 1 / r) input of NeRF++ (lotd_cfg.py:136-194): resolution `min_res_xyz` x aspect on xyz and `min_res_w`
    on the fourth axis, every level per_level_scale finer, Hash once a level exceeds the table, levels added while the budget lasts."""
    hashmap_size = 2 ** log2_hashmap_size
    st = np.array([stretch] * (dim - 1) if np.isscalar(stretch) else list(stretch), dtype=np.float64)
    base = np.concatenate([min_res_xyz * st / st.min(), np.array([min_res_w], dtype=np.float32)])
    num, res, types = 0, [], []
    for i in range(max_layers):
        r = np.ceil(base).astype(np.int64)
        cells = int(r.prod())
        if cells > hashmap_size and i >= min_dense_layers:
            t, n = "Hash", hashmap_size * n_feats
        else:
            t, n = "Dense", cells * n_feats
        if num + n > target_num_params:
            break
        res.append(r.tolist()); types.append(t)
        num += n
        base = base * per_level_scale
    return dict(lod_res=res, lod_n_feats=[n_feats] * len(res), lod_types=types, hashmap_size=hashmap_size)


def tau_to_alpha(tau):
    return 1 - torch.exp(-tau)                     # graphics/nerf/nerf_utils.py:23-24


def ray_box_intersect(rays_o, rays_d, r):
    """far intersection depth of every ray with every box [-r, r]^3; NaN where the ray misses (fields_distant/nerf/renderer_mixin.py:53-85)"""
    o, d, r = rays_o.unsqueeze(1), rays_d.unsqueeze(1), r[..., None]
    t_min, t_max = (-r - o) / d, (r - o) / d
    t_near = torch.minimum(t_min, t_max).max(dim=-1).values
    t_far = torch.maximum(t_min, t_max).min(dim=-1).values
    t_far[~((t_far > t_near) & (t_far > 0))] = math.nan
    return t_far


def shell_radii(radius_scale_min, radius_scale_max, max_steps):
    """1 / r uniform between 1 / radius_scale_min and 1 / radius_scale_max (`inverse_proportional`, renderer_mixin.py:181-188).  Evaluated on
    the host in fp32 so that every device sees the same radii (the reference's `torch.arange(..., device=cuda)`)."""
    a, b = 1. / radius_scale_min, 1. / radius_scale_max
    return torch.arange(a, b, (b - a) / max_steps, dtype=torch.float32)


class LoTDNeRF(nn.Module):
    """sigma, rgb = f(x in [-1,1]^D, v, h_appear): LoTD encoding (+ identity embedding of x) -> density MLP; radiance MLP on (SH(v), h, h_appear)
    (lotd_nerf.py:32-190 with n_extra_feat_from_output = 0, extra_pos_embed identity: the shipped Distant configuration)."""

    def __init__(self, encoding_cfg: dict = None, density_decoder_cfg: dict = None, radiance_decoder_cfg: dict = None, extra_pos_embed_cfg=dict(type="identity"),
                 aabb=None, bounding_size=2.0, dtype=torch.half, device=None, generator=None):
        super().__init__()
        self.dtype = dtype
        ec = dict(encoding_cfg or {})
        input_ch = ec.pop("input_ch", 3)
        self.space = AABBSpace(bounding_size, aabb=aabb, device=device)
        auto = ec.get("lotd_auto_compute_cfg")
        if ec.get("lotd_cfg") is None and auto is not None and auto.get("type") == "ngp4d":
            a = dict(auto); a.pop("type")
            stretch = (self.space.radius3d * 2).tolist() if ec.get("lotd_use_cuboid", False) else 1.0
            ec["lotd_cfg"] = auto_ngp4d_cfg(dim=input_ch, stretch=stretch, **a)
            ec.pop("lotd_auto_compute_cfg")
        ec.pop("lotd_use_cuboid", None); ec.pop("anneal_cfg", None); ec.pop("space_cfg", None)
        self.encoding = LoTDEncoding(input_ch, **ec, dtype=dtype, device=device, generator=generator)
        if extra_pos_embed_cfg is not None and extra_pos_embed_cfg.get("type", "identity") != "identity":
            raise RuntimeError("extra_pos_embed_cfg: only the identity embedding is built")
        self.n_extra_embed = input_ch if extra_pos_embed_cfg is not None else 0
        dc = dict(D=1, W=64, output_activation="softplus")
        dc.update(density_decoder_cfg or {})
        dc.pop("type", None)
        self.density_decoder = MLP(self.encoding.out_features + self.n_extra_embed, 1, **dc, dtype=dtype, device=device, generator=generator)
        rc = dict(use_pos=False, use_view_dirs=True, use_nablas=False, dir_embed_cfg=dict(type="spherical", degree=4), D=2, W=64)
        rc.update(radiance_decoder_cfg or {})
        self.rgb_decoder = RadianceNet(n_extra_feat=self.encoding.out_features, dtype=dtype, device=device, generator=generator, **rc)
        self.use_view_dirs, self.use_h_appear = self.rgb_decoder.use_view_dirs, self.rgb_decoder.use_h_appear

    @property
    def device(self):
        return self.encoding.flattened_params.device

    def _density(self, x):
        h = self.encoding(x)
        inp = torch.cat([h, x.to(h.dtype)], dim=-1) if self.n_extra_embed else h
        return self.density_decoder(inp)[..., 0], h

    def forward_density(self, x):
        return dict(sigma=self._density(x)[0])

    @torch.no_grad()
    def query_density(self, x):
        return self._density(x)[0]

    def forward(self, x, *, v=None, h_appear=None):
        sigma, h = self._density(x)
        rgb = self.rgb_decoder(x, v=v, n=None, h_extra=h, h_appear=h_appear)["rgb"]
        return dict(sigma=sigma, rgb=rgb)


class LoTDNeRFDistant(LoTDNeRF):
    """+ the shell sampler and `ray_query` of NeRFRendererMixinDistant (query_mode `march`, sample_mode `box`, interval `inverse_proportional`)."""

    def __init__(self, *args, ray_query_cfg: dict = None, radius_scale_min=1.0, radius_scale_max=100.0, include_inf_distance=True, **kw):
        super().__init__(*args, **kw)
        self.ray_query_cfg = dict(ray_query_cfg or dict(query_mode="march", query_param=dict(march_cfg=dict(sample_mode="box", max_steps=64))))
        self.radius_scale_min, self.radius_scale_max, self.include_inf_distance = radius_scale_min, radius_scale_max, include_inf_distance

    def _ray_marching(self, rays_o, rays_d, t_min, t_max=None, perturb=False, max_steps=256, sample_mode="box", interval_type="inverse_proportional"):
        """-> (ridx_hit, samples [M,4], depth_samples [M], deltas [M], ridx [M], pack_infos [Rh,2]) or Nones   (renderer_mixin.py:170-288)"""
        if sample_mode not in ("box", "fixed_cuboid_shells") or interval_type != "inverse_proportional":
            raise RuntimeError(f"sample_mode={sample_mode!r} / interval_type={interval_type!r} is not built (box / inverse_proportional as shipped)")
        n, dev, dtype = rays_o.shape[0], rays_o.device, rays_o.dtype
        r_reci = shell_radii(self.radius_scale_min, self.radius_scale_max, max_steps).to(dev).expand(n, -1)
        if perturb:
            step = (1. / self.radius_scale_max - 1. / self.radius_scale_min) / max_steps
            r_reci = (r_reci + torch.rand_like(r_reci) * step).clamp(1e-5)
        r = r_reci.reciprocal()
        r_ext = torch.cat([r, torch.full([n, 1], 1.0e10 if self.include_inf_distance else self.radius_scale_max, device=dev, dtype=dtype)], dim=-1)
        o_n, d_n = self.space.normalize_rays(rays_o, rays_d)
        t_ext = ray_box_intersect(o_n, d_n, r_ext)
        deltas, t = t_ext.diff(dim=-1), t_ext[:, :-1]
        x = torch.addcmul(o_n.unsqueeze(-2), d_n.unsqueeze(-2), t.unsqueeze(-1))
        x4 = torch.cat([x * r_reci.unsqueeze(-1), r_reci.unsqueeze(-1) * 2. - 1], dim=-1)
        valid = ~(torch.isnan(t) | (t < t_min[:, None]))
        ridx, pidx = valid.nonzero().long().t()
        if ridx.numel() == 0:
            return (None,) * 6
        pack_infos = get_pack_infos_from_n(valid.sum(-1))
        ridx_hit = pack_infos[..., 1].nonzero().long()[..., 0].contiguous()
        return ridx_hit, x4[ridx, pidx], t[ridx, pidx], deltas[ridx, pidx], ridx, pack_infos[ridx_hit].contiguous().long()

    def ray_query(self, ray_input=None, ray_tested=None, config=dict(), return_buffer=True, return_details=False, render_per_obj_individual=False):
        """-> {'volume_buffer': packed buffer with t, sigma, opacity_alpha (, rgb), 'details'}   (renderer_mixin.py:290-381, 383-440)"""
        cfg = dict(config)
        qp = dict(cfg.get("query_param", None) or self.ray_query_cfg.get("query_param", {}))
        with_rgb, perturb = cfg.get("with_rgb", True), cfg.get("perturb", False)
        empty = dict(volume_buffer=dict(type="empty", rays_inds_hit=[]), details={})
        if ray_tested["num_rays"] == 0:
            return empty
        rays_o, rays_d, near, rays_inds = itemgetter("rays_o", "rays_d", "near", "rays_inds")(ray_tested)
        dtype = rays_o.dtype
        ridx_hit, samples, depth, deltas, ridx, pack_infos = self._ray_marching(rays_o, rays_d, near, None, perturb=perturb, **qp.get("march_cfg", {}))
        if ridx_hit is None:
            return empty
        old = pack_infos.clone()
        if qp.get("compression", True):
            with torch.no_grad():
                alphas = tau_to_alpha(self.forward_density(samples)["sigma"].float() * deltas)
            nidx, pack_infos, pidx = packed_volume_render_compression(alphas, pack_infos)
            if nidx.numel() == 0:
                return empty
            ridx_hit, samples, depth, deltas, ridx = ridx_hit[nidx], samples[pidx], depth[pidx], deltas[pidx], ridx[pidx]
        vb = dict(type="packed", rays_inds_hit=rays_inds[ridx_hit], pack_infos_hit=pack_infos, t=depth.to(dtype))
        if with_rgb:
            view_dirs = rays_d / rays_d.detach().norm(dim=-1).clamp_min(1.0e-10).unsqueeze(-1)
            ha = ray_tested.get("rays_h_appear", None)
            out = self.forward(samples, v=view_dirs[ridx], h_appear=None if ha is None else ha[ridx])
            vb["rgb"] = out["rgb"].to(dtype)
        else:
            out = self.forward_density(samples)
        vb["sigma"] = out["sigma"].to(dtype)
        vb["opacity_alpha"] = tau_to_alpha(out["sigma"].float() * deltas).to(dtype)
        return dict(volume_buffer=vb, details={"march.num_per_ray": old[:, 1], "render.num_per_ray": pack_infos[:, 1]})
