"""Joint rendering of several objects along the same rays -- the collect / sort / integrate core of `app.renderers.BufferComposeRenderer`
(reference: app/renderers/buffer_compose_renderer.py:110-869; collect :644-681, per-ray sort :683-694, integration :696-714) and the batched
ray test of shared (batched) models (nr3d_lib/models/spatial/batched.py:95-145).

Every object is queried in ITS OWN frame (rays transformed by the object's pose and scale, as the reference's scene graph does with
`obj.world_transform`), returns the packed `volume_buffer` of `ray_query`, and the buffers are merged per ray:

  reference                                              here
  ------------------------------------------------------ -------------------------------------------------------------------------
  get_pack_infos_from_n + interleave_linstep scatter     same pack arithmetic; the result is kept
  of t / alpha / rgb / nablas per buffer (~6 ATen each)   as ONE gather index into the concatenation of the buffers
  packed_sort (serial quicksort per ray, pack_ops_cuda    warp-cooperative stable rank sort per ray (csrc/pack_ops.cu:k_packed_sort_rank): sorted depths
  .cu:2671-2720) + 4 index ops                            + the permutation, composed with the gather index -> one index op per payload
  packed_alpha_to_vw + 5 packed_sum + packed_div          the fused compositing kernel (graphics/neus_fused.py:composite), written straight into the
  + scatter into the image                                whole-image buffers

Gradients reach every object's buffers through the (differentiable) index ops.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from .graphics import neus_fused
from .graphics.pack_ops import get_pack_infos_from_n, packed_sort_inplace
from .graphics.raytest import ray_box_intersection_fast_float_nocheck

__all__ = ["collect_and_sort", "compose_render", "BufferComposeRenderer", "BatchedBlockSpace", "ObjectPose"]


def collect_and_sort(buffers: Sequence[dict], total_num_rays: int):
    """buffers: packed volume buffers (`rays_inds_hit` [P_i] unique & sorted, `pack_infos_hit` [P_i,2], `t`, `opacity_alpha`, optional `rgb`,
    `nablas_in_world`) over the SAME `total_num_rays` rays.  -> None if no buffer holds a sample, else dict(rays_inds_hit, pack_infos_hit,
    src_index [S] (position of every sorted sample in the concatenation of the buffers), buffer_of [S], t [S])."""
    bufs = [b for b in buffers if b is not None and b.get("type", "packed") != "empty"]
    if not bufs:
        return None
    dev = bufs[0]["t"].device
    with torch.no_grad():
        visible = torch.zeros(total_num_rays, dtype=torch.int64, device=dev)
        for b in bufs:
            visible.index_add_(0, b["rays_inds_hit"], b["pack_infos_hit"][:, 1])
        sparse = get_pack_infos_from_n(visible)
        rays_hit = visible.nonzero().long()[..., 0]
        if rays_hit.numel() == 0:
            return None
        pack_infos = sparse[rays_hit].contiguous()
        S = int(sum(b["t"].numel() for b in bufs))
        cursor = sparse[:, 0].clone()
        slot = torch.empty(S, dtype=torch.int64, device=dev)          # slot[j] = where sample j of the concatenation lands before sorting
        owner = torch.empty(S, dtype=torch.int64, device=dev)
        base = 0
        for i, b in enumerate(bufs):
            n = b["t"].numel()
            ri, lens = b["rays_inds_hit"], b["pack_infos_hit"][:, 1].contiguous()
            # pack p of this buffer occupies `lens[p]` consecutive slots of its ray's pack, after the samples of the buffers before it
            # (the reference's interleave_linstep(current_pack_indices_buffer[rays_inds], lens, 1), buffer_compose_renderer.py:668-671)
            k_in_pack = torch.arange(n, device=dev) - torch.repeat_interleave(lens.cumsum(0) - lens, lens, output_size=n)
            tgt = torch.repeat_interleave(cursor[ri], lens, output_size=n) + k_in_pack
            src = torch.repeat_interleave(b["pack_infos_hit"][:, 0], lens, output_size=n) + k_in_pack
            slot[base + src] = tgt
            owner[base:base + n] = i
            cursor.index_add_(0, ri, lens)
            base += n
        t_cat = torch.cat([b["t"].detach().reshape(-1).float() for b in bufs])
        unsorted_src = torch.empty(S, dtype=torch.int64, device=dev)
        unsorted_src[slot] = torch.arange(S, device=dev)              # unsorted_src[k] = concatenation index of the sample in slot k
        t_total = t_cat[unsorted_src].contiguous()
        order = packed_sort_inplace(t_total, pack_infos, return_idx=True)            # in place: t_total is sorted per ray afterwards
        src_index = unsorted_src[order]
    return dict(rays_inds_hit=rays_hit, pack_infos_hit=pack_infos, src_index=src_index, buffer_of=owner[src_index], t=t_total, n_buffers=len(bufs), _bufs=bufs)


def compose_render(buffers: Sequence[dict], total_num_rays: int, *, with_rgb=True, with_normal=True, training=True, depth_use_normalized_vw=True):
    """-> (rendered dict of whole-image buffers, total volume buffer | None)  (buffer_compose_renderer.py:644-714)"""
    dev = next((b["t"].device for b in buffers if b is not None and b.get("type", "packed") != "empty"), None)
    col = collect_and_sort(buffers, total_num_rays)
    if col is None:
        z = lambda *s: torch.zeros(*s, device=dev)
        r = dict(mask_volume=z(total_num_rays), depth_volume=z(total_num_rays))
        if with_rgb:
            r["rgb_volume"] = z(total_num_rays, 3)
        if with_normal:
            r["normals_volume"] = z(total_num_rays, 3)
        return r, None
    bufs, idx = col.pop("_bufs"), col["src_index"]
    alpha = torch.cat([b["opacity_alpha"].reshape(-1) for b in bufs])[idx]
    rgb = nab = None
    if with_rgb:
        rgb = torch.cat([b["rgb"].reshape(-1, 3) if "rgb" in b else b["t"].new_zeros(b["t"].numel(), 3) for b in bufs])[idx]
    if with_normal:
        nab = torch.cat([b["nablas_in_world"].reshape(-1, 3) if "nablas_in_world" in b else b["t"].new_zeros(b["t"].numel(), 3) for b in bufs])[idx]
        if not training:
            nab = F.normalize(nab.clamp(-1, 1), dim=-1)
    vw, m, d, c, nn_ = neus_fused.composite(alpha, col["t"], col["pack_infos_hit"], rgb=rgb, nablas=nab, normalize_depth=depth_use_normalized_vw,
                                            ray_index=col["rays_inds_hit"], n_rays=total_num_rays)
    rendered = dict(mask_volume=m, depth_volume=d)
    if with_rgb:
        rendered["rgb_volume"] = c
    if with_normal:
        rendered["normals_volume"] = nn_
    total = dict(type="packed", rays_inds_hit=col["rays_inds_hit"], pack_infos_hit=col["pack_infos_hit"], t=col["t"], opacity_alpha=alpha, vw=vw,
                 buffer_of=col["buffer_of"], src_index=idx)
    if rgb is not None:
        total["rgb"] = rgb
    if nab is not None:
        total["nablas"] = nab
    return rendered, total


class ObjectPose:
    """world -> object frame of one drawable: x_obj = R^T (x_world - t) / s   (the scene graph's `world_transform` + `scale`, app/resources/nodes.py)"""

    def __init__(self, rotation=None, translation=None, scale=1.0, device=None):
        self.R = torch.eye(3, device=device) if rotation is None else torch.as_tensor(rotation, dtype=torch.float, device=device)
        self.t = torch.zeros(3, device=device) if translation is None else torch.as_tensor(translation, dtype=torch.float, device=device)
        self.s = float(scale)

    def rays_to_object(self, rays_o, rays_d):
        # depths along the ray are preserved: d_obj = R^T d / s, so x_obj(t) = o_obj + t d_obj  (buffer_compose_renderer.py:330-345)
        return (rays_o - self.t) @ self.R / self.s, rays_d @ self.R / self.s

    def normals_to_world(self, n):
        return n @ self.R.t()


class BufferComposeRenderer:
    """N single-object NeuS models in one scene (buffer_compose_renderer.py `query_single` + the joint integration).  `objects`: list of
    (model: LoTDNeuSModel, pose: ObjectPose).  Shared / batched (conditional, permutohedral) foreground models are not built."""

    def __init__(self, config: dict = None):
        cfg = dict(near=0.01, far=None, with_rgb=True, with_normal=True, perturb=False, depth_use_normalized_vw=True)
        cfg.update(config or {})
        self.config, self.training = cfg, True

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def render(self, objects, rays_o, rays_d, rays_h_appear=None, return_buffer=False) -> Dict:
        cfg, n = self.config, rays_o.shape[0]
        buffers = []
        for model, pose in objects:
            o, d = pose.rays_to_object(rays_o, rays_d)
            extra = {} if rays_h_appear is None else dict(rays_h_appear=rays_h_appear)
            rt = model.ray_test(o.contiguous(), d.contiguous(), near=cfg["near"], far=cfg["far"], **extra)
            qcfg = dict(model.ray_query_cfg)
            qcfg.update(with_rgb=cfg["with_rgb"], with_normal=cfg["with_normal"], perturb=cfg["perturb"])
            vb = model.ray_query(ray_tested=rt, config=qcfg, return_buffer=True)["volume_buffer"]
            if vb["type"] != "empty" and "nablas" in vb:
                vb["nablas_in_world"] = pose.normals_to_world(vb["nablas"])          # rotate_volume_buffer_nablas
            buffers.append(vb)
        rendered, total = compose_render(buffers, n, with_rgb=cfg["with_rgb"], with_normal=cfg["with_normal"], training=self.training,
                                         depth_use_normalized_vw=cfg["depth_use_normalized_vw"])
        ret = dict(rendered=rendered)
        if return_buffer:
            ret["volume_buffer"], ret["per_object"] = total, buffers
        return ret


class BatchedBlockSpace:
    """B axis-aligned blocks that share one normalisation (nr3d_lib/models/spatial/batched.py:20-145): the ray test of batched (shared) models."""

    def __init__(self, bounding_size: float = 2.0, device=None):
        h = bounding_size / 2.
        self.aabb = torch.tensor([[-h, -h, -h], [h, h, h]], dtype=torch.float, device=device)

    center = property(lambda self: (self.aabb[1] + self.aabb[0]) / 2.)
    radius3d = property(lambda self: (self.aabb[1] - self.aabb[0]) / 2.)

    def cur_batch__normalize_rays(self, rays_o, rays_d):
        return (rays_o - self.center) / self.radius3d, rays_d / self.radius3d

    def cur_batch__ray_test(self, rays_o, rays_d, near=None, far=None, return_rays=True, normalized=False, compact_batch=False, **extra_ray_data):
        """rays [B, N, 3] (every ray in every object's frame) -> the (ray, object) pairs that hit, ORDERED BY RAY (batched.py:95-145):
        dict(num_rays, rays_inds [M], rays_bidx [M], full_bidx_map, rays_full_bidx, near, far, rays_o, rays_d)"""
        assert rays_o.dim() == rays_d.dim() == 3
        if not normalized:
            rays_o, rays_d = self.cur_batch__normalize_rays(rays_o, rays_d)
        with torch.no_grad():
            B, dev = rays_o.shape[0], rays_o.device
            near_, far_ = ray_box_intersection_fast_float_nocheck(rays_o, rays_d, -1., 1.)
            if near is not None:
                near_.clamp_min_(near)
            if far is not None:
                far_.clamp_max_(far)
            mask = (far_ > near_) & (far_ > (0 if near is None else near))
            if far is not None:
                mask &= near_ < far
        if not compact_batch:
            ridx, bidx = mask.t().nonzero(as_tuple=True)
            full_map, full_bidx = torch.arange(B, device=dev, dtype=torch.long), bidx
        else:
            full_map = mask.any(dim=-1).nonzero().long()[..., 0]
            ridx, bidx = mask[full_map].t().nonzero(as_tuple=True)
            full_bidx = full_map[bidx]
        inds = (full_bidx, ridx)
        ret = dict(num_rays=ridx.numel(), rays_inds=ridx, rays_bidx=bidx, full_bidx_map=full_map, rays_full_bidx=full_bidx, near=near_[inds], far=far_[inds])
        ret.update({k: v[inds] if isinstance(v, torch.Tensor) else v for k, v in extra_ray_data.items()})
        if return_rays:
            ret.update(rays_o=rays_o[inds], rays_d=rays_d[inds])
        return ret
