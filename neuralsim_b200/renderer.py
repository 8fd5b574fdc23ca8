"""Single-object volume renderer -- the call path of `app.renderers.SingleVolumeRenderer` for one close-range NeuS
model (reference: app/renderers/single_volume_renderer.py:73-102,136-460,495-620): ray_test -> model.ray_query ->
volume integration -> rendered dict, chunked over rays for full images.

The reference resolves models / cameras through its scene graph (app/resources, out of scope here); this renderer
takes the model and the rays directly.  `render()` is what bench.py times.
"""
from __future__ import annotations

from typing import Dict

import torch

from .fields.neus import LoTDNeuSModel, volume_integration


class _LazyZeros(dict):
    """The `rendered` dict of one chunk: zero images that are only allocated if somebody reads them (the fused integration replaces
    them wholesale, so a chunk that hits the object never pays for the four zero-fills)."""

    def __init__(self, n, device, with_rgb, with_normal):
        super().__init__()
        self._n, self._device = n, device
        self._shapes = dict(depth_volume=(n,), mask_volume=(n,))
        if with_rgb:
            self._shapes["rgb_volume"] = (n, 3)
        if with_normal:
            self._shapes["normals_volume"] = (n, 3)

    def __missing__(self, k):
        if k not in self._shapes:
            raise KeyError(k)
        v = self[k] = torch.zeros(*self._shapes[k], device=self._device)
        return v

    def __contains__(self, k):
        return k in self._shapes or dict.__contains__(self, k)

    def materialise(self):
        for k in self._shapes:
            self[k]
        return dict(self)


class SingleVolumeRenderer:
    def __init__(self, config: dict = None):
        cfg = dict(near=0.01, far=None, with_rgb=True, with_normal=True, perturb=False, rayschunk=0, depth_use_normalized_vw=True)
        cfg.update(config or {})
        self.config = cfg
        self.training = True

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def _volume_integration(self, volume_buffer, rendered, fresh=False):
        return volume_integration(volume_buffer, rendered, training=self.training,
                                  depth_use_normalized_vw=self.config["depth_use_normalized_vw"], nablas_key="nablas_in_world", fresh=fresh)

    def ray_query(self, model: LoTDNeuSModel, rays_o, rays_d, rays_h_appear=None, near=None, far=None, return_buffer=True,
                  return_details=False, distant_model=None) -> Dict:
        """One chunk of rays: -> dict(rendered={rgb_volume, depth_volume, mask_volume, normals_volume}, volume_buffer, details).
        distant_model (fields/distant.py:LoTDNeRFDistant): the NeRF++ background of every shipped config -- queried from the close-range far
        bound on, its buffer merged per ray with the close-range one before the integration (single_volume_renderer.py:273-375)."""
        cfg = self.config
        near = cfg["near"] if near is None else near
        far = cfg["far"] if far is None else far
        n, device = rays_o.shape[0], rays_o.device
        extra = {} if rays_h_appear is None else dict(rays_h_appear=rays_h_appear)
        ray_tested = model.ray_test(rays_o, rays_d, near=near, far=far, **extra)
        rendered = _LazyZeros(n, device, cfg["with_rgb"], cfg["with_normal"])     # buffers nobody overwrites are materialised on first use
        ret = dict(rendered=rendered, ray_tested=ray_tested)
        qcfg = dict(model.ray_query_cfg)
        qcfg.update(with_rgb=cfg["with_rgb"], with_normal=cfg["with_normal"], perturb=cfg["perturb"])
        raw = model.ray_query(ray_tested=ray_tested, config=qcfg, return_buffer=True, return_details=return_details)
        vb = raw["volume_buffer"]
        if vb["type"] != "empty" and "nablas" in vb:
            # obj -> world rotation is the identity for a single static object (single_volume_renderer.py:262-276)
            vb["nablas_in_world"] = vb["nablas"]
        if distant_model is not None:
            from .compose import compose_render
            near_dv = rays_o.new_full([n], 0. if near is None else float(near))
            if ray_tested["num_rays"] > 0:
                near_dv[ray_tested["rays_inds"]] = ray_tested["far"]          # behind the close-range box (single_volume_renderer.py:286-290)
            dv_rt = dict(rays_o=rays_o.detach(), rays_d=rays_d.detach(), near=near_dv, far=None, num_rays=n, rays_inds=torch.arange(n, device=device),
                         rays_h_appear=rays_h_appear)
            dv = distant_model.ray_query(ray_tested=dv_rt, config=dict(with_rgb=cfg["with_rgb"], perturb=cfg["perturb"]), return_buffer=True)["volume_buffer"]
            # both buffers are depth-sorted per ray: their merge (merge_two_packs_sorted) == the per-ray sort of their union
            merged, total = compose_render([vb, dv], n, with_rgb=cfg["with_rgb"], with_normal=cfg["with_normal"], training=self.training,
                                           depth_use_normalized_vw=cfg["depth_use_normalized_vw"])
            ret["rendered"] = merged
            if return_buffer:
                ret["volume_buffer"], ret["cr_volume_buffer"], ret["dv_volume_buffer"] = (total if total is not None else vb), vb, dv
            if return_details:
                ret["details"] = raw.get("details", {})
            return ret
        if vb["type"] != "empty":
            self._volume_integration(vb, rendered, fresh=True)
        ret["rendered"] = rendered.materialise()
        if return_buffer:
            ret["volume_buffer"] = vb
        if return_details:
            ret["details"] = raw.get("details", {})
        return ret

    def render(self, model: LoTDNeuSModel, rays_o, rays_d, rays_h_appear=None, near=None, far=None, rayschunk=None,
               return_buffer=False, return_details=False, distant_model=None) -> Dict:
        """Whole batch / image; with `rayschunk` > 0 the rays are processed in chunks (batchify_query, models/utils.py:441)."""
        chunk = self.config["rayschunk"] if rayschunk is None else rayschunk
        n = rays_o.shape[0]
        if not chunk or chunk >= n:
            return self.ray_query(model, rays_o, rays_d, rays_h_appear, near, far, return_buffer, return_details, distant_model)
        outs = []
        for s in range(0, n, chunk):
            e = min(s + chunk, n)
            ha = None if rays_h_appear is None else rays_h_appear[s:e]
            outs.append(self.ray_query(model, rays_o[s:e], rays_d[s:e], ha, near, far, False, False, distant_model)["rendered"])
        return dict(rendered={k: torch.cat([o[k] for o in outs], 0) for k in outs[0]})
