"""Fused per-ray NeuS stages (csrc/neus_fused.cu): each function is ONE kernel launch and is numerically the same
computation as the chain of `nr3d_lib` calls named in its docstring, which stays available in `graphics.neus` /
`graphics.pack_ops` (the unfused chain is what the parity tests compare these against).

  upsample_cdf            neus_packed_sdf_to_(upsample_)alpha -> packed_alpha_to_vw -> packed_cumsum(exclusive) -> normalise
  sample_cdf_uniform      packed_sample_cdf(perturb=False)           (reference: graphics/raysample.py:38-61)
  neus_alpha_compress     neus_packed_sdf_to_alpha (autograd) + packed_volume_render_compression's selector pass
  composite               packed_alpha_to_vw + packed_sum/packed_div + products of the volume integration
                          (reference: app/renderers/single_volume_renderer.py:73-102), with its adjoint
"""
from __future__ import annotations

import ctypes

import torch

from .. import _lib as L

__all__ = ["upsample_cdf", "sample_cdf_uniform", "neus_alpha_compress", "composite"]

_U_CACHE = {}


def _f32c(t):
    return t.detach().contiguous().float()


@torch.no_grad()
def upsample_cdf(sdf, depth, pack_infos, inv_s: float, use_estimate_alpha=False, early_stop_eps=1e-4, alpha_thre=0.0):
    sdf, depth = _f32c(sdf), _f32c(depth)
    cdf = torch.empty_like(sdf)
    L.check(L.lib().nsb_neus_upsample_cdf(L.ptr(sdf, "f32"), L.ptr(depth, "f32"), L.ptr(pack_infos, "i64"), L.c_i64(pack_infos.shape[0]),
                                          L.c_f32(inv_s), ctypes.c_int(1 if use_estimate_alpha else 0), L.c_f32(early_stop_eps),
                                          L.c_f32(alpha_thre), L.ptr(cdf), L.stream_ptr()), "neus_upsample_cdf")
    return cdf


@torch.no_grad()
def sample_cdf_uniform(bins, cdfs, pack_infos, num_to_sample: int):
    key = (num_to_sample, bins.device)
    u = _U_CACHE.get(key)
    if u is None:
        u = _U_CACHE[key] = torch.linspace(0., 1., num_to_sample + 2, device=bins.device, dtype=torch.float32)[1:-1].contiguous()
    P = pack_infos.shape[0]
    out = torch.empty(P, num_to_sample, device=bins.device, dtype=torch.float32)
    L.check(L.lib().nsb_packed_invert_cdf_shared_u(L.ptr(bins, "f32"), L.ptr(cdfs, "f32"), L.ptr(u, "f32"), L.ptr(pack_infos, "i64"),
                                                   L.c_i64(P), L.c_i32(num_to_sample), L.ptr(out), L.stream_ptr()),
            "packed_invert_cdf_shared_u")
    return out


class _NeusAlpha(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, inv_s, pack_infos, early_stop_eps, alpha_thre):
        sdf_c, inv_c = _f32c(sdf), _f32c(inv_s).reshape(1)
        P = pack_infos.shape[0]
        alpha = torch.empty_like(sdf_c)
        sel = torch.empty(sdf_c.shape[0], dtype=torch.bool, device=sdf_c.device)
        steps = torch.empty(P, dtype=torch.int64, device=sdf_c.device)
        L.check(L.lib().nsb_neus_alpha_forward(L.ptr(sdf_c, "f32"), L.ptr(pack_infos, "i64"), L.c_i64(P), L.ptr(inv_c, "f32"),
                                               L.c_f32(early_stop_eps), L.c_f32(alpha_thre), L.ptr(alpha), L.ptr(sel), L.ptr(steps),
                                               L.stream_ptr()), "neus_alpha_forward")
        ctx.save_for_backward(sdf_c, inv_c, pack_infos)
        ctx.inv_shape = inv_s.shape
        ctx.mark_non_differentiable(sel, steps)
        return alpha, sel, steps

    @staticmethod
    def backward(ctx, g_alpha, _gs, _gn):
        sdf_c, inv_c, pack_infos = ctx.saved_tensors
        g = g_alpha.contiguous().float()
        d_sdf = torch.empty_like(sdf_c)
        d_inv = torch.zeros(1, device=sdf_c.device, dtype=torch.float32)
        L.check(L.lib().nsb_neus_alpha_backward(L.ptr(sdf_c, "f32"), L.ptr(pack_infos, "i64"), L.c_i64(pack_infos.shape[0]),
                                                L.ptr(inv_c, "f32"), L.ptr(g, "f32"), L.ptr(d_sdf), L.ptr(d_inv), L.stream_ptr()),
                "neus_alpha_backward")
        return (d_sdf if ctx.needs_input_grad[0] else None, d_inv.reshape(ctx.inv_shape) if ctx.needs_input_grad[1] else None,
                None, None, None)


def neus_alpha_compress(sdf, inv_s, pack_infos, early_stop_eps=1e-4, alpha_thre=0.0):
    """-> (alpha [S] (differentiable wrt sdf, inv_s), nidx_useful, pack_infos_useful, pidx_useful)."""
    if not isinstance(inv_s, torch.Tensor):
        inv_s = torch.tensor(float(inv_s), device=sdf.device)
    alpha, sel, steps = _NeusAlpha.apply(sdf, inv_s, pack_infos, early_stop_eps, alpha_thre)
    pidx = sel.nonzero()[..., 0]
    nidx = (steps > 0).nonzero()[..., 0]
    kept = steps[nidx]
    cs = kept.cumsum(0)
    return alpha, nidx, torch.stack([cs - kept, kept], 1), pidx


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, t, rgb, nablas, pack_infos, normalize_depth, early_stop_eps, alpha_thre):
        a, tt = _f32c(alpha), _f32c(t)
        r = None if rgb is None else _f32c(rgb)
        nb = None if nablas is None else _f32c(nablas)
        P, dev = pack_infos.shape[0], a.device
        vw = torch.empty_like(a)
        mask, depth = torch.empty(P, device=dev), torch.empty(P, device=dev)
        rgb_o = torch.empty(P, 3, device=dev) if r is not None else None
        nab_o = torch.empty(P, 3, device=dev) if nb is not None else None
        L.check(L.lib().nsb_composite_forward(L.ptr(a, "f32"), L.ptr(tt, "f32"), L.ptr(r, "f32", allow_none=True),
                                              L.ptr(nb, "f32", allow_none=True), L.ptr(pack_infos, "i64"), L.c_i64(P), L.c_f32(early_stop_eps),
                                              L.c_f32(alpha_thre), ctypes.c_int(1 if normalize_depth else 0), L.ptr(vw), L.ptr(mask),
                                              L.ptr(depth), L.ptr(rgb_o, allow_none=True), L.ptr(nab_o, allow_none=True), L.stream_ptr()),
                "composite_forward")
        ctx.save_for_backward(a, tt, r, nb, vw, pack_infos, mask, depth)
        ctx.cfg = (normalize_depth, early_stop_eps, alpha_thre)
        ctx.set_materialize_grads(False)
        empty = a.new_empty(0)
        return vw, mask, depth, (rgb_o if rgb_o is not None else empty), (nab_o if nab_o is not None else empty)

    @staticmethod
    def backward(ctx, g_vw, g_mask, g_depth, g_rgb, g_nab):
        a, tt, r, nb, vw, pack_infos, mask, depth = ctx.saved_tensors
        normalize_depth, eps, thre = ctx.cfg
        P = pack_infos.shape[0]

        def opt(g, present=True):
            return None if (g is None or not present) else g.contiguous().float()
        g_vw, g_mask, g_depth = opt(g_vw), opt(g_mask), opt(g_depth)
        g_rgb, g_nab = opt(g_rgb, r is not None), opt(g_nab, nb is not None)
        d_alpha = torch.empty_like(a)
        d_rgb = torch.empty_like(r) if r is not None else None
        d_nab = torch.empty_like(nb) if nb is not None else None
        P_ = L.ptr
        L.check(L.lib().nsb_composite_backward(
            P_(a, "f32"), P_(tt, "f32"), P_(r, allow_none=True), P_(nb, allow_none=True), P_(vw, "f32"), P_(pack_infos, "i64"), L.c_i64(P),
            L.c_f32(eps), L.c_f32(thre), ctypes.c_int(1 if normalize_depth else 0), P_(mask), P_(depth), P_(g_mask, allow_none=True),
            P_(g_depth, allow_none=True), P_(g_rgb, allow_none=True), P_(g_nab, allow_none=True), P_(g_vw, allow_none=True), P_(d_alpha),
            P_(d_rgb, allow_none=True), P_(d_nab, allow_none=True), L.stream_ptr()), "composite_backward")
        return d_alpha, None, d_rgb, d_nab, None, None, None, None


def composite(alpha, t, pack_infos, rgb=None, nablas=None, normalize_depth=True, early_stop_eps=1e-4, alpha_thre=0.0):
    """-> (vw [K], mask [P], depth [P], rgb [P,3] | None, normals [P,3] | None); differentiable wrt alpha, rgb, nablas."""
    vw, mask, depth, rgb_o, nab_o = _Composite.apply(alpha, t, rgb, nablas, pack_infos, normalize_depth, early_stop_eps, alpha_thre)
    return vw, mask, depth, (rgb_o if rgb is not None else None), (nab_o if nablas is not None else None)
