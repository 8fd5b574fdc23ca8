"""The NeuS step with every data-dependent size kept ON THE DEVICE: no host read, no `.item()`, no `nonzero()` -- so a whole
fwd+bwd step can be captured in ONE CUDA graph and replayed with a single launch (`StaticFrame`).

Same kernels, same values as `graphics.neus._query_fused` + `fields.neus.volume_integration` (reference:
nr3d_lib/graphics/neus/neus_ray_query.py:732-1104, app/renderers/single_volume_renderer.py:73-102,136-460): the only difference
is where the sizes live.  The reference reads ~25 sizes back per `ray_query`; `_query_fused` reads three (+ one
in the backward); here the three scans leave their totals in a device block `cnt` (layout: include/neuralsim_b200.h, nsb_query_counts),
every buffer is allocated at a fixed CAPACITY and every kernel processes `min(capacity, *count)` items (nsb_bind_device_counts).
Capacities: rays -> R (the chunk), boundary samples -> R (n_coarse + 1 + sum n_fine), marched / merged samples -> `march_cap`,
samples kept by the compression -> `kept_cap`.  If a frame needs more than a capacity the step renders nothing and raises bit 0 / 1 of
cnt[20]; `StaticFrame.check()` reads that flag (one D2H, whenever the caller wants it) and `StaticFrame` re-captures with larger arenas.

tests/test_static_gpu.py: images bit-equal to the host-sized path, gradients equal up to the order of the fp32 atomics.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.nn.functional as F

from .. import _lib as L
from .raysample import batch_sample_step_linear
from .pack_ops import get_pack_infos_from_batch
from . import neus_fused as NF

__all__ = ["render_static", "StaticFrame", "CNT_SLOTS"]

CNT_SLOTS = dict(n_rays=0, pairs=2, marched_raw=3, hit_raw=4, kept_raw=6, kept_rays_raw=7, nonzero=9, marched=12, hit=13, fine0=14,
                 boundary=18, kept=19, overflow=20, kept_rays=21, merged0=22, rays_if_kept_fits=26)
_NULL = ctypes.c_void_p(0)
# "auto": small batches march once and copy (nsb_ray_marching_record + nsb_march_compact) when the per-ray record fits this many bytes;
# "0": always the two-round march; "1": always the recorded march.  Same samples bit for bit either way.
MARCH_ONEPASS = os.environ.get("NSB_MARCH_ONEPASS", "auto")
MARCH_ONEPASS_MAX_BYTES = 64 << 20


def march_onepass(n_rays, max_steps):
    if MARCH_ONEPASS == "0":
        return False
    return MARCH_ONEPASS == "1" or 4 * int(n_rays) * int(max_steps) <= MARCH_ONEPASS_MAX_BYTES


def _slot(cnt, k):
    return ctypes.c_void_p(cnt.data_ptr() + 8 * k)


def _call(fn, what, cnt, k0, k1, *args):
    """one count-aware launch: bind cnt[k0] (and cnt[k1]) to this thread, launch, clear"""
    lib = L.lib()
    lib.nsb_bind_device_counts(_slot(cnt, k0), _slot(cnt, k1) if k1 is not None else _NULL)
    try:
        rc = fn(*args)
    finally:
        lib.nsb_bind_device_counts(_NULL, _NULL)
    L.check(rc, what)


def _scan(counts, cnt, slot, *, first=None, info2=None, index=None, pack=None, src=None, nz_src=None, ws=None):
    """nsb_scan_counts with the totals left in cnt[slot], cnt[slot + 1] (no host hand-off)"""
    if ws is None:
        ws = torch.zeros(NF._scan_ws_bytes(), dtype=torch.uint8, device=counts.device)
    P = L.ptr
    L.check(L.lib().nsb_scan_counts(P(counts, "i32"), L.c_i64(counts.shape[0]), P(first, allow_none=True), P(info2, allow_none=True),
                                    P(index, allow_none=True), P(pack, allow_none=True), P(src, "i64", allow_none=True), P(nz_src, allow_none=True),
                                    _slot(cnt, slot), None, L.c_i64(0), P(ws), L.stream_ptr()), "scan_counts")


def _query_counts(cnt, phase, n_coarse1, num_fine, march_cap, kept_cap):
    nf = (ctypes.c_int32 * max(len(num_fine), 1))(*[int(n) for n in num_fine])
    L.check(L.lib().nsb_query_counts(ctypes.c_void_p(cnt.data_ptr()), L.c_i32(phase), L.c_i32(n_coarse1), nf, L.c_i32(len(num_fine)),
                                     L.c_i64(march_cap), L.c_i64(kept_cap), L.stream_ptr()), "query_counts")


def _sdf_launch(meta, grid16, dec, rays_o, rays_d, t, sdf, *, ridx=None, packs=None, ml, collect, cnt, slot, timer="lotd_gather"):
    """the fused SDF query on rays with a device-resident count: mode 1 (ridx[n], count = samples) or mode 2 (packs, count = packs)"""
    P = L.ptr
    mode = 2 if packs is not None else 1
    with L.KERNEL_TIMER.time(timer, t.numel()):
        _call(L.lib().nsb_fused_sdf_collect, "fused_sdf", cnt, slot, None,
              meta.c_ref, P(grid16, "f16"), ctypes.byref(dec), None, P(rays_o, "f32"), P(rays_d, "f32"),
              P(ridx, "i64") if mode == 1 else None, P(t, "f32"), L.c_i64(t.numel()),
              P(packs[0], "i64") if mode == 2 else None, P(packs[1], "i64", allow_none=True) if mode == 2 else None,
              L.c_i64(packs[0].shape[0] if mode == 2 else 0), L.c_i32(mode), L.c_i32(ml), P(sdf),
              ctypes.byref(collect) if collect is not None else None, L.stream_ptr())
    return sdf


# ---------------------------------------------------------------------------------------------------------------- autograd pieces
class _StaticSDF(torch.autograd.Function):
    """boundary SDF query with grad (fields/networks.py:_FusedSDF with device-resident sizes): the backward compacts the samples with a
    non-zero cotangent by flag -> scan -> index list, all on the device, and runs k_sdf_bwd_tc over that list."""

    @staticmethod
    def forward(ctx, st, ridx, t, packs, count_slot, grid, W1, b1, W2, b2):
        sdf = torch.empty(t.numel(), dtype=torch.float32, device=t.device)
        _sdf_launch(st.meta, st.grid16, st.dec, st.rays_o, st.rays_d, t, sdf, ridx=ridx if packs is None else None, packs=packs, ml=st.ml,
                    collect=st.collect, cnt=st.cnt, slot=count_slot, timer="fused_sdf_fwd")
        ctx.st, ctx.ridx, ctx.t = st, ridx, t
        ctx.shapes = (grid.shape, W1.shape, b1.shape, W2.shape, b2.shape)
        return sdf

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_sdf):
        st, dev = ctx.st, d_sdf.device
        gs, w1s, b1s, w2s, b2s = ctx.shapes
        d_grid = torch.zeros(gs, dtype=torch.float32, device=dev)
        ks = [int(torch.Size(x).numel()) for x in (w1s, b1s, w2s, b2s)]
        small = torch.zeros(sum(ks), dtype=torch.float32, device=dev)
        d_W1, d_b1 = small[:ks[0]].view(w1s), small[ks[0]:ks[0] + ks[1]].view(b1s)
        d_W2, d_b2 = small[ks[0] + ks[1]:ks[0] + ks[1] + ks[2]].view(w2s), small[ks[0] + ks[1] + ks[2]:].view(b2s)
        d_sdf = d_sdf.contiguous().float()
        n = d_sdf.numel()
        P, cnt = L.ptr, st.cnt
        flag = torch.empty(n, dtype=torch.int32, device=dev)
        _call(L.lib().nsb_flag_nonzero, "flag_nonzero", cnt, CNT_SLOTS["boundary"], None, P(d_sdf, "f32"), L.c_i64(n), P(flag), L.stream_ptr())
        keep = torch.empty(n, dtype=torch.int64, device=dev)
        _scan(flag, cnt, CNT_SLOTS["nonzero"], index=keep, ws=st.ws[3])
        with L.KERNEL_TIMER.time("fused_sdf_bwd", n):
            _call(L.lib().nsb_fused_sdf_bwd_indexed, "fused_sdf_bwd", cnt, CNT_SLOTS["nonzero"], None,
                  st.meta.c_ref, P(st.grid16, "f16"), ctypes.byref(st.dec), None, P(st.rays_o, "f32"), P(st.rays_d, "f32"), P(ctx.ridx, "i64"),
                  P(ctx.t, "f32"), P(d_sdf, "f32"), P(keep, "i64"), L.c_i64(n), L.c_i32(st.ml), P(d_grid), P(d_W1), P(d_b1), P(d_W2), P(d_b2),
                  L.stream_ptr())
        return None, None, None, None, None, d_grid, d_W1, d_b1, d_W2, d_b2


class _StaticAlpha(torch.autograd.Function):
    """graphics/neus_fused.py:_NeusAlpha over the live packs cnt[0]"""

    @staticmethod
    def forward(ctx, sdf, inv_s, pack_infos, cnt, early_stop_eps, alpha_thre):
        sdf_c, inv_c = sdf.detach().contiguous().float(), inv_s.detach().contiguous().float().reshape(1)
        Pn, dev = pack_infos.shape[0], sdf_c.device
        alpha = torch.empty_like(sdf_c)
        sel = torch.empty(sdf_c.shape[0], dtype=torch.bool, device=dev)
        steps = torch.empty(Pn, dtype=torch.int32, device=dev)
        P = L.ptr
        _call(L.lib().nsb_neus_alpha_forward, "neus_alpha_forward", cnt, CNT_SLOTS["n_rays"], None, P(sdf_c, "f32"), P(pack_infos, "i64"), L.c_i64(Pn),
              P(inv_c, "f32"), L.c_f32(early_stop_eps), L.c_f32(alpha_thre), P(alpha), P(sel), P(steps), L.stream_ptr())
        ctx.save_for_backward(sdf_c, inv_c, pack_infos)
        ctx.cnt, ctx.inv_shape = cnt, inv_s.shape
        ctx.mark_non_differentiable(sel, steps)
        return alpha, sel, steps

    @staticmethod
    def backward(ctx, g_alpha, _gs, _gn):
        sdf_c, inv_c, pack_infos = ctx.saved_tensors
        g = g_alpha.contiguous().float()
        d_sdf = torch.empty_like(sdf_c)
        d_inv = torch.zeros(1, device=sdf_c.device, dtype=torch.float32)
        P = L.ptr
        _call(L.lib().nsb_neus_alpha_backward, "neus_alpha_backward", ctx.cnt, CNT_SLOTS["n_rays"], None, P(sdf_c, "f32"), P(pack_infos, "i64"),
              L.c_i64(pack_infos.shape[0]), P(inv_c, "f32"), P(g, "f32"), P(d_sdf), P(d_inv), L.stream_ptr())
        return (d_sdf if ctx.needs_input_grad[0] else None, d_inv.reshape(ctx.inv_shape) if ctx.needs_input_grad[1] else None, None, None, None, None)


class _StaticGather(torch.autograd.Function):
    """out = the kernel-made gather src[pidx[:K]]; backward scatters the K live rows into zeros"""

    @staticmethod
    def forward(ctx, src, pidx, gathered, cnt):
        ctx.save_for_backward(pidx)
        ctx.n, ctx.cnt = src.shape[0], cnt
        return gathered

    @staticmethod
    def backward(ctx, g):
        pidx, = ctx.saved_tensors
        g = g.contiguous().float()
        d = torch.zeros(ctx.n, dtype=torch.float32, device=g.device)
        P = L.ptr
        _call(L.lib().nsb_scatter_f32, "scatter_f32", ctx.cnt, CNT_SLOTS["kept"], None, P(g, "f32"), P(pidx, "i64"), L.c_i64(pidx.shape[0]), P(d), L.stream_ptr())
        return d, None, None, None


class _StaticColor(torch.autograd.Function):
    """fields/fused_color.py:_FusedColor over the K = cnt[19] kept samples (buffers sized kept_cap)"""

    @staticmethod
    def forward(ctx, st, ridx, t, view_dirs, h_appear, keep_acts, *params):
        n, dev = t.numel(), t.device
        sdf = torch.empty(n, dtype=torch.float32, device=dev)
        nab = torch.empty(n, 3, dtype=torch.float32, device=dev)
        rgb = torch.empty(n, 3, dtype=torch.float32, device=dev)
        x = torch.empty(n, 3, dtype=torch.float32, device=dev)
        acts = torch.empty(4, int(L.lib().nsb_color_tile_bytes(L.c_i64(n))), dtype=torch.uint8, device=dev) if keep_acts else None
        ap = [L.ptr(acts[k]) if keep_acts else None for k in range(4)]
        P = L.ptr
        with L.KERNEL_TIMER.time("fused_color_fwd", n):
            _call(L.lib().nsb_fused_color_fwd, "fused_color_fwd", st.cnt, CNT_SLOTS["kept"], None,
                  st.meta.c_ref, P(st.grid16, "f16"), ctypes.byref(st.net), None, P(st.rays_o, "f32"), P(st.rays_d, "f32"), P(ridx, "i64"), P(t, "f32"),
                  P(view_dirs, "f32"), P(h_appear, "f32", allow_none=True), L.c_i64(n), L.c_i32(st.ml), P(sdf), P(nab), P(rgb), P(x), *ap,
                  ctypes.byref(st.collect) if st.collect is not None else None, L.stream_ptr())
        ctx.st, ctx.ridx, ctx.t, ctx.n = st, ridx, t, n
        ctx.held = (acts, rgb)
        ctx.shapes = [p.shape for p in params]
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(x)
        return sdf, nab, rgb, x

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_sdf, g_nab, g_rgb, _gx):
        acts, rgb = ctx.held
        if acts is None:
            raise RuntimeError("static colour query: backward through a forward that ran without grad")
        st, dev, n = ctx.st, rgb.device, ctx.n
        sizes = [int(torch.Size(s).numel()) for s in ctx.shapes[1:]]
        small = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        grads, o = [torch.zeros(ctx.shapes[0], dtype=torch.float32, device=dev)], 0
        for sh, k in zip(ctx.shapes[1:], sizes):
            grads.append(small[o:o + k].view(sh))
            o += k
        if g_sdf is None and g_nab is None and g_rgb is None:
            return (None,) * 6 + tuple(grads)
        c = lambda g: None if g is None else g.contiguous().float()
        g_sdf, g_nab, g_rgb = c(g_sdf), c(g_nab), c(g_rgb)
        dh = torch.empty(n, 32, dtype=torch.float32, device=dev)
        P = L.ptr
        with L.KERNEL_TIMER.time("fused_color_bwd", n):
            _call(L.lib().nsb_fused_color_bwd, "fused_color_bwd", st.cnt, CNT_SLOTS["kept"], None,
                  st.meta.c_ref, P(st.grid16, "f16"), ctypes.byref(st.net), None, P(st.rays_o, "f32"), P(st.rays_d, "f32"), P(ctx.ridx, "i64"),
                  P(ctx.t, "f32"), L.c_i64(n), L.c_i32(st.ml), P(acts[0]), P(acts[1]), P(acts[2]), P(acts[3]), P(rgb), P(g_sdf, allow_none=True),
                  P(g_nab, allow_none=True), P(g_rgb, allow_none=True), P(dh), *[P(g) for g in grads], L.stream_ptr())
        return (None,) * 6 + tuple(grads)


class _StaticComposite(torch.autograd.Function):
    """graphics/neus_fused.py:_Composite over the cnt[21] rays that keep samples, written straight into whole-image buffers"""

    @staticmethod
    def forward(ctx, alpha, t, rgb, nablas, pack_infos, ray_index, n_rays, cnt, normalize_depth, early_stop_eps, alpha_thre):
        a, tt = alpha.detach().contiguous().float(), t.detach().contiguous().float()
        r = None if rgb is None else rgb.detach().contiguous().float()
        nb = None if nablas is None else nablas.detach().contiguous().float()
        Pn, dev = pack_infos.shape[0], a.device
        vw = torch.empty_like(a)
        cols = 2 + (3 if r is not None else 0) + (3 if nb is not None else 0)
        buf = torch.zeros(cols * n_rays, device=dev)
        mask, depth = buf[:n_rays], buf[n_rays:2 * n_rays]
        rgb_o = buf[2 * n_rays:5 * n_rays].view(n_rays, 3) if r is not None else None
        o3 = 5 * n_rays if r is not None else 2 * n_rays
        nab_o = buf[o3:o3 + 3 * n_rays].view(n_rays, 3) if nb is not None else None
        P = L.ptr
        _call(L.lib().nsb_composite_forward, "composite_forward", cnt, CNT_SLOTS["kept_rays"], None, P(a, "f32"), P(tt, "f32"), P(r, "f32", allow_none=True),
              P(nb, "f32", allow_none=True), P(pack_infos, "i64"), L.c_i64(Pn), L.c_f32(early_stop_eps), L.c_f32(alpha_thre),
              ctypes.c_int(1 if normalize_depth else 0), P(ray_index, "i64"), P(vw), P(mask), P(depth), P(rgb_o, allow_none=True),
              P(nab_o, allow_none=True), L.stream_ptr())
        ctx.save_for_backward(a, tt, r, nb, vw, pack_infos, mask, depth, ray_index)
        ctx.cfg = (normalize_depth, early_stop_eps, alpha_thre, cnt)
        ctx.set_materialize_grads(False)
        empty = a.new_empty(0)
        return vw, mask, depth, (rgb_o if rgb_o is not None else empty), (nab_o if nab_o is not None else empty)

    @staticmethod
    def backward(ctx, g_vw, g_mask, g_depth, g_rgb, g_nab):
        a, tt, r, nb, vw, pack_infos, mask, depth, ray_index = ctx.saved_tensors
        normalize_depth, eps, thre, cnt = ctx.cfg

        def opt(g, present=True):
            return None if (g is None or not present) else g.contiguous().float()
        g_vw, g_mask, g_depth = opt(g_vw), opt(g_mask), opt(g_depth)
        g_rgb, g_nab = opt(g_rgb, r is not None), opt(g_nab, nb is not None)
        d_alpha = torch.empty_like(a)
        d_rgb = torch.empty_like(r) if r is not None else None
        d_nab = torch.empty_like(nb) if nb is not None else None
        P = L.ptr
        _call(L.lib().nsb_composite_backward, "composite_backward", cnt, CNT_SLOTS["kept_rays"], None,
              P(a, "f32"), P(tt, "f32"), P(r, allow_none=True), P(nb, allow_none=True), P(vw, "f32"), P(pack_infos, "i64"), L.c_i64(pack_infos.shape[0]),
              L.c_f32(eps), L.c_f32(thre), ctypes.c_int(1 if normalize_depth else 0), P(mask), P(depth), P(g_mask, allow_none=True),
              P(g_depth, allow_none=True), P(g_rgb, allow_none=True), P(g_nab, allow_none=True), P(g_vw, allow_none=True), P(ray_index, "i64"),
              P(d_alpha), P(d_rgb, allow_none=True), P(d_nab, allow_none=True), L.stream_ptr())
        return (d_alpha, None, d_rgb, d_nab) + (None,) * 7


class _State:
    """what the kernels of one static step share"""
    __slots__ = ("meta", "grid16", "dec", "net", "held", "rays_o", "rays_d", "ml", "collect", "cnt", "ws")


def _fp16_images(model):
    """fp16 images of the masters, re-cast INSIDE the step (a captured graph must not rely on a host-side version check)"""
    s, b = model.implicit_surface, model.radiance_net.blocks.layers
    d = s.decoder.layers
    ps = [s.encoding.flattened_params, d[0].weight, d[0].bias, d[1].weight, d[1].bias, b[0].weight, b[0].bias, b[1].weight, b[1].bias, b[2].weight, b[2].bias]
    # one multi-tensor cast for the ten small tensors (a launch each otherwise: at 4096 rays per step the step is launch-bound), one for the table
    t = [torch.empty(p.shape, dtype=torch.half, device=p.device) for p in ps]
    with torch.no_grad():
        t[0].copy_(ps[0].detach())
        torch._foreach_copy_(t[1:], [p.detach() for p in ps[1:]])
    dec = L.SdfDecoderC(t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), d[0].out_features, float(d[0].activation.beta))
    r3 = s.radius3d_original
    fk = (r3.data_ptr(), r3._version, float(s.sdf_scale))
    if getattr(model, "_fac_cache", (None,))[0] != fk:
        model._fac_cache = (fk, (s.sdf_scale / r3).float().tolist())
    net = L.ColorNetC(*[x.data_ptr() for x in t[1:]], d[0].out_features, b[0].out_features, b[0].in_features, b[0].in_features - 54,
                      float(d[0].activation.beta), (ctypes.c_float * 3)(*model._fac_cache[1]))
    return t, dec, net, ps


def static_supported(model, cfg):
    qp = dict(model.ray_query_cfg.get("query_param", {}) or {})
    occ = getattr(getattr(model, "accel", None), "occ", None)
    return (getattr(model, "_color_fusable", lambda: False)() and model.use_view_dirs and occ is not None and occ.occ_grid.dim() == 3
            and occ.occ_grid.numel() * 4 // 32 <= 96 * 1024 and qp.get("num_coarse", 0) > 0 and len(qp.get("upsample_inv_s_factors", (1, 4, 16))) <= 4
            and qp.get("coarse_step_cfg", {}).get("step_mode", "linear") == "linear" and (cfg.get("with_rgb", True) or cfg.get("with_normal", True)))


def render_static(model, rays_o, rays_d, rays_h_appear=None, *, near=None, far=None, march_cap, kept_cap, coherent=False, with_rgb=True, with_normal=True,
                  perturb=False, training=None, depth_use_normalized_vw=True, cnt=None):
    """One chunk of rays, ray test -> query -> integration, without a host read.  -> (rendered dict of whole-chunk images, cnt int64[32]).
    `coherent`: image-ordered rays (the boundary / fine queries then walk the samples ray-tiled) -- a host decision here (the host-sized
    path measures it in the ray-test kernel)."""
    P, lib = L.ptr, L.lib()
    if perturb:
        raise RuntimeError("render_static: perturb=True is not built (random streams of capacity-sized draws differ from the reference's); "
                           "use the host-sized path (SingleVolumeRenderer.render)")
    training = model.training if training is None else training
    R, dev = rays_o.shape[0], rays_o.device
    qp = dict(model.ray_query_cfg.get("query_param", {}) or {})
    nc1 = int(qp["num_coarse"]) + 1
    factors = list(qp.get("upsample_inv_s_factors", (1, 4, 16)))
    n_stage = len(factors)
    num_fine = qp.get("num_fine", 8)
    num_fine = [num_fine] * n_stage if isinstance(num_fine, int) else list(num_fine)
    num_fine = [n // 2 * 2 + 1 for n in num_fine]
    nf_tot = int(sum(num_fine))
    upsample_inv_s = qp.get("upsample_inv_s", 64.) / model.upsample_s_divisor
    use_est = bool(qp.get("upsample_use_estimate_alpha", False))
    nablas_has_grad = bool(qp.get("nablas_has_grad", False))
    mc = dict(qp.get("march_cfg", {}))
    fac = mc.pop("step_size_factor", 1.0)
    step_size, dt_gamma = mc.get("step_size", 1e-3) * fac, mc.get("dt_gamma", 0.0) * fac
    max_steps, max_step_size = int(mc.get("max_steps", 512)), mc.get("max_step_size", 1e10)
    march_cap, kept_cap = int(march_cap), int(kept_cap)
    S_cap = R * (nc1 + nf_tot)
    if cnt is None:
        cnt = torch.zeros(32, dtype=torch.int64, device=dev)
    else:
        cnt.zero_()
    st = _State()
    st.cnt = cnt
    wsb = (NF._scan_ws_bytes() + 255) // 256 * 256
    st.ws = torch.zeros(4, wsb, dtype=torch.uint8, device=dev)          # the zeroed workspaces of the step's four scans, one fill
    with torch.no_grad():
        t16, st.dec, st.net, masters = _fp16_images(model)
        st.held, st.grid16 = t16, t16[0]
        st.meta = model.implicit_surface.encoding.meta
        st.ml = model.implicit_surface._ml(model.max_level)
        st.collect = model.accel.occ.collect_struct() if training else None
        # ---------------- ray test (fields/space.py:_ray_test_fused without the host read)
        sp = model.space
        if getattr(sp, "_host_cr", None) is None or sp._host_cr[0] != (sp.aabb.data_ptr(), sp.aabb._version):
            c, r = sp.center.tolist(), sp.radius3d.tolist()
            sp._host_cr = ((sp.aabb.data_ptr(), sp.aabb._version), (ctypes.c_float * 3)(*c), (ctypes.c_float * 3)(*r))
        o_n, d_n = torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev)
        nr, fr = torch.empty(R, device=dev), torch.empty(R, device=dev)
        flag = torch.empty(R, dtype=torch.int32, device=dev)
        L.check(lib.nsb_ray_test_aabb(P(rays_o.contiguous(), "f32"), P(rays_d.contiguous(), "f32"), L.c_i64(R), sp._host_cr[1], sp._host_cr[2],
                                      ctypes.c_int(0 if near is None else 1), L.c_f32(0. if near is None else near),
                                      ctypes.c_int(0 if far is None else 1), L.c_f32(0. if far is None else far), P(o_n), P(d_n), P(nr), P(fr), P(flag),
                                      _slot(cnt, CNT_SLOTS["pairs"]), L.stream_ptr()), "ray_test_aabb")
        rays_inds = torch.empty(R, dtype=torch.int64, device=dev)
        _scan(flag, cnt, CNT_SLOTS["n_rays"], index=rays_inds, ws=st.ws[0])
        ha = rays_h_appear.detach().contiguous().float() if (rays_h_appear is not None and model.use_h_appear) else None
        if ha is None and model.use_h_appear:             # LiDAR-style rays carry no appearance code: the (dropped) radiance head reads zeros
            ha = torch.zeros(R, model.radiance_net.blocks.layers[0].in_features - 54, device=dev)
        n_ha = ha.shape[1] if ha is not None else 0
        rbuf = torch.zeros(R * (8 + n_ha), device=dev)     # the compacted rays in ONE zero-filled allocation (rows beyond the live count stay 0)
        o_c, d_c = rbuf[:3 * R].view(R, 3), rbuf[3 * R:6 * R].view(R, 3)
        n_c, f_c = rbuf[6 * R:7 * R], rbuf[7 * R:8 * R]
        ha_c = rbuf[8 * R:].view(R, n_ha) if ha is not None else None
        _call(lib.nsb_gather_rays, "gather_rays", cnt, CNT_SLOTS["n_rays"], None, P(rays_inds, "i64"), L.c_i64(R), P(o_n), P(d_n), P(nr), P(fr), P(o_c), P(d_c),
              P(n_c), P(f_c), P(ha, allow_none=True), P(ha_c, allow_none=True), L.c_i32(0 if ha is None else ha.shape[1]), L.stream_ptr())
        st.rays_o, st.rays_d = o_c, d_c
        view_dirs = (d_c / d_c.norm(dim=-1).clamp_min(1.0e-10).unsqueeze(-1)).contiguous()
        # ---------------- coarse samples + march
        coarse = batch_sample_step_linear(n_c, f_c, nc1, prefix_shape=[R], perturb=perturb).contiguous()
        occ_grid = model.accel.occ.occ_grid
        res = occ_grid.shape[-3:]
        g8 = occ_grid.contiguous().view(torch.uint8)
        bits = torch.empty((occ_grid.numel() + 31) // 32, dtype=torch.int32, device=dev)
        L.check(lib.nsb_pack_occ_bits(P(g8, "u8"), L.c_i64(occ_grid.numel()), P(bits), L.stream_ptr()), "pack_occ_bits")
        roi = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=dev) if getattr(model, "_static_roi", None) is None else model._static_roi
        model._static_roi = roi
        margs = (L.c_i64(R), P(o_c, "f32"), P(d_c, "f32"), P(n_c, "f32"), P(f_c, "f32"), P(roi, "f32"), None, L.c_i32(res[0]), L.c_i32(res[1]), L.c_i32(res[2]),
                 P(g8, "u8"), L.c_f32(step_size), L.c_f32(max_step_size), L.c_f32(dt_gamma), ctypes.c_uint32(max_steps))
        num_steps = torch.empty(R, dtype=torch.int32, device=dev)
        # small batches: a march costs the latency of its longest ray -> march ONCE, recording the samples per ray, and copy (csrc/march.cu)
        rec_t = torch.empty(R * max_steps, dtype=torch.float32, device=dev) if march_onepass(R, max_steps) else None
        with L.KERNEL_TIMER.time("march", R):
            if rec_t is not None:
                _call(lib.nsb_ray_marching_record, "ray_marching_record", cnt, CNT_SLOTS["n_rays"], None, L.c_i64(R), P(o_c, "f32"), P(d_c, "f32"), P(n_c, "f32"),
                      P(f_c, "f32"), P(roi, "f32"), L.c_i32(res[0]), L.c_i32(res[1]), L.c_i32(res[2]), P(g8, "u8"), L.c_f32(step_size), L.c_f32(max_step_size),
                      L.c_f32(dt_gamma), ctypes.c_uint32(max_steps), P(num_steps), P(rec_t), P(bits), L.stream_ptr())
            else:
                _call(lib.nsb_ray_marching_listed, "ray_marching", cnt, CNT_SLOTS["n_rays"], None, *margs, None, P(num_steps), None, None, None, None, None, None,
                      L.c_i64(0), P(bits), L.stream_ptr())
        info2 = torch.empty(R, 2, dtype=torch.int32, device=dev)
        ridx_hit = torch.empty(R, dtype=torch.int64, device=dev)
        pack_infos = torch.empty(R, 2, dtype=torch.int64, device=dev)
        _scan(num_steps, cnt, CNT_SLOTS["marched_raw"], info2=info2, index=ridx_hit, pack=pack_infos, ws=st.ws[1])
        _query_counts(cnt, 0, nc1, num_fine, march_cap, kept_cap)
        depth = torch.empty(march_cap, dtype=torch.float32, device=dev)
        ridx32 = torch.empty(march_cap, dtype=torch.int32, device=dev)
        with L.KERNEL_TIMER.time("march", R):
            if rec_t is not None:
                _call(lib.nsb_march_compact, "march_compact", cnt, CNT_SLOTS["hit"], None, P(rec_t, "f32"), ctypes.c_uint32(max_steps), P(info2, "i32"),
                      P(ridx_hit, "i64"), L.c_i64(R), P(depth), P(ridx32), L.stream_ptr())
            else:
                _call(lib.nsb_ray_marching_listed, "ray_marching", cnt, CNT_SLOTS["hit"], None, *margs, P(info2), None, P(depth), None, P(ridx32), None, None,
                      P(ridx_hit, "i64"), L.c_i64(R), P(bits), L.stream_ptr())
        ridx = ridx32.long()
        # ---------------- up-sampling (no grad)
        from . import neus as GN
        fine_stages = None
        if GN.use_persistent_upsample(R):                    # ONE persistent per-ray kernel (csrc/ray_upsample.cu): small batches (graphics/neus.py)
            fine_all, _ovf = NF.upsample_rays(st.meta, st.grid16, st.dec, ridx_hit, pack_infos, depth, o_c, d_c, [upsample_inv_s * f for f in factors], num_fine,
                                              max_level=st.ml, max_steps=max_steps, use_estimate_alpha=use_est, collect=st.collect, count=(cnt, CNT_SLOTS["hit"]))
            factors_loop = []
        else:
            factors_loop = factors
        sdf = torch.empty(march_cap if factors_loop else 1, dtype=torch.float32, device=dev)
        if factors_loop:
            fine_stages = []
            _sdf_launch(st.meta, st.grid16, st.dec, o_c, d_c, depth, sdf, ridx=ridx, ml=st.ml, collect=st.collect, cnt=cnt, slot=CNT_SLOTS["marched"])
        for i, factor in enumerate(factors_loop):
            cdf = torch.empty(march_cap, dtype=torch.float32, device=dev)
            _call(lib.nsb_neus_upsample_cdf, "neus_upsample_cdf", cnt, CNT_SLOTS["hit"], None, P(sdf, "f32"), P(depth, "f32"), P(pack_infos, "i64"), L.c_i64(R),
                  L.c_f32(upsample_inv_s * factor), ctypes.c_int(1 if use_est else 0), L.c_f32(1e-4), L.c_f32(0.0), P(cdf), L.stream_ptr())
            nf = num_fine[i]
            fine = torch.empty(R, nf, dtype=torch.float32, device=dev)
            u = NF._U_CACHE.get((nf, dev))
            if u is None:
                u = NF._U_CACHE[(nf, dev)] = torch.linspace(0., 1., nf + 2, device=dev, dtype=torch.float32)[1:-1].contiguous()
            _call(lib.nsb_packed_invert_cdf_shared_u, "packed_invert_cdf_shared_u", cnt, CNT_SLOTS["hit"], None, P(depth, "f32"), P(cdf, "f32"), P(u, "f32"),
                  P(pack_infos, "i64"), L.c_i64(R), L.c_i32(nf), P(fine), L.stream_ptr())
            fine_stages.append(fine)
            if i < n_stage - 1:
                sdf_fine = torch.empty(R * nf, dtype=torch.float32, device=dev)
                if coherent:
                    _sdf_launch(st.meta, st.grid16, st.dec, o_c, d_c, fine.view(-1), sdf_fine, packs=(get_pack_infos_from_batch(R, nf, device=dev), ridx_hit),
                                ml=st.ml, collect=st.collect, cnt=cnt, slot=CNT_SLOTS["hit"])
                else:
                    _sdf_launch(st.meta, st.grid16, st.dec, o_c, d_c, fine.view(-1), sdf_fine, ridx=ridx_hit.unsqueeze(-1).expand(R, nf).reshape(-1).contiguous(),
                                ml=st.ml, collect=st.collect, cnt=cnt, slot=CNT_SLOTS["fine0"] + i)
                dep_m = torch.empty(march_cap, dtype=torch.float32, device=dev)
                sdf_m = torch.empty(march_cap, dtype=torch.float32, device=dev)
                pim = torch.empty_like(pack_infos)
                _call(lib.nsb_merge_sorted_vals, "merge_sorted_vals", cnt, CNT_SLOTS["hit"], None, P(depth, "f32"), P(sdf, "f32"), P(pack_infos, "i64"),
                      P(fine, "f32"), P(sdf_fine, "f32"), L.c_i64(R), L.c_i32(nf), P(dep_m), P(sdf_m), P(pim), L.stream_ptr())
                depth, sdf, pack_infos = dep_m, sdf_m, pim
        if fine_stages is not None:
            fine_all = (torch.cat(fine_stages, dim=-1) if n_stage > 1 else fine_stages[0]).contiguous()
        d1 = torch.empty(S_cap, dtype=torch.float32, device=dev)
        mid = torch.empty(S_cap, dtype=torch.float32, device=dev)
        ridx_all = torch.empty(S_cap, dtype=torch.int64, device=dev)
        pinfo = torch.empty(R, 2, dtype=torch.int64, device=dev)
        rl = (ctypes.c_int32 * n_stage)(*num_fine)
        _call(lib.nsb_assemble_boundary, "assemble_boundary", cnt, CNT_SLOTS["n_rays"], CNT_SLOTS["hit"], P(coarse, "f32"), L.c_i64(R), L.c_i32(nc1),
              P(ridx_hit, "i64"), L.c_i64(R), P(fine_all, "f32"), L.c_i32(nf_tot), rl, L.c_i32(n_stage), P(d1), P(mid), P(ridx_all), P(pinfo), L.stream_ptr())
    # ---------------- boundary SDF (grad) -> alpha -> compression
    s, b = model.implicit_surface, model.radiance_net.blocks.layers
    dl = s.decoder.layers
    sdf_b = _StaticSDF.apply(st, ridx_all, d1, (pinfo, None) if coherent else None, CNT_SLOTS["n_rays"] if coherent else CNT_SLOTS["boundary"],
                             s.encoding.flattened_params, dl[0].weight, dl[0].bias, dl[1].weight, dl[1].bias)
    inv_s = model.forward_inv_s()
    if not isinstance(inv_s, torch.Tensor):
        inv_s = torch.tensor(float(inv_s), device=dev)
    alpha, sel, steps = _StaticAlpha.apply(sdf_b, inv_s, pinfo, cnt, 1e-4, 0.0)
    with torch.no_grad():
        first = torch.empty(R, dtype=torch.int32, device=dev)
        nidx = torch.empty(R, dtype=torch.int64, device=dev)
        pinfo_kept = torch.empty(R, 2, dtype=torch.int64, device=dev)
        rays_inds_hit = torch.empty(R, dtype=torch.int64, device=dev)
        _scan(steps, cnt, CNT_SLOTS["kept_raw"], first=first, index=nidx, pack=pinfo_kept, src=rays_inds, nz_src=rays_inds_hit, ws=st.ws[2])
        _query_counts(cnt, 1, nc1, num_fine, march_cap, kept_cap)
        pidx, ridx_k = torch.empty(kept_cap, dtype=torch.int64, device=dev), torch.empty(kept_cap, dtype=torch.int64, device=dev)
        t_k, alpha_c = torch.empty(kept_cap, dtype=torch.float32, device=dev), torch.empty(kept_cap, dtype=torch.float32, device=dev)
        _call(lib.nsb_compact_samples, "compact_samples", cnt, CNT_SLOTS["rays_if_kept_fits"], None, P(sel.view(torch.uint8), "u8"), P(pinfo, "i64"), P(first, "i32"),
              P(steps, "i32"), L.c_i64(R), P(ridx_all, "i64"), P(mid, "f32"), P(alpha.detach(), "f32"), P(pidx), P(ridx_k), P(t_k), P(alpha_c), L.stream_ptr())
    alpha_k = _StaticGather.apply(alpha, pidx, alpha_c, cnt) if alpha.requires_grad else alpha_c
    # ---------------- colour / normal query on the kept samples
    params = (s.encoding.flattened_params, dl[0].weight, dl[0].bias, dl[1].weight, dl[1].bias, b[0].weight, b[0].bias, b[1].weight, b[1].bias,
              b[2].weight, b[2].bias)
    keep_acts = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    _sdf_k, nab, rgb, x = _StaticColor.apply(st, ridx_k, t_k, view_dirs, ha_c, keep_acts, *params)
    if not nablas_has_grad:
        nab = nab.detach()
    nab_i = nab if with_normal else None
    if nab_i is not None and not training:
        nab_i = F.normalize(nab_i.clamp(-1, 1), dim=-1)
    vw, m, d, c, nn_ = _StaticComposite.apply(alpha_k, t_k, rgb if with_rgb else None, nab_i, pinfo_kept, rays_inds_hit, R, cnt,
                                              bool(depth_use_normalized_vw), 1e-4, 0.0)
    rendered = dict(mask_volume=m, depth_volume=d)
    if with_rgb:
        rendered["rgb_volume"] = c
    if with_normal:
        rendered["normals_volume"] = nn_
    buffers = dict(opacity_alpha=alpha_k, t=t_k, rgb=rgb, nablas=nab, net_x=x, vw=vw, pack_infos_hit=pinfo_kept, rays_inds_hit=rays_inds_hit, ridx=ridx_k,
                   pack_infos_boundary=pinfo, march_pack_infos=pack_infos)
    return rendered, cnt, buffers


def sliced_volume_buffer(buffers, cnt):
    """The reference's packed `volume_buffer` (renderer_mixin.py:263-303) out of a static step's capacity-sized buffers: ONE host read of the
    counts, then views.  (Losses that read the buffer -- eikonal on `nablas` -- can use it; gradients flow through the views.)"""
    c = cnt.tolist()
    K, Pu = c[CNT_SLOTS["kept"]], c[CNT_SLOTS["kept_rays"]]
    if c[CNT_SLOTS["overflow"]]:
        raise RuntimeError(f"static step: arena overflow (flags {c[CNT_SLOTS['overflow']]}: 1 = marched samples, 2 = kept samples)")
    if K == 0:
        return dict(type="empty", rays_inds_hit=[])
    vb = dict(type="packed", rays_inds_hit=buffers["rays_inds_hit"][:Pu], pack_infos_hit=buffers["pack_infos_hit"][:Pu])
    for k in ("opacity_alpha", "t", "rgb", "nablas", "net_x", "vw"):
        vb[k] = buffers[k][:K]
    return vb


# ---------------------------------------------------------------------------------------------------------------- one-launch step
class StaticFrame:
    """fwd (+ loss + bwd) of one fixed-size ray batch as ONE CUDA graph launch.

        frame = StaticFrame(model, n_rays, loss_fn=lambda rendered: ..., near=0.01)
        loss = frame.step(rays_o, rays_d, rays_h_appear)        # device tensors (or pinned host tensors): copied into the graph's inputs
        frame.rendered["rgb_volume"], p.grad                    # static outputs / accumulated gradients
        frame.check()                                           # optional: one D2H of the counts; re-captures with larger arenas on overflow

    The first call probes the sizes with the host-sized path (SingleVolumeRenderer.ray_query, no grad), sizes the arenas with `slack`,
    warms up and captures.  Gradients are accumulated into `p.grad` (kept in place; `zero_grads=True` or a `pre_hook` zeroes them inside the graph).
    Capture precondition (PyTorch): no autograd graph of an EARLIER backward on the default stream may still be referenced (a kept loss / rendered
    tensor): it pins the parameters' AccumulateGrad nodes to the default stream, which cannot take part in a capture."""

    def __init__(self, model, n_rays, loss_fn=None, *, near=None, far=None, with_rgb=True, with_normal=True, slack=1.5, march_cap=None, kept_cap=None,
                 coherent=None, use_graph=True, zero_grads=False, h_appear_dim=None, pre_hook=None):
        self.model, self.n_rays, self.loss_fn = model, int(n_rays), loss_fn
        self.near, self.far, self.with_rgb, self.with_normal, self.slack = near, far, with_rgb, with_normal, float(slack)
        self.march_cap, self.kept_cap, self.coherent = march_cap, kept_cap, coherent
        self.use_graph, self.zero_grads, self.pre_hook = use_graph, zero_grads, pre_hook
        dev = model.device
        self.device = dev
        self.rays_o = torch.zeros(self.n_rays, 3, device=dev)
        self.rays_d = torch.zeros(self.n_rays, 3, device=dev)
        na = h_appear_dim if h_appear_dim is not None else (model.radiance_net.blocks.layers[0].in_features - 54 if model.use_h_appear else 0)
        self.h_appear = torch.zeros(self.n_rays, na, device=dev) if na > 0 else None
        self.cnt = torch.zeros(32, dtype=torch.int64, device=dev)
        self.graph, self.loss, self.rendered, self.buffers, self._occ_captured = None, None, None, None, None
        self.captures = 0

    # -- sizes
    @torch.no_grad()
    def _probe(self):
        """sizes of this batch from the host-sized path (two host reads): M (merged marched samples), K (kept), coherence"""
        from ..renderer import SingleVolumeRenderer
        r = SingleVolumeRenderer(dict(near=self.near, far=self.far, with_rgb=self.with_rgb, with_normal=self.with_normal))
        r.train(self.model.training)
        out = r.ray_query(self.model, self.rays_o, self.rays_d, self.h_appear, return_buffer=True, return_details=True)
        det, vb = out.get("details", {}), out["volume_buffer"]
        qp = dict(self.model.ray_query_cfg.get("query_param", {}) or {})
        nf = qp.get("num_fine", 8)
        nf = [nf] * len(qp.get("upsample_inv_s_factors", (1, 4, 16))) if isinstance(nf, int) else list(nf)
        nf = [n // 2 * 2 + 1 for n in nf]
        M = int(det["march.num_per_ray"].sum()) if "march.num_per_ray" in det else 0
        n_hit = int(det["march.num_per_ray"].shape[0]) if "march.num_per_ray" in det else 0
        K = int(vb["t"].shape[0]) if vb.get("type") == "packed" else 0
        return M + n_hit * sum(nf[:-1]), K, bool(out["ray_tested"].get("rays_coherent", False))

    def _size(self, grow=1.0):
        need_m, need_k, coh = self._probe()
        if self.coherent is None:
            self.coherent = coh
        floor = 4096 + 8 * min(self.n_rays, 65536)
        m = int(max(need_m * self.slack, floor) * grow)
        k = int(max(need_k * self.slack, floor) * grow)
        self.march_cap = max(self.march_cap or 0, m)
        self.kept_cap = max(self.kept_cap or 0, k)

    # -- the step
    def _run(self):
        if self.pre_hook is not None:
            self.pre_hook()
        if self.zero_grads:
            for p in self.model.parameters():
                if p.grad is not None:
                    p.grad.zero_()
        cv = getattr(self.model, "ctrl_var", None)
        if cv is not None and hasattr(cv, "mix_weight"):
            if getattr(cv, "_w_dev", None) is None:
                cv._w_dev = torch.zeros((), device=self.device)
            cv._use_w_dev = True                             # inv_s annealing weight from a device scalar (refreshed in step())
        try:
            rendered, _, buffers = render_static(self.model, self.rays_o, self.rays_d, self.h_appear, near=self.near, far=self.far, march_cap=self.march_cap,
                                                 kept_cap=self.kept_cap, coherent=bool(self.coherent), with_rgb=self.with_rgb, with_normal=self.with_normal, cnt=self.cnt)
        finally:
            if cv is not None:
                cv._use_w_dev = False
        loss = None
        if self.loss_fn is not None:
            loss = self.loss_fn(rendered)
            if loss.requires_grad:
                loss.backward()
            loss = loss.detach()
        return rendered, buffers, loss

    def capture(self):
        if self.march_cap is None or self.kept_cap is None or self.coherent is None:
            self._size()
        if not self.use_graph:
            self.graph = None
            return self
        if not bool(self.rays_d.any()):
            # the warm-up and the capture RUN the step on the input buffers; a zero direction never leaves the marcher's voxel-skipping loop
            # (the reference's kernel has the same loop, occ_grid/helpers_march.h:58-69) -- refuse instead of hanging the device
            raise RuntimeError("StaticFrame.capture: the input buffers hold no rays yet; call step(rays_o, rays_d, ...) or fill frame.rays_o / frame.rays_d first")
        was = L.KERNEL_TIMER.enabled
        L.KERNEL_TIMER.enabled = False                      # events cannot be recorded into a capture
        import gc
        gc.collect()                                        # autograd graphs of earlier (default-stream) backward passes that only the cycle collector
        #                                                     frees keep the parameters' AccumulateGrad nodes on the default stream -> capture error
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                          # warm-up on a side stream (allocator, lazy module state)
                    self._run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # capture on the warm-up's stream: the parameters' AccumulateGrad nodes were created there.  (A backward that ran on the default
            # stream BEFORE this and whose graph is still referenced -- a kept loss / rendered tensor -- pins those nodes to the default
            # stream and invalidates the capture: drop such references first.)
            self._occ_captured = self.model.accel.occ.occ_grid
            with torch.cuda.graph(g, stream=side):
                self.rendered, self.buffers, self.loss = self._run()
            self.graph = g
            self.captures += 1
        finally:
            L.KERNEL_TIMER.enabled = was
        return self

    def step(self, rays_o, rays_d, rays_h_appear=None):
        """copy the batch into the graph's inputs (H2D if the tensors are on the host) and launch.  -> loss (device scalar) or None"""
        self.rays_o.copy_(rays_o, non_blocking=True)
        self.rays_d.copy_(rays_d, non_blocking=True)
        if self.h_appear is not None and rays_h_appear is not None:
            self.h_appear.copy_(rays_h_appear, non_blocking=True)
        cv = getattr(self.model, "ctrl_var", None)
        if cv is not None and hasattr(cv, "mix_weight"):
            if getattr(cv, "_w_dev", None) is None:
                cv._w_dev = torch.zeros((), device=self.device)
            cv._w_dev.fill_(cv.mix_weight())                  # the variance schedule's host-side weight of THIS iteration
        if self.graph is None and (self.use_graph or self.march_cap is None):
            self.capture()
        if self.graph is not None:
            occ = self.model.accel.occ.occ_grid
            if self._occ_captured is not None and occ.data_ptr() != self._occ_captured.data_ptr():
                # somebody RE-ASSIGNED the grid (the reference's EMA does, ema_single.py:190): the graph reads the tensor it captured -> refresh it
                self._occ_captured.copy_(occ)
            self.graph.replay()
        else:
            self.rendered, self.buffers, self.loss = self._run()
        return self.loss

    def counts(self):
        """host copy of the step's sizes (one D2H + sync)"""
        c = self.cnt.tolist()
        return {k: c[v] for k, v in CNT_SLOTS.items()}

    def check(self, retry=True):
        """True if the last step fitted its arenas.  Otherwise the arenas are re-sized from this batch, the graph is re-captured and --
        with `retry` -- the step is run again (gradients of the overflowed step were those of an empty render: nothing accumulated)."""
        if int(self.cnt[CNT_SLOTS["overflow"]]) == 0:
            return True
        self.march_cap = self.kept_cap = None
        self._size(grow=1.25)
        self.graph = None
        if self.use_graph:
            self.capture()
        if retry:
            if self.graph is not None:
                self.graph.replay()
            else:
                self.rendered, self.buffers, self.loss = self._run()
        return False

    def volume_buffer(self):
        return sliced_volume_buffer(self.buffers, self.cnt)
