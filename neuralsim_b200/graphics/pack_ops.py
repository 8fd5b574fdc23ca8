"""Packed-tensor operators with autograd -- the API of `nr3d_lib.graphics.pack_ops`
(reference: nr3d_lib/nr3d_lib/graphics/pack_ops/pack_ops.py) on top of neuralsim_b200's CUDA kernels.

A packed tensor stores per-ray (per-"pack") variable-length data back to back; `pack_infos[P,2]` (int64) holds
(first index, length).  Gradient rules are the adjoints the reference implements (pack_ops.py:97-392).
"""
from __future__ import annotations

import torch

from ..bindings import _pack_ops as _backend

__all__ = [
    "packed_sum", "packed_mean", "packed_cumsum", "packed_cumprod", "packed_diff", "packed_backward_diff", "packed_add",
    "packed_sub", "packed_mul", "packed_div", "packed_matmul", "packed_gt", "packed_geq", "packed_lt", "packed_leq", "packed_eq",
    "packed_neq", "packed_searchsorted", "packed_invert_cdf", "packed_alpha_to_vw", "packed_volume_render_compression",
    "packed_sort", "packed_sort_inplace", "interleave_arange_simple", "interleave_arange", "interleave_linstep",
    "interleave_linspace", "merge_two_packs_sorted_aligned", "merge_two_packs_sorted_a_includes_b", "merge_two_packs_sorted",
    "merge_two_batch_a_includes_b", "get_pack_infos_from_boundary", "get_pack_infos_from_first", "get_pack_infos_from_n",
    "get_pack_infos_from_batch", "mark_pack_boundaries",
]


# ------------------------------------------------------------------ pack_infos builders (no kernels involved)
@torch.no_grad()
def get_pack_infos_from_n(n_per_pack):
    return torch.stack([n_per_pack.cumsum(0) - n_per_pack, n_per_pack], 1)


@torch.no_grad()
def get_pack_infos_from_first(first_inds, numel):
    return torch.stack([first_inds, first_inds.diff(append=first_inds.new_tensor([numel]))], 1)


@torch.no_grad()
def get_pack_infos_from_boundary(boundary):
    return get_pack_infos_from_first(boundary.nonzero().long()[..., 0], boundary.numel())


@torch.no_grad()
def get_pack_infos_from_batch(n_batches, batch_data_size, device=None):
    first = torch.arange(0, n_batches * batch_data_size, batch_data_size, device=device, dtype=torch.long)
    return torch.stack([first, torch.full_like(first, batch_data_size)], 1)


def mark_pack_boundaries(pack_ids):
    return _backend.mark_pack_boundaries_cuda(pack_ids.contiguous()).bool()


# ------------------------------------------------------------------ reductions and scans
class _Sum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, pack_infos):
        ctx.save_for_backward(pack_infos)
        return _backend.packed_sum(feats, pack_infos)

    @staticmethod
    def backward(ctx, g):
        (pack_infos,) = ctx.saved_tensors
        return g.repeat_interleave(pack_infos[:, 1], dim=0), None


def packed_sum(feats, pack_infos):
    feats = feats.contiguous()
    return _Sum.apply(feats, pack_infos) if feats.requires_grad else _backend.packed_sum(feats, pack_infos)


def packed_mean(feats, pack_infos):
    n = pack_infos[:, 1] + 1e-8
    return packed_sum(feats, pack_infos) / (n if feats.dim() == 1 else n.unsqueeze(-1))


class _Cumsum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, pack_infos, exclusive, reverse):
        ctx.save_for_backward(pack_infos)
        ctx.flags = (exclusive, reverse)
        return _backend.packed_cumsum(feats, pack_infos, exclusive, reverse)

    @staticmethod
    def backward(ctx, g):
        (pack_infos,) = ctx.saved_tensors
        exclusive, reverse = ctx.flags
        return _backend.packed_cumsum(g.contiguous(), pack_infos, exclusive, not reverse), None, None, None


def packed_cumsum(feats, pack_infos, exclusive=False, reverse=False):
    feats = feats.contiguous()
    if feats.requires_grad:
        return _Cumsum.apply(feats, pack_infos, exclusive, reverse)
    return _backend.packed_cumsum(feats, pack_infos, exclusive, reverse)


def packed_cumprod(feats, pack_infos, exclusive=False, reverse=False):
    return _backend.packed_cumprod(feats.contiguous(), pack_infos, exclusive, reverse)


class _Diff(torch.autograd.Function):
    """Forward difference inside each pack; the adjoint is (minus) the backward difference with edge fix-ups."""

    @staticmethod
    def forward(ctx, feats, pack_infos, appends, last_fill):
        ctx.save_for_backward(pack_infos)
        ctx.flags = (appends is not None, last_fill is not None)
        return _backend.packed_diff(feats, pack_infos, appends, last_fill)

    @staticmethod
    def backward(ctx, g):
        (pack_infos,) = ctx.saved_tensors
        has_append, has_fill = ctx.flags
        g = g.contiguous()
        first, n = pack_infos[:, 0], pack_infos[:, 1]
        last = first + n - 1
        gf = None
        if ctx.needs_input_grad[0]:
            gf = -1 * _backend.packed_backward_diff(g, pack_infos, None, g[first].contiguous())
            if not has_append:
                prev = g[(last - 1).clamp_min(0)]
                gf[last] = torch.where((n > 1).view(-1, *([1] * (g.dim() - 1))), prev, torch.zeros_like(prev))
        ga = g[last] if (has_append and ctx.needs_input_grad[2]) else None
        gl = g[last] if (has_fill and ctx.needs_input_grad[3]) else None
        return gf, None, ga, gl


def packed_diff(feats, pack_infos, pack_appends=None, pack_last_fill=None):
    feats = feats.contiguous()
    if feats.requires_grad or (pack_appends is not None and pack_appends.requires_grad):
        return _Diff.apply(feats, pack_infos, pack_appends, pack_last_fill)
    return _backend.packed_diff(feats, pack_infos, pack_appends, pack_last_fill)


class _BackwardDiff(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, pack_infos, prepends, first_fill):
        ctx.save_for_backward(pack_infos)
        ctx.flags = (prepends is not None, first_fill is not None)
        return _backend.packed_backward_diff(feats, pack_infos, prepends, first_fill)

    @staticmethod
    def backward(ctx, g):
        (pack_infos,) = ctx.saved_tensors
        has_prepend, has_fill = ctx.flags
        g = g.contiguous()
        first, n = pack_infos[:, 0], pack_infos[:, 1]
        last = first + n - 1
        gf = None
        if ctx.needs_input_grad[0]:
            gf = -1 * _backend.packed_diff(g, pack_infos, None, (-g[last]).contiguous())
            if not has_prepend:
                nxt = g[(first + 1).clamp_max(g.shape[0] - 1)]
                gf[first] = torch.where((n > 1).view(-1, *([1] * (g.dim() - 1))), -nxt, torch.zeros_like(nxt))
        gp = -g[first] if (has_prepend and ctx.needs_input_grad[2]) else None
        gl = g[first] if (has_fill and ctx.needs_input_grad[3]) else None
        return gf, None, gp, gl


def packed_backward_diff(feats, pack_infos, pack_prepends=None, pack_first_fill=None):
    feats = feats.contiguous()
    if feats.requires_grad:
        return _BackwardDiff.apply(feats, pack_infos, pack_prepends, pack_first_fill)
    return _backend.packed_backward_diff(feats, pack_infos, pack_prepends, pack_first_fill)


# ------------------------------------------------------------------ per-pack broadcast arithmetic
def _bcast(grad_shape, g):
    return g


class _Arith(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, other, pack_infos, op):
        ctx.op = op
        ctx.save_for_backward(feats, other, pack_infos)
        return getattr(_backend, f"packed_{op}")(feats, other, pack_infos)

    @staticmethod
    def backward(ctx, g):
        feats, other, pack_infos = ctx.saved_tensors
        g = g.contiguous()
        op = ctx.op
        gi = go = None
        if op in ("add", "sub"):
            gi = g if ctx.needs_input_grad[0] else None
            if ctx.needs_input_grad[1]:
                go = _backend.packed_sum(g, pack_infos)
                go = -go if op == "sub" else go
        elif op == "mul":
            if ctx.needs_input_grad[0]:
                gi = _backend.packed_mul(g, other, pack_infos)
            if ctx.needs_input_grad[1]:
                go = _backend.packed_sum((g * feats).contiguous(), pack_infos)
        else:  # div
            if ctx.needs_input_grad[0]:
                gi = _backend.packed_div(g, other, pack_infos)
            if ctx.needs_input_grad[1]:
                go = _backend.packed_sum(_backend.packed_div((-g * feats).contiguous(), (other * other).contiguous(), pack_infos), pack_infos)
        return gi, go, None, None


def _arith(op):
    def fn(feats, other, pack_infos):
        feats, other, pack_infos = feats.contiguous(), other.contiguous(), pack_infos.contiguous()
        if feats.requires_grad or other.requires_grad:
            return _Arith.apply(feats, other, pack_infos, op)
        return getattr(_backend, f"packed_{op}")(feats, other, pack_infos)
    fn.__name__ = f"packed_{op}"
    return fn


packed_add, packed_sub, packed_mul, packed_div = _arith("add"), _arith("sub"), _arith("mul"), _arith("div")


def packed_matmul(feats, other, pack_infos):
    return (feats.unsqueeze(-2) * torch.repeat_interleave(other, pack_infos[:, 1], dim=0)).sum(-1)


def _cmp(op):
    def fn(feats, other, pack_infos):
        return getattr(_backend, f"packed_{op}")(feats.contiguous(), other.contiguous(), pack_infos.contiguous())
    fn.__name__ = f"packed_{op}"
    return fn


packed_gt, packed_geq, packed_lt, packed_leq, packed_eq, packed_neq = (_cmp(o) for o in ("gt", "geq", "lt", "leq", "eq", "neq"))


# ------------------------------------------------------------------ search / sample / sort
@torch.no_grad()
def packed_searchsorted(bins, vals, pack_infos):
    return _backend.packed_searchsorted(bins.contiguous(), vals.contiguous(), pack_infos)


@torch.no_grad()
def packed_invert_cdf(bins, cdfs, u_vals, pack_infos):
    return _backend.packed_invert_cdf(bins.contiguous(), cdfs.contiguous(), u_vals.contiguous(), pack_infos)


@torch.no_grad()
def packed_sort_inplace(vals, pack_infos, return_idx=True):
    return _backend.packed_sort_qsort(vals.contiguous(), pack_infos, return_idx)


def packed_sort(vals, pack_infos):
    indices = packed_sort_inplace(vals.detach().clone(), pack_infos, return_idx=True)
    return vals[indices], indices


# ------------------------------------------------------------------ volume rendering weights
class _AlphaToVW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alphas, pack_infos, early_stop_eps, alpha_thre):
        w = _backend.packed_alpha_to_vw_forward(alphas, pack_infos, early_stop_eps, alpha_thre, False)[0]
        ctx.save_for_backward(alphas, pack_infos, w)
        ctx.cfg = (early_stop_eps, alpha_thre)
        return w

    @staticmethod
    def backward(ctx, gw):
        alphas, pack_infos, w = ctx.saved_tensors
        return _backend.packed_alpha_to_vw_backward(w, gw.contiguous(), alphas, pack_infos, *ctx.cfg), None, None, None


def packed_alpha_to_vw(alpha, pack_infos, early_stop_eps=1e-4, alpha_thre=0.0):
    alpha = alpha.contiguous()
    if alpha.requires_grad:
        return _AlphaToVW.apply(alpha, pack_infos, early_stop_eps, alpha_thre)
    return _backend.packed_alpha_to_vw_forward(alpha, pack_infos, early_stop_eps, alpha_thre, False)[0]


@torch.no_grad()
def packed_volume_render_compression(alpha, pack_infos, early_stop_eps=1e-4, alpha_thre=0.0):
    """-> (indices of packs that keep samples, their compacted pack_infos, indices of the kept samples)."""
    _, info, sel = _backend.packed_alpha_to_vw_forward(alpha.contiguous(), pack_infos, early_stop_eps, alpha_thre, True)
    pidx = sel.nonzero().long()[..., 0]
    nidx = (info[:, 1] > 0).nonzero()[..., 0]
    return nidx, info[nidx].long(), pidx


# ------------------------------------------------------------------ producers
@torch.no_grad()
def interleave_arange_simple(stop, return_idx=True):
    out, nidx = _backend.interleave_arange(stop.contiguous(), return_idx)
    return (out, nidx) if return_idx else out


@torch.no_grad()
def interleave_linstep(start, num_steps, step_size, return_idx=True):
    out, nidx = _backend.interleave_linstep(start.contiguous().float(), num_steps.contiguous(), step_size, return_idx)
    if start.dtype in (torch.int64, torch.int32):
        out = out.round().to(start.dtype)
    return (out, nidx) if return_idx else out


@torch.no_grad()
def interleave_arange(start, stop, step_size, return_idx=True):
    return interleave_linstep(start, stop.subtract(start).div(step_size).ceil().long(), step_size, return_idx)


@torch.no_grad()
def interleave_linspace(start, stop, num_steps, return_idx=True):
    step = (stop - start) / (num_steps - 1)
    if not isinstance(num_steps, torch.Tensor):
        num_steps = torch.full(start.shape, num_steps, device=start.device, dtype=torch.long)
    return interleave_linstep(start, num_steps, step, return_idx)


# ------------------------------------------------------------------ merging sorted packs / batches
def merge_two_packs_sorted_aligned(vals_a, pack_infos_a, vals_b, pack_infos_b, b_sorted=True, return_val=False):
    pidx_a, pidx_b, pack_infos = _backend.try_merge_two_packs_sorted_aligned(
        vals_a.detach().contiguous(), pack_infos_a, vals_b.detach().contiguous(), pack_infos_b, b_sorted)
    if not return_val:
        return pidx_a, pidx_b, pack_infos
    val = vals_a.new_empty([vals_a.numel() + vals_b.numel()])
    val[pidx_a], val[pidx_b] = vals_a, vals_b
    return val, pack_infos


def _offsets_of(pack_infos, idx):
    """flat element indices of the packs `idx` + the pack each element came from (device-side ragged arange)."""
    n = pack_infos[idx, 1].contiguous()
    local, which = interleave_arange_simple(n, return_idx=True)
    return local + pack_infos[idx, 0][which], local, which


def merge_two_packs_sorted_a_includes_b(vals_a, pack_infos_a, nidx_a, vals_b, pack_infos_b, nidx_b, b_sorted=True, return_val=False):
    if nidx_a.numel() == nidx_b.numel() and torch.equal(nidx_a, nidx_b):
        return merge_two_packs_sorted_aligned(vals_a, pack_infos_a, vals_b, pack_infos_b, b_sorted, return_val)
    with torch.no_grad():
        where_b = torch.searchsorted(nidx_a, nidx_b)
        only_a = torch.ones(nidx_a.numel(), dtype=torch.bool, device=vals_a.device)
        only_a[where_b] = False
        only_a = only_a.nonzero().long()[..., 0]
        n_per = pack_infos_a[:, 1].clone()
        n_per.index_add_(0, where_b, pack_infos_b[:, 1])
        pack_infos = get_pack_infos_from_n(n_per)
        pidx_a = pack_infos_a.new_full([vals_a.numel()], -1)
        ia, _, _ = _offsets_of(pack_infos_a, where_b)
        pinfo_a_u = get_pack_infos_from_n(pack_infos_a[where_b, 1].contiguous())
        pa_u, pb_u, pinfo_u = merge_two_packs_sorted_aligned(vals_a[ia], pinfo_a_u, vals_b, pack_infos_b, b_sorted)
        shift = pack_infos[where_b, 0] - pinfo_u[:, 0]
        pidx_a[ia] = pa_u + torch.repeat_interleave(shift, pinfo_a_u[:, 1])
        pidx_b = pb_u + torch.repeat_interleave(shift, pack_infos_b[:, 1])
        if only_a.numel() > 0:
            src, local, which = _offsets_of(pack_infos_a, only_a)
            pidx_a[src] = local + pack_infos[only_a, 0][which]
    if not return_val:
        return pidx_a, pidx_b, pack_infos
    val = vals_a.new_zeros([vals_a.numel() + vals_b.numel()])
    val[pidx_a], val[pidx_b] = vals_a, vals_b
    return val, pack_infos


def merge_two_packs_sorted(vals_a, pack_infos_a, nidx_a, vals_b, pack_infos_b, nidx_b, return_val=False):
    """Merge two sorted packed buffers whose pack ids (nidx, sorted & unique) only partly overlap
    (close-range + distant-view buffers, single_volume_renderer.py:337-375)."""
    if nidx_a.numel() == nidx_b.numel() and torch.equal(nidx_a, nidx_b):
        return merge_two_packs_sorted_aligned(vals_a, pack_infos_a, vals_b, pack_infos_b, True, return_val)
    with torch.no_grad():
        u, inv = torch.unique(torch.cat([nidx_a, nidx_b]), return_inverse=True)
        inv_a, inv_b = inv[:nidx_a.numel()], inv[nidx_a.numel():]
        n_per = pack_infos_a.new_zeros([u.numel()])
        n_per.index_add_(0, inv_a, pack_infos_a[:, 1])
        n_per.index_add_(0, inv_b, pack_infos_b[:, 1])
        pack_infos = get_pack_infos_from_n(n_per)
        has_a = torch.zeros(u.numel(), dtype=torch.bool, device=u.device)
        has_b = torch.zeros_like(has_a)
        has_a[inv_a], has_b[inv_b] = True, True
        both = has_a & has_b
        in_a, in_b = both[inv_a].nonzero()[..., 0], both[inv_b].nonzero()[..., 0]
        ex_a, ex_b = (~both[inv_a]).nonzero()[..., 0], (~both[inv_b]).nonzero()[..., 0]
        pidx_a = pack_infos_a.new_full([vals_a.numel()], -1)
        pidx_b = pack_infos_b.new_full([vals_b.numel()], -1)
        if in_a.numel() > 0:
            ia, _, wa = _offsets_of(pack_infos_a, in_a)
            ib, _, wb = _offsets_of(pack_infos_b, in_b)
            pa_u, pb_u, pinfo_u = merge_two_packs_sorted_aligned(
                vals_a[ia], get_pack_infos_from_n(pack_infos_a[in_a, 1].contiguous()),
                vals_b[ib], get_pack_infos_from_n(pack_infos_b[in_b, 1].contiguous()))
            shift = pack_infos[inv_a[in_a], 0] - pinfo_u[:, 0]
            pidx_a[ia] = pa_u + shift[wa]
            pidx_b[ib] = pb_u + shift[wb]
        if ex_a.numel() > 0:
            src, local, which = _offsets_of(pack_infos_a, ex_a)
            pidx_a[src] = local + pack_infos[inv_a[ex_a], 0][which]
        if ex_b.numel() > 0:
            src, local, which = _offsets_of(pack_infos_b, ex_b)
            pidx_b[src] = local + pack_infos[inv_b[ex_b], 0][which]
    if not return_val:
        return pidx_a, pidx_b, pack_infos
    val = vals_a.new_zeros([vals_a.numel() + vals_b.numel()])
    val[pidx_a], val[pidx_b] = vals_a, vals_b
    return val, pack_infos


def merge_two_batch_a_includes_b(vals_a, nidx_a, vals_b, nidx_b, a_sorted=True, return_val=False):
    """Rows of the batch `b` ([Nb, Wb], ids nidx_b) are merged into the matching rows of `a` ([Na, Wa]); rows of `a`
    without a partner are kept.  -> destination indices of every element + pack_infos of the merged packs."""
    device = vals_a.device
    n_a, wa, wb = nidx_a.numel(), vals_a.shape[-1], vals_b.shape[-1]
    same = n_a == nidx_b.numel() and torch.equal(nidx_a, nidx_b)
    where_b = torch.arange(n_a, device=device) if same else torch.searchsorted(nidx_a, nidx_b)
    n_per = nidx_a.new_full([n_a], wa)
    n_per[where_b] = wa + wb
    pack_infos = get_pack_infos_from_n(n_per)
    first = pack_infos[:, 0:1]
    rank = torch.cat([vals_a[where_b], vals_b], 1).detach().argsort(dim=-1, stable=True).argsort(dim=-1)
    if a_sorted:
        pidx_a = first + torch.arange(wa, device=device)[None, :]
    else:
        pidx_a = first + vals_a.detach().argsort(-1).argsort(-1)
    pidx_a[where_b] = first[where_b] + rank[:, :wa]
    pidx_b = first[where_b] + rank[:, wa:]
    if not return_val:
        return pidx_a, pidx_b, pack_infos
    vals = vals_a.new_empty([vals_a.numel() + vals_b.numel()])
    vals[pidx_a], vals[pidx_b] = vals_a, vals_b
    return vals, pack_infos
