"""ctypes front-end of oracle/march.c with the `nr3d_lib.bindings._occ_grid` call surface
(csrc/occ_grid/src/occ_grid.cpp:21-33, include/occ_grid/cpp_api.h:14-65).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_march.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "march.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.nsb_oracle_ray_marching.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def ray_marching(rays_o, rays_d, t_min, t_max, roi, grid_binary, step_size, max_step_size, dt_gamma, max_steps,
                 return_gidx=True, batch_inds=None):
    """-> [packed_info i32[R,2], t_starts[M], t_ends[M], ridx i32[M], gidx i32[M](, bidx)] (squeezed t's).
    Two passes with an int32 cumsum in between, as ray_marching.cu:179-241."""
    lib = _load()
    o = np.ascontiguousarray(rays_o.detach().numpy(), dtype=np.float32)
    d = np.ascontiguousarray(rays_d.detach().numpy(), dtype=np.float32)
    tn = np.ascontiguousarray(t_min.detach().numpy(), dtype=np.float32)
    tx = np.ascontiguousarray(t_max.detach().numpy(), dtype=np.float32)
    r = np.ascontiguousarray(roi.detach().numpy(), dtype=np.float32)
    g = np.ascontiguousarray(grid_binary.detach().numpy().astype(np.uint8))
    bi = None if batch_inds is None else np.ascontiguousarray(batch_inds.detach().numpy(), dtype=np.int32)
    res = g.shape[-3:]
    R = o.shape[0]
    f, i32, u8 = ctypes.c_float, ctypes.c_int32, ctypes.c_uint8
    num = np.zeros(R, dtype=np.int32)
    args = (R, _p(o, f), _p(d, f), _p(tn, f), _p(tx, f), _p(r, f), _p(bi, i32), int(res[0]), int(res[1]), int(res[2]),
            _p(g, u8), f(step_size), f(max_step_size), f(dt_gamma), ctypes.c_uint32(max_steps))
    lib.nsb_oracle_ray_marching(*args, None, _p(num, i32), None, None, None, None, None)
    cum = np.cumsum(num, dtype=np.int32)
    info = np.ascontiguousarray(np.stack([cum - num, num], 1).astype(np.int32))
    M = int(cum[-1]) if R else 0
    t0 = np.zeros(M, np.float32); t1 = np.zeros(M, np.float32)
    ridx = np.zeros(M, np.int32); gidx = np.zeros(M, np.int32)
    bidx = np.zeros(M, np.int32) if bi is not None else None
    lib.nsb_oracle_ray_marching(*args, _p(info, i32), None, _p(t0, f), _p(t1, f), _p(ridx, i32),
                                _p(gidx, i32) if return_gidx else None, _p(bidx, i32))
    out = [torch.from_numpy(info), torch.from_numpy(t0), torch.from_numpy(t1), torch.from_numpy(ridx),
           torch.from_numpy(gidx)]
    if bi is not None:
        out.append(torch.from_numpy(bidx))
    return out
