"""Drop-in for `nr3d_lib.bindings._shencoder` (reference: nr3d_lib/externals/shencoder/bindings.cpp, shencoder.h).
Caller-allocated outputs, as in the reference; launches on the *current* stream."""
from __future__ import annotations

from .. import _lib as L


def sh_encode_forward(inputs, outputs, B, D, C, calc_grad_inputs, dy_dx):
    if D != 3:
        raise RuntimeError("SH encoder only support input dim == 3")
    L.check(L.lib().nsb_sh_encode_forward(L.ptr(inputs, "f32", "inputs"), L.ptr(outputs, "f32", "outputs"), L.c_i64(B), L.c_i32(C),
                                          L.ptr(dy_dx, "f32", "dy_dx") if calc_grad_inputs else None, L.stream_ptr()),
            "sh_encode_forward")


def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
    L.check(L.lib().nsb_sh_encode_backward(L.ptr(grad, "f32", "grad"), L.ptr(dy_dx, "f32", "dy_dx"), L.c_i64(B), L.c_i32(C),
                                           L.ptr(grad_inputs, "f32", "grad_inputs"), L.stream_ptr()), "sh_encode_backward")
