"""Host-side mirror of the reference's native-extension modules (`nr3d_lib.bindings._*`).

`neuralsim_b200.bindings._lotd / _pack_ops / _occ_grid / _shencoder` export the functions, argument orders and
error behaviour of the pybind11 modules built by /root/reference/nr3d_lib/setup.py, implemented by ctypes calls
into libneuralsim_b200.so.  `install_as_nr3d_lib_bindings()` registers them under the reference's module names so
that an unmodified `nr3d_lib` Python tree imports them.
"""
from . import _lotd, _pack_ops, _occ_grid, _shencoder  # noqa: F401


def install_as_nr3d_lib_bindings():
    import sys
    import types
    pkg = sys.modules.get("nr3d_lib.bindings")
    if pkg is None:
        pkg = types.ModuleType("nr3d_lib.bindings")
        pkg.__path__ = []
        sys.modules["nr3d_lib.bindings"] = pkg
    for name, mod in (("_lotd", _lotd), ("_pack_ops", _pack_ops), ("_occ_grid", _occ_grid), ("_shencoder", _shencoder)):
        sys.modules[f"nr3d_lib.bindings.{name}"] = mod
        setattr(pkg, name, mod)
