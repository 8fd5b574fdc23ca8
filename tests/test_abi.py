"""The C-ABI library builds, loads, and exports every symbol include/neuralsim_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from neuralsim_b200 import build
    build.build_library()
    from neuralsim_b200 import _lib
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "neuralsim_b200.h")).read()
    names = sorted(set(re.findall(r"\b(nsb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_meta_create_matches_oracle(lib):
    from neuralsim_b200.bindings import _lotd
    from oracle import lotd as olotd
    cfg = olotd.gen_ngp_cfg()
    m = _lotd.LoDMeta(3, cfg["lod_res"], cfg["lod_n_feats"], cfg["lod_types"], cfg["hashmap_size"])
    o = olotd.LoDMeta(3, **cfg)
    assert m.n_params == o.n_params == 12131648 and m.n_encoded_dims == 32 and m.n_pseudo_levels == 16
    for a in ("level_offsets", "level_sizes", "level_n_feats", "level_types", "level_res_multidim", "map_levels", "map_cnt", "level_n_params"):
        assert getattr(m, a) == getattr(o, a), a
    cub = _lotd.LoDMeta(3, [[4, 6, 8], [8, 12, 16]], [2, 4], ["Dense", "Hash"], 4096)     # cuboid + mixed widths -> pseudo levels
    oc = olotd.LoDMeta(3, [[4, 6, 8], [8, 12, 16]], [2, 4], ["Dense", "Hash"], 4096)
    assert cub.n_pseudo_levels == oc.n_pseudo_levels == 3 and cub.level_offsets == oc.level_offsets and cub.level_res == [0, 0]


def test_errors_surface_as_runtime_error(lib):
    from neuralsim_b200.bindings import _lotd
    with pytest.raises(RuntimeError, match="n_input_dim"):
        _lotd.LoDMeta(5, [16], [2], ["Dense"], None)
    with pytest.raises(RuntimeError, match="greatest common divisor"):
        _lotd.LoDMeta(3, [16], [3], ["Dense"], None)
    assert lib.nsb_version() >= 100


def test_no_cpu_path():
    """Product ops refuse CPU tensors instead of falling back."""
    import torch
    from neuralsim_b200.bindings import _pack_ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _pack_ops.packed_sum(torch.zeros(4), torch.tensor([[0, 4]]))


def test_product_never_imports_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "neuralsim_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "from .. import oracle" in src:
                    bad.append(f)
    assert not bad, bad


def test_every_python_module_imports():
    """A syntax / import error anywhere in the package must fail the CPU suite, not the first GPU call."""
    import importlib
    import pkgutil

    import neuralsim_b200
    for m in pkgutil.walk_packages(neuralsim_b200.__path__, "neuralsim_b200."):
        if not m.name.rsplit(".", 1)[-1].startswith("lib"):          # libneuralsim_b200.so is the C ABI, not a Python extension
            importlib.import_module(m.name)
    import bench  # noqa: F401
    import __graft_entry__  # noqa: F401
