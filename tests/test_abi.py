"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/*.h declares; the ctypes table mirrors the header one to one.  No
compute call is made (no GPU here)."""
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(dmh_[a-z0-9_]+)\s*\(", txt))
    return names


@pytest.fixture(scope="module")
def lib():
    from depthmodelhardening_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from depthmodelhardening_b200 import _lib
    declared = header_symbols()
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))


def test_built_for_sm100a_only(lib):
    from depthmodelhardening_b200 import _lib
    assert lib.dmh_build_arch() == 100
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode == 0:
        archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
        assert archs == {"100a"}, archs


def test_error_channel_without_gpu(lib):
    rc = lib.dmh_ssim_fwd(None, None, 1, 3, 8, 8, None, None)
    assert rc == 1
    assert b"null pointer" in lib.dmh_last_error()
    assert lib.dmh_photo_tiles(320, 1024) == 32 * 20
    assert lib.dmh_smooth_workspace_floats(2, 8, 8) > 0


def test_cpu_tensor_is_rejected(lib):
    import torch
    from depthmodelhardening_b200 import layers
    with pytest.raises(RuntimeError, match="CUDA-only"):
        layers.SSIM()(torch.rand(1, 3, 8, 8), torch.rand(1, 3, 8, 8))


def test_oracle_not_imported_by_product():
    """The product package must never import oracle/ (parity would be void)."""
    pkg = os.path.join(ROOT, "depthmodelhardening_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
