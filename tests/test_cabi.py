"""The C-ABI library builds, loads and exports every symbol include/msb200.h declares.
No compute calls here (no GPU in the CPU test tier)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from motifscan_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "msb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in msb200.h but not exported"


def test_binding_covers_header(lib):
    from motifscan_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_gpu_fails_loudly(lib):
    """Without a device the library must refuse, not fall back to anything."""
    from motifscan_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    rc = lib.msb_ctx_create(0, None, ctypes.byref(h))
    assert rc == _lib.MSB_ECUDA
    assert b"no CPU path" in lib.msb_last_error()
    from motifscan_b200.motif.cscore import c_scan_motif
    with pytest.raises(_lib.MsbError):
        c_scan_motif([[[1.0], [0.0], [0.0], [0.0]]], [0.5], ["ACGT"], 3, 1)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under motifscan_b200/ may reference it."""
    pkg = os.path.join(ROOT, "motifscan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_shim_module_resolves_without_a_gpu():
    """shim/motifscan/motif/cscore.py is the file a maintainer drops over the reference's extension: under
    the reference's module name it exposes exactly c_score and c_scan_motif (cscore.c:479-482)."""
    import sys
    shim = os.path.join(ROOT, "shim")
    saved = {k: v for k, v in sys.modules.items() if k == "motifscan" or k.startswith("motifscan.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, shim)
    try:
        import motifscan.motif.cscore as mod
        from motifscan_b200.motif import cscore as ours
        assert mod.c_score is ours.c_score and mod.c_scan_motif is ours.c_scan_motif
    finally:
        sys.path.remove(shim)
        for k in [k for k in sys.modules if k == "motifscan" or k.startswith("motifscan.")]:
            del sys.modules[k]
        sys.modules.update(saved)
