"""The C-ABI library loads and exports every symbol include/umr_b200.h declares (no compute calls
without a GPU), and the product path fails loudly when the library is missing."""
import ctypes
import os
import re

import pytest

from umr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "umr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(umr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = header_functions()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_scalar_entry_points_without_gpu():
    lib = _lib.load()
    assert lib.umr_version() >= 201
    assert lib.umr_error_string(0) == b"ok"
    assert b"not supported" in lib.umr_error_string(-1)
    assert lib.umr_raster_workspace_bytes(16, 1280, 256, 1) >= 16 * 1280 * (128 + 16 + 16) + 16 * 64 * 1280 * 2
    assert lib.umr_raster_workspace_bytes(0, 5, 64, 1) == 0
    # pair buffer: 1540 bytes per 32-record block + per-tile headers
    assert lib.umr_raster_pair_buffer_bytes(2, 64, 1, 1000) >= 1000 * 1540 + 2 * 64 * 4
    assert lib.umr_raster_pair_buffer_bytes(0, 64, 1, 10) == 0
    assert lib.umr_launch_count() >= 0


def test_params_struct_matches_header():
    p = _lib.UmrRasterParams()
    # 4 bytes padding before the pointers; shared_textures, tile_mode, color_channels, background_extra = 16
    assert ctypes.sizeof(p) == 5 * 4 + 6 * 4 + 5 * 4 + 3 * 4 + 4 + 2 * 8 + 8 + 8 + 16
    lib = _lib.load()
    assert lib.umr_sizeof_raster_params() == ctypes.sizeof(p)
    assert lib.umr_sizeof_project_params() == ctypes.sizeof(_lib.UmrProjectParams())


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libumr_b200.so")
    with pytest.raises(_lib.UmrLibraryError):
        _lib.load()


def test_product_never_imports_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "umr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(import|from)\s+(softras|build_oracle|oracle)\b", s, flags=re.M) or "oracle/" in s and f.endswith(".py") and "import" in s and re.search(r"sys\.path.*oracle", s):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
