"""The C-ABI library loads without a GPU and exports every symbol include/hirest_b200.h declares."""
import ctypes
import os
import re

from hirest_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(headers=("hirest_b200.h", "hirest_b200_debug.h")):
    text = "".join(open(os.path.join(ROOT, "include", h)).read() for h in headers)
    return sorted(set(re.findall(r"HB_API[^;(]*?\b(hb_\w+)\s*\(", text)))


def test_boundary_header_has_no_tuning_switches():
    """A/B knobs do not belong in the drop-in boundary; they live in hirest_b200_debug.h (hb_debug_set)."""
    syms = _declared_symbols(("hirest_b200.h",))
    assert not [s for s in syms if s.startswith("hb_set_") or s.startswith("hb_debug")]
    assert "hb_debug_set" in _declared_symbols(("hirest_b200_debug.h",))
    lib = _lib.load()
    assert lib.hb_debug_set(b"no_such_key", 1) == -22 and lib.hb_debug_set(b"attention_version", 7) == -22
    assert lib.hb_debug_set(b"attention_version", 3) == 0


def test_header_declares_the_path():
    syms = _declared_symbols()
    for must in ("hb_vit_encode", "hb_text_encode", "hb_pool_normalize", "hb_similarity", "hb_linear", "hb_init"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_python_binding_covers_every_symbol():
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_error_strings_without_device():
    lib = _lib.load()
    assert lib.hb_strerror(0) == b"ok"
    assert b"invalid" in lib.hb_strerror(-22)
    assert lib.hb_launch_count() == 0


def test_struct_layouts_match_header():
    # 7 ints + 1 float / 6 ints + 1 float + 1 int, natural alignment; pointer-only weight structs
    assert ctypes.sizeof(_lib.HbVitConfig) == 32
    assert ctypes.sizeof(_lib.HbTextConfig) == 32
    assert ctypes.sizeof(_lib.HbVitWeights) == 21 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.HbTextWeights) == 17 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.HbProfileSummary) == 6 * 8 * 3
