"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol that
include/fsb200.h declares; the ctypes binding covers the same set (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "fsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fasterseg_b200 import _lib, build
    build.build()
    syms = _header_symbols()
    assert len(syms) >= 15
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(dll, s), "libfsb200.so does not export " + s
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms
    assert _lib.lib().fsb_abi_version() == _lib.ABI_VERSION == 2


def test_conv_desc_layout_matches_header():
    from fasterseg_b200._lib import ConvDesc
    assert ctypes.sizeof(ConvDesc) == 18 * 4
    src = open(os.path.join(ROOT, "include", "fsb200.h")).read()
    body = src[src.index("typedef struct fsb_conv_desc {"):src.index("} fsb_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in re.findall(r"(?:int32_t|uint32_t)\s+([^;]+);", body):
        names += [n.strip() for n in decl.split(",")]
    assert names == [f[0] for f in ConvDesc._fields_]


def test_packed_bytes_is_pure_host_math():
    from fasterseg_b200 import _lib
    d = _lib.ConvDesc(1, 8, 8, 64, 64, 3, 1, 1, 1, 0, 0, 8, 8, 64, 64, 0)
    assert _lib.lib().fsb_conv_packed_bytes(ctypes.byref(d)) == 9 * 64 * 64 * 2
    d = _lib.ConvDesc(1, 8, 8, 48, 19, 1, 1, 0, 1, 0, 0, 8, 8, 48, 19, 0)
    assert _lib.lib().fsb_conv_packed_bytes(ctypes.byref(d)) == 1 * 32 * 64 * 2


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from fasterseg_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.FsbError):
        _lib.lib()


def test_native_dp_api_is_inert_until_initialised():
    """csrc/dp.cu: NCCL is resolved with dlopen on first use; before fsb_dp_init the library is single-process."""
    from fasterseg_b200 import _lib, engine
    lib = _lib.lib()
    assert lib.fsb_dp_world() == 1 and not engine.dp_native()
    assert lib.fsb_dp_allreduce_f32(None, 0, None) == 0            # no communicator: a no-op, not an error
    assert lib.fsb_dp_init(bytes(128), 0, 1) == 0 and lib.fsb_dp_world() == 1   # world 1: nothing to create
    assert lib.fsb_dp_init(bytes(128), 2, 2) != 0 and b"bad arguments" in lib.fsb_last_error_string()
    assert lib.fsb_dp_enable(0) == 0 and lib.fsb_dp_enable(1) == 0 and lib.fsb_dp_shutdown() == 0
    ident = (ctypes.c_char * 128)()
    if lib.fsb_dp_unique_id(ident) == 0:      # needs libnccl.so.2 (bundled with torch); no GPU required for the id
        assert any(bytes(ident))


def test_every_entry_point_is_documented_for_the_reference_maintainer():
    """INTEGRATION.md section 3 maps each export (or its fwd/bwd family) to the reference call site it replaces; the header itself
    cites reference file:line in the comment block above each family"""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    undocumented = []
    for s in _header_symbols():
        stem = re.sub(r"_(fwd|bwd|apply|reduce|sel|f16|f32|nchw|bytes|rows)$", "", s)
        if s not in doc and stem not in doc:
            undocumented.append(s)
    assert not undocumented, undocumented
    header = open(os.path.join(ROOT, "include", "fsb200.h")).read()
    cites = re.findall(r"[a-z_/]+\.py:\d+", header)
    assert len(cites) >= 30, "the header must cite the reference lines its entry points replace"
