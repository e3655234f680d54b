"""CPU: the C-ABI library loads and exports every symbol include/mirfold.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mirfold.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mirfold_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_all_header_symbols():
    from mir_prefer_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build first: python __graft_entry__.py"
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.EXPORTS) == syms


def test_struct_layouts_match_header():
    from mir_prefer_b200 import _lib
    assert ctypes.sizeof(_lib.Hit) == 24
    assert ctypes.sizeof(_lib.DuplexQuery) == 40
    assert ctypes.sizeof(_lib.DuplexVerdict) == 48
    assert ctypes.sizeof(_lib.Stats) == 7 * 8 + 6 * 8 + 2 * 4 + 8
    assert _lib.Result.stats.offset == 64


def test_no_device_fails_loudly_not_silently():
    """Without a GPU mirfold_open must return MIRFOLD_ERR_NO_DEVICE -- there is no CPU fallback."""
    import torch
    import mir_prefer_b200 as mp
    if torch.cuda.is_available():
        return
    try:
        mp.MirFold()
    except mp.MirfoldError as e:
        assert e.code == -1
    else:
        raise AssertionError("MirFold() succeeded without a CUDA device")


def test_strerror_and_version():
    from mir_prefer_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.mirfold_version()
    assert lib.mirfold_strerror(0) == b"ok"
    assert b"fallback" in lib.mirfold_strerror(-1)
