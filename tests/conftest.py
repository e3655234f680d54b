import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_HAVE_DEVICE = None


def have_device():
    """True if libmirfold can open a CUDA device.  A missing library is NOT "no device": that must fail loudly."""
    global _HAVE_DEVICE
    if _HAVE_DEVICE is None:
        import mir_prefer_b200 as mp
        from mir_prefer_b200 import _lib
        assert os.path.exists(_lib.LIB_PATH), "libmirfold.so missing; run python __graft_entry__.py (no CPU fallback)"
        try:
            mp.MirFold().close()
            _HAVE_DEVICE = True
        except mp.MirfoldError as e:
            if e.code != -1:     # anything but MIRFOLD_ERR_NO_DEVICE is a real failure
                raise
            _HAVE_DEVICE = False
    return _HAVE_DEVICE


@pytest.fixture(autouse=True)
def _gpu_tests_need_a_device(request):
    if request.node.get_closest_marker("gpu") is not None and not have_device():
        pytest.skip("no CUDA device (libmirfold has no CPU path)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def mf():
    """Session-wide libmirfold context.  GPU tests must fail loudly if the extension is missing."""
    import mir_prefer_b200 as mp
    from mir_prefer_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libmirfold.so missing; run python __graft_entry__.py (no CPU fallback)"
    if not have_device():
        pytest.skip("no CUDA device (libmirfold has no CPU path)")
    ctx = mp.MirFold()
    yield ctx
    ctx.close()


def golden_cases():
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".out"):
            name, ltag, _ = f.rsplit(".", 2)
            out.append((name, int(ltag[1:])))
    return out
