import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (built here by nvcc cross-compilation if missing)."""
    from volpick_b200 import _lib, build

    if not os.path.exists(_lib.LIB_PATH):
        build.build_library()
    return _lib.load()


@pytest.fixture(scope="session")
def oracle_c():
    import ctypes as C

    from volpick_b200 import build

    lib = C.CDLL(build.build_oracle_c())
    lib.vpo_window_starts.restype = C.c_int64
    lib.vpo_window_starts.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
    lib.vpo_stack.restype = C.c_int
    lib.vpo_stack.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                              C.c_int, C.c_void_p, C.c_int64]
    lib.vpo_picks.restype = C.c_int64
    lib.vpo_picks.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int64]
    return lib


def _state_dict(kind, name="volpick"):
    from oracle import nets
    from volpick_b200 import weights_io

    _, wpath = weights_io.find_weights(kind, name)
    return nets.state_dict_from_numpy(weights_io.load_weights(wpath))


@pytest.fixture(scope="session")
def sd_eqt():
    return _state_dict("eqtransformer")


@pytest.fixture(scope="session")
def sd_pn():
    return _state_dict("phasenet")


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    return {fn[:-4]: np.load(os.path.join(d, fn), allow_pickle=False) for fn in sorted(os.listdir(d)) if fn.endswith(".npz")}
