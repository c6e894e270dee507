"""CPU: the C-ABI library loads and exports every symbol include/deepfm_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

from recommender_tensorflow_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "deepfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s


def test_binding_covers_header():
    assert set(declared_symbols()) == set(_lib.EXPORTS)


def test_version_call():
    lib = _lib.load()
    assert b"sm_100a" in lib.dfm_version()


def test_create_rejects_bad_config_without_gpu():
    # argument validation happens before any CUDA call (same messages as trainers/deep_fm.py:31-34)
    lib = _lib.load()
    cfg = _lib.Config()
    cfg.n_cat = 0
    cfg.n_num = 0
    h = ctypes.c_void_p()
    assert lib.dfm_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"At least 1 feature column" in lib.dfm_last_error(None)
