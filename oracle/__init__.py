"""ORACLE — test infrastructure only (never imported by the product package).

CPU restatement of the reference's DeepFM / wide&deep train step.
Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and --impl reference legs.

Parity status: the integer path (hashing, bucketize, vocab, identity) is pinned on upstream
TensorFlow known-answer vectors (tests/golden/hash_kat.json).  The floating-point path
(forward, backward, optimizers) is **parity unpinned**: the reference has no tests or golden
vectors, and TensorFlow 1.12 cannot be installed here, so it restates SURVEY.md Appendix A.2/A.3
and is cross-checked only against torch autograd (tests/test_oracle_model.py).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "c", "oracle_hash.c")
_SO = os.path.join(_HERE, "_build", "liboracle_hash.so")


def build_c(force=False):
    """Compile the plain-C part of the oracle with gcc (called from __graft_entry__.build())."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, _SRC])
    return _SO


_lib = None


def clib():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_fingerprint64.restype = ctypes.c_uint64
        lib.oracle_fingerprint64.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        _lib = lib
    return _lib
