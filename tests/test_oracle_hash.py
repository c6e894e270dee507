"""CPU: pin the oracle's integer path on the upstream-TF known-answer vectors and cross-check the
two independent restatements (C and pure Python) of FarmHash Fingerprint64."""
import ctypes
import json
import os

import numpy as np

from oracle import clib, transforms
from oracle.farmhash import fingerprint64

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hash_kat.json")))


def c_fp(b):
    return int(clib().oracle_fingerprint64(b, len(b)))


def test_fingerprint_kat():
    for k, v in KAT["fingerprint64"].items():
        assert fingerprint64(k.encode()) == v
        assert c_fp(k.encode()) == v


def test_hash_bucket_kat():
    for case in KAT["hash_bucket_strings"]:
        keys = np.array([k.encode() for k in case["keys"]], dtype=object)
        for use_c in (True, False):
            assert transforms.hash_strings(keys, case["num_buckets"], use_c).tolist() == case["ids"]
    for case in KAT["hash_bucket_int32"]:
        for use_c in (True, False):
            assert transforms.hash_int32(case["keys"], case["num_buckets"], use_c).tolist() == case["ids"]


def test_c_equals_python_all_lengths():
    rng = np.random.default_rng(1)
    for n in list(range(0, 140)) + [191, 192, 193, 255, 256, 257, 1000]:
        for _ in range(3):
            b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            assert c_fp(b) == fingerprint64(b), n


def test_empty_and_ignore_values():
    assert transforms.hash_strings(np.array([b"", b"x"], dtype=object), 7).tolist()[0] == -1
    assert transforms.hash_int32([-1, 5, -7], 11)[0] == -1
    assert transforms.hash_int32([-7], 11)[0] == fingerprint64(b"-7") % 11


def test_bucketize_vocab_identity():
    b = transforms.bucketize(np.array([7, 15, 24, 25, 65, 73], np.int32), list(range(15, 66, 10)))
    assert b.tolist() == [0, 1, 1, 2, 6, 6]
    v = transforms.vocab_lookup(np.array([b"F", b"M", b"X", b""], dtype=object), ["F", "M"], 1)
    assert v.tolist() == [0, 1, 2, -1]
    assert transforms.identity([0, 1, -1], 2).tolist() == [0, 1, -1]
    try:
        transforms.identity([2], 2)
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
