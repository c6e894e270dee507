"""ORACLE (test infrastructure): feature-column transforms raw column -> int32 ids.

Restates the TF-1.12 feature-column transforms the reference declares in
trainers/ml_100k.py:18-39:

  hash        ids = Fingerprint64(key) mod N; int keys go through AsString first
              (`_HashedCategoricalColumn`), '' / -1 -> empty bag (-1 here)
  bucketized  cast to float32, id = #{boundaries <= x}   (Bucketize / std::upper_bound)
  vocab       index in list, else vocab_size + Fingerprint64 mod num_oov; '' -> empty
  identity    id = value; -1 -> empty; outside [0, N) is a runtime error in TF

A column spec is a plain dict:
  {"name", "kind": "hash"|"bucketized"|"vocab"|"identity", "source": raw feature key,
   "dtype": "int32"|"float32"|"string", "num_buckets", "boundaries", "vocab", "num_oov"}
String columns are numpy object arrays of `bytes`.
"""
import ctypes

import numpy as np

from . import clib
from .farmhash import fingerprint64


def pack_strings(values):
    """object array of bytes -> (uint8 data, int32 offsets[B+1]) Arrow-style."""
    lens = np.fromiter((len(v) for v in values), dtype=np.int64, count=len(values))
    offsets = np.zeros(len(values) + 1, dtype=np.int32)
    np.cumsum(lens, out=offsets[1:])
    data = np.frombuffer(b"".join(values), dtype=np.uint8).copy()
    if data.size == 0:
        data = np.zeros(1, dtype=np.uint8)
    return data, offsets


def unpack_strings(col):
    """(uint8 data, int32 offsets) -> list of bytes; anything else -> list of its elements."""
    if isinstance(col, tuple):
        data, offs = col
        raw = np.asarray(data, dtype=np.uint8).tobytes()
        return [raw[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    return list(col)


def hash_strings(values, num_buckets, use_c=True):
    values = unpack_strings(values)
    out = np.empty(len(values), dtype=np.int32)
    if use_c:
        data, offsets = pack_strings(values)
        clib().oracle_hash_bucket_strings(
            data.ctypes.data_as(ctypes.c_void_p), offsets.ctypes.data_as(ctypes.c_void_p),
            ctypes.c_int64(len(values)), ctypes.c_uint64(num_buckets),
            out.ctypes.data_as(ctypes.c_void_p))
    else:
        for i, v in enumerate(values):
            out[i] = -1 if len(v) == 0 else fingerprint64(v) % num_buckets
    return out


def hash_int32(keys, num_buckets, use_c=True):
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    out = np.empty(keys.shape[0], dtype=np.int32)
    if use_c:
        clib().oracle_hash_bucket_int32(
            keys.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(keys.shape[0]),
            ctypes.c_uint64(num_buckets), out.ctypes.data_as(ctypes.c_void_p))
    else:
        for i, v in enumerate(keys.tolist()):
            out[i] = -1 if v == -1 else fingerprint64(str(v).encode()) % num_buckets
    return out


def bucketize(x, boundaries):
    x = np.asarray(x).astype(np.float32)
    b = np.asarray(boundaries, dtype=np.float32)
    return (b[None, :] <= x[:, None]).sum(axis=1).astype(np.int32)


def vocab_lookup(values, vocab, num_oov):
    values = unpack_strings(values)
    table = {v if isinstance(v, bytes) else v.encode(): i for i, v in enumerate(vocab)}
    out = np.empty(len(values), dtype=np.int32)
    for i, v in enumerate(values):
        if len(v) == 0:
            out[i] = -1
        elif v in table:
            out[i] = table[v]
        elif num_oov > 0:
            out[i] = len(vocab) + fingerprint64(v) % num_oov
        else:
            out[i] = -1  # default_value=-1 -> dropped
    return out


def identity(x, num_buckets):
    x = np.asarray(x).astype(np.int64)
    bad = (x != -1) & ((x < 0) | (x >= num_buckets))
    if bad.any():
        raise ValueError("identity column value out of range [0, %d)" % num_buckets)
    return x.astype(np.int32)


def transform_column(spec, features, use_c=True):
    raw = features[spec.get("source", spec["name"])]
    kind = spec["kind"]
    if kind == "hash":
        if spec.get("dtype", "string") == "string":
            return hash_strings(raw, spec["num_buckets"], use_c)
        return hash_int32(raw, spec["num_buckets"], use_c)
    if kind == "bucketized":
        return bucketize(raw, spec["boundaries"])
    if kind == "vocab":
        return vocab_lookup(raw, spec["vocab"], spec.get("num_oov", 0))
    if kind == "identity":
        return identity(raw, spec["num_buckets"])
    raise ValueError(kind)


def transform(cat_specs, features, use_c=True):
    """-> ids [B, n_slots] int32 in the order of cat_specs (caller passes model order).  A multivalent column
    (spec["width"] = M > 1, raw feature [B, M] with -1 / '' padding) contributes M consecutive slots."""
    cols = []
    for s in cat_specs:
        w = int(s.get("width", 1))
        ids = transform_column(s, features if w == 1 else {s.get("source", s["name"]): _flat(features[s.get("source", s["name"])])}, use_c)
        cols.append(ids.reshape(-1, w))
    return np.concatenate(cols, axis=1)


def _flat(raw):
    if isinstance(raw, tuple):
        return raw
    a = np.asarray(raw)
    return a.reshape(-1)


def num_buckets(spec):
    if spec["kind"] == "bucketized":
        return len(spec["boundaries"]) + 1
    if spec["kind"] == "vocab":
        return len(spec["vocab"]) + spec.get("num_oov", 0)
    return int(spec["num_buckets"])
