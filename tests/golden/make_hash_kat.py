"""Consistency check of tests/golden/hash_kat.json: the bucket ids must follow from the
fingerprints (id = fingerprint mod num_buckets).  The vectors themselves are transcribed upstream
TensorFlow test expectations (no TensorFlow available to regenerate them)."""
import json
import os

d = json.load(open(os.path.join(os.path.dirname(__file__), "hash_kat.json")))
fp = d["fingerprint64"]
case = d["hash_bucket_strings"][0]
assert [fp[k] % case["num_buckets"] for k in case["keys"]] == case["ids"]
print("hash_kat.json is self-consistent")
