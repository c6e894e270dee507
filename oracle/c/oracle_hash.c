/*
 * ORACLE (test infrastructure, NOT product code): CPU restatement of the integer
 * side of the reference's input path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * What it restates
 * ----------------
 * The reference (trainers/ml_100k.py:19-35) declares its categorical columns with
 * tf.feature_column.categorical_column_with_hash_bucket / bucketized_column /
 * categorical_column_with_vocabulary_list / categorical_column_with_identity.  The
 * arithmetic behind them lives in TensorFlow 1.12 (environment.yml:11), which is NOT
 * vendored under /root/reference and is not installable here:
 *   - string_to_hash_bucket_fast(s, N) = farmhash::Fingerprint64(s) mod N
 *       (tensorflow/core/kernels/string_to_hash_bucket_op.h, core/platform/default/fingerprint.h;
 *        FarmHash 1.1 `farmhashna::Hash64`, published algorithm restated below);
 *   - integer keys are first rendered by AsString (decimal, '-' for negatives);
 *   - Bucketize = std::upper_bound over the float32 boundaries;
 *   - vocabulary lookup = index in list, else vocab_size + Fingerprint64 mod num_oov.
 *
 * Pinning: checked (tests/test_oracle_hash.py) against the upstream TensorFlow
 * known-answer vectors listed in SURVEY.md §8c (key lengths 1..16).  Key lengths
 * > 16 bytes have no recalled vector: "KAT-unpinned", cross-checked only against the
 * independent pure-Python restatement in oracle/farmhash.py.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>

static const uint64_t k0 = 0xc3a5c85c97cb3127ULL;
static const uint64_t k1 = 0xb492b66fbe98f273ULL;
static const uint64_t k2 = 0x9ae16a3b2f90404fULL;

static inline uint64_t fetch64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint64_t fetch32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rot(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
static inline uint64_t shiftmix(uint64_t v) { return v ^ (v >> 47); }

static inline uint64_t hashlen16(uint64_t u, uint64_t v, uint64_t mul) {
    uint64_t a = (u ^ v) * mul;
    a ^= (a >> 47);
    uint64_t b = (v ^ a) * mul;
    b ^= (b >> 47);
    b *= mul;
    return b;
}

static uint64_t hashlen0to16(const uint8_t *s, size_t len) {
    if (len >= 8) {
        uint64_t mul = k2 + len * 2;
        uint64_t a = fetch64(s) + k2;
        uint64_t b = fetch64(s + len - 8);
        uint64_t c = rot(b, 37) * mul + a;
        uint64_t d = (rot(a, 25) + b) * mul;
        return hashlen16(c, d, mul);
    }
    if (len >= 4) {
        uint64_t mul = k2 + len * 2;
        uint64_t a = fetch32(s);
        return hashlen16(len + (a << 3), fetch32(s + len - 4), mul);
    }
    if (len > 0) {
        uint8_t a = s[0], b = s[len >> 1], c = s[len - 1];
        uint32_t y = (uint32_t)a + ((uint32_t)b << 8);
        uint32_t z = (uint32_t)len + ((uint32_t)c << 2);
        return shiftmix(y * k2 ^ z * k0) * k2;
    }
    return k2;
}

static uint64_t hashlen17to32(const uint8_t *s, size_t len) {
    uint64_t mul = k2 + len * 2;
    uint64_t a = fetch64(s) * k1;
    uint64_t b = fetch64(s + 8);
    uint64_t c = fetch64(s + len - 8) * mul;
    uint64_t d = fetch64(s + len - 16) * k2;
    return hashlen16(rot(a + b, 43) + rot(c, 30) + d, a + rot(b + k2, 18) + c, mul);
}

static uint64_t hashlen33to64(const uint8_t *s, size_t len) {
    uint64_t mul = k2 + len * 2;
    uint64_t a = fetch64(s) * k2;
    uint64_t b = fetch64(s + 8);
    uint64_t c = fetch64(s + len - 8) * mul;
    uint64_t d = fetch64(s + len - 16) * k2;
    uint64_t y = rot(a + b, 43) + rot(c, 30) + d;
    uint64_t z = hashlen16(y, a + rot(b + k2, 18) + c, mul);
    uint64_t e = fetch64(s + 16) * mul;
    uint64_t f = fetch64(s + 24);
    uint64_t g = (y + fetch64(s + len - 32)) * mul;
    uint64_t h = (z + fetch64(s + len - 24)) * mul;
    return hashlen16(rot(e + f, 43) + rot(g, 30) + h, e + rot(f + a, 18) + g, mul);
}

typedef struct { uint64_t first, second; } u128;

static inline u128 weak32(const uint8_t *s, uint64_t a, uint64_t b) {
    uint64_t w = fetch64(s), x = fetch64(s + 8), y = fetch64(s + 16), z = fetch64(s + 24);
    a += w;
    b = rot(b + a + z, 21);
    uint64_t c = a;
    a += x;
    a += y;
    b += rot(a, 44);
    u128 r = { a + z, b + c };
    return r;
}

uint64_t oracle_fingerprint64(const uint8_t *s, size_t len) {
    const uint64_t seed = 81;
    if (len <= 32) return len <= 16 ? hashlen0to16(s, len) : hashlen17to32(s, len);
    if (len <= 64) return hashlen33to64(s, len);
    uint64_t x = seed;
    uint64_t y = seed * k1 + 113;
    uint64_t z = shiftmix(y * k2 + 113) * k2;
    u128 v = {0, 0}, w = {0, 0};
    x = x * k2 + fetch64(s);
    const uint8_t *end = s + ((len - 1) / 64) * 64;
    const uint8_t *last64 = end + ((len - 1) & 63) - 63;
    do {
        x = rot(x + y + v.first + fetch64(s + 8), 37) * k1;
        y = rot(y + v.second + fetch64(s + 48), 42) * k1;
        x ^= w.second;
        y += v.first + fetch64(s + 40);
        z = rot(z + w.first, 33) * k1;
        v = weak32(s, v.second * k1, x + w.first);
        w = weak32(s + 32, z + w.second, y + fetch64(s + 16));
        uint64_t t = z; z = x; x = t;
        s += 64;
    } while (s != end);
    uint64_t mul = k1 + ((z & 0xff) << 1);
    s = last64;
    w.first += ((len - 1) & 63);
    v.first += w.first;
    w.first += v.first;
    x = rot(x + y + v.first + fetch64(s + 8), 37) * mul;
    y = rot(y + v.second + fetch64(s + 48), 42) * mul;
    x ^= w.second * 9;
    y += v.first * 9 + fetch64(s + 40);
    z = rot(z + w.first, 33) * mul;
    v = weak32(s, v.second * mul, x + w.first);
    w = weak32(s + 32, z + w.second, y + fetch64(s + 16));
    { uint64_t t = z; z = x; x = t; }
    return hashlen16(hashlen16(v.first, w.first, mul) + shiftmix(y) * k0 + z,
                     hashlen16(v.second, w.second, mul) + x, mul);
}

/* hashed column over an Arrow-style string column; '' -> -1 (empty bag) */
void oracle_hash_bucket_strings(const uint8_t *data, const int32_t *offsets, int64_t n,
                                uint64_t num_buckets, int32_t *out) {
    for (int64_t i = 0; i < n; ++i) {
        size_t len = (size_t)(offsets[i + 1] - offsets[i]);
        out[i] = len == 0 ? -1
                          : (int32_t)(oracle_fingerprint64(data + offsets[i], len) % num_buckets);
    }
}

/* hashed column over int32 keys: AsString (decimal) then Fingerprint64 mod N; -1 -> empty */
void oracle_hash_bucket_int32(const int32_t *keys, int64_t n, uint64_t num_buckets, int32_t *out) {
    char buf[16];
    for (int64_t i = 0; i < n; ++i) {
        if (keys[i] == -1) { out[i] = -1; continue; }
        int len = snprintf(buf, sizeof buf, "%d", keys[i]);
        out[i] = (int32_t)(oracle_fingerprint64((const uint8_t *)buf, (size_t)len) % num_buckets);
    }
}

/* Bucketize: number of boundaries <= x (std::upper_bound), x already float32 */
void oracle_bucketize_f32(const float *x, int64_t n, const float *bounds, int nb, int32_t *out) {
    for (int64_t i = 0; i < n; ++i) {
        int c = 0;
        for (int j = 0; j < nb; ++j) c += (bounds[j] <= x[i]);
        out[i] = c;
    }
}
