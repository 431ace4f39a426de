"""TEST INFRASTRUCTURE - CPU oracle.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.

FarmHash ``Fingerprint64`` (== ``farmhashna::Hash64``), the hash behind
``tf.strings.to_hash_bucket_fast`` which the reference calls at
deepctr/layers/utils.py:103-107.  TensorFlow (unpinned; CI uses 1.15.5 / 2.10 / 2.15 / 2.20) and
google/farmhash are NOT under /root/reference, so this restates the published algorithm
(farmhash.cc, namespace farmhashna).  PARITY UNPINNED for bucket values: the reference's tests
hold no hash-bucket golden vector; only the vocabulary-file path is pinned.
Pure-Python integers (arbitrary precision) masked to 64 bits - slow, small cases only.
"""
M64 = (1 << 64) - 1
K0 = 0xc3a5c85c97cb3127
K1 = 0xb492b66fbe98f273
K2 = 0x9ae16a3b2f90404f


def _rot(v, s):
    return ((v >> s) | (v << (64 - s))) & M64 if s else v


def _shift_mix(v):
    return v ^ (v >> 47)


def _f64(s, i):
    return int.from_bytes(s[i:i + 8], "little")


def _f32(s, i):
    return int.from_bytes(s[i:i + 4], "little")


def _hash_len16(u, v, mul):
    a = ((u ^ v) * mul) & M64
    a ^= a >> 47
    b = ((v ^ a) * mul) & M64
    b ^= b >> 47
    return (b * mul) & M64


def _len0to16(s):
    n = len(s)
    if n >= 8:
        mul = (K2 + n * 2) & M64
        a = (_f64(s, 0) + K2) & M64
        b = _f64(s, n - 8)
        c = (_rot(b, 37) * mul + a) & M64
        d = ((_rot(a, 25) + b) * mul) & M64
        return _hash_len16(c, d, mul)
    if n >= 4:
        mul = (K2 + n * 2) & M64
        a = _f32(s, 0)
        return _hash_len16((n + (a << 3)) & M64, _f32(s, n - 4), mul)
    if n > 0:
        a, b, c = s[0], s[n >> 1], s[n - 1]
        y = (a + (b << 8)) & 0xFFFFFFFF
        z = (n + (c << 2)) & 0xFFFFFFFF
        return (_shift_mix(((y * K2) & M64) ^ ((z * K0) & M64)) * K2) & M64
    return K2


def _len17to32(s):
    n = len(s)
    mul = (K2 + n * 2) & M64
    a = (_f64(s, 0) * K1) & M64
    b = _f64(s, 8)
    c = (_f64(s, n - 8) * mul) & M64
    d = (_f64(s, n - 16) * K2) & M64
    return _hash_len16((_rot((a + b) & M64, 43) + _rot(c, 30) + d) & M64,
                       (a + _rot((b + K2) & M64, 18) + c) & M64, mul)


def _len33to64(s):
    n = len(s)
    mul = (K2 + n * 2) & M64
    a = (_f64(s, 0) * K2) & M64
    b = _f64(s, 8)
    c = (_f64(s, n - 8) * mul) & M64
    d = (_f64(s, n - 16) * K2) & M64
    y = (_rot((a + b) & M64, 43) + _rot(c, 30) + d) & M64
    z = _hash_len16(y, (a + _rot((b + K2) & M64, 18) + c) & M64, mul)
    e = (_f64(s, 16) * mul) & M64
    f = _f64(s, 24)
    g = ((y + _f64(s, n - 32)) * mul) & M64
    h = ((z + _f64(s, n - 24)) * mul) & M64
    return _hash_len16((_rot((e + f) & M64, 43) + _rot(g, 30) + h) & M64,
                       (e + _rot((f + a) & M64, 18) + g) & M64, mul)


def fingerprint64(s):
    """s: bytes (len <= 64 supported; longer strings are outside the hot path)."""
    if isinstance(s, str):
        s = s.encode("utf-8")
    n = len(s)
    if n <= 16:
        return _len0to16(s)
    if n <= 32:
        return _len17to32(s)
    if n <= 64:
        return _len33to64(s)
    raise NotImplementedError("fingerprint64: strings longer than 64 bytes are not restated")


def as_string(x):
    """tf.as_string for ints / passthrough for str/bytes (deepctr/layers/utils.py:91-95)."""
    if isinstance(x, bytes):
        return x
    if isinstance(x, str):
        return x.encode("utf-8")
    return str(int(x)).encode("ascii")


def hash_bucket(x, num_buckets, mask_zero=False):
    """deepctr/layers/utils.py:101-110."""
    s = as_string(x)
    nb = num_buckets - 1 if mask_zero else num_buckets
    h = fingerprint64(s) % nb
    if mask_zero:
        return 0 if s == b"0" else h + 1
    return h
