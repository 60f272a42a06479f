"""Seeded byte texts for the wavelet-tree / FM-index parity tests.  The literal fixtures are the reference's
own test inputs (test/test_cases/*.txt, test/wt_byte_test.config, test/csa_byte_test.config), restated here
because /root/reference does not exist on the GPU box."""
import numpy as np

FIXTURES = {
    "100a.txt": b"a" * 100,
    "abc_abc_abc.txt": b"abc_abc_abc\n",
    "abc_abc_abc2.txt": b"abc abc abc\n",
    "all_symbols.txt": bytes(range(256)),
    "example01.txt": b"abracadabra\n",
    "one_byte.txt": b"\n",
}


def text_catalogue(zero_free=False, large=True):
    """yields (name, bytes).  zero_free: only texts usable for csa construction (construct.hpp:34-46)"""
    rng = np.random.default_rng(12345)
    for k, v in FIXTURES.items():
        if zero_free and 0 in v:
            continue
        yield k, v
    lo = 1 if zero_free else 0
    yield "two_symbols", rng.integers(65, 67, 5000, dtype=np.uint8).tobytes()
    yield "dna", rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 30000).tobytes()
    yield "skewed", np.clip(rng.geometric(0.3, 50000) + lo, lo, 255).astype(np.uint8).tobytes()
    yield "fibonacci", _fib(17)
    yield "runs", np.repeat(rng.integers(lo, 256, 300, dtype=np.uint8), rng.integers(1, 300, 300)).tobytes()
    yield "uniform", rng.integers(lo, 256, 100000, dtype=np.uint8).tobytes()
    if large:
        yield "uniform_3M", rng.integers(lo, 256, 3_000_000, dtype=np.uint8).tobytes()
        yield "english_like", _markov(rng, 1_000_000)


def _fib(k):
    a, b = b"a", b"ab"
    for _ in range(k):
        a, b = b, b + a
    return b


def _markov(rng, n):
    """order-1 Markov text over ~60 symbols with a skewed stationary distribution (repetitive like natural text)"""
    sym = np.frombuffer(b" etaoinshrdlcumwfgypbvkjxqz,.\nETAOINSHRDLCUMWFGYPBVK0123456789", dtype=np.uint8)
    k = len(sym)
    w = 1.0 / np.arange(1, k + 1)
    trans = np.stack([rng.permutation(w) for _ in range(k)])
    trans /= trans.sum(1, keepdims=True)
    cdf = np.cumsum(trans, 1)
    u = rng.random(n)
    out = np.empty(n, np.int64)
    s = 0
    for i in range(n):
        s = int(np.searchsorted(cdf[s], u[i]))
        if s >= k:
            s = k - 1
        out[i] = s
    return sym[out].tobytes()


def wt_queries(text, rng, nq):
    """(i, c) pairs: half the symbols drawn from the text, half uniform over 0..255 (absent symbols included)"""
    t = np.frombuffer(text, dtype=np.uint8)
    n = len(t)
    i = rng.integers(0, n + 1, nq, dtype=np.uint64)
    c = rng.integers(0, 256, nq, dtype=np.uint8)
    if n:
        c[::2] = t[rng.integers(0, n, len(c[::2]))]
    i[0], i[-1] = 0, n
    return i, c
