"""Seeded input catalogue shared by the parity tests, the golden-vector generator and bench.py.

Bit-vector shapes follow the list the reference *intended* to test
(/root/reference/test/rank_support_test.config:2-32 and test/bit_vector_generator.cpp:25-78; because of
the generator bug described in SURVEY.md §4 the reference really only exercises MAT-SELECT, so the shapes
are generated here).
"""
import numpy as np


def pack_bits(bits):
    """bool/0-1 array -> uint64 words, LSB first (int_vector<1> layout, int_vector.hpp:1900-1904)"""
    bits = np.asarray(bits, dtype=np.uint8)
    n = len(bits)
    pad = (-n) % 64
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, np.uint8)])
    if len(bits) == 0:
        return np.zeros(0, np.uint64)
    return np.packbits(bits, bitorder="little").view(np.uint64).copy()


def unpack_bits(words, nbits):
    return np.unpackbits(np.asarray(words, dtype=np.uint64).view(np.uint8), bitorder="little")[:nbits]


def random_words(nbits, seed, dirty_tail=False):
    """util::set_random_bits semantics (util.hpp:466-485): every word is one rng draw; the tail of the last
    word stays random when dirty_tail (SDSL leaves those bits unspecified)."""
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 2**64, (nbits + 63) // 64, dtype=np.uint64)
    if not dirty_tail and nbits % 64:
        w[-1] &= np.uint64((1 << (nbits % 64)) - 1)
    return w


def bernoulli_words(nbits, density, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros((nbits + 63) // 64, dtype=np.uint64)
    step = 1 << 24
    for lo in range(0, nbits, step):
        hi = min(nbits, lo + step)
        bits = rng.random(hi - lo) < density
        w = pack_bits(bits)
        out[lo // 64 : lo // 64 + len(w)] = w
    return out


def crafted(name):
    """-> (words, nbits) for the CRAFTED-* ids of test/bit_vector_generator.cpp"""
    rng = np.random.default_rng(abs(hash(name)) % (2**32) if False else sum(map(ord, name)))
    if name == "CRAFTED-32":
        b = np.zeros(32, np.uint8)
        b[[1, 4, 7, 18, 24, 26, 30, 31]] = 1
        return pack_bits(b), 32
    n = 1000000
    if name in ("CRAFTED-SPARSE-0", "CRAFTED-SPARSE-1"):
        dv = int(name[-1])
        b = np.full(n, dv, np.uint8)
        b[rng.integers(0, n, n // 1000)] = 1 - dv
        return pack_bits(b), n
    if name in ("CRAFTED-BLOCK-0", "CRAFTED-BLOCK-1"):
        dv = int(name[-1])
        b = np.full(n, dv, np.uint8)
        for x, ln in zip(rng.integers(0, n, n // 1000), rng.integers(0, 1000, n // 1000)):
            b[x : min(n, x + ln)] = 1 - dv
        return pack_bits(b), n
    if name == "CRAFTED-MAT-SELECT":
        b = np.zeros(n, np.uint8)
        ones = 4030 + int(rng.integers(0, 80))
        b[rng.choice(n, ones, replace=False)] = 1
        return pack_bits(b), n
    raise KeyError(name)


def bitvector_catalogue(large=True):
    """yields (case_id, words, nbits)"""
    for n, v in [(0, 0), (1, 0), (1, 1), (7, 1), (8, 0), (9, 1), (10, 0), (11, 1), (12, 0), (13, 1), (14, 0), (15, 1)]:
        yield f"const.{n}.{v}", pack_bits(np.full(n, v, np.uint8)), n
    for n, seed in [(8, 17), (16, 42), (32, 111), (64, 222), (128, 73), (256, 4887), (512, 432), (1024, 898), (2048, 5432), (4096, 793), (8192, 1043)]:
        yield f"rand.{n}.{seed}", random_words(n, seed), n
    # sizes around our own block / superblock boundaries (224-bit sector blocks, 512-bit SDSL superblocks)
    for n in (63, 65, 223, 224, 225, 447, 448, 449, 511, 513, 4095, 4097, 100000 - 1, 100000, 100001):
        yield f"rand.{n}.7", random_words(n, 7 + n, dirty_tail=True), n
    yield "CRAFTED-32", *crafted("CRAFTED-32")
    if large:
        yield "const.1000000.0", pack_bits(np.zeros(1000000, np.uint8)), 1000000
        yield "const.1000000.1", pack_bits(np.ones(1000000, np.uint8)), 1000000
        yield "rand.1000000.815", random_words(1000000, 815), 1000000
        for name in ("CRAFTED-SPARSE-0", "CRAFTED-SPARSE-1", "CRAFTED-BLOCK-0", "CRAFTED-BLOCK-1", "CRAFTED-MAT-SELECT"):
            yield name, *crafted(name)
        # long select superblocks: 4096 ones spread over more than log^4 n bits
        n = (1 << 24) + 333
        w = bernoulli_words(n, 0.004, 99)
        half = (n // 128) * 64
        w[: half // 64] = bernoulli_words(half, 0.45, 98)
        yield "mixed.2^24", w, n


def rank_queries(nbits, seed, n):
    """all positions for small vectors, else n uniform ones (always including 0 and nbits)"""
    if nbits + 1 <= n:
        return np.arange(nbits + 1, dtype=np.uint64)
    rng = np.random.default_rng(seed)
    q = rng.integers(0, nbits + 1, n, dtype=np.uint64)
    q[0], q[-1] = 0, nbits
    return q


def select_queries(m, seed, n):
    if m == 0:
        return np.zeros(0, np.uint64)
    if m <= n:
        return np.arange(1, m + 1, dtype=np.uint64)
    rng = np.random.default_rng(seed)
    q = rng.integers(1, m + 1, n, dtype=np.uint64)
    q[0], q[-1] = 1, m
    return q
