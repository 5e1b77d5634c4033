"""Seeded synthetic texts and query sets.

Counter-based SplitMix64, so the same (seed, index) gives the same symbol whether it is produced
here with numpy or by the device generator kernel (csrc/synth.cu). The reference's own generator
(benchmarks/input_generator.hpp:52-63) draws uniform ranks from std::mt19937 through
std::uniform_int_distribution, whose stream is libstdc++-specific; only its role is kept:
i.i.d. uniform ranks in [0, sigma).
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """SplitMix64 finaliser over a uint64 array of counters (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _stream(seed: int, start: int, count: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = np.uint64((seed * 0xD1342543DE82EF95 + 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF)
        ctr = np.arange(start, start + count, dtype=np.uint64) + base
    return splitmix64(ctr)


def uniform_below(seed: int, start: int, count: int, bound: int) -> np.ndarray:
    """count values in [0, bound) from stream `seed` at counters start.. (top-32-bit multiply-shift)."""
    hi = _stream(seed, start, count) >> np.uint64(32)
    return (hi * np.uint64(bound)) >> np.uint64(32)


def random_text(n: int, sigma: int, seed: int, chunk: int = 1 << 24) -> np.ndarray:
    """n i.i.d. uniform ranks in [0, sigma) as uint8 (one byte per symbol, like seqan3 alphabets)."""
    out = np.empty(n, dtype=np.uint8)
    for s in range(0, n, chunk):
        c = min(chunk, n - s)
        out[s:s + c] = uniform_below(seed, s, c, sigma).astype(np.uint8)
    return out


def random_lengths(Q: int, m_lo: int, m_hi: int, seed: int) -> np.ndarray:
    return (uniform_below(seed ^ 0x5EED, 0, Q, m_hi - m_lo + 1) + np.uint64(m_lo)).astype(np.uint64)


def offsets_from_lengths(lens: np.ndarray) -> np.ndarray:
    off = np.zeros(lens.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return off


def random_queries(Q: int, m_lo: int, m_hi: int, sigma: int, seed: int):
    """BASELINE query set: Q independent uniform random queries, lengths uniform in [m_lo, m_hi]
    (as the reference's benchmarks draw them, benchmarks/just_k/main.cpp:60)."""
    lens = random_lengths(Q, m_lo, m_hi, seed)
    off = offsets_from_lengths(lens)
    ranks = random_text(int(off[-1]), sigma, seed ^ 0xC0FFEE)
    return ranks, off


def stress_queries(text: np.ndarray, Q: int, m_lo: int, m_hi: int, sigma: int, seed: int,
                   low_sigma: int | None = None):
    """Parity-stress set: ~50 % windows sampled from the text (a share of them ending exactly at n),
    ~25 % uniform random, ~25 % low-period. Random queries longer than log_sigma(n) almost never
    occur in the text, so the reference's own protocol (test_main.cpp:32-34) never sees non-empty
    multi-part results; this set does."""
    n = int(text.size)
    lens = random_lengths(Q, m_lo, m_hi, seed)
    lens = np.minimum(lens, np.uint64(n))
    off = offsets_from_lengths(lens)
    kind = uniform_below(seed ^ 0xA11CE, 0, Q, 8)
    r1 = uniform_below(seed ^ 0xB0B, 0, Q, 1 << 30)
    r2 = uniform_below(seed ^ 0xD00D, 0, Q, 1 << 30)
    ranks = random_text(int(off[-1]), low_sigma or sigma, seed ^ 0xFACADE)
    for i in range(Q):
        m = int(lens[i])
        o = int(off[i])
        kd = int(kind[i])
        if kd < 4:                                   # sampled from the text
            if kd == 0 and (int(r2[i]) & 1):
                start = n - m - (int(r1[i]) % 3)     # ending at / just before n
                start = max(start, 0)
            else:
                start = int(r1[i]) % (n - m + 1)
            ranks[o:o + m] = text[start:start + m]
        elif kd < 6:                                 # low period (1..4)
            period = 1 + int(r1[i]) % 4
            ranks[o:o + m] = np.resize(ranks[o:o + period], m)
        # else: uniform random (already filled)
    return ranks, off


def low_entropy_text(seed: int, sigma: int = 4) -> np.ndarray:
    """The low-entropy text of SURVEY.md Appendix B: random over ranks {0,1}, then period-6, period-2
    and constant stretches. Buckets become long and multi-part paths return non-empty results."""
    a = random_text(6000, 2, seed)
    i = np.arange(3000)
    b = ((i // 3) & 1).astype(np.uint8)
    c = (i & 1).astype(np.uint8)
    d = np.zeros(2000, dtype=np.uint8)
    t = np.concatenate([a, b, c, d])
    return np.minimum(t, sigma - 1).astype(np.uint8)
