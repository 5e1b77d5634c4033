"""CPU (no GPU): the C-ABI library loads and exports every declared symbol, the host-side logic (scheme table,
fast_pow, hash, choose_best_k, shard geometry) matches the golden vectors and the oracle, and computing entry
points fail loudly without a device."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_cases, load_golden


@pytest.fixture(scope="module")
def kb():
    from kmer_index_b200 import build
    build.build()
    import kmer_index_b200
    return kmer_index_b200


def test_library_exports_every_declared_symbol(kb):
    from kmer_index_b200 import _capi
    header = open(os.path.join(ROOT, "include", "kmer_b200.h")).read()
    declared = set(re.findall(r"\b(kmer_b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"kmer_b200_config_default"} - {"kmer_b200_config_default"}  # keep all
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = sorted(declared - exported)
    assert not missing, f"declared in include/kmer_b200.h but not exported: {missing}"
    assert declared == set(_capi.SYMBOLS), sorted(declared ^ set(_capi.SYMBOLS))
    assert _capi.lib().kmer_b200_abi_version() == 2


def test_no_oracle_or_cpu_fallback_in_product(kb):
    """The product never imports/links the oracle, and there is no torch/numpy compute fallback."""
    for root, _, files in os.walk(os.path.join(ROOT, "kmer_index_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("intersection oracle", ""), f"{f} mentions the oracle"
    out = subprocess.run(["ldd", kb._capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "kmer_oracle" not in out and "kmer_ref" not in out


def test_compute_fails_loudly_without_gpu(kb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(kb.KmerB200Error) as e:
        kb.KmerIndex(np.zeros(100, np.uint8), 4, [5])
    assert e.value.code == -2


def test_fast_pow_hash_choose_best_k(kb, oracle_mod):
    s = load_golden("scalars")
    for bi, b in enumerate(s["bases"]):
        for ei, e in enumerate(s["exps"]):
            assert kb.fast_pow(int(b), int(e)) == int(s["fast_pow"][bi, ei]), (b, e)
    o_ = 0
    for ln, want in zip(s["cbk_intervals"], s["cbk_out"]):
        iv = s["cbk_flat"][o_:o_ + ln]
        o_ += int(ln)
        assert kb.choose_best_k(iv.tolist(), 4) == want.tolist()
    rng = np.random.default_rng(5)
    for sigma, k in [(4, 12), (4, 16), (15, 8), (27, 5), (5, 13), (4, 31), (27, 13)]:
        r = rng.integers(0, sigma, k, dtype=np.uint8)
        assert kb.kmer_hash(r, sigma) == oracle_mod.Oracle.hash(r, sigma)


@pytest.mark.parametrize("name", golden_cases())
def test_scheme_table_matches_golden(kb, name):
    g = load_golden(name)
    flat, o_ = g["scheme_flat"], 0
    for m, ln, multi in zip(g["scheme_m"], g["scheme_len"], g["scheme_multi"]):
        ks, use_multi = kb.scheme_for_ks(g["ks"].tolist(), int(m))
        assert ks == flat[o_:o_ + ln].tolist() and use_multi == bool(multi), (name, int(m), ks)
        o_ += int(ln)


@pytest.mark.parametrize("ks", [[12], [5, 7, 9, 11, 13], [9, 10], [3], [16], [10, 11, 12], [29, 27, 25, 23], [13, 5, 11, 7, 9]])
def test_scheme_table_matches_oracle_full_range(kb, oracle_mod, ks):
    text = np.zeros(64, dtype=np.uint8)
    with oracle_mod.Oracle(text, 4, ks) as o:
        for m in list(range(1, 400)) + [999, 5000, 9998, 9999]:
            assert kb.scheme_for_ks(ks, m) == o.scheme(m), (ks, m)


def test_shard_geometry():
    from kmer_index_b200 import sharded
    n, world, halo = 1000, 4, 63
    shards = [sharded.shard_range(n, world, r, halo) for r in range(world)]
    assert shards[0].begin == 0 and shards[-1].end == n and shards[-1].halo == 0
    for a, b in zip(shards, shards[1:]):
        assert a.end == b.begin and a.halo == min(halo, n - a.end) and a.length == a.end - a.begin + a.halo
    assert sum(s.end - s.begin for s in shards) == n


def _pack_stream_reference(ranks: np.ndarray, bits: int) -> np.ndarray:
    """symbol s in bits [64 - bits (s % spw + 1), 64 - bits (s % spw)) of word s // spw"""
    spw = 64 // bits
    n_words = (ranks.size + spw - 1) // spw
    padded = np.zeros(n_words * spw, dtype=np.uint64)
    padded[:ranks.size] = ranks
    shifts = (np.uint64(64) - np.uint64(bits) * (np.arange(spw, dtype=np.uint64) + np.uint64(1)))
    return np.bitwise_or.reduce(padded.reshape(n_words, spw) << shifts, axis=1) if n_words else np.zeros(0, dtype=np.uint64)


@pytest.mark.parametrize("sigma,bits", [(4, 2), (3, 2), (5, 4), (16, 4), (27, 8), (200, 8)])
def test_host_stream_pack_matches_the_bit_layout(kb, sigma, bits):
    """The host side of large kmer_b200_search_batch calls: 1-byte ranks -> b-bit MSB-first words, query boundaries
    ignored (csrc/host_pack.cpp, AVX2 for 2 and 4 bits). Every length class around the vector loops' steps, unaligned
    sources, and the rank check."""
    rng = np.random.default_rng(5)
    sizes = [0, 1, 7, 31, 32, 33, 127, 128, 129, 511, 4096, 4097, 100_003, 1_000_001]
    for n in sizes:
        buf = rng.integers(0, sigma, n + 3, dtype=np.uint8)
        for shift in (0, 3):
            ranks = buf[shift:shift + n]
            got = kb.host_pack_stream(ranks, sigma)
            assert got.size == (n + 64 // bits - 1) // (64 // bits)
            assert np.array_equal(got, _pack_stream_reference(ranks, bits)), (sigma, n, shift)
    if sigma < 256:
        for n, at in [(1, 0), (200, 199), (4096, 77), (100_003, 100_002), (100_003, 50_000)]:
            bad = rng.integers(0, sigma, n, dtype=np.uint8)
            bad[at] = sigma
            with pytest.raises(kb.KmerB200Error) as e:
                kb.host_pack_stream(bad, sigma)
            assert e.value.code == -4


def test_host_pool_serves_several_callers_at_once(kb):
    """The library's host thread pool is shared by every caller in the process (the devices of a multi-device handle search
    their slices of a batch at the same time, each with its own producer thread): concurrent kmer_b200_host_pack_stream
    calls must neither deadlock nor mix their batches up."""
    import threading
    rng = np.random.default_rng(11)
    inputs = [rng.integers(0, 4, 3_000_017 + 1000 * t, dtype=np.uint8) for t in range(6)]
    want = [_pack_stream_reference(x, 2) for x in inputs]
    errors = []

    def worker(t):
        try:
            for _ in range(8):
                got = kb.host_pack_stream(inputs[t], 4)
                if not np.array_equal(got, want[t]):
                    errors.append(f"caller {t}: wrong words")
                    return
        except Exception as e:   # noqa: BLE001
            errors.append(f"caller {t}: {e!r}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(len(inputs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
    assert not any(th.is_alive() for th in threads), "a caller is stuck in the pool"
    assert not errors, errors
