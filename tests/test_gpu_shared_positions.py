"""Shared-positions multi-k index (the reference's outlook, thesis/content/04_outlook_and_conclusion.tex:25-45):
ONE position array sorted by the largest k serves every k. Same results as the per-k index -- against the golden fixtures
of the compiled reference, the oracle, and ground truth -- from about 1 / len(ks) of the element memory."""
import numpy as np
import pytest

from conftest import assert_results_equal, golden_cases, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kb():
    import kmer_index_b200
    return kmer_index_b200


def _multi_k_golden():
    return [name for name in golden_cases() if load_golden(name)["ks"].size > 1]


@pytest.mark.parametrize("name", _multi_k_golden())
def test_shared_positions_match_golden_and_oracle(kb, oracle_mod, name):
    g = load_golden(name)
    ks = g["ks"].tolist()
    with kb.KmerIndex(g["text"], int(g["sigma"]), ks, shared_positions=True) as ix, \
            oracle_mod.Oracle(g["text"], int(g["sigma"]), ks) as o:
        got = ix.search_batch(g["q"], g["q_off"]).as_tuple()
        assert_results_equal(got, (g["r_off"], g["r_pos"], g["r_status"]), skip=g["ub"].astype(bool), label=name)
        assert_results_equal(got, o.search(g["q"], g["q_off"]), label=name + " vs oracle")
        flat, o_ = g["scheme_flat"], 0
        for m, ln, multi in zip(g["scheme_m"], g["scheme_len"], g["scheme_multi"]):
            sk, use_multi = ix.scheme(int(m))
            assert sk == flat[o_:o_ + ln].tolist() and use_multi == bool(multi), (name, int(m))
            o_ += int(ln)


def _tail_queries(text, k_max, m_lo, m_hi):
    """Every query length at every start among the last positions of the text: the k-mer starts that the largest k's
    array does not hold (a shorter k still indexes them) plus a few before them."""
    qs = []
    n = text.size
    for m in range(m_lo, m_hi + 1):
        for start in range(max(0, n - m - k_max - 2), n - m + 1):
            qs.append(text[start:start + m])
    off = np.zeros(len(qs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([q.size for q in qs])
    return np.concatenate(qs).astype(np.uint8), off


CASES = [
    # (label, sigma, ks, n, Q, m_lo, m_hi)
    ("c3", 4, [5, 7, 9, 11, 13], 2_000_000, 20_000, 4, 40),
    ("c3_unordered", 4, [9, 13, 5, 11, 7], 300_000, 6000, 1, 45),
    ("dna4_3_9_16", 4, [3, 9, 16], 200_000, 6000, 1, 70),
    ("dna4_mixed", 4, [9, 21], 300_000, 6000, 5, 70),          # 64-bit hashes, sparse directory + sorted hashes kept
    ("dna4_12_24", 4, [12, 24], 100_000, 6000, 5, 80),
    ("dna15_multi", 15, [10, 11, 12], 300_000, 6000, 5, 40),
    ("aa27_k9_10", 27, [9, 10], 300_000, 6000, 5, 45),         # k-mers wider than one packed window
    ("aa27_5_9_12", 27, [5, 9, 12], 50_000, 4000, 1, 40),
    ("dna5_k13_18_27", 5, [13, 18, 27], 200_000, 6000, 10, 90),
]


@pytest.mark.parametrize("label,sigma,ks,n,Q,m_lo,m_hi", CASES)
@pytest.mark.parametrize("qkind", ["random", "stress", "tail"])
def test_shared_positions_match_oracle(kb, oracle_mod, label, sigma, ks, n, Q, m_lo, m_hi, qkind):
    from kmer_index_b200 import synth
    text = synth.random_text(n, sigma, 200 + len(label))
    if qkind == "random":
        q, off = synth.random_queries(Q, m_lo, m_hi, sigma, 1234 + len(label))
    elif qkind == "stress":
        q, off = synth.stress_queries(text, Q, m_lo, m_hi, sigma, 4321 + len(label))
    else:
        q, off = _tail_queries(text, max(ks), m_lo, min(m_hi, 3 * max(ks)))
    with kb.KmerIndex(text, sigma, ks, shared_positions=True) as ix, oracle_mod.Oracle(text, sigma, ks) as o:
        want = o.search(q, off)
        got = ix.search_batch(q, off).as_tuple()
        assert_results_equal(got, want, label=f"{label}/{qkind}")
        if qkind != "random":
            assert want[1].size > 0
        # only the largest k owns arrays
        owner = int(np.argmax(ks))
        for e in range(len(ks)):
            info = ix.element_info(e)
            assert info.k == ks[e] and info.n_kmers == n - ks[e] + 1
            assert (info.device_bytes > 0) == (e == owner)


@pytest.mark.parametrize("sigma,ks,n,m_hi", [(4, [5, 7, 9, 11, 13], 200_000, 45), (4, [9, 21], 100_000, 70),
                                             (15, [6, 8], 100_000, 30), (27, [5, 9, 12], 50_000, 40)])
def test_shared_positions_correct_mode_matches_ground_truth(kb, oracle_mod, sigma, ks, n, m_hi):
    from kmer_index_b200 import synth
    text = synth.random_text(n, sigma, 31)
    text[1000:1400] = np.resize(np.array([0, 1, 1], dtype=np.uint8), 400)  # a repetitive stretch
    text[-60:] = np.resize(np.array([1, 0], dtype=np.uint8), 60)            # ... and one that runs into the end of the text
    q, off = synth.stress_queries(text, 3000, 1, m_hi, sigma, 5150)
    tq, toff = _tail_queries(text, max(ks), 1, min(m_hi, 30))
    off = np.concatenate([off, toff[1:] + off[-1]])
    q = np.concatenate([q, tq])
    with kb.KmerIndex(text, sigma, ks, mode=kb.MODE_CORRECT, shared_positions=True) as ix:
        got = ix.search_batch(q, off).as_tuple()
    assert_results_equal(got, oracle_mod.Oracle.truth(text, q, off), label=f"correct shared {ks}")


def test_shared_positions_exact_lengths_only(kb, oracle_mod):
    """A batch with nothing but lengths that ARE one of the ks: whole slabs of a view, sorted per query."""
    from kmer_index_b200 import synth
    n, ks = 1_500_000, [5, 7, 9, 11, 13]
    text = synth.random_text(n, 4, 41)
    with kb.KmerIndex(text, 4, ks, shared_positions=True) as ix, oracle_mod.Oracle(text, 4, ks) as o:
        for m in ks:
            q, off = synth.stress_queries(text, 500, m, m, 4, 50 + m)
            assert_results_equal(ix.search_batch(q, off).as_tuple(), o.search(q, off), label=f"exact m={m}")


def test_shared_positions_heavy_slabs_and_repetitive_text(kb, oracle_mod):
    """Low-entropy text: slabs of tens of thousands of positions, results that are whole slabs in non-position order."""
    from kmer_index_b200 import synth
    text = synth.low_entropy_text(9)
    ks = [5, 7, 9, 11, 13]
    q, off = synth.stress_queries(text, 4000, 1, 45, 4, 77)
    with kb.KmerIndex(text, 4, ks, shared_positions=True) as ix, oracle_mod.Oracle(text, 4, ks) as o:
        assert_results_equal(ix.search_batch(q, off).as_tuple(), o.search(q, off), label="low entropy shared")


def test_shared_positions_memory_and_save_load(kb, oracle_mod, tmp_path):
    from kmer_index_b200 import synth
    n, ks = 1_000_000, [5, 7, 9, 11, 13]
    text = synth.random_text(n, 4, 5)
    q, off = synth.stress_queries(text, 5000, 1, 45, 4, 6)
    path = str(tmp_path / "shared.kmerb200")
    with kb.KmerIndex(text, 4, ks) as plain, kb.KmerIndex(text, 4, ks, shared_positions=True) as shared:
        want = plain.search_batch(q, off).as_tuple()
        assert_results_equal(shared.search_batch(q, off).as_tuple(), want, label="shared vs per-k arrays")
        per_k = sum(plain.element_info(e).device_bytes for e in range(len(ks)))
        one = sum(shared.element_info(e).device_bytes for e in range(len(ks)))
        assert one == plain.element_info(ks.index(13)).device_bytes
        assert one * 2 < per_k          # five position arrays and five directories against one of each
        shared.save(path)
        with pytest.raises(kb.KmerB200Error):
            shared.element_arrays(0)    # a view owns no arrays
    with kb.KmerIndex.load(path) as ld:
        assert ld.ks == ks
        assert_results_equal(ld.search_batch(q, off).as_tuple(), want, label="loaded shared index")
        assert ld.element_info(0).device_bytes == 0 and ld.element_info(4).device_bytes > 0
    # one k: the flag means nothing
    with kb.KmerIndex(text, 4, [12], shared_positions=True) as ix, kb.KmerIndex(text, 4, [12]) as ref:
        assert_results_equal(ix.search_batch(q, off).as_tuple(), ref.search_batch(q, off).as_tuple(), label="single k")
        assert ix.element_info(0).device_bytes == ref.element_info(0).device_bytes
    # not combinable with shards or parts
    with pytest.raises(kb.KmerB200Error):
        kb.KmerIndex(text[:500_000], 4, ks, shared_positions=True, n_total=n, halo=63)
    with pytest.raises(kb.KmerB200Error):
        kb.KmerIndex(text, 4, [10, 12], shared_positions=True, key_part=0, key_parts=2)
