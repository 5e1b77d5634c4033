"""Smallest end-to-end case for compute-sanitizer (profiles/README.md): every kernel family runs once."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmer_index_b200 as kb  # noqa: E402
from kmer_index_b200 import synth  # noqa: E402

text = synth.random_text(30_011, 4, 3)
text[5000:9000] = 0                                    # a heavy bucket -> warp-per-query launch
q, off = synth.stress_queries(text, 700, 1, 70, 4, 4, low_sigma=2)
for ks, kw in (([12], {}), ([5, 7, 9, 11, 13], {}), ([20], {}), ([12], {"aux_elements": False}), ([12], {"mode": kb.MODE_CORRECT})):
    with kb.KmerIndex(text, 4, ks, **kw) as ix:
        for _ in range(2):
            r = ix.search_batch(q, off)
        print(ks, kw, "hits", r.positions.size, flush=True)
# round 2: lean count kernel (single k, dna4, dense directory), shared-positions views, a key-range part, the FASTA parser
q16, off16 = synth.stress_queries(text, 700, 16, 64, 4, 8)
with kb.KmerIndex(text, 4, [10]) as ix:
    print("lean", ix.search_batch(q16, off16).positions.size, flush=True)
with kb.KmerIndex(text, 4, [5, 7, 9, 11, 13], shared_positions=True) as ix:
    print("shared positions", ix.search_batch(q, off).positions.size, flush=True)
with kb.KmerIndex(text, 4, [10], key_part=1, key_parts=3) as ix:
    print("key-range part", ix.search_batch(q16, off16).positions.size, flush=True)
fasta = b">r1 first\n" + bytes(b"ACGT"[c] for c in text[:5000]) + b"\n>r2\n" + bytes(b"ACGT"[c] for c in text[5000:7000]) + b"\n"
with kb.parse_sequences(fasta, "dna4") as recs:
    with recs.index([12]) as ix:
        print("fasta", len(recs), ix.search_batch(q16, off16).positions.size, flush=True)
t15 = synth.random_text(20_003, 15, 5)
q15, off15 = synth.stress_queries(t15, 300, 1, 30, 15, 6)
with kb.KmerIndex(t15, 15, [8]) as ix:
    print("dna15", ix.search_batch(q15, off15).positions.size)
print("sanitize case done")
if os.environ.get("KMER_B200_GUARD"):
    # canary zones around every device allocation of the library, verified at free time (kmer_b200.h)
    seen = kb.guard_violations()
    kb.guard_selftest()                 # two deliberate out-of-bounds stores: both must be counted
    print("guard violations", seen, "selftest ok", kb.guard_violations() - seen == 2)
