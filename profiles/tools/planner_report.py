#!/usr/bin/env python
"""which plan answers which query length, and what it costs on the GPU path.

For BASELINE config 3 (multi_kmer_index<dna4,{5,7,9,11,13}>, 100 Mbp) and config 2 (kmer_index<dna4,12>, 100 Mbp):
the plan table of both modes (kmer_b200_plan_table: seed element, lookups, expected candidates from the bucket
statistics measured on the index) next to the measured time of a batch of queries of exactly that length
(half random, half windows of the text), device-resident, count pass + write pass. The reference's "bad" lengths
(m = n k + 1, thesis/content/03_measuring_performance.tex:60-62) are the rows to look at.
Writes markdown to stdout."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch

    import kmer_index_b200 as kb
    from kmer_index_b200 import synth
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    n, Q_MAX = 100_000_000, 200_000
    text = synth.random_text(n, 4, 205)
    d_text = torch.from_numpy(text).to(dev)
    for label, ks, lens in (("config 3: multi_kmer_index<dna4,{5,7,9,11,13}>", [5, 7, 9, 11, 13], list(range(4, 41))),
                            ("config 2: kmer_index<dna4,12>", [12], [12, 13, 23, 24, 25, 35, 36, 37, 47, 48, 49, 60, 61, 72, 73, 96, 97, 100])):
        print(f"\n### {label}, {n // 10**6} Mbp, up to {Q_MAX} queries per length (half random, half text windows; fewer where a "
              "query has ~n / 4^m hits, so that the batch's positions stay below 4e8)\n")
        print("| m | queries | reference-exact plan | seed k | lookups | exp. candidates | us / query | CORRECT plan | seed k | exp. candidates | us / query | hits equal |")
        print("|---|---|---|---|---|---|---|---|---|---|---|---|")
        with kb.KmerIndex(None, 4, ks, stream=stream.cuda_stream, text_device_ptr=d_text.data_ptr(), n=n) as ix:
            ref_rows = {r["m"]: r for r in ix.plan_table(min(lens), max(lens), mode=kb.MODE_REFERENCE_EXACT)}
            cor_rows = {r["m"]: r for r in ix.plan_table(min(lens), max(lens), mode=kb.MODE_CORRECT)}
            for m in lens:
                Q = max(1000, min(Q_MAX, int(4e8 / (n / 4.0 ** m + 1))))
                q1, off = synth.random_queries(Q // 2, m, m, 4, 1000 + m)
                starts = synth.uniform_below(77 + m, 0, Q - Q // 2, n - m).astype(np.int64)
                q2 = text[(starts[:, None] + np.arange(m)[None, :]).ravel()]
                q = np.concatenate([q1, q2])
                off = np.arange(0, (Q + 1) * m, m, dtype=np.uint64)[:Q + 1]
                d_q = torch.from_numpy(q).to(dev)
                d_off = torch.from_numpy(off.view(np.int64)).to(dev)
                out = {}
                for mode in (kb.MODE_REFERENCE_EXACT, kb.MODE_CORRECT):
                    ts = []
                    hits = 0
                    for rep in range(4):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        res = ix.search_batch_device(d_q.data_ptr(), d_off.data_ptr(), Q, m, mode=mode)
                        e1.record(stream)
                        torch.cuda.synchronize()
                        hits = res.n_positions
                        res.free()
                        if rep:
                            ts.append(e0.elapsed_time(e1))
                    out[mode] = (float(np.median(ts)), hits)
                r, c = ref_rows[m], cor_rows[m]
                print(f"| {m} | {Q} | {r['kind']} | {r['seed_k']} | {r['lookups']} | {r['candidates']:.3g} | {out[0][0] * 1e3 / Q:.4f} | "
                      f"{c['kind']} | {c['seed_k']} | {c['candidates']:.3g} | {out[1][0] * 1e3 / Q:.4f} | {'yes' if out[0][1] == out[1][1] else 'no: %d vs %d' % (out[0][1], out[1][1])} |")


if __name__ == "__main__":
    main()
