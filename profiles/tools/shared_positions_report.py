#!/usr/bin/env python
"""the shared-positions multi-k index against per-k arrays on BASELINE config 3
(multi_kmer_index<dna4,{5,7,9,11,13}>, 100 Mbp): device bytes of the elements, build time, and the search time of
a device-resident batch per query length (the price: results seeded from a shorter k are sorted per query) and for
the config's own mix of lengths 4-40. Checks that both indices return identical results. Markdown to stdout."""
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
    n, ks = 100_000_000, [5, 7, 9, 11, 13]
    text = synth.random_text(n, 4, 205)
    d_text = torch.from_numpy(text).to(dev)

    def timed(fn, reps=3):
        ts = []
        out = None
        for rep in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = fn()
            e1.record(stream)
            torch.cuda.synchronize()
            if rep:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), out

    def build(shared):
        return kb.KmerIndex(None, 4, ks, stream=stream.cuda_stream, text_device_ptr=d_text.data_ptr(), n=n,
                            shared_positions=shared)

    b_ms = {}
    for shared in (False, True):
        def once():
            build(shared).close()
        b_ms[shared], _ = timed(once)
    plain, shared = build(False), build(True)
    bytes_plain = sum(plain.element_info(e).device_bytes for e in range(len(ks)))
    bytes_shared = sum(shared.element_info(e).device_bytes for e in range(len(ks)))
    print(f"### config 3: multi_kmer_index<dna4,{{5,7,9,11,13}}>, {n // 10**6} Mbp, one B200\n")
    print("| | per-k arrays | shared positions |")
    print("|---|---|---|")
    print(f"| element bytes on the device (positions + directories) | {bytes_plain / 1e6:.1f} MB | {bytes_shared / 1e6:.1f} MB "
          f"({bytes_plain / bytes_shared:.2f}x less) |")
    print(f"| build, text resident | {b_ms[False]:.2f} ms | {b_ms[True]:.2f} ms |")

    def search_pair(q, off, max_len, label, count_only=False):
        Q = off.size - 1
        d_q = torch.from_numpy(q).to(dev)
        d_off = torch.from_numpy(off.view(np.int64)).to(dev)
        row = []
        fp = []
        for ix in (plain, shared):
            def once():
                fn = ix.count_batch_device if count_only else ix.search_batch_device
                return fn(d_q.data_ptr(), d_off.data_ptr(), Q, max_len)
            keep = []

            def run():
                for r in keep:
                    r.free()
                keep.clear()
                keep.append(once())
                return keep[0]
            ms, res = timed(run)
            hits = res.n_positions
            offs = torch.as_tensor(res.offsets(), device=dev)
            if count_only:
                fp.append((int(offs[-1].item()), int(offs.sum().item() & ((1 << 62) - 1)), 0))
            else:
                pos = torch.as_tensor(res.positions(), device=dev) if hits else torch.empty(0, dtype=torch.int32, device=dev)
                # order-sensitive fingerprint of the position stream: sum of pos * (rank in stream + 1) mod 2^61
                chunk, acc = 1 << 26, 0
                for c0 in range(0, pos.numel(), chunk):
                    p = pos[c0:c0 + chunk].to(torch.int64) & 0xFFFFFFFF
                    w = (torch.arange(c0, c0 + p.numel(), device=dev, dtype=torch.int64) % 1000003) + 1
                    acc = (acc + int((p * w).sum().item())) % (1 << 61)
                fp.append((int(offs[-1].item()), int(offs.sum().item() & ((1 << 62) - 1)), acc))
            row.append((ms, hits))
            for r in keep:
                r.free()
            keep.clear()
        same = "yes" if fp[0] == fp[1] else f"NO {fp}"
        print(f"| {label} | {Q} | {row[0][1]:.3e} | {row[0][0]:.2f} | {row[1][0]:.2f} | {row[1][0] / max(row[0][0], 1e-9):.1f}x | {same} |")

    print("\n| query length | queries | hits | per-k arrays, ms | shared positions, ms | ratio | identical results |")
    print("|---|---|---|---|---|---|---|")
    for m in (4, 5, 6, 7, 9, 11, 12, 13, 14, 20, 26, 33, 40):
        Q = max(1000, min(200_000, int(2e8 / (n / 4.0 ** m + 1))))
        q1, _ = synth.random_queries(Q // 2, m, m, 4, 1000 + m)
        starts = synth.uniform_below(77 + m, 0, Q - Q // 2, n - m).astype(np.int64)
        q2 = text[(starts[:, None] + np.arange(m)[None, :]).ravel()]
        q = np.concatenate([q1, q2])
        off = np.arange(0, (Q + 1) * m, m, dtype=np.uint64)[:Q + 1]
        search_pair(q, off, m, str(m))
    # the config's own mix: 10^6 random queries of lengths 4-40 (1.4e10 positions: counted, not materialised, for the
    # per-length rows above already time the write + sort), then a 10^5 sample of it materialised
    q, off = synth.random_queries(1_000_000, 4, 40, 4, 99)
    search_pair(q, off, 40, "4-40 mix, count only", count_only=True)
    q, off = synth.random_queries(20_000, 4, 40, 4, 98)
    search_pair(q, off, 40, "4-40 mix (2e4 queries), positions written")
    plain.close()
    shared.close()


if __name__ == "__main__":
    main()
