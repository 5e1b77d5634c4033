import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import kmer_index_b200 as kb
from kmer_index_b200 import sharded, synth
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
n, k, parts, Q = 1_000_000_000, 16, 2, 30_000_000
text = torch.empty(n, dtype=torch.uint8, device=dev)
kb._capi.check(kb._capi.lib().kmer_b200_synth_ranks_device(text.data_ptr(), n, 0, 4, 7, None))
torch.cuda.synchronize()
q, off = synth.random_queries(Q, 16, 64, 4, 5)
d_q = torch.from_numpy(q).to(dev); d_off = torch.from_numpy(off.view(np.int64)).to(dev)
def timed(ix, label):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = ix.count_batch_device(d_q.data_ptr(), d_off.data_ptr(), Q, 64); e1.record(); torch.cuda.synchronize()
        h = r.n_positions; r.free()
    print(label, e0.elapsed_time(e1), "ms hits", h, flush=True)
whole = kb.KmerIndex(None, 4, [k], text_device_ptr=text.data_ptr(), n=n, stream=stream or None)
timed(whole, "whole index"); whole.close()
idx = [kb.KmerIndex(None, 4, [k], text_device_ptr=text.data_ptr(), n=n, key_part=r, key_parts=parts, stream=stream or None) for r in range(parts)]
ps = [ix.element_part(0) for ix in idx]
bufs, first = [], [0]
dir_full = torch.empty(4 ** k + 1, dtype=torch.int32, device=dev)
for r, (ix, p) in enumerate(zip(idx, ps)):
    buf = torch.empty(p.n_kmers, dtype=torch.int32, device=dev)
    buf.copy_(sharded._dev_view(p.d_positions, p.n_kmers, dev)); bufs.append(buf)
    ix.export_directory(0, first[-1], p.key_hi - p.key_lo + (1 if r == parts - 1 else 0), dir_full.data_ptr() + 4 * p.key_lo)
    first.append(first[-1] + p.n_kmers)
torch.cuda.synchronize()
idx[0].adopt_element_parts(0, bufs, first, dir_full)
timed(idx[0], "positions in 2 local parts")
