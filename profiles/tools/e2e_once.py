"""One config-5 search through the host C ABI (pinned buffers): the command whose ncu launch list shows what the host
pipeline launches per chunk (align_stream_kernel, the lengths' prefix sum, the search passes).
usage: python profiles/tools/e2e_once.py [text_symbols] [queries]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import kmer_index_b200 as kb
from kmer_index_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
dev = torch.device("cuda", 0)
L = _capi.lib()
text = torch.empty(n, dtype=torch.uint8, device=dev)
_capi.check(L.kmer_b200_synth_ranks_device(text.data_ptr(), n, 0, 4, 205, None))
g = torch.Generator(device=dev)
g.manual_seed(1239)
lens = torch.randint(16, 65, (Q,), generator=g, device=dev, dtype=torch.int64)
off = torch.zeros(Q + 1, dtype=torch.int64, device=dev)
torch.cumsum(lens, 0, out=off[1:])
n_sym = int(off[-1].item())
q = torch.empty(n_sym, dtype=torch.uint8, device=dev)
_capi.check(L.kmer_b200_synth_ranks_device(q.data_ptr(), n_sym, 0, 4, 1239 ^ 0xC0FFEE, None))
torch.cuda.synchronize()
h_q, h_off = q.cpu().pin_memory(), off.cpu().pin_memory()
del q, off, lens
ix = kb.KmerIndex(None, 4, [16], text_device_ptr=text.data_ptr(), n=n)
r = ix.search_batch(h_q.numpy(), h_off.numpy().view(np.uint64), copy=False)
print("hits", int(r.offsets[-1]), ix.last_search_host_path(), ix.last_search_transfer(), flush=True)
r.free()
ix.close()
