"""Multi-GPU layer: one process per GPU, text sharded by position range with a halo.

Shard r owns the match starts [begin, end) and additionally holds `halo` symbols of the next shard so
matches (and k-mers) that start inside the shard can be completed locally. Every rank searches ALL
queries on its shard ("queries are broadcast"); the only cross-rank steps are

  1. presence: the reference decides "a part is absent => empty result" and "rest too short => throw"
     on the WHOLE text (kmer_index.hpp:216-227, :119-122), so the per-shard presence masks are OR-ed
     across ranks (all_gather + fold) before the candidate phase, and
  2. result merge: shards are disjoint ascending position ranges and each per-shard list is ascending,
     so the merged sorted list of a query is the concatenation of the shard lists in rank order. Counts
     are gathered, rank 0 computes the write offsets and places every shard's payload.

torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing; the search itself is
libkmer_b200.so on every rank.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    begin: int     # global position of the first owned symbol
    end: int       # one past the last owned match start
    length: int    # symbols held locally = (end - begin) + halo
    halo: int      # symbols borrowed from the next shard (0 on the last shard)


def shard_range(n: int, world: int, rank: int, halo: int) -> Shard:
    """Position-range partition of a text of n symbols; the halo is clipped at the end of the text."""
    per = -(-n // world)
    begin = min(rank * per, n)
    end = min(begin + per, n)
    h = 0 if rank == world - 1 else min(halo, n - end)
    return Shard(begin, end, (end - begin) + h, h)


# ----------------------------------------------------------------------------------------------------------
# pure tensor logic (device-agnostic: exercised on CPU with gloo in tests/test_sharded_cpu.py)
# ----------------------------------------------------------------------------------------------------------
def fold_presence(gathered):
    """OR over ranks of the gathered presence masks, shape [world, Q] -> [Q]."""
    acc = gathered[0].clone()
    for r in range(1, gathered.shape[0]):
        acc |= gathered[r]
    return acc


def all_gather_fold(mask, world: int, dist):
    """all_gather the per-shard presence masks [Q] and OR them (NCCL has no bitwise-OR reduction)."""
    import torch
    flat = torch.empty(world * mask.numel(), dtype=mask.dtype, device=mask.device)
    dist.all_gather_into_tensor(flat, mask)
    return fold_presence(flat.view(world, mask.numel()))


def merge_offsets(counts):
    """counts[world, Q] (hits of query q on shard r) -> (global offsets [Q+1], write base [world, Q]):
    shard r's hits of query q go to global_offsets[q] + sum_{r' < r} counts[r', q]."""
    import torch
    per_query = counts.sum(dim=0)
    offsets = torch.zeros(counts.shape[1] + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(per_query, 0, out=offsets[1:])
    before = torch.cumsum(counts, 0) - counts
    return offsets, before + offsets[:-1].unsqueeze(0)


def place_shard(final, shard_offsets, shard_positions, write_base):
    """Copy one shard's CSR payload into the merged array: element i of query q goes to
    write_base[q] + (i - shard_offsets[q])."""
    import torch
    total = int(shard_positions.numel())
    if total == 0:
        return
    idx = torch.arange(total, dtype=torch.int64, device=shard_positions.device)
    q = torch.searchsorted(shard_offsets, idx, right=True) - 1
    final[write_base[q] + (idx - shard_offsets[q])] = shard_positions


def merge_to_rank0(offsets, positions, world: int, rank: int, dist):
    """Gather every shard's (offsets[Q+1], positions) on rank 0 and merge; returns (offsets, positions) on
    rank 0 and (None, None) elsewhere. offsets: int64, positions: any 32-bit integer dtype."""
    import torch
    Q = offsets.numel() - 1
    counts = (offsets[1:] - offsets[:-1]).to(torch.int32)
    if rank == 0:
        all_counts = [torch.empty_like(counts) for _ in range(world)]
        dist.gather(counts, all_counts, dst=0)
        cm = torch.stack(all_counts).to(torch.int64)
        g_off, base = merge_offsets(cm)
        final = torch.empty(int(g_off[-1].item()), dtype=positions.dtype, device=positions.device)
        place_shard(final, offsets, positions, base[0])
        for r in range(1, world):
            tot = int(cm[r].sum().item())
            buf = torch.empty(tot, dtype=positions.dtype, device=positions.device)
            if tot:
                dist.recv(buf, src=r)
            off_r = torch.zeros(Q + 1, dtype=torch.int64, device=positions.device)
            torch.cumsum(cm[r], 0, out=off_r[1:])
            place_shard(final, off_r, buf, base[r])
        return g_off, final
    dist.gather(counts, None, dst=0)
    if positions.numel():
        dist.send(positions, dst=0)
    return None, None


# ----------------------------------------------------------------------------------------------------------
# GPU paths used by bench.py
# ----------------------------------------------------------------------------------------------------------
def _global_presence(ix, q_ptr, off_ptr, Q, max_len, world, dev):
    import torch
    import torch.distributed as dist
    present = torch.empty(Q, dtype=torch.int64, device=dev)
    ix.presence_batch_device(q_ptr, off_ptr, Q, max_len, present.data_ptr())
    max_parts = max_len // min(ix.ks) + 1
    narrow = present.to(torch.uint8) if max_parts <= 8 else present   # fewer bytes over NVLink
    return all_gather_fold(narrow, world, dist).to(torch.int64)


def search_device(ix, q_ptr: int, off_ptr: int, Q: int, max_len: int, world: int, dev, count_only: bool = False) -> int:
    """Whole-job search with queries resident in HBM; returns the total number of hits (on rank 0 when sharded)."""
    import torch
    if world == 1:
        fn = ix.count_batch_device if count_only else ix.search_batch_device
        res = fn(q_ptr, off_ptr, Q, max_len)
        hits = res.n_positions
        res.free()
        return hits
    import torch.distributed as dist
    rank = dist.get_rank()
    present = _global_presence(ix, q_ptr, off_ptr, Q, max_len, world, dev)
    res = ix.search_batch_device_global(q_ptr, off_ptr, Q, max_len, present.data_ptr())
    offsets = torch.as_tensor(res.offsets(), device=dev)
    positions = (torch.as_tensor(res.positions(), device=dev) if res.n_positions
                 else torch.empty(0, dtype=torch.int32, device=dev))
    g_off, final = merge_to_rank0(offsets, positions, world, rank, dist)
    hits = int(g_off[-1].item()) if rank == 0 else 0
    torch.cuda.current_stream().synchronize()
    res.free()
    return hits


def search_host(ix, h_q: np.ndarray, h_off: np.ndarray, world: int, dev):
    """Whole-job search through host buffers. world == 1: the plain C-ABI host call. Sharded: H2D on every rank,
    device search + NCCL merge, D2H of the merged result on rank 0."""
    import torch
    if world == 1:
        return ix.search_batch(h_q, h_off, copy=False)
    import torch.distributed as dist
    from . import BatchResult
    rank = dist.get_rank()
    Q = h_off.size - 1
    d_q = torch.from_numpy(h_q).to(dev, non_blocking=True)
    d_off = torch.from_numpy(h_off.view(np.int64)).to(dev, non_blocking=True)
    max_len = int((d_off[1:] - d_off[:-1]).max().item()) if Q else 0
    present = _global_presence(ix, d_q.data_ptr(), d_off.data_ptr(), Q, max_len, world, dev)
    res = ix.search_batch_device_global(d_q.data_ptr(), d_off.data_ptr(), Q, max_len, present.data_ptr())
    offsets = torch.as_tensor(res.offsets(), device=dev)
    positions = (torch.as_tensor(res.positions(), device=dev) if res.n_positions
                 else torch.empty(0, dtype=torch.int32, device=dev))
    status = torch.as_tensor(res.status(), device=dev).clone()
    g_off, final = merge_to_rank0(offsets, positions, world, rank, dist)
    out = None
    if rank == 0:
        out = BatchResult(g_off.cpu().numpy().view(np.uint64), final.cpu().numpy().view(np.uint32), status.cpu().numpy())
    else:
        out = BatchResult(np.zeros(Q + 1, np.uint64), np.zeros(0, np.uint32), status.cpu().numpy())
    torch.cuda.current_stream().synchronize()
    res.free()
    return out
