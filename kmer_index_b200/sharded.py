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


def _same_stream(ix):
    """The multi-GPU paths interleave library launches (on the index's stream) with torch / NCCL work (on torch's current
    stream) without further synchronisation: both must be the SAME stream. Create the index with
    stream=torch.cuda.current_stream().cuda_stream from inside a `with torch.cuda.stream(s)` / after set_stream(s)."""
    import torch
    cur = torch.cuda.current_stream().cuda_stream
    if getattr(ix, "stream", None) != cur or not cur:
        raise ValueError("multi-GPU search: the index must be created on torch's current (non-default) CUDA stream "
                         f"(index stream {getattr(ix, 'stream', None)}, current stream {cur})")


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


def sparse_lists(offsets, hit_ids=None):
    """(ascending ids of the queries with hits, their counts) of one shard's CSR offsets. `hit_ids` (optional, any
    order, may contain queries whose list turned out empty): the ids as the count pass lists them
    (kmer_b200_result_hit_queries) -- without it they are found by scanning all Q counts."""
    import torch
    if hit_ids is None:
        counts = offsets[1:] - offsets[:-1]
        qids = torch.nonzero(counts).squeeze(1)
        return qids, counts[qids]
    qids = torch.sort(hit_ids.to(torch.int64) & 0xFFFFFFFF).values
    cnts = offsets[qids + 1] - offsets[qids]
    keep = cnts > 0
    return qids[keep], cnts[keep]


def collect_counts_on_rank0(qids, cnts, world: int, rank: int, dist, dev):
    """Phase 1 of the exchange: only queries that have hits on a shard travel, as (query id, count) pairs. Rank 0
    returns [(ids, counts, number of positions)] for ranks 1..world-1, the other ranks return None."""
    import torch
    meta = torch.stack([torch.tensor(qids.numel(), dtype=torch.int64, device=dev), cnts.sum().to(torch.int64)])
    all_meta = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_meta, meta)
    all_meta = all_meta.view(world, 2).cpu()
    if rank != 0:
        if qids.numel():
            dist.send(torch.cat([qids, cnts]), dst=0)
        return None
    lists = []
    for r in range(1, world):
        nq, npos = int(all_meta[r, 0]), int(all_meta[r, 1])
        if nq == 0:
            lists.append((qids[:0], cnts[:0], 0))
            continue
        buf = torch.empty(2 * nq, dtype=torch.int64, device=dev)
        dist.recv(buf, src=r)
        lists.append((buf[:nq], buf[nq:], npos))
    return lists


def collect_positions_on_rank0(positions, lists, world: int, rank: int, dist):
    """Phase 2: the positions behind the pairs of phase 1 (a shard's lists back to back, ascending query id).
    Rank 0 returns [(ids, counts, positions)]."""
    import torch
    if rank != 0:
        if positions.numel():
            dist.send(positions, dst=0)
        return None
    others = []
    for r, (q_r, c_r, npos) in enumerate(lists, start=1):
        pos_r = torch.empty(npos, dtype=positions.dtype, device=positions.device)
        if npos:
            dist.recv(pos_r, src=r)
        others.append((q_r, c_r, pos_r))
    return others


def collect_on_rank0(qids, cnts, positions, world: int, rank: int, dist):
    """Both phases back to back."""
    lists = collect_counts_on_rank0(qids, cnts, world, rank, dist, positions.device)
    return collect_positions_on_rank0(positions, lists, world, rank, dist)


def place_lists(final, starts, cnts, payload):
    """final[starts[i] + j] = j-th element of list i, the lists lying back to back in `payload`."""
    import torch
    if payload.numel() == 0:
        return
    dev = payload.device
    seg_off = torch.cumsum(cnts, 0) - cnts
    seg = torch.repeat_interleave(torch.arange(cnts.numel(), device=dev), cnts)
    idx = torch.arange(payload.numel(), dtype=torch.int64, device=dev)
    final[starts[seg] + (idx - seg_off[seg])] = payload


def merge_lists(offsets, positions, own, others):
    """Rank 0: this shard's full CSR (offsets[Q+1], positions) + the other shards' sparse lists -> merged CSR.
    Shards are disjoint ascending position ranges, so a query's merged list is the concatenation in rank order."""
    import torch
    Q = offsets.numel() - 1
    dev = offsets.device
    # global offsets = this shard's own CSR offsets + the other shards' hits in earlier queries: two passes over Q
    # (a scan and an add); everything else touches only the queries that have hits somewhere
    delta = torch.zeros(Q + 1, dtype=torch.int64, device=dev)
    for q_r, c_r, _ in others:
        if q_r.numel():
            delta.index_add_(0, q_r + 1, c_r)
    g_off = torch.cumsum(delta, 0)
    g_off += offsets
    final = torch.empty(int(g_off[-1].item()), dtype=positions.dtype, device=dev)
    for q_r, _, _ in others:
        if q_r.numel():
            delta[q_r + 1] = 0
    filled = delta                                               # reuse: hits of query q already placed
    for q_r, c_r, p_r in [(own[0], own[1], positions)] + list(others):   # rank order == ascending position order
        if p_r.numel() == 0:
            continue
        start = g_off[q_r] + filled[q_r]
        filled.index_add_(0, q_r, c_r)
        place_lists(final, start, c_r, p_r)
    return g_off, final


def merge_to_rank0(offsets, positions, world: int, rank: int, dist, hit_ids=None):
    """Gather every shard's (offsets[Q+1], positions) on rank 0 and merge; returns (offsets, positions) on
    rank 0 and (None, None) elsewhere. offsets: int64, positions: a 32-bit integer dtype."""
    qids, cnts = sparse_lists(offsets, hit_ids)
    others = collect_on_rank0(qids, cnts, positions, world, rank, dist)
    if rank != 0:
        return None, None
    return merge_lists(offsets, positions, (qids, cnts), others)


# ----------------------------------------------------------------------------------------------------------
# GPU paths used by bench.py
# ----------------------------------------------------------------------------------------------------------
def _global_presence(ix, q_ptr, off_ptr, Q, max_len, world, dev):
    """Per-query flags "part j occurs somewhere in the text", combined over all shards. Returns (tensor, format)
    for KmerIndex.search_batch_device_global."""
    import torch
    import torch.distributed as dist
    max_parts = max(1, max_len // min(ix.ks))
    if max_parts <= 8 and world <= 15:
        # nibble per part, SUM all-reduce acts as OR (NCCL has no bitwise OR); 4 bytes per query over NVLink
        present = torch.empty(Q, dtype=torch.int32, device=dev)
        ix.presence_batch_device(q_ptr, off_ptr, Q, max_len, present.data_ptr(), fmt=1)
        dist.all_reduce(present, op=dist.ReduceOp.SUM)
        return present, 1
    present = torch.empty(Q, dtype=torch.int64, device=dev)
    ix.presence_batch_device(q_ptr, off_ptr, Q, max_len, present.data_ptr(), fmt=0)
    return all_gather_fold(present, world, dist), 0


def _search_shard(ix, q_ptr, off_ptr, Q, max_len, world, dev):
    """This shard's result with the whole-text presence rule applied: count pass -> one all-reduce of the
    presence flags -> epilogue + write pass. Falls back to the two-pass form for queries with > 8 indexed parts."""
    import torch
    import torch.distributed as dist
    max_parts = max(1, max_len // min(ix.ks))
    if max_parts <= 8 and world <= 15:
        present = torch.empty(Q, dtype=torch.int32, device=dev)
        pending = ix.search_sharded_begin(q_ptr, off_ptr, Q, max_len, present.data_ptr())
        dist.all_reduce(present, op=dist.ReduceOp.SUM)   # nibble per part: SUM over <= 15 shards acts as OR
        return ix.search_sharded_finish(pending, present.data_ptr(), Q)
    present, fmt = _global_presence(ix, q_ptr, off_ptr, Q, max_len, world, dev)
    return ix.search_batch_device_global(q_ptr, off_ptr, Q, max_len, present.data_ptr(), fmt=fmt)


def _hit_ids(res, dev):
    """The ids of the queries the shard's count pass found hits for (device tensor), or None."""
    import torch
    listed = res.hit_queries()
    if listed is None:
        return None
    listed = torch.as_tensor(listed, device=dev)
    return listed[1:1 + int(listed[0].item())]


def search_merged(ix, q_ptr: int, off_ptr: int, Q: int, max_len: int, world: int, dev, trace=None):
    """Sharded search of a device-resident batch with the merge on rank 0. Returns (res, offsets, positions):
    `res` is this shard's DeviceResult (keep it alive while the tensors are used, then free() it); offsets /
    positions are the MERGED CSR on rank 0 (None elsewhere).

    Fused form (<= 8 indexed parts per query, <= 15 shards): count pass -> all-reduce of the presence flags ->
    ranks > 0 finish and send their sparse lists -> rank 0 adds their counts to its own BEFORE its offsets scan
    (kmer_b200_search_sharded_add_counts), so its scan yields the merged offsets, its write pass lands its own hits
    in place, and only the other shards' lists are copied. Otherwise (or when the batch needs the segment sort):
    every shard finishes on its own and rank 0 merges the finished CSRs."""
    _same_stream(ix)
    import torch
    import torch.distributed as dist
    from . import KmerB200Error
    rank = dist.get_rank()
    max_parts = max(1, max_len // min(ix.ks))
    fused = max_parts <= 8 and world <= 15

    def finished(res):
        offsets = torch.as_tensor(res.offsets(), device=dev)
        positions = (torch.as_tensor(res.positions(), device=dev) if res.n_positions
                     else torch.empty(0, dtype=torch.int32, device=dev))
        return offsets, positions

    if not fused:
        res = _search_shard(ix, q_ptr, off_ptr, Q, max_len, world, dev)
        if trace:
            trace()
        offsets, positions = finished(res)
        g_off, final = merge_to_rank0(offsets, positions, world, rank, dist, hit_ids=_hit_ids(res, dev))
        return res, g_off, final

    present = torch.empty(Q, dtype=torch.int32, device=dev)
    pending = ix.search_sharded_begin(q_ptr, off_ptr, Q, max_len, present.data_ptr())
    try:
        dist.all_reduce(present, op=dist.ReduceOp.SUM)   # nibble per part: SUM over <= 15 shards acts as OR
        if rank != 0:
            # the sparse counts leave before this shard's own scan and write pass, the positions after them
            counts, listed = ix.search_sharded_peek(pending, present.data_ptr(), Q)
            counts, listed = torch.as_tensor(counts, device=dev), torch.as_tensor(listed, device=dev)
            qids = torch.sort(listed[1:1 + int(listed[0].item())].to(torch.int64) & 0xFFFFFFFF).values
            cnts = counts[qids]
            keep = cnts > 0
            qids, cnts = qids[keep], cnts[keep]
            collect_counts_on_rank0(qids, cnts, world, rank, dist, dev)
    except Exception:
        ix.search_sharded_abort(pending)   # the exchange failed before the finish: do not leak the pending search
        raise
    if rank != 0:
        res = ix.search_sharded_finish(pending, present.data_ptr(), Q)
        if trace:
            trace()
        _offsets, positions = finished(res)
        collect_positions_on_rank0(positions, None, world, rank, dist)
        return res, None, None
    empty64 = torch.empty(0, dtype=torch.int64, device=dev)
    try:
        lists = collect_counts_on_rank0(empty64, empty64, world, rank, dist, dev)
    except Exception:
        ix.search_sharded_abort(pending)
        raise
    within = []
    try:
        for q_r, c_r, _ in lists:                        # rank order
            w = torch.empty(q_r.numel(), dtype=torch.int64, device=dev)
            q_c, c_c = q_r.contiguous(), c_r.contiguous()
            ix.search_sharded_add_counts(pending, present.data_ptr(), q_c.data_ptr(), c_c.data_ptr(), q_c.numel(),
                                         w.data_ptr())
            within.append(w)
    except KmerB200Error as e:
        if e.code != -5 or within:
            ix.search_sharded_abort(pending)
            raise
        within = None   # the batch needs the segment sort over this shard's own lists: merge the finished CSRs
    res = ix.search_sharded_finish(pending, present.data_ptr(), Q)
    if trace:
        trace()
    offsets, positions = finished(res)
    others = collect_positions_on_rank0(positions, lists, world, rank, dist)
    if within is None:
        g_off, final = merge_lists(offsets, positions, sparse_lists(offsets, _hit_ids(res, dev)), others)
        return res, g_off, final
    # offsets and n_positions are the merged ones and this shard's own hits are in place
    for (q_r, c_r, p_r), w in zip(others, within):
        place_lists(positions, offsets[q_r] + w, c_r, p_r)
    return res, offsets, positions


def search_device(ix, q_ptr: int, off_ptr: int, Q: int, max_len: int, world: int, dev, count_only: bool = False) -> int:
    """Whole-job search with queries resident in HBM; returns the total number of hits (on rank 0 when sharded)."""
    import torch
    if world == 1:
        fn = ix.count_batch_device if count_only else ix.search_batch_device
        res = fn(q_ptr, off_ptr, Q, max_len)
        hits = res.n_positions
        res.free()
        return hits
    import os
    import torch.distributed as dist
    rank = dist.get_rank()
    trace = os.environ.get("KMER_B200_TRACE") and rank == 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if trace else None
    if trace:
        ev[0].record()
    res, g_off, _final = search_merged(ix, q_ptr, off_ptr, Q, max_len, world, dev,
                                       trace=(lambda: ev[1].record()) if trace else None)
    hits = int(g_off[-1].item()) if rank == 0 else 0
    if trace:
        ev[2].record()
    torch.cuda.current_stream().synchronize()
    if trace:
        print(f"[trace] shard search {ev[0].elapsed_time(ev[1]):.2f} ms, merge {ev[1].elapsed_time(ev[2]):.2f} ms", flush=True)
    res.free()
    return hits


# ----------------------------------------------------------------------------------------------------------
# replicated index, partitioned build: every rank sorts the k-mers of ONE key-range part (1/world of the work),
# the parts are all-gathered over NVLink into the whole index on every rank, and each rank then answers its own
# slice of the query batch with no per-query exchange at all
# ----------------------------------------------------------------------------------------------------------
def _dev_view(ptr: int, n: int, dev):
    """int32 torch view of n 32-bit words of device memory owned by the library."""
    import torch

    class _V:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_V(), device=dev) if n else torch.empty(0, dtype=torch.int32, device=dev)


def assemble_replicated(ix, world: int, rank: int, dist, dev, timing=None):
    """`ix` was built with key_part=rank, key_parts=world on the torch current stream: all-gather every element's
    part (positions behind the earlier parts' positions, directory entries offset by the earlier parts' k-mer counts)
    and hand the whole arrays to the index (kmer_b200_adopt_element). Afterwards ix is a complete, replicated index.
    timing: a list that receives (label, CUDA event) marks -- the caller turns consecutive marks into milliseconds."""
    _same_stream(ix)
    import torch

    def mark(label):
        if timing is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream())
            timing.append((label, ev))

    mark("start")
    for e, k in enumerate(ix.ks):
        part = ix.element_part(e)
        n_kmers = ix.n - k + 1
        key_space = ix.sigma ** k
        # the part's directory as bucket sizes (one byte per hash): a quarter of the bytes to all-gather, unless some
        # bucket is too large for a byte on some rank
        width_mine = part.key_hi - part.key_lo
        sizes = torch.empty(max(width_mine, 1), dtype=torch.uint8, device=dev)
        n_large = ix.export_bucket_sizes(e, sizes.data_ptr())
        meta = torch.tensor([part.n_kmers, part.key_lo, part.key_hi, n_large], dtype=torch.int64, device=dev)
        all_meta = torch.empty(4 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_meta, meta)
        all_meta = all_meta.view(world, 4).cpu()
        mark("wait_for_parts")   # bucket sizes exported, the slowest rank's part is built
        counts = [int(x) for x in all_meta[:, 0]]
        los = [int(x) for x in all_meta[:, 1]]
        his = [int(x) for x in all_meta[:, 2]]
        any_large = int(all_meta[:, 3].sum()) > 0
        bases = [sum(counts[:r]) for r in range(world)]
        assert sum(counts) == n_kmers, (counts, n_kmers)
        pos_full = torch.empty(n_kmers, dtype=torch.int32, device=dev)
        dir_full = torch.empty(key_space + 1, dtype=torch.int32, device=dev)
        if counts[rank]:
            pos_full[bases[rank]:bases[rank] + counts[rank]].copy_(_dev_view(part.d_positions, counts[rank], dev))
        last = rank == world - 1
        n_dir = his[rank] - los[rank] + (1 if last else 0)
        # exchange. Measured on 8 B200s (25 GB received per rank as 4-byte directory entries + positions): grouped
        # point-to-point sends reach 385 GB/s per rank, NCCL's all-gather (NVSwitch) about twice that -- so the arrays
        # travel as all-gathers of EQUAL slices: directory slices are equal by construction (equal key ranges), position
        # parts are padded to the largest part and compacted afterwards.
        widths = {his[r] - los[r] for r in range(world)}
        equal = len(widths) == 1 and his[-1] == key_space and los[0] == 0
        pad = max(counts)
        staged = pos_mine = pos_work = None
        if pad:
            staged = torch.empty(world * pad, dtype=torch.int32, device=dev)
            pos_mine = torch.empty(pad, dtype=torch.int32, device=dev)
            pos_mine[:counts[rank]].copy_(pos_full[bases[rank]:bases[rank] + counts[rank]])
        if equal and not any_large:
            # both all-gathers are enqueued back to back on NCCL's stream; the prefix sum over the sizes (this stream)
            # runs while the positions are still arriving
            sizes_full = torch.empty(key_space, dtype=torch.uint8, device=dev)
            sizes_work = dist.all_gather_into_tensor(sizes_full, sizes, async_op=True)
            if pad:
                pos_work = dist.all_gather_into_tensor(staged, pos_mine, async_op=True)
            sizes_work.wait()
            mark("gather_sizes")
            ix.directory_from_sizes(sizes_full.data_ptr(), key_space, dir_full.data_ptr())   # prefix sum on every rank
            mark("directory_prefix_sum")
            del sizes_full
        elif equal:
            ix.export_directory(e, bases[rank], n_dir, dir_full.data_ptr() + 4 * los[rank])
            width = his[0] - los[0]
            mine = dir_full[los[rank]:los[rank] + width].clone()
            dist.all_gather_into_tensor(dir_full[:key_space], mine)
            dir_full[key_space:].fill_(n_kmers - (1 << 32) if n_kmers >= (1 << 31) else n_kmers)   # uint32 bit pattern
        else:
            ix.export_directory(e, bases[rank], n_dir, dir_full.data_ptr() + 4 * los[rank])
            ops = []
            for r in range(world):
                if r == rank:
                    continue
                hi_r = his[r] + (1 if r == world - 1 else 0)
                if n_dir:
                    ops.append(dist.P2POp(dist.isend, dir_full[los[rank]:los[rank] + n_dir], r))
                if hi_r > los[r]:
                    ops.append(dist.P2POp(dist.irecv, dir_full[los[r]:hi_r], r))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        del sizes
        if pad:
            if pos_work is not None:
                pos_work.wait()
            else:
                dist.all_gather_into_tensor(staged, pos_mine)
            mark("gather_positions")
            for r in range(world):
                if r != rank and counts[r]:
                    pos_full[bases[r]:bases[r] + counts[r]].copy_(staged[r * pad:r * pad + counts[r]])
            del staged, pos_mine
        ix.adopt_element(e, pos_full, dir_full)
        mark("compact_and_adopt")


# ----------------------------------------------------------------------------------------------------------
# peer positions: the directory is replicated (it travels as one byte per bucket), the position array stays where it was
# sorted -- every rank maps the other ranks' parts into its address space (CUDA IPC) and the search reads the candidates
# of a foreign bucket over NVLink. The build moves 1/4 of the bytes of the replicated form; a query still touches no
# collective.
# ----------------------------------------------------------------------------------------------------------
class PeerPositions:
    """Per-element position buffers of all ranks, mapped into every rank (kmer_b200_peer_buffer_*: plain device
    allocations exported and opened through CUDA IPC). Created once (like a communicator) and reused by every build; a
    buffer is re-made -- collectively -- when some rank's part outgrows it."""

    def __init__(self, world: int, rank: int, dist, dev):
        self.world, self.rank, self.dist, self.dev = world, rank, dist, dev
        self.ptrs = {}      # element -> [device pointer per rank]; ptrs[e][rank] is this rank's own buffer
        self.capacity = {}  # element -> entries per buffer

    def _release(self, e: int):
        import torch
        from . import peer_buffer_release
        if e not in self.ptrs:
            return
        torch.cuda.synchronize()
        self.dist.barrier()          # nobody reads a buffer that is about to go away
        for r, p in enumerate(self.ptrs[e]):
            if r != self.rank:
                peer_buffer_release(self.dev.index, p, True)
        self.dist.barrier()          # every mapping is closed before the owner frees
        peer_buffer_release(self.dev.index, self.ptrs[e][self.rank], False)
        del self.ptrs[e], self.capacity[e]

    def ensure(self, e: int, need: int):
        """Collective: every rank passes the same `need` (the largest part over all ranks)."""
        import torch
        from . import peer_buffer_create, peer_buffer_open
        if self.capacity.get(e, 0) >= need:
            return
        self._release(e)
        cap = int(need * 1.02) + 1024
        mine, handle = peer_buffer_create(self.dev.index, cap * 4)
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle)
        self.ptrs[e] = [mine if r == self.rank else peer_buffer_open(self.dev.index, handles[r]) for r in range(self.world)]
        self.capacity[e] = cap
        torch.cuda.synchronize()
        self.dist.barrier()

    def close(self):
        for e in list(self.ptrs):
            self._release(e)


def assemble_peer(ix, world: int, rank: int, dist, dev, peers: PeerPositions, timing=None):
    """`ix` was built with key_part=rank, key_parts=world on the torch current stream. Every rank copies its part's
    positions into its shared buffer, the directory is all-gathered as bucket sizes and prefix-summed on every rank,
    and the index adopts (own part, the other ranks' mapped parts, whole directory). Needs dense directories whose
    buckets fit a byte on every rank or ships 4-byte directory entries instead."""
    _same_stream(ix)
    import torch

    def mark(label):
        if timing is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream())
            timing.append((label, ev))

    mark("start")
    for e, k in enumerate(ix.ks):
        part = ix.element_part(e)
        n_kmers = ix.n - k + 1
        key_space = ix.sigma ** k
        width_mine = part.key_hi - part.key_lo
        sizes = torch.empty(max(width_mine, 1), dtype=torch.uint8, device=dev)
        n_large = ix.export_bucket_sizes(e, sizes.data_ptr())
        meta = torch.tensor([part.n_kmers, part.key_lo, part.key_hi, n_large], dtype=torch.int64, device=dev)
        all_meta = torch.empty(4 * world, dtype=torch.int64, device=dev)
        # also the fence between the previous index and this one: no rank overwrites its shared buffer before every
        # rank's stream has passed its last search
        dist.all_gather_into_tensor(all_meta, meta)
        all_meta = all_meta.view(world, 4).cpu()
        mark("wait_for_parts")
        counts = [int(x) for x in all_meta[:, 0]]
        los = [int(x) for x in all_meta[:, 1]]
        his = [int(x) for x in all_meta[:, 2]]
        any_large = int(all_meta[:, 3].sum()) > 0
        bases = [sum(counts[:r]) for r in range(world)]
        assert sum(counts) == n_kmers, (counts, n_kmers)
        peers.ensure(e, max(counts))
        if counts[rank]:
            _dev_view(peers.ptrs[e][rank], counts[rank], dev).copy_(_dev_view(part.d_positions, counts[rank], dev))
        dir_full = torch.empty(key_space + 1, dtype=torch.int32, device=dev)
        widths = {his[r] - los[r] for r in range(world)}
        equal = len(widths) == 1 and his[-1] == key_space and los[0] == 0
        if equal and not any_large:
            sizes_full = torch.empty(key_space, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(sizes_full, sizes)
            mark("gather_sizes")
            ix.directory_from_sizes(sizes_full.data_ptr(), key_space, dir_full.data_ptr())
            mark("directory_prefix_sum")
            del sizes_full
        else:
            last = rank == world - 1
            n_dir = his[rank] - los[rank] + (1 if last else 0)
            ix.export_directory(e, bases[rank], n_dir, dir_full.data_ptr() + 4 * los[rank])
            ops = []
            for r in range(world):
                if r == rank:
                    continue
                hi_r = his[r] + (1 if r == world - 1 else 0)
                if n_dir:
                    ops.append(dist.P2POp(dist.isend, dir_full[los[rank]:los[rank] + n_dir], r))
                if hi_r > los[r]:
                    ops.append(dist.P2POp(dist.irecv, dir_full[los[r]:hi_r], r))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            mark("gather_directory")
        del sizes
        # every rank's copy into its shared buffer precedes its contribution to the directory exchange in stream
        # order, so once the exchange has completed here the other ranks' parts are in place
        ix.adopt_element_parts(e, [int(p) for p in peers.ptrs[e]], bases + [n_kmers], dir_full)
        mark("adopt")


# ----------------------------------------------------------------------------------------------------------
# partitioned index, routed queries: rank r keeps index part r (the k-mers whose hash lies in the r-th slice of the key
# space) and the whole packed text; a query is answered by the owner of the hash of its first k symbols. Three
# all-to-all exchanges move queries out and results back; nothing is replicated except a presence bitmap.
# ----------------------------------------------------------------------------------------------------------
def share_presence(ix, world: int, rank: int, dist, dev):
    """Every rank exports the presence bits of its key-range part; the parts are summed (disjoint bits: a sum is an OR)
    into the whole bitmap, which is attached to the index. Once per index."""
    _same_stream(ix)
    import torch
    for e in range(len(ix.ks)):
        bits = torch.zeros(ix.presence_words(e), dtype=torch.int64, device=dev)
        ix.presence_export(e, bits.data_ptr())
        dist.all_reduce(bits, op=dist.ReduceOp.SUM)   # the parts' words are disjoint (part boundaries are multiples of 64 hashes)
        ix.presence_attach(e, bits)


def search_routed(ix, q_ptr: int, off_ptr: int, Q: int, max_len: int, plan_queries: int, world: int, rank: int, dist, dev,
                  timing=None):
    """Routed search of this rank's slice of the batch (Q queries, device-resident) on a partitioned index prepared with
    share_presence(). plan_queries: the largest slice over all ranks (every rank must pass the same value). Returns the
    DeviceResult of the slice (offsets / positions / status in batch order)."""
    _same_stream(ix)
    import torch
    slack = 0.0
    while True:
        plan = ix.route_plan(plan_queries, max_len, world, slack)
        send = torch.empty(world * plan.block_bytes, dtype=torch.uint8, device=dev)
        status = torch.empty(max(Q, 1), dtype=torch.uint8, device=dev)
        sent = ix.route_queries(q_ptr, off_ptr, Q, plan, send.data_ptr(), status.data_ptr())
        # a send block that overflowed anywhere makes every rank re-plan with room for the largest block seen
        need = torch.tensor([max(sent)], dtype=torch.int64, device=dev)
        dist.all_reduce(need, op=dist.ReduceOp.MAX)
        need = int(need.item())
        if need <= plan.capacity:
            break
        slack = 1.05 * need * world / max(plan_queries, 1) + 0.05
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    ret_send = torch.empty(world * plan.return_block_bytes, dtype=torch.uint8, device=dev)
    ret_recv = torch.empty_like(ret_send)
    res, pos_splits = ix.search_routed(recv.data_ptr(), plan, ret_send.data_ptr())
    dist.all_to_all_single(ret_recv, ret_send)
    recv_splits = [int(x) for x in ret_recv.view(world, plan.return_block_bytes)[:, 8:16].contiguous().view(torch.int64).flatten().cpu()]
    pos_send = (torch.as_tensor(res.positions(), device=dev) if res.n_positions else torch.empty(0, dtype=torch.int32, device=dev))
    pos_recv = torch.empty(sum(recv_splits), dtype=torch.int32, device=dev)
    dist.all_to_all_single(pos_recv, pos_send, output_split_sizes=recv_splits, input_split_sizes=pos_splits)
    out = ix.unroute(send.data_ptr(), ret_recv.data_ptr(), plan, sent, pos_recv.data_ptr(), recv_splits, Q, status.data_ptr())
    torch.cuda.current_stream().synchronize()
    res.free()
    return out


def search_routed_host(ix, h_q, h_off, plan_queries: int, max_len: int, world: int, rank: int, dist, dev):
    """Routed search of this rank's slice given as pinned host tensors (uint8 ranks, int64 offsets): H2D of the slice,
    routed search, D2H of the slice's result into pinned memory."""
    import torch
    from . import BatchResult
    Q = h_off.numel() - 1
    d_q = h_q.to(dev, non_blocking=True)
    d_off = h_off.to(dev, non_blocking=True)
    res = search_routed(ix, d_q.data_ptr(), d_off.data_ptr(), Q, max_len, plan_queries, world, rank, dist, dev)
    h_goff = _pinned("r_offsets", Q + 1, torch.int64)
    h_status = _pinned("r_status", Q, torch.uint8)
    h_pos = _pinned("r_positions", res.n_positions, torch.int32)
    h_goff.copy_(torch.as_tensor(res.offsets(), device=dev), non_blocking=True)
    h_status.copy_(torch.as_tensor(res.status(), device=dev), non_blocking=True)
    if res.n_positions:
        h_pos.copy_(torch.as_tensor(res.positions(), device=dev), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    res.free()
    return BatchResult(h_goff.numpy().view(np.uint64), h_pos.numpy().view(np.uint32), h_status.numpy())


def fingerprint_of(res, Q: int, dev, first_query_id: int = 0):
    """fingerprint() of a finished DeviceResult (frees it)."""
    import torch
    off = torch.as_tensor(res.offsets(), device=dev)
    pos = (torch.as_tensor(res.positions(), device=dev).to(torch.int64) & 0xFFFFFFFF if res.n_positions
           else torch.empty(0, dtype=torch.int64, device=dev))
    status = torch.as_tensor(res.status(), device=dev)
    counts = off[1:] - off[:-1]
    qid = torch.repeat_interleave(torch.arange(first_query_id + 1, first_query_id + Q + 1, device=dev), counts)
    checksum = int((pos * qid).sum().item()) if pos.numel() else 0
    hist = torch.bincount(status.to(torch.int64), minlength=4)[:4]
    out = {"hits": int(off[-1].item()), "checksum": checksum, "status_hist": [int(x) for x in hist.cpu()]}
    torch.cuda.current_stream().synchronize()
    res.free()
    return out


def fingerprint(ix, q_ptr: int, off_ptr: int, Q: int, max_len: int, world: int, dev, first_query_id: int = 0):
    """{hits, status histogram, position checksum} of a device-resident batch: sum over all hits of
    position * (global query id + 1) modulo 2^64 -- a property of the result alone, so every N and every multi-GPU
    mode must print the same values. Position-range sharding: computed on rank 0 from the merged CSR (other ranks
    return zeros). Verification code outside any timed region (torch ops are fine here)."""
    import torch
    if world == 1:
        res = ix.search_batch_device(q_ptr, off_ptr, Q, max_len)
        off = torch.as_tensor(res.offsets(), device=dev)
        pos = (torch.as_tensor(res.positions(), device=dev).to(torch.int64) & 0xFFFFFFFF if res.n_positions
               else torch.empty(0, dtype=torch.int64, device=dev))
        status = torch.as_tensor(res.status(), device=dev)
    else:
        import torch.distributed as dist
        res, off, pos = search_merged(ix, q_ptr, off_ptr, Q, max_len, world, dev)
        status = torch.as_tensor(res.status(), device=dev)
        if dist.get_rank() != 0:
            torch.cuda.current_stream().synchronize()
            res.free()
            return {"hits": 0, "checksum": 0, "status_hist": [0, 0, 0, 0]}
        pos = pos.to(torch.int64) & 0xFFFFFFFF
    counts = off[1:] - off[:-1]
    qid = torch.repeat_interleave(torch.arange(first_query_id + 1, first_query_id + Q + 1, device=dev), counts)
    checksum = int((pos * qid).sum().item()) if pos.numel() else 0
    hist = torch.bincount(status.to(torch.int64), minlength=4)[:4]
    out = {"hits": int(off[-1].item()), "checksum": checksum, "status_hist": [int(x) for x in hist.cpu()]}
    torch.cuda.current_stream().synchronize()
    res.free()
    return out


_pinned_cache: dict = {}


def _pinned(name: str, n: int, dtype):
    """Grow-only pinned host buffers for the merged result (pinned allocation is slow; reuse across calls)."""
    import torch
    buf = _pinned_cache.get(name)
    if buf is None or buf.numel() < n or buf.dtype != dtype:
        buf = torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
        _pinned_cache[name] = buf
    return buf[:n]


def _upload_sliced(h, world: int, rank: int, dev, dist):
    """H2D of this rank's 1/world slice of a pinned host tensor (identical on all ranks) + NCCL all-gather."""
    import torch
    n = h.numel()
    per = -(-n // world)
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    mine = torch.zeros(per, dtype=h.dtype, device=dev) if hi - lo < per else torch.empty(per, dtype=h.dtype, device=dev)
    mine[:hi - lo].copy_(h[lo:hi], non_blocking=True)
    full = torch.empty(per * world, dtype=h.dtype, device=dev)
    dist.all_gather_into_tensor(full, mine)
    return full[:n]


def search_host(ix, h_q, h_off, world: int, dev):
    """Whole-job search through host buffers. world == 1: the plain C-ABI host call (numpy arrays). Sharded:
    h_q (uint8) / h_off (int64) are pinned torch tensors; H2D on every rank, device search + NCCL merge, D2H of
    the merged result into pinned memory on rank 0."""
    import torch
    if world == 1:
        return ix.search_batch(h_q, h_off, copy=False)
    import torch.distributed as dist
    from . import BatchResult
    rank = dist.get_rank()
    Q = h_off.numel() - 1
    # "queries are broadcast": every rank uploads 1/world of the batch over its own PCIe link and the slices
    # are all-gathered over NVLink (8 links x ~50 GB/s into the box instead of one; 8 ranks pulling the same
    # 5 GB each would be slower still)
    d_q = _upload_sliced(h_q, world, rank, dev, dist)
    d_off = _upload_sliced(h_off, world, rank, dev, dist)
    max_len = int((d_off[1:] - d_off[:-1]).max().item()) if Q else 0
    res, g_off, final = search_merged(ix, d_q.data_ptr(), d_off.data_ptr(), Q, max_len, world, dev)
    status = torch.as_tensor(res.status(), device=dev)
    h_status = _pinned("status", Q, torch.uint8)
    h_status.copy_(status, non_blocking=True)
    if rank == 0:
        h_goff = _pinned("offsets", Q + 1, torch.int64)
        h_pos = _pinned("positions", final.numel(), torch.int32)
        h_goff.copy_(g_off, non_blocking=True)
        h_pos.copy_(final, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = BatchResult(h_goff.numpy().view(np.uint64), h_pos.numpy().view(np.uint32), h_status.numpy())
    else:
        torch.cuda.current_stream().synchronize()
        out = BatchResult(np.zeros(Q + 1, np.uint64), np.zeros(0, np.uint32), h_status.numpy())
    res.free()
    return out
