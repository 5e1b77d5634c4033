"""CPU, world_size 2 over gloo: the cross-rank half of the sharded search (presence OR, count gather, offset
computation, payload placement) merges per-shard sorted hit lists into exactly the whole-text result."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, halo, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from kmer_index_b200 import sharded, synth
    from oracle.bindings import Oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    text = synth.random_text(n, 4, 99)
    text[500:560] = 1                      # a run that straddles the shard boundary region
    q, off = synth.stress_queries(text, 400, 2, halo + 1, 4, 17)
    sh = sharded.shard_range(n, world, rank, halo)
    local = text[sh.begin:sh.begin + sh.length]
    # per-shard result: true occurrences that START in the owned range (what every shard's device search reports)
    l_off, l_pos, _ = Oracle.truth(local, q, off)
    keep = l_pos < (sh.end - sh.begin)
    csum = np.concatenate([[0], np.cumsum(keep.astype(np.int64))])
    counts = csum[l_off[1:].astype(np.int64)] - csum[l_off[:-1].astype(np.int64)]
    s_off = np.zeros(off.size, np.int64)
    np.cumsum(counts, out=s_off[1:])
    s_pos = (l_pos[keep].astype(np.int64) + sh.begin).astype(np.int32)
    # presence: bit 0 = "query occurs in this shard"; OR over ranks must equal "occurs anywhere"
    present = torch.from_numpy((counts > 0).astype(np.uint8))
    anywhere = sharded.all_gather_fold(present, world, dist)
    g_off, g_pos = sharded.merge_to_rank0(torch.from_numpy(s_off), torch.from_numpy(s_pos), world, rank, dist)
    # the same with the hit list the device count pass provides: unordered, as int32, and with some queries
    # listed whose lists are empty (the presence rule can empty a list after the count pass)
    listed = np.flatnonzero(counts > 0)
    extra = np.flatnonzero(counts == 0)[:7]
    hit_ids = torch.from_numpy(np.random.default_rng(rank).permutation(np.concatenate([listed, extra])).astype(np.int32))
    h_off, h_pos = sharded.merge_to_rank0(torch.from_numpy(s_off), torch.from_numpy(s_pos), world, rank, dist, hit_ids=hit_ids)
    if rank == 0:
        assert torch.equal(h_off, g_off) and torch.equal(h_pos, g_pos)
        w_off, w_pos, _ = Oracle.truth(text, q, off)
        ok = (np.array_equal(g_off.numpy().astype(np.uint64), w_off)
              and np.array_equal(g_pos.numpy().astype(np.uint32), w_pos)
              and np.array_equal(anywhere.numpy().astype(bool), (w_off[1:] - w_off[:-1]) > 0)
              and w_pos.size > 0)
        open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,halo", [(2, 1000, 31), (2, 4097, 63), (3, 1500, 15)])
def test_merge_over_gloo(tmp_path, oracle_mod, world, n, halo):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, halo, str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "ok").read() == "1"


def test_merge_offsets_and_place():
    import torch

    from kmer_index_b200 import sharded
    counts = torch.tensor([[1, 0, 2], [0, 3, 1]], dtype=torch.int64)
    g_off, base = sharded.merge_offsets(counts)
    assert g_off.tolist() == [0, 1, 4, 7]
    assert base.tolist() == [[0, 1, 4], [1, 1, 6]]
    final = torch.zeros(7, dtype=torch.int32)
    sharded.place_shard(final, torch.tensor([0, 1, 1, 3]), torch.tensor([10, 30, 31], dtype=torch.int32), base[0])
    sharded.place_shard(final, torch.tensor([0, 0, 3, 4]), torch.tensor([20, 21, 22, 32], dtype=torch.int32), base[1])
    assert final.tolist() == [10, 20, 21, 22, 30, 31, 32]


def test_sparse_merge_helpers():
    """sparse_lists / merge_lists / place_lists without a process group: three shards of a 5-query batch."""
    import torch

    from kmer_index_b200 import sharded
    # per-shard CSRs (positions already global and ascending per query; shards own ascending position ranges)
    off = [torch.tensor([0, 2, 2, 3, 3, 3]), torch.tensor([0, 0, 0, 1, 1, 3]), torch.tensor([0, 1, 1, 1, 1, 2])]
    pos = [torch.tensor([1, 5, 7], dtype=torch.int32), torch.tensor([12, 14, 19], dtype=torch.int32),
           torch.tensor([21, 28], dtype=torch.int32)]
    lists = [sharded.sparse_lists(o) for o in off]
    assert lists[0][0].tolist() == [0, 2] and lists[0][1].tolist() == [2, 1]
    # the hit list of the count pass: unordered, int32, may name queries whose list is empty
    q1, c1 = sharded.sparse_lists(off[1], torch.tensor([4, 1, 2], dtype=torch.int32))
    assert q1.tolist() == [2, 4] and c1.tolist() == [1, 2]
    others = [(lists[r][0], lists[r][1], pos[r]) for r in (1, 2)]
    g_off, final = sharded.merge_lists(off[0], pos[0], lists[0], others)
    assert g_off.tolist() == [0, 3, 3, 5, 5, 8]
    assert final.tolist() == [1, 5, 21, 7, 12, 14, 19, 28]
    # the merged-finish form: offsets already merged, own hits in place, the others placed at offsets[q] + within
    merged = torch.tensor([1, 5, 0, 7, 0, 0, 0, 0], dtype=torch.int32)
    within = [torch.tensor([1, 0]), torch.tensor([2, 2])]   # what add_counts returns when called in rank order
    for (q_r, c_r, p_r), w in zip(others, within):
        sharded.place_lists(merged, g_off[q_r] + w, c_r, p_r)
    assert merged.tolist() == final.tolist()
