"""Worker of tests/test_gpu_multi.py: run under torchrun with one process per GPU (NCCL). Both multi-GPU modes are
compared bit for bit with the CPU oracle on rank 0:
  * position-range shards (+ halo), every rank searches the whole batch, presence flags all-reduced, merged CSR
    assembled on rank 0 (kmer_index_b200.sharded.search_merged) -- the real send/recv + add_counts + place path;
  * peer positions: the same parts, only the directory replicated, positions read from the owners over NVLink
  * replicated index: every rank sorts one key-range part, parts all-gathered (sharded.assemble_replicated), every
    rank answers its slice of the batch;
  * partitioned index, routed queries (sharded.search_routed): every rank keeps its key-range part, queries travel to
    the owner of their first k-mer and results travel back.
Prints one line 'MULTI_GPU_PARITY_OK <world>' from rank 0 on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import kmer_index_b200 as kb
    from conftest import assert_results_equal
    from kmer_index_b200 import sharded, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    peer_buffers = sharded.PeerPositions(world, rank, dist, dev)
    cases = [(4, [16], 2_000_000, 16, 64, False), (4, [12], 600_000, 13, 64, True), (4, [5, 7, 9, 11, 13], 300_000, 4, 40, False)]
    for sigma, ks, n, m_lo, m_hi, heavy in cases:
        text = synth.random_text(n, sigma, 31)
        q, off = synth.stress_queries(text, 20_000, m_lo, m_hi, sigma, 32)
        if heavy:   # a long constant run: candidate lists far beyond 2048 entries (the warp-per-query launch)
            text[50_000:120_000] = 0
            lens = (off[1:] - off[:-1]).astype(np.int64)
            for i in range(0, 900, 3):
                q[int(off[i]):int(off[i]) + int(lens[i])] = 0
        Q = off.size - 1
        want = None
        if rank == 0:
            from oracle import bindings
            bindings.build()
            with bindings.Oracle(text, sigma, ks) as o:
                want = o.search(q, off)
            assert want[1].size > 0
        d_q = torch.from_numpy(q).to(dev)
        d_off = torch.from_numpy(off.view(np.int64)).to(dev)
        torch.cuda.synchronize()

        # ---- position-range shards, merge on rank 0 (two batches: the second runs with auxiliary elements in place)
        sh = sharded.shard_range(n, world, rank, halo=max(m_hi, max(ks)) - 1)
        ix = kb.KmerIndex(text[sh.begin:sh.begin + sh.length], sigma, ks, shard_begin=sh.begin, n_total=n, halo=sh.halo,
                          stream=sptr, device=local)
        for attempt in range(2):
            res, g_off, final = sharded.search_merged(ix, d_q.data_ptr(), d_off.data_ptr(), Q, m_hi, world, dev)
            torch.cuda.synchronize()
            if rank == 0:
                got = (g_off.cpu().numpy().astype(np.uint64), final.cpu().numpy().view(np.uint32),
                       torch.as_tensor(res.status(), device=dev).cpu().numpy())
                assert_results_equal(got, want, label=f"position-range x{world} {ks} attempt {attempt}")
            res.free()
        fp = sharded.fingerprint(ix, d_q.data_ptr(), d_off.data_ptr(), Q, m_hi, world, dev)
        ix.close()

        # ---- replicated index from key-range parts; every rank answers its slice
        ix = kb.KmerIndex(text, sigma, ks, stream=sptr, device=local, key_part=rank, key_parts=world)
        sharded.assemble_replicated(ix, world, rank, dist, dev)
        lo, hi = rank * Q // world, (rank + 1) * Q // world
        mine = ix.search_batch(q[int(off[lo]):int(off[hi])], off[lo:hi + 1] - off[lo]).as_tuple()
        fp_r = sharded.fingerprint(ix, d_q.data_ptr() + int(off[lo]), (d_off[lo:hi + 1] - d_off[lo]).contiguous().data_ptr(),
                                   hi - lo, m_hi, 1, dev, lo)
        ix.close()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        t = torch.tensor([fp_r["hits"], fp_r["checksum"]] + fp_r["status_hist"], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        if rank == 0:
            g_off = np.concatenate([[0]] + [np.asarray(p[0][1:], dtype=np.uint64) + np.uint64(sum(int(x[0][-1]) for x in parts[:i]))
                                            for i, p in enumerate(parts)]).astype(np.uint64)
            got = (g_off, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]))
            assert_results_equal(got, want, label=f"replicated x{world} {ks}")
            # the fingerprints bench.py prints are equal across the modes (and equal to the oracle's)
            counts = (want[0][1:] - want[0][:-1]).astype(np.int64)
            qid = np.repeat(np.arange(1, Q + 1, dtype=np.uint64), counts)
            w_sum = int((want[1].astype(np.uint64) * qid).sum(dtype=np.uint64))
            w_sum = w_sum - (1 << 64) if w_sum >= (1 << 63) else w_sum
            assert fp["hits"] == int(want[0][-1]) == int(t[0]) and fp["checksum"] == w_sum == int(t[1]), (fp, w_sum, t)
            assert fp["status_hist"] == [int(x) for x in t[2:6]] == [int((want[2] == s).sum()) for s in range(4)]
        # ---- peer positions: whole directory everywhere, the position array left in the ranks' parts and read over
        # NVLink; built twice on the same shared buffers (the second build must not disturb a search still running)
        for attempt in range(2):
            ix = kb.KmerIndex(text, sigma, ks, stream=sptr, device=local, key_part=rank, key_parts=world)
            sharded.assemble_peer(ix, world, rank, dist, dev, peer_buffers)
            mine = ix.search_batch(q[int(off[lo]):int(off[hi])], off[lo:hi + 1] - off[lo]).as_tuple()
            fp_p = sharded.fingerprint(ix, d_q.data_ptr() + int(off[lo]), (d_off[lo:hi + 1] - d_off[lo]).contiguous().data_ptr(),
                                       hi - lo, m_hi, 1, dev, lo)
            ix.close()
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            tp = torch.tensor([fp_p["hits"], fp_p["checksum"]], dtype=torch.int64, device=dev)
            dist.all_reduce(tp)
            if rank == 0:
                g_off = np.concatenate([[0]] + [np.asarray(p[0][1:], dtype=np.uint64) + np.uint64(sum(int(x[0][-1]) for x in parts[:i]))
                                                for i, p in enumerate(parts)]).astype(np.uint64)
                got = (g_off, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]))
                assert_results_equal(got, want, label=f"peer positions x{world} {ks} attempt {attempt}")
                assert int(tp[0]) == fp["hits"] and int(tp[1]) == fp["checksum"]
        # ---- partitioned index, routed queries (single-k indices): three all-to-all exchanges, nothing replicated
        if len(ks) == 1 and m_lo >= ks[0]:
            ix = kb.KmerIndex(text, sigma, ks, stream=sptr, device=local, key_part=rank, key_parts=world)
            sharded.share_presence(ix, world, rank, dist, dev)
            sl_off = (d_off[lo:hi + 1] - d_off[lo]).contiguous()
            for attempt in range(2):
                out = sharded.search_routed(ix, d_q.data_ptr() + int(off[lo]), sl_off.data_ptr(), hi - lo, m_hi, -(-Q // world),
                                            world, rank, dist, dev)
                mine = (torch.as_tensor(out.offsets(), device=dev).cpu().numpy().astype(np.uint64),
                        torch.as_tensor(out.positions(), device=dev).cpu().numpy().view(np.uint32) if out.n_positions
                        else np.zeros(0, np.uint32), torch.as_tensor(out.status(), device=dev).cpu().numpy())
                out.free()
                parts = [None] * world
                dist.all_gather_object(parts, mine)
                if rank == 0:
                    g_off = np.concatenate([[0]] + [np.asarray(p[0][1:], dtype=np.uint64) + np.uint64(sum(int(x[0][-1]) for x in parts[:i]))
                                                    for i, p in enumerate(parts)]).astype(np.uint64)
                    got = (g_off, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]))
                    assert_results_equal(got, want, label=f"routed x{world} {ks} attempt {attempt}")
            ix.close()
        dist.barrier()
    peer_buffers.close()
    if rank == 0:
        print(f"MULTI_GPU_PARITY_OK {world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
