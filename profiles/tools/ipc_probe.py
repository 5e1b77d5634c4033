#!/usr/bin/env python
"""Can one torchrun rank map another rank's device buffer (CUDA IPC through torch's storage sharing) and read it from
a kernel? Prints the per-rank result and the P2P read bandwidth. Used to decide on the peer-positions multi-GPU mode."""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 28
    mine = torch.full((n,), rank + 1, dtype=torch.int32, device=dev)
    handle = mine.untyped_storage()._share_cuda_()
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    peers = []
    for r in range(world):
        if r == rank:
            peers.append(mine)
            continue
        st = torch.UntypedStorage._new_shared_cuda(*handles[r])
        peers.append(torch.empty(0, dtype=torch.int32, device=dev).set_(st, 0, (n,)))
    torch.cuda.synchronize()
    dist.barrier()
    ok = all(int(peers[r][12345].item()) == r + 1 and int(peers[r][-1].item()) == r + 1 for r in range(world))
    other = peers[(rank + 1) % world]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        s = other.sum()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    idx = torch.randint(0, n, (1 << 24,), device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        g = other[idx]
    torch.cuda.synchronize()
    dg = (time.perf_counter() - t0) / 5
    print(f"rank {rank}: peers readable {ok}; streaming read of a peer {n * 4 / dt / 1e9:.0f} GB/s; "
          f"random 4-byte reads from a peer {idx.numel() / dg / 1e9:.2f} G/s (ptr {other.data_ptr():#x})", flush=True)
    dist.barrier()
    del peers, other
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
